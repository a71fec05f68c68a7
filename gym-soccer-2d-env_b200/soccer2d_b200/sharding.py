"""Episode sharding over ranks and the one collective of the path.

Episodes are independent, so the step has no exchange: rank r of R holds a contiguous range of GLOBAL env ids
and the RNG is keyed on the global id, which makes a sharded run reproduce the single-process episodes exactly.
The only collective is a sum of the episode statistics (what utils/info_collector_callback.py:37-53 tallies in
the reference), issued outside the step loop.
"""
from __future__ import annotations

import torch

STAT_KEYS = ("episodes", "goals", "outs", "timeouts", "episode_steps", "env_steps", "return_sum")


def shard_range(rank: int, world: int, total_envs: int) -> tuple[int, int]:
    """(env_id_offset, num_envs) of `rank`: contiguous, sizes differ by at most one, covers [0, total_envs)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(total_envs), int(world))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def allreduce_stats(stats: dict, device=None, group=None) -> dict:
    """Sum of per-rank statistics dicts over the process group.  Integer counters travel as int64 (exact),
    the return sum as float64.  Without an initialised process group the input is returned unchanged.
    With the NCCL backend the tensors live on `device`; with gloo on the CPU."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return dict(stats)
    on_gpu = dist.get_backend(group) == "nccl"
    dev = device if on_gpu else "cpu"
    int_keys = [k for k in STAT_KEYS if k != "return_sum"]
    counts = torch.tensor([int(stats[k]) for k in int_keys], dtype=torch.int64, device=dev)
    ret = torch.tensor([float(stats["return_sum"])], dtype=torch.float64, device=dev)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(ret, op=dist.ReduceOp.SUM, group=group)
    out = {k: int(v) for k, v in zip(int_keys, counts.cpu().tolist())}
    out["return_sum"] = float(ret.item())
    return out


class StatsFuture:
    """Result of an asynchronous statistics all-reduce: `.result()` waits (for the side stream / the collective) and
    returns the summed dict; until then neither the host nor the stepping stream is held up."""

    def __init__(self, counts, ret, works=(), event=None):
        self._counts, self._ret, self._works, self._event, self._out = counts, ret, list(works), event, None

    def done(self) -> bool:
        if self._out is not None:
            return True
        if any(not w.is_completed() for w in self._works):
            return False
        return self._event is None or self._event.query()

    def result(self) -> dict:
        if self._out is None:
            for w in self._works:
                w.wait()
            if self._event is not None:
                self._event.synchronize()
            int_keys = [k for k in STAT_KEYS if k != "return_sum"]
            self._out = {k: int(v) for k, v in zip(int_keys, self._counts.cpu().tolist())}
            self._out["return_sum"] = float(self._ret.cpu().item())
        return self._out


def allreduce_stats_async(counts: torch.Tensor, ret: torch.Tensor, group=None, stream=None) -> StatsFuture:
    """Sum over the ranks of `counts` (int64 [6]: episodes, goals, outs, timeouts, episode_steps, env_steps) and `ret`
    (float64 [1]: return sum), in place, without blocking: with NCCL the collective is enqueued behind `stream` (a side
    stream: the stepping stream is not involved) and the call returns at once; with gloo it runs on a background thread."""
    import torch.distributed as dist

    works, event = [], None
    if dist.is_available() and dist.is_initialized():
        if counts.is_cuda and stream is not None:
            with torch.cuda.stream(stream):
                works = [dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group, async_op=True),
                         dist.all_reduce(ret, op=dist.ReduceOp.SUM, group=group, async_op=True)]
        else:
            works = [dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group, async_op=True),
                     dist.all_reduce(ret, op=dist.ReduceOp.SUM, group=group, async_op=True)]
    if counts.is_cuda:
        event = torch.cuda.Event()
        event.record(stream if stream is not None else torch.cuda.current_stream(counts.device))
    return StatsFuture(counts, ret, works, event)
