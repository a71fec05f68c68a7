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
