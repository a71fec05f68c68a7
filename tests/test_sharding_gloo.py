"""CPU-only, world_size 2 over gloo: the N>1 host logic - shard ranges, shard-invariance of the episodes (with the
C oracle standing in for the kernels), and the statistics all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H


def test_shard_range_covers_everything():
    from soccer2d_b200 import shard_range
    for total, world in [(10, 1), (10, 2), (10, 3), (1 << 20, 8), (7, 8), (4_000_000, 8)]:
        spans = [shard_range(r, world, total) for r in range(world)]
        assert spans[0][0] == 0 and sum(n for _, n in spans) == total
        for (a, n), (b, _) in zip(spans[:-1], spans[1:]):
            assert a + n == b
        assert max(n for _, n in spans) - min(n for _, n in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(2, 2, 10)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, k, launches, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as OL
    from soccer2d_b200 import _abi, allreduce_stats, shard_range
    off, n = shard_range(rank, world, total)
    cfg = H.make_config(n, "discrete", seed=17, env_id_offset=off, change_ball_velocity=1, max_steps=25)
    sim = OL.OracleSim(cfg, "f64")
    sim.reset()
    rng = np.random.default_rng(123)  # every rank draws the GLOBAL action block and takes its slice
    for _ in range(launches):
        act = H.random_actions(rng, "discrete", total, k)
        sim.step(np.ascontiguousarray(act[off:off + n]), k)
    st = sim.stats(_abi.Stats())
    local = {key: getattr(st, key) for key in ("episodes", "goals", "outs", "timeouts", "episode_steps", "env_steps", "return_sum")}
    total_stats = allreduce_stats(local)
    # the asynchronous form (what a training loop uses between launches): same sum, returned as a future
    import torch
    from soccer2d_b200 import allreduce_stats_async
    keys = ("episodes", "goals", "outs", "timeouts", "episode_steps", "env_steps")
    fut = allreduce_stats_async(torch.tensor([int(local[k_]) for k_ in keys], dtype=torch.int64),
                                torch.tensor([float(local["return_sum"])], dtype=torch.float64))
    assert fut.result() == total_stats and fut.done()
    q.put((rank, off, n, sim.obs.copy(), local, total_stats))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_one_and_allreduce_sums():
    import oracle_lib as OL
    from soccer2d_b200 import _abi
    total, k, launches, world = 301, 4, 30, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, k, launches, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the single-process run
    cfg = H.make_config(total, "discrete", seed=17, change_ball_velocity=1, max_steps=25)
    sim = OL.OracleSim(cfg, "f64")
    sim.reset()
    rng = np.random.default_rng(123)
    for _ in range(launches):
        sim.step(H.random_actions(rng, "discrete", total, k), k)
    st = sim.stats(_abi.Stats())
    assert np.array_equal(np.concatenate([g[3] for g in got]), sim.obs)
    assert [g[1:3] for g in got] == [(0, 151), (151, 150)]
    for g in got:
        red = g[5]
        assert red == got[0][5]  # every rank holds the same reduced result
        for key in ("episodes", "goals", "outs", "timeouts", "episode_steps", "env_steps"):
            assert red[key] == getattr(st, key) == sum(x[4][key] for x in got)
        assert red["return_sum"] == pytest.approx(st.return_sum, rel=1e-12)
    assert got[0][5]["episodes"] > 0
