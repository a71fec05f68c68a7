"""The host-buffer path of the C ABI (s2d_step_host, s2d_bind_pipeline_slot / s2d_submit_host / s2d_wait_host): packed
output block (one device-to-host copy per step), three slots in flight, and the ordering between the pipeline's
internal streams and work the caller enqueues on its own stream.  The reference's counterpart of all of this is the
pair of blocking `Queue.get()` calls in Soccer2DEnv.step (soccer_2d_env.py:245,252)."""
import ctypes as C

import numpy as np
import pytest
import torch

import helpers as H
from soccer2d_b200 import Soccer2DVecEnv, _abi

pytestmark = pytest.mark.gpu

KW = dict(device="cuda:0", seed=12, use_continuous_action=False, change_ball_velocity=True, max_steps=30)


def host_steps(env, acts):
    out = []
    for act in acts:
        o, r, d, res = env.step_host(act)
        out.append((o.copy(), r.copy(), d.copy(), res.copy()))
    return out


@pytest.mark.parametrize("n,slots", [(5000, 3), (777, 4), (4096, 2)])
def test_three_slots_in_flight_one_copy_per_step(n, slots):
    """`slots` steps in flight; the outputs of every step come back with ONE cudaMemcpyAsync (packed block, N not a
    multiple of the 256-byte section alignment) and equal the synchronous call's bit for bit."""
    k = 4
    a, b = Soccer2DVecEnv(n, substeps=k, terminal_obs=True, **KW), Soccer2DVecEnv(n, substeps=k, terminal_obs=True, **KW)
    assert np.array_equal(a.reset(), b.reset())
    rng = np.random.default_rng(0)
    acts = [torch.from_numpy(H.random_actions(rng, "discrete", n, k)).pin_memory() for _ in range(13)]
    want = host_steps(a, acts)
    assert a.pipeline_info()["d2h_copies_last_step"] == 1
    b.enable_pipeline(slots=slots)
    assert b.pipeline_info()["slots"] == slots
    tickets, got = [], []
    for i, act in enumerate(acts):
        if i >= slots:  # the slot about to be re-used must have been collected
            o, r, d, res = b.wait_host(tickets[i - slots])
            got.append((o.copy(), r.copy(), d.copy(), res.copy()))
        tickets.append(b.submit_host(act))
    assert b.pipeline_info()["d2h_copies_last_step"] == 1
    for t in tickets[max(0, len(acts) - slots):]:
        o, r, d, res = b.wait_host(t)
        got.append((o.copy(), r.copy(), d.copy(), res.copy()))
    assert len(got) == len(want)
    for w, g in zip(want, got):
        for x, y in zip(w, g):
            assert np.array_equal(x, y)
    torch.cuda.synchronize()
    assert torch.equal(a.state, b.state)
    sa, sb = a.stats(), b.stats()
    assert {k_: v for k_, v in sa.items() if k_ != "return_sum"} == {k_: v for k_, v in sb.items() if k_ != "return_sum"}
    assert abs(sa["return_sum"] - sb["return_sum"]) <= 1e-9 * max(1.0, abs(sa["return_sum"]))
    a.close()
    b.close()


def test_separate_host_pointers_still_work():
    """host outputs that are NOT laid out like the device block fall back to one copy per output"""
    n, k = 1000, 2
    env, ref = Soccer2DVecEnv(n, substeps=k, **KW), Soccer2DVecEnv(n, substeps=k, **KW)
    env.reset(), ref.reset()
    act = torch.from_numpy(H.random_actions(np.random.default_rng(1), "discrete", n, k)).pin_memory()
    want = ref.step_host(act)
    obs = torch.empty((n, 10), dtype=torch.float32).pin_memory()
    rew = torch.empty(n, dtype=torch.float32).pin_memory()
    done, res = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    stream = torch.cuda.current_stream().cuda_stream
    _abi.check(env.lib.s2d_step_host(env.handle, k, act.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(),
                                     res.data_ptr(), stream), env.handle)
    torch.cuda.synchronize()
    assert env.pipeline_info()["d2h_copies_last_step"] == 4
    assert np.array_equal(obs.numpy(), want[0]) and np.array_equal(rew.numpy(), want[1])
    assert np.array_equal(done.numpy().view(bool), want[2]) and np.array_equal(res.numpy(), want[3])
    # and with only some outputs requested
    want = ref.step_host(act)
    _abi.check(env.lib.s2d_step_host(env.handle, k, act.data_ptr(), None, rew.data_ptr(), None, None, stream), env.handle)
    torch.cuda.synchronize()
    assert env.pipeline_info()["d2h_copies_last_step"] == 1 and np.array_equal(rew.numpy(), want[1])
    env.close()
    ref.close()


def test_caller_stream_work_is_ordered_against_slots_in_flight():
    """reset / step / masked reset / statistics enqueued on the caller's stream while slots are in flight neither race
    on the shared state nor on slot 0's buffers: the mixed sequence equals the same sequence run synchronously."""
    n, k = 20000, 8
    a, b = Soccer2DVecEnv(n, substeps=k, **KW), Soccer2DVecEnv(n, substeps=k, **KW)
    a.reset(), b.reset()
    b.enable_pipeline(slots=3)
    rng = np.random.default_rng(2)
    host = [torch.from_numpy(H.random_actions(rng, "discrete", n, k)).pin_memory() for _ in range(9)]
    devact = [h.cuda() for h in host]
    mask = torch.from_numpy((rng.uniform(size=n) < 0.3)).cuda()
    # the synchronous run
    want = []
    for i in range(9):
        if i % 3 == 2:
            o, r, d, res = a.step_torch(devact[i])
            want.append((o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy(), res.cpu().numpy()))
            if i == 5:
                a.reset_torch(mask)
        else:
            o, r, d, res = a.step_host(host[i])
            want.append((o.copy(), r.copy(), d.copy(), res.copy()))
    torch.cuda.synchronize()
    # the same sequence, pipelined: nothing waits on the host between submissions and caller-stream launches
    got, tickets = [None] * 9, {}
    for i in range(9):
        if i % 3 == 2:
            o, r, d, res = b.step_torch(devact[i])  # two slots are still in flight here
            got[i] = (o.clone(), r.clone(), d.clone(), res.clone())
            if i == 5:
                b.reset_torch(mask)
            for j in (i - 2, i - 1):
                o, r, d, res = b.wait_host(tickets[j])
                got[j] = (o.copy(), r.copy(), d.copy(), res.copy())
        else:
            tickets[i] = b.submit_host(host[i])
    torch.cuda.synchronize()
    for i, (w, g) in enumerate(zip(want, got)):
        for x, y in zip(w, g):
            y = y.cpu().numpy() if isinstance(y, torch.Tensor) else y
            assert np.array_equal(x, y), i
    assert torch.equal(a.state, b.state)
    assert a.stats()["episodes"] == b.stats()["episodes"] > 0
    a.close()
    b.close()


def test_rebinding_actions_reaches_the_pipeline():
    """s2d_bind after s2d_bind_pipeline_slot (vec_env.bind_actions) updates slot 0's parameter block too: the next
    submission to slot 0 copies into - and the kernel reads - the NEW action tensor (stale pkp[0])."""
    n, k = 3000, 2
    env, ref = Soccer2DVecEnv(n, substeps=k, **KW), Soccer2DVecEnv(n, substeps=k, **KW)
    env.reset(), ref.reset()
    env.enable_pipeline(slots=2)
    rng = np.random.default_rng(3)
    act = [torch.from_numpy(H.random_actions(rng, "discrete", n, k)).pin_memory() for _ in range(3)]
    old = env.actions
    env.wait_host(env.submit_host(act[0]))          # slot 0, old tensor
    ref.step_host(act[0])
    assert torch.equal(old.cpu(), act[0])
    fresh = torch.full_like(old, 255)
    env.bind_actions(fresh)
    old.fill_(7)                                     # would be read by a stale parameter block
    env.wait_host(env.submit_host(act[1]))          # slot 1
    ref.step_host(act[1])
    o, r, d, res = env.wait_host(env.submit_host(act[2]))  # slot 0 again: must use `fresh`
    want = ref.step_host(act[2])
    assert torch.equal(fresh.cpu(), act[2]) and bool((old == 7).all())
    assert np.array_equal(o, want[0]) and np.array_equal(r, want[1]) and np.array_equal(d, want[2])
    env.close()
    ref.close()


def test_clear_makes_garbage_buffers_usable():
    """s2d_clear: a native caller binding cudaMalloc'd (non-zero) memory gets the same episodes as one that zeroed it."""
    n = 2048
    ref = Soccer2DVecEnv(n, **KW)
    want = ref.reset().copy()
    env = Soccer2DVecEnv(n, **KW)
    env.state.fill_(0xA5)
    env.stats_buf.fill_(0x5A)
    _abi.check(env.lib.s2d_clear(env.handle, torch.cuda.current_stream().cuda_stream), env.handle)
    assert np.array_equal(env.reset(), want)
    assert env.stats()["episodes"] == 0
    act = H.random_actions(np.random.default_rng(4), "discrete", n, 1)
    for _ in range(40):
        a, b = env.step_host(act), ref.step_host(act)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert env.stats() == ref.stats()
    env.close()
    ref.close()


def test_output_layout_contract():
    lib = _abi.load()
    for n, scn in [(1, "reachball"), (777, "reachball"), (1 << 20, "reachball"), (1000, "fullgame")]:
        env_dim = 120 if scn == "fullgame" else 10
        cfg = H.make_config(n, "command" if scn == "fullgame" else "discrete",
                            scenario=_abi.SCENARIO_FULLGAME if scn == "fullgame" else _abi.SCENARIO_REACHBALL)
        lay = _abi.OutputLayout()
        assert lib.s2d_output_layout(C.byref(cfg), C.byref(lay)) == 0
        assert lay.obs == 0 and lay.reward >= n * env_dim * 4 and lay.done >= lay.reward + 4 * n and lay.result >= lay.done + n
        assert all(x % 256 == 0 for x in (lay.reward, lay.done, lay.result, lay.terminal_obs, lay.bytes))
        assert lay.step_bytes == lay.result + n <= lay.bytes <= n * (env_dim * 4 + 6) + 4 * 256
        assert lay.bytes_with_terminal_obs >= lay.terminal_obs + n * env_dim * 4


def test_async_statistics_do_not_block_and_equal_the_synchronous_ones():
    """allreduce_stats_async: device-side reduction of the 256 partial accumulators on a side stream (+ the NCCL sum when
    a process group exists), returned as a future; launches enqueued AFTER the call are not in the snapshot."""
    n, k = 50000, 8
    env = Soccer2DVecEnv(n, substeps=k, **KW)
    env.reset_torch()
    rng = np.random.default_rng(5)
    acts = [torch.from_numpy(H.random_actions(rng, "discrete", n, k)).cuda() for _ in range(6)]
    for a in acts[:3]:
        env.step_torch(a)
    fut = env.allreduce_stats_async()
    for a in acts[3:]:          # keep stepping: none of this may leak into the snapshot
        env.step_torch(a)
    snap = fut.result()
    after = env.stats()
    ref = Soccer2DVecEnv(n, substeps=k, **KW)
    ref.reset_torch()
    for a in acts[:3]:
        ref.step_torch(a)
    want = ref.stats()
    assert {x: snap[x] for x in want if x != "return_sum"} == {x: want[x] for x in want if x != "return_sum"}
    assert abs(snap["return_sum"] - want["return_sum"]) <= 1e-9 * max(1.0, abs(want["return_sum"]))
    assert after["episodes"] > snap["episodes"] > 0 and after["env_steps"] == 2 * snap["env_steps"]
    env.close()
    ref.close()
