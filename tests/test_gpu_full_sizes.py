"""BASELINE configs[2] and [3] at their full sizes (4 M shoot-on-goal envs; 256 K 11v11 matches): the oracle does not
finish those in seconds, so the checks are size-independent properties - fused launches equal single-cycle launches
bit for bit, a shard reproduces its slice of the global run, physical invariants hold, the statistics add up - plus
the oracle on a strided sample of the very same envs (global ids), which ties the big run to the checked small ones."""
import numpy as np
import pytest
import torch

import helpers as H
import oracle_lib as OL
from soccer2d_b200 import Soccer2DVecEnv

pytestmark = pytest.mark.gpu


def _commands(shape, gen):
    a = torch.zeros(shape + (4,), device="cuda")
    cmd = torch.randint(0, 5, shape, device="cuda", generator=gen)
    u = lambda: torch.rand(shape, device="cuda", generator=gen)  # noqa: E731
    a[..., 0] = cmd.float()
    a[..., 1] = torch.where(cmd == 4, u() * 100 - 50, u() * 100)
    a[..., 2] = torch.where(cmd == 4, u() * 60 - 30, u() * 360 - 180)
    a[..., 3] = 100.0
    return a


def test_shoot_4m_envs_properties():
    n, k, launches = 1 << 22, 4, 6
    kw = dict(scenario="shoot", device="cuda:0", seed=3, change_ball_velocity=True)
    fused = Soccer2DVecEnv(n, substeps=k, **kw)
    single = Soccer2DVecEnv(n, substeps=1, **kw)
    half = n // 2
    shard = Soccer2DVecEnv(half, substeps=k, env_id_offset=half, **kw)  # rank 1 of 2
    for e in (fused, single, shard):
        e.reset_torch()
    assert torch.equal(fused.obs, single.obs) and torch.equal(fused.obs[half:], shard.obs)
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(launches):
        act = torch.randint(0, 24, (n, k), dtype=torch.uint8, device="cuda", generator=g)
        fused.step_torch(act)
        shard.step_torch(act[half:].contiguous())
        for j in range(k):
            single.step_torch(act[:, j:j + 1].contiguous())
    assert torch.equal(fused.state, single.state) and torch.equal(fused.obs, single.obs)
    assert torch.equal(fused.obs[half:], shard.obs) and torch.equal(fused.reward[half:], shard.reward)
    f, u = fused.state_planes()
    sp = fused.cfg.sp
    assert bool(torch.isfinite(f).all())
    assert float(torch.hypot(f[0, :, 2], f[0, :, 3]).max()) <= sp.player_speed_max * sp.player_decay * (1 + 1e-6)
    assert float(torch.hypot(f[2, :, 2], f[2, :, 3]).max()) <= sp.ball_speed_max * sp.ball_decay * (1 + 1e-6)
    assert float(f[1, :, 1].min()) >= 0.0 and float(f[1, :, 1].max()) <= sp.stamina_max
    st = fused.stats()
    assert st["env_steps"] == n * k * launches and st["episodes"] == st["goals"] + st["outs"] + st["timeouts"]
    assert int((u[:, 2] - 1).sum()) == st["episodes"]
    for e in (fused, single, shard):
        e.close()


def test_fullgame_256k_matches_properties_and_oracle_sample():
    n, k, launches = 1 << 18, 4, 5
    kw = dict(scenario="fullgame", device="cuda:0", seed=17, half_time_cycles=8)
    fused = Soccer2DVecEnv(n, substeps=k, **kw)
    single = Soccer2DVecEnv(n, substeps=1, **kw)
    quarter = n // 4
    shard = Soccer2DVecEnv(quarter, substeps=k, env_id_offset=3 * quarter, **kw)  # rank 3 of 4
    # the oracle plays one match in 1024 of the same global run: global ids 5, 1029, ...
    sample = np.arange(5, n, 1024)
    sims = []
    for gid in sample[:64]:
        one = Soccer2DVecEnv(1, substeps=k, env_id_offset=int(gid), **kw)  # (only its config is used)
        sims.append(OL.OracleSim(one.cfg, "f32"))
        one.close()
    for e in (fused, single, shard):
        e.reset_torch()
    want = np.stack([s.reset()[0] for s in sims])
    assert np.array_equal(fused.obs[sample[:64]].cpu().numpy(), want)
    g = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(launches):
        act = _commands((n, k, 22), g)
        fused.step_torch(act)
        shard.step_torch(act[3 * quarter:].contiguous())
        for j in range(k):
            single.step_torch(act[:, j:j + 1].contiguous())
        host = act[sample[:64]].cpu().numpy()
        for s, a in zip(sims, host):
            s.step(a.reshape(1, -1), k)
        assert np.array_equal(fused.obs[sample[:64]].cpu().numpy(), np.stack([s.obs[0] for s in sims]))
        assert np.array_equal(fused.reward[sample[:64]].cpu().numpy(), np.array([s.reward[0] for s in sims]))
    assert torch.equal(fused.state, single.state) and torch.equal(fused.obs, single.obs)
    assert torch.equal(fused.obs[3 * quarter:], shard.obs) and torch.equal(fused.done[3 * quarter:], shard.done)
    pl = fused.fullgame_planes()
    sp = fused.cfg.sp
    assert bool(torch.isfinite(pl["pa"]).all()) and bool(torch.isfinite(pl["ball"]).all())
    assert float(torch.hypot(pl["pa"][..., 2], pl["pa"][..., 3]).max()) <= sp.player_speed_max * sp.player_decay * (1 + 1e-6)
    assert float(pl["pb"][..., 1].min()) >= 0.0 and float(pl["pb"][..., 1].max()) <= sp.stamina_max
    assert float(pl["pb"][..., 0].abs().max()) <= 180.0
    assert int(pl["ej"][:, 0].min()) >= 0 and int(pl["ej"][:, 1].min()) >= 0
    st = fused.stats()
    assert st["env_steps"] == n * k * launches
    assert st["episodes"] == n * (k * launches // 16) == st["goals"] + st["outs"] + st["timeouts"]
    for e in (fused, single, shard):
        e.close()


def test_reachball_2_to_the_27_envs_addresses_in_64_bits():
    """134 M episodes on one GPU (11 GB of state, 5 GB of observations): byte offsets exceed 2^32.  The last thousand
    envs of the big handle equal a small handle that owns the same global env ids, bit for bit, after fused launches."""
    n, k, tail = 1 << 27, 4, 1000
    kw = dict(device="cuda:0", seed=2, substeps=k, use_continuous_action=False, action_space_size=16, change_ball_velocity=True)
    big = Soccer2DVecEnv(n, **kw)
    small = Soccer2DVecEnv(tail, env_id_offset=n - tail, **kw)
    big.reset_torch()
    small.reset_torch()
    assert torch.equal(big.obs[n - tail:], small.obs)
    g = torch.Generator(device="cuda").manual_seed(0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for launch in range(3):
        act = torch.randint(0, 16, (n, k), dtype=torch.uint8, device="cuda", generator=g)
        ev[0].record()
        big.step_torch(act)
        ev[1].record()
        small.step_torch(act[n - tail:].contiguous())
        assert torch.equal(big.obs[n - tail:], small.obs) and torch.equal(big.reward[n - tail:], small.reward)
    torch.cuda.synchronize()
    assert big.stats()["env_steps"] == n * k * 3
    rate = n * k / (ev[0].elapsed_time(ev[1]) * 1e-3)
    assert rate > 2e10, rate  # K = 4: still HBM-bound, four cycles per 222 bytes
    big.close()
    small.close()
