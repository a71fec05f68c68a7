"""CPU-only property tests (hypothesis) on the kernel SOURCE compiled for the host (tests/emu): invariants the
domain offers, for arbitrary seeds / parameters / action streams (speed <= speed_max,
stamina in [0, stamina_max], bounded observations, done => result, determinism, K-fusion, shard invariance)."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

import emu_lib as EL
import helpers as H
import oracle_lib as OL
from soccer2d_b200 import _abi

SETTINGS = dict(max_examples=20, deadline=None, suppress_health_check=[HealthCheck.too_slow])


@settings(**SETTINGS)
@given(seed=st.integers(0, 2**63 - 1), mode=st.sampled_from(["discrete", "continuous", "turning", "command"]),
       noise=st.booleans(), max_steps=st.integers(1, 60), min_dist=st.floats(0.2, 8.0),
       player_decay=st.floats(0.2, 0.6), inertia=st.floats(1.0, 8.0))
def test_invariants_hold_for_any_seed_and_parameters(seed, mode, noise, max_steps, min_dist, player_decay, inertia):
    n = 48
    cfg = H.make_config(n, mode, seed=seed, noise=int(noise), change_ball_velocity=1, max_steps=max_steps,
                        min_distance_to_ball=float(np.float32(min_dist)),
                        sp=dict(player_decay=float(np.float32(player_decay)), inertia_moment=float(np.float32(inertia))))
    emu = EL.EmuSim(cfg)
    emu.reset()
    rng = np.random.default_rng(seed % 2**32)
    sp = cfg.sp
    vmax = sp.player_speed_max * (1 + 1.5 * sp.player_rand * noise) * (1 + 1e-5)
    for _ in range(80):
        emu.step(H.random_actions(rng, mode, n))
        s = emu.get_state()
        assert np.isfinite(s).all() and np.isfinite(emu.obs).all() and np.isfinite(emu.reward).all()
        assert (np.hypot(s[:, 2], s[:, 3]) <= vmax * sp.player_decay + 1e-6).all()        # player speed after decay
        assert (np.hypot(s[:, 11], s[:, 12]) <= sp.ball_speed_max * (1 + 1.5 * sp.ball_rand * noise) + 1e-5).all()
        assert (s[:, 5] >= 0).all() and (s[:, 5] <= sp.stamina_max).all()                 # stamina
        assert (s[:, 6] >= sp.effort_min - 1e-6).all() and (s[:, 6] <= sp.effort_max + 1e-6).all()
        assert (s[:, 7] >= sp.recover_min - 1e-6).all() and (s[:, 7] <= sp.recover_init + 1e-6).all()
        assert (np.abs(s[:, 4]) <= 180.0).all()                                           # body direction normalised
        assert (np.abs(emu.obs[:, 0:2]) <= 1.0 + 1e-6).all() and (np.abs(emu.obs[:, 7]) <= 0.5 + 1e-6).all()
        assert (s[:, 16] <= max_steps).all() and (s[:, 16] >= 0).all()                    # step_number after auto-reset
        assert ((emu.result != 0) == (emu.done != 0)).all()                               # done <=> a result
    st6 = emu.stats6
    assert st6[0] == st6[1] + st6[2] + st6[3]


@settings(**SETTINGS)
@given(seed=st.integers(0, 2**32 - 1), k=st.integers(2, 9), noise=st.booleans(), scenario=st.sampled_from(["reachball", "shoot"]))
def test_fusion_and_sharding_do_not_change_results(seed, k, noise, scenario):
    """K fused cycles == K single cycles, and a shard of the envs == the same slice of the whole run (RNG keyed on the
    global env id and the server cycle), bit for bit - with noise on as well."""
    n, lo, hi = 40, 13, 29
    scn = _abi.SCENARIO_SHOOT if scenario == "shoot" else _abi.SCENARIO_REACHBALL
    kw = dict(scenario=scn, seed=seed, noise=int(noise), max_steps=17)
    if scenario == "reachball":
        kw["change_ball_velocity"] = 1
    fused, single = EL.EmuSim(H.make_config(n, "discrete", **kw)), EL.EmuSim(H.make_config(n, "discrete", **kw))
    shard = EL.EmuSim(H.make_config(hi - lo, "discrete", env_id_offset=lo, **kw))
    for e in (fused, single, shard):
        e.reset()
    rng = np.random.default_rng(seed)
    n_act = 24 if scenario == "shoot" else 16
    for _ in range(6):
        act = rng.integers(0, n_act, size=(n, k)).astype(np.uint8)
        fused.step(act, k)
        shard.step(np.ascontiguousarray(act[lo:hi]), k)
        rsum = np.zeros(n, np.float32)
        for j in range(k):
            single.step(np.ascontiguousarray(act[:, j:j + 1]))
            rsum += single.reward
        assert np.array_equal(fused.obs, single.obs) and np.array_equal(fused.get_state(), single.get_state())
        assert np.array_equal(fused.obs[lo:hi], shard.obs) and np.array_equal(fused.get_state()[lo:hi], shard.get_state())
    assert np.array_equal(fused.stats6[:5], single.stats6[:5])


@settings(max_examples=10, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(seed=st.integers(0, 2**32 - 1), mode=st.sampled_from(["discrete", "continuous", "turning"]))
def test_kernel_source_equals_fp32_oracle_for_any_seed(seed, mode):
    n = 32
    cfg = H.make_config(n, mode, seed=seed, change_ball_velocity=1, max_steps=25)
    emu, sim = EL.EmuSim(cfg), OL.OracleSim(cfg, "f32")
    assert np.array_equal(emu.reset(), sim.reset())
    rng = np.random.default_rng(seed)
    for _ in range(60):
        act = H.random_actions(rng, mode, n)
        emu.step(act)
        sim.step(act)
        assert np.array_equal(emu.obs, sim.obs) and np.array_equal(emu.reward, sim.reward) and np.array_equal(emu.done, sim.done)
    assert np.array_equal(emu.get_state(), sim.get_state())
