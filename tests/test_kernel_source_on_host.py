"""CPU-only: the __device__ functions of csrc/*.cuh (the kernel SOURCE), compiled for the host by tests/emu, must
reproduce the fp32 oracle bit for bit.  This catches arithmetic / logic divergences before any GPU time is spent;
the -m gpu tests then check the real kernels, memory layout and launch plumbing on the B200."""
import numpy as np
import pytest

import emu_lib as EL
import helpers as H
import oracle_lib as OL
from test_gpu_parity_cases import HAND_PLACED_STATES


@pytest.mark.parametrize("mode,k", [("discrete", 1), ("discrete", 16), ("continuous", 1), ("continuous", 4), ("turning", 1), ("turning", 3)])
def test_kernel_source_bit_exact(mode, k):
    n = 257
    cfg = H.make_config(n, mode, seed=7, change_ball_velocity=1, max_steps=60 if k == 1 else 200)
    emu, sim = EL.EmuSim(cfg), OL.OracleSim(cfg, "f32")
    assert np.array_equal(emu.reset(), sim.reset())
    assert np.array_equal(emu.get_state(), sim.get_state())
    rng = np.random.default_rng(0)
    episodes = 0
    for t in range(600 // k):
        act = H.random_actions(rng, mode, n, k)
        emu.step(act, k)
        sim.step(act, k)
        assert np.array_equal(emu.done, sim.done) and np.array_equal(emu.result, sim.result)
        assert np.array_equal(emu.obs, sim.obs) and np.array_equal(emu.reward, sim.reward)
        d = emu.done.astype(bool)
        assert np.array_equal(emu.term_obs[d], sim.term_obs[d])
        episodes += int(d.sum())
    assert np.array_equal(emu.get_state(), sim.get_state())
    assert episodes > n
    from soccer2d_b200 import _abi
    st = sim.stats(_abi.Stats())
    assert [st.episodes, st.goals, st.outs, st.timeouts, st.episode_steps] == [int(x) for x in emu.stats6[:5]]
    assert st.return_sum == pytest.approx(emu.stats6[5], rel=1e-12)


def test_default_constants_fold_to_the_same_bits():
    """DefaultSP (immediates in the constant-folded kernels) == make_cycle_consts(default ServerParam)."""
    assert EL.lib().emu_check_default_consts() == 0


@pytest.mark.parametrize("mode", ["discrete", "continuous", "turning"])
def test_kernel_source_runtime_params_path(mode):
    """Non-default ServerParam -> the RuntimeSP kernels (constants from the constant bank), incl. back dashes
    (min_dash_power < 0 is irrelevant at power 100, so exercise slowness, a 45-degree dash_angle_step, other decays)."""
    n = 129
    sp = dict(dash_angle_step=45.0, slowness_on_top_for_left_team=1.25, player_decay=0.5, ball_decay=0.9,
              side_dash_rate=0.5, back_dash_rate=0.6, stamina_inc_max=30.0, player_speed_max=0.8, inertia_moment=3.0)
    cfg = H.make_config(n, mode, seed=3, change_ball_velocity=1, max_steps=80, sp=sp)
    emu, sim = EL.EmuSim(cfg), OL.OracleSim(cfg, "f32")
    assert EL.lib().emu_uses_default_sp(emu.h) == 0
    assert np.array_equal(emu.reset(), sim.reset())
    rng = np.random.default_rng(0)
    for t in range(300):
        act = H.random_actions(rng, mode, n)
        emu.step(act)
        sim.step(act)
        assert np.array_equal(emu.done, sim.done) and np.array_equal(emu.result, sim.result)
        assert np.array_equal(emu.obs, sim.obs) and np.array_equal(emu.reward, sim.reward)
    assert np.array_equal(emu.get_state(), sim.get_state())


@pytest.mark.parametrize("step", [1.0, 22.5, 0.0])
def test_kernel_source_sincos_memo_hits_and_misses(step):
    """ReachBall Discrete(n) launches of K >= 4 cycles take the dash's sin / cos from the whole-degree memo
    (s2d_math.cuh: sincos_deg_memo; the emulation applies step_kernel's rule).  dash_angle_step = 1 only hits;
    22.5 and 0 (no snapping) make half of the 16 directions fractional, those miss and take the polynomial.  The
    oracle only knows the polynomial: equal bits either way."""
    n, k = 131, 8
    kw = dict(sp=dict(dash_angle_step=step)) if step != 1.0 else {}
    cfg = H.make_config(n, "discrete", seed=5, change_ball_velocity=1, max_steps=50, **kw)
    emu, sim = EL.EmuSim(cfg), OL.OracleSim(cfg, "f32")
    assert np.array_equal(emu.reset(), sim.reset())
    rng = np.random.default_rng(1)
    for t in range(40):
        act = H.random_actions(rng, "discrete", n, k)
        emu.step(act, k)
        sim.step(act, k)
        assert np.array_equal(emu.done, sim.done) and np.array_equal(emu.result, sim.result)
        assert np.array_equal(emu.obs, sim.obs) and np.array_equal(emu.reward, sim.reward)
    assert np.array_equal(emu.get_state(), sim.get_state())


@pytest.mark.parametrize("collision_model", [0, 1])
def test_kernel_source_hand_placed_states(collision_model):
    cases = HAND_PLACED_STATES + [[0, 0, 0.2, 0, 0, 8000, 1, 1, 130600, 1.0, 0, -0.5, 0, 1, 5, 0, 3, 3, 1],
                                  [0, 0, 0.3, 0.1, 0, 8000, 1, 1, 130600, 0.5, 0.2, 0, 0, 1, 5, 0, 3, 3, 1]]
    n = len(cases)
    cfg = H.make_config(n, "continuous", seed=1, min_distance_to_ball=0.05, max_steps=100000, collision_model=collision_model)
    emu, sim = EL.EmuSim(cfg), OL.OracleSim(cfg, "f32")
    emu.reset()
    sim.reset()
    for i, c in enumerate(cases):
        emu.set_state(i, c + [0])
        sim.set_state(i, np.array(c + [0], dtype=np.float64))
    rng = np.random.default_rng(4)
    flags = 0
    for _ in range(60):
        act = H.random_actions(rng, "continuous", n)
        act[6:8] = 0.0
        emu.step(act)
        sim.step(act)
        assert np.array_equal(emu.obs, sim.obs) and np.array_equal(emu.reward, sim.reward)
        assert np.array_equal(emu.done, sim.done)
        g = emu.get_state()
        assert np.array_equal(g, sim.get_state())
        flags |= int(np.bitwise_or.reduce(g[:, 19].astype(int)))
    assert flags & 1 and flags & 2


def test_kernel_source_masked_reset_no_auto_reset():
    n = 64
    cfg = H.make_config(n, "discrete", seed=4, auto_reset=0, max_steps=10, change_ball_velocity=1)
    emu, sim = EL.EmuSim(cfg), OL.OracleSim(cfg, "f32")
    emu.reset()
    sim.reset()
    rng = np.random.default_rng(1)
    for t in range(40):
        act = H.random_actions(rng, "discrete", n)
        emu.step(act)
        sim.step(act)
        assert np.array_equal(emu.get_state(), sim.get_state())
        if t % 7 == 6:
            mask = emu.done.copy()
            emu.reset(mask)
            sim.reset(mask)
            assert np.array_equal(emu.obs[mask.astype(bool)], sim.obs[mask.astype(bool)])
            assert np.array_equal(emu.get_state(), sim.get_state())


# ---- command mode (proto PlayerAction vocabulary) and the SHOOT scenario ---------------------------------

def _same(emu, sim):
    assert np.array_equal(emu.done, sim.done) and np.array_equal(emu.result, sim.result)
    assert np.array_equal(emu.obs, sim.obs) and np.array_equal(emu.reward, sim.reward)
    d = emu.done.astype(bool)
    assert np.array_equal(emu.term_obs[d], sim.term_obs[d])


@pytest.mark.parametrize("default_sp", [True, False])
def test_kernel_source_command_mode_reachball(default_sp):
    """none / dash / turn / kick / go-to-point with out-of-range arguments, ReachBall scoring.  min_distance 0.3 so
    that episodes get close enough for kicks and collisions to happen."""
    n = 193
    sp = {} if default_sp else dict(inertia_moment=4.0, kick_power_rate=0.03, kickable_margin=0.9, dash_angle_step=0.0)
    cfg = H.make_config(n, "command", seed=9, change_ball_velocity=1, max_steps=120, min_distance_to_ball=0.3,
                        goto_dist_thr=0.4, sp=sp)
    emu, sim = EL.EmuSim(cfg), OL.OracleSim(cfg, "f32")
    assert EL.lib().emu_uses_default_sp(emu.h) == int(default_sp)
    assert np.array_equal(emu.reset(), sim.reset())
    rng = np.random.default_rng(0)
    flags = 0
    for t in range(500):
        act = H.random_commands(rng, n)
        if t % 2:  # every other step chase the ball so that kicks find it kickable
            act = np.where(rng.uniform(size=(n, 1, 1)) < 0.7, H.chase_and_shoot(sim.obs), act).astype(np.float32)
        emu.step(act)
        sim.step(act)
        _same(emu, sim)
        g = emu.get_state()
        assert np.array_equal(g, sim.get_state())
        flags |= int(np.bitwise_or.reduce(g[:, 19].astype(int)))
    assert flags & 4, "no kick was ever applied"


@pytest.mark.parametrize("mode", ["discrete", "command"])
def test_kernel_source_shoot(mode):
    from soccer2d_b200 import _abi
    n = 160
    cfg = H.make_config(n, mode, scenario=_abi.SCENARIO_SHOOT, seed=21, change_ball_velocity=0, max_steps=150)
    emu, sim = EL.EmuSim(cfg), OL.OracleSim(cfg, "f32")
    assert np.array_equal(emu.reset(), sim.reset())
    rng = np.random.default_rng(1)
    for t in range(700):
        if mode == "discrete":
            act = rng.integers(0, 24, size=(n, 1)).astype(np.uint8)
        else:
            act = H.chase_and_shoot(sim.obs, rng, kick_prob=0.9)
            rnd = H.random_commands(rng, n)
            act = np.where(rng.uniform(size=(n, 1, 1)) < 0.1, rnd, act).astype(np.float32)
        emu.step(act)
        sim.step(act)
        _same(emu, sim)
    assert np.array_equal(emu.get_state(), sim.get_state())
    st = sim.stats(_abi.Stats())
    assert st.episodes > 0
    if mode == "command":  # the scripted policy scores, misses and runs out of time
        assert st.goals > 20 and st.outs > 0 and st.timeouts >= 0


def test_kernel_source_shoot_goal_line_cases():
    """Hand-placed balls about to cross the goal line: inside the post, outside it, own goal, side line; and a
    kick from the edge of the kickable area."""
    from soccer2d_b200 import _abi
    line = 52.5 + 0.085
    # px py vx vy body stamina effort recovery capacity | bx by bvx bvy | mem_pb mem_bg ep_ret | step cycle episode
    base = [0, 0, 0, 0, 0, 8000, 1, 1, 130600]
    cases = [
        base + [line - 0.5, 6.9, 1.0, 0.1, 50, 1, 0, 3, 3, 1],      # crosses inside the post -> Goal
        base + [line - 0.5, 7.0, 1.0, 0.2, 50, 1, 0, 3, 3, 1],      # crosses just outside -> Out
        base + [line - 0.5, 0.0, 0.4, 0.0, 50, 1, 0, 3, 3, 1],      # does not reach the line this cycle
        base + [-line + 0.5, 0.0, -1.0, 0.0, 50, 104, 0, 3, 3, 1],  # own goal -> Out
        base + [10, 33.9, 0.0, 0.5, 40, 45, 0, 3, 3, 1],            # over the side line -> Out
        [20, 10, 0, 0, 30, 8000, 1, 1, 130600, 20 + 1.08, 10, 0, 0, 1.08, 34, 0, 3, 3, 1],   # kickable (edge)
        [20, 10, 0, 0, 30, 8000, 1, 1, 130600, 20 + 1.09, 10, 0, 0, 1.09, 34, 0, 3, 3, 1],   # just not kickable
        [51, 0, 0, 0, 0, 8000, 1, 1, 130600, 51.5, 0, 0, 0, 0.5, 1, 0, 3, 3, 1],             # kick straight in
    ]
    n = len(cases)
    cfg = H.make_config(n, "command", scenario=_abi.SCENARIO_SHOOT, seed=2, max_steps=50)
    emu, sim = EL.EmuSim(cfg), OL.OracleSim(cfg, "f32")
    emu.reset()
    sim.reset()
    for i, c in enumerate(cases):
        emu.set_state(i, c + [0])
        sim.set_state(i, np.array(c + [0], dtype=np.float64))
    act = np.zeros((n, 1, 4), np.float32)
    act[5:, 0, :3] = [3, 100, 0]  # Kick(100, 0)
    emu.step(act)
    sim.step(act)
    _same(emu, sim)
    assert np.array_equal(emu.get_state(), sim.get_state())
    assert list(sim.result[:5]) == [1, 2, 0, 2, 2]
    st = sim.get_state()
    assert int(st[5, 19]) & 4 and not int(st[6, 19]) & 4  # kicked flag: edge of the kickable area
    for _ in range(3):
        emu.step(act)
        sim.step(act)
        _same(emu, sim)
    assert sim.stats(_abi.Stats()).goals >= 2  # case 0 and the straight kick


# ---- noise (player_rand / ball_rand / kick_rand from the counter-based RNG) ---------------------------------

@pytest.mark.parametrize("scenario,mode", [("reachball", "discrete"), ("reachball", "turning"), ("reachball", "command"),
                                           ("shoot", "command")])
def test_kernel_source_with_noise_bit_exact(scenario, mode):
    from soccer2d_b200 import _abi
    n = 150
    scn = _abi.SCENARIO_SHOOT if scenario == "shoot" else _abi.SCENARIO_REACHBALL
    kw = dict(min_distance_to_ball=0.3, goto_dist_thr=0.4) if mode == "command" and scenario == "reachball" else {}
    cfg = H.make_config(n, mode, scenario=scn, seed=31, noise=1, change_ball_velocity=1, max_steps=100, **kw)
    emu, sim, quiet = EL.EmuSim(cfg), OL.OracleSim(cfg, "f32"), OL.OracleSim(H.make_config(
        n, mode, scenario=scn, seed=31, noise=0, change_ball_velocity=1, max_steps=100, **kw), "f32")
    assert np.array_equal(emu.reset(), sim.reset())
    quiet.reset()
    assert not np.array_equal(sim.obs, quiet.obs)  # the reset's idle cycle already moves the ball noisily
    rng = np.random.default_rng(0)
    flags = 0
    for t in range(400):
        if mode == "command":
            act = H.chase_and_shoot(sim.obs, rng, kick_prob=0.9)
            act = np.where(rng.uniform(size=(n, 1, 1)) < 0.15, H.random_commands(rng, n), act).astype(np.float32)
        else:
            act = H.random_actions(rng, mode, n)
        emu.step(act)
        sim.step(act)
        _same(emu, sim)
        if t % 40 == 39:
            g = emu.get_state()
            assert np.array_equal(g, sim.get_state())
            flags |= int(np.bitwise_or.reduce(g[:, 19].astype(int)))
    if mode == "command":
        assert flags & 4  # kicks (and their noise) happened


def test_noise_statistics_and_f64_agreement():
    """fp32 spec vs f64 truth with noise on (the uniforms are identical integers, so the 1e-5 bound still holds over
    a few hundred cycles), and the size of the velocity noise: |dv| <= rand * |v| per axis."""
    n = 256
    cfg = H.make_config(n, "discrete", seed=8, noise=1, change_ball_velocity=1)
    t, s = OL.OracleSim(cfg, "f64"), OL.OracleSim(cfg, "f32")
    t.reset()
    s.reset()
    rng = np.random.default_rng(2)
    for _ in range(150):
        act = H.random_actions(rng, "discrete", n)
        t.step(act)
        s.step(act)
        assert np.array_equal(t.done, s.done)
        assert H.obs_close(t.obs, s.obs) < H.TOL
    # ball only: with noise off the ball decays along a straight line; with noise on its direction wanders a little
    quiet = OL.OracleSim(H.make_config(n, "discrete", seed=8, noise=0, change_ball_velocity=1), "f64")
    loud = OL.OracleSim(H.make_config(n, "discrete", seed=8, noise=1, change_ball_velocity=1), "f64")
    quiet.reset()
    loud.reset()
    q, l = quiet.get_state(), loud.get_state()
    vq, vl = np.hypot(q[:, 11], q[:, 12]), np.hypot(l[:, 11], l[:, 12])
    moving = vq > 0.1
    rel = np.abs(vl[moving] - vq[moving]) / vq[moving]
    assert rel.max() <= 0.05 * np.sqrt(2) + 1e-6 and rel.mean() > 0.005  # ball_rand = 0.05 per axis
