"""CPU-only: pins the C oracle (oracle/s2d_oracle.c) against the Python oracle (oracle/soccer2d_oracle.py, itself
pinned to the reference by tests/test_oracle_contract.py), and bounds the fp32 spec against the f64 truth."""
import ctypes as C
import math

import numpy as np
import pytest

import helpers as H
import oracle_lib as OL
from oracle import soccer2d_oracle as O
from soccer2d_b200 import _abi


def test_philox_c_matches_known_answers():
    L = OL.lib("f64")
    out = (C.c_uint32 * 4)()
    # counter = (env_lo, env_hi, index, purpose<<24|sub), key = seed -> Random123 known answers
    L.s2do_probe_philox(0, 0, 0, 0, 0, out)
    assert tuple(out) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    L.s2do_probe_philox(0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFF, 0xFF, 0xFFFFFF, out)
    assert tuple(out) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    L.s2do_probe_philox(0x299F31D0A4093822, 0x85A308D3243F6A88, 0x13198A2E, 0x03, 0x707344, out)
    assert tuple(out) == (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)
    for seed, env, idx, purpose, sub in [(7, 123456789012, 55, 2, 0), (2**63 + 5, 3, 9, 1, 31)]:
        L.s2do_probe_philox(seed, env, idx, purpose, sub, out)
        assert tuple(out) == O.rng_block(seed, env, idx, purpose, sub)


def test_fp32_math_spec_close_to_libm():
    L = OL.lib("f32")
    s, c = C.c_double(), C.c_double()
    rng = np.random.default_rng(0)
    xs = np.concatenate([np.arange(-720, 721, 0.5), rng.uniform(-400, 400, 4000)]).astype(np.float32)
    worst = 0.0
    for x in xs:
        L.s2do_probe_sincos_deg(float(x), C.byref(s), C.byref(c))
        r = math.radians(float(x))
        worst = max(worst, abs(s.value - math.sin(r)), abs(c.value - math.cos(r)))
    assert worst < 2.5e-7
    # exact at the multiples of 90 degrees (what integer body directions and Discrete(16) actions hit)
    for d, (es, ec) in {0: (0, 1), 90: (1, 0), 180: (0, -1), -90: (-1, 0), 270: (-1, 0), 360: (0, 1), -180: (0, -1)}.items():
        L.s2do_probe_sincos_deg(float(d), C.byref(s), C.byref(c))
        assert (s.value, c.value) == (es, ec)
    worst = 0.0
    pts = rng.uniform(-60, 60, (6000, 2)).astype(np.float32)
    pts[:50, 0] = 0.0
    pts[50:100, 1] = 0.0
    for y, x in pts:
        got = L.s2do_probe_atan2_deg(float(y), float(x))
        want = O.atan2_deg(float(y), float(x))
        d = abs(got - want)
        worst = max(worst, min(d, 360 - d))
    assert worst < 3.0e-5  # degrees; the f32 ulp at 180 is 1.5e-5
    assert L.s2do_probe_atan2_deg(0.0, 0.0) == 0.0
    worst = 0.0
    for a, b in rng.uniform(-1, 1, (2000, 2)):
        want = math.exp(a) / (math.exp(a) + math.exp(b))
        worst = max(worst, abs(L.s2do_probe_softmax_first(float(np.float32(a)), float(np.float32(b))) - want))
    assert worst < 2.0e-7


def _py_oracles(cfg, kw, n):
    pc = O.ReachBallConfig(seed=int(cfg.seed), sp=O.ServerParam().as_f32(), **kw)  # S2DServerParam is float
    return [O.ReachBallOracle(pc, env_id=int(cfg.env_id_offset) + i, auto_reset=bool(cfg.auto_reset)) for i in range(n)]


@pytest.mark.parametrize("mode", ["discrete", "continuous", "turning"])
def test_c_f64_equals_python_oracle(mode):
    """Same algorithm, two restatements (C and Python): trajectories must agree to rounding noise."""
    n, steps = 6, 450
    kw = dict(change_ball_position=True, change_ball_velocity=True, use_continuous_action=mode != "discrete",
              use_turning=mode == "turning", max_steps=60, min_distance_to_ball=3.0)
    cfg = H.make_config(n, mode, seed=11, env_id_offset=1000, change_ball_velocity=1, max_steps=60,
                        min_distance_to_ball=3.0)
    sim = OL.OracleSim(cfg, "f64")
    py = _py_oracles(cfg, kw, n)
    obs = sim.reset().copy()
    for i, o in enumerate(py):
        assert np.allclose(o.reset(), obs[i], rtol=0, atol=1e-12)
    rng = np.random.default_rng(3)
    results = set()
    for _ in range(steps):
        act = H.random_actions(rng, mode, n)
        obs, rew, done, res = sim.step(act)
        for i, o in enumerate(py):
            a = act[i, 0]
            po, pr, pd, pres, _ = o.step(a if mode != "continuous" else [a])
            assert pd == bool(done[i]) and pres == int(res[i])
            assert pr == pytest.approx(float(rew[i]), rel=1e-11, abs=1e-11)
            assert H.obs_close(po, obs[i], 0) < 1e-11
            results.add(pres)
    assert results == {0, 1, 2, 3} or mode != "discrete" or results >= {0, 1, 3}
    st = sim.stats(_abi.Stats())
    assert st.episodes > 0 and st.env_steps == n * steps
    assert st.episodes == st.goals + st.outs + st.timeouts


def test_c_f64_k_substeps_equals_single_steps():
    n, k, launches = 16, 8, 40
    cfg = H.make_config(n, "discrete", seed=2, change_ball_velocity=1, max_steps=50)
    a, b = OL.OracleSim(cfg, "f64"), OL.OracleSim(cfg, "f64")
    a.reset()
    b.reset()
    rng = np.random.default_rng(0)
    for _ in range(launches):
        act = H.random_actions(rng, "discrete", n, k)
        oa, ra, da, _ = a.step(act, k)
        rsum = np.zeros(n)
        dany = np.zeros(n, np.uint8)
        for j in range(k):
            ob, rb, db, _ = b.step(np.ascontiguousarray(act[:, j:j + 1]), 1)
            rsum += rb
            dany |= db
        assert np.array_equal(oa, ob) and np.array_equal(da, dany)
        assert np.allclose(ra, rsum, rtol=0, atol=1e-9)
    assert np.array_equal(a.get_state(), b.get_state())


@pytest.mark.parametrize("mode", ["discrete", "continuous", "turning"])
def test_fp32_spec_tracks_f64_truth_1000_cycles(mode):
    """The fp32 arithmetic the GPU uses (C f32 build) against the double-precision truth over 1 000 cycles:
    flags bit-exact, floats within the north-star tolerance (1e-5 relative to each quantity's scale)."""
    n = 192
    cfg = H.make_config(n, mode, seed=2024, change_ball_velocity=1)
    t, s = OL.OracleSim(cfg, "f64"), OL.OracleSim(cfg, "f32")
    assert H.obs_close(t.reset(), s.reset()) < H.TOL
    rng = np.random.default_rng(5)
    episodes = 0
    for _ in range(1000):
        act = H.random_actions(rng, mode, n)
        ot, rt, dt, rest = t.step(act)
        os_, rs, ds, ress = s.step(act)
        assert np.array_equal(dt, ds) and np.array_equal(rest, ress)
        assert H.obs_close(ot, os_) < H.TOL
        assert np.abs(rt - rs).max() < H.TOL * 100.0  # rewards are differences of distances (scale 100 m)
        episodes += int(dt.sum())
    assert episodes > n  # every env went through several resets on the way
    st, ss = t.get_state(), s.get_state()
    assert np.array_equal(st[:, 16:], ss[:, 16:])  # step_number, cycle, episode, collision flags
    assert H.state_err(st, ss) < H.TOL


def test_fp32_flag_flips_are_rare_and_near_thresholds():
    """On a big batch a threshold test (dist < min_distance, |x| > 52.5) can land within fp32 rounding of
    its threshold; count how often the fp32 spec and the f64 truth disagree on `done`."""
    n = 4096
    cfg = H.make_config(n, "continuous", seed=5, change_ball_velocity=1)
    t, s = OL.OracleSim(cfg, "f64"), OL.OracleSim(cfg, "f32")
    t.reset()
    s.reset()
    rng = np.random.default_rng(1)
    alive = np.ones(n, bool)
    for _ in range(400):
        act = H.random_actions(rng, "continuous", n)
        _, _, dt, _ = t.step(act)
        _, _, ds, _ = s.step(act)
        alive &= dt == ds
    assert (~alive).sum() <= 4  # <= 1e-3 of the envs, ~2.5e-6 per env-step


@pytest.mark.parametrize("mode", ["discrete", "command"])
def test_shoot_c_f64_equals_python_oracle(mode):
    """The Shoot scenario has no counterpart in the reference; its two restatements (C and Python) must agree."""
    n, steps = 5, 500
    cfg = H.make_config(n, mode, scenario=_abi.SCENARIO_SHOOT, seed=4, env_id_offset=50, max_steps=120, goto_dist_thr=0.5)
    sim = OL.OracleSim(cfg, "f64")
    pc = O.ShootConfig(seed=4, max_steps=120, sp=O.ServerParam().as_f32())
    py = [O.ShootOracle(pc, env_id=50 + i) for i in range(n)]
    obs = sim.reset().copy()
    for i, o in enumerate(py):
        assert np.allclose(o.reset(), obs[i], rtol=0, atol=1e-12)
    rng = np.random.default_rng(6)
    results = set()
    kicks = 0
    for _ in range(steps):
        if mode == "discrete":
            act = rng.integers(0, 24, size=(n, 1)).astype(np.uint8)
        else:
            act = H.chase_and_shoot(sim.obs, rng, kick_prob=0.9)
            act = np.where(rng.uniform(size=(n, 1, 1)) < 0.1, H.random_commands(rng, n), act).astype(np.float32)
        obs, rew, done, res = sim.step(act)
        for i, o in enumerate(py):
            a = int(act[i, 0]) if mode == "discrete" else act[i, 0]
            po, pr, pd, pres, _ = o.step(a)
            assert pd == bool(done[i]) and pres == int(res[i])
            assert pr == pytest.approx(float(rew[i]), rel=1e-10, abs=1e-10)
            assert H.obs_close(po, obs[i], 0) < 1e-10
            results.add(pres)
            kicks += o.player.kicked
    if mode == "command":
        assert {0, 1} <= results and kicks > 20


def _place(sim, i, px, py, vx, vy, body, bx, by, bvx, bvy):
    st = sim.get_state(i)
    st[0:5] = [px, py, vx, vy, body]
    st[9:13] = [bx, by, bvx, bvy]
    sim.set_state(i, st)


def test_body_actions_do_what_librcsc_intends():
    """S2D_CMD_TURN_TO_POINT / _BALL / _ANGLE, KICK_ONE_STEP, STOP_BALL (include/soccer2d.h): properties of the
    lowered turn / kick, in the double-precision build."""
    cfg = H.make_config(8, "command", scenario=_abi.SCENARIO_SHOOT, auto_reset=0, max_steps=10 ** 6)
    sim = OL.OracleSim(cfg, "f64")
    sim.reset()
    decay, bdecay = 0.4, 0.94
    #          px,  py,  vx,   vy,  body,   bx,   by,  bvx,  bvy
    _place(sim, 0, 0.0, 0.0, 0.0, 0.0, 30.0, 20.0, 20.0, 0.0, 0.0)    # standing: turn to a point
    _place(sim, 1, 0.0, 0.0, 0.1, 0.05, 30.0, 20.0, 20.0, 0.0, 0.0)   # moving: turn to a point, n = 1
    _place(sim, 2, 0.0, 0.0, 0.3, -0.1, 0.0, 10.0, 5.0, 1.0, 0.5)     # moving: turn to the ball, n = 3
    _place(sim, 3, 1.0, 1.0, 0.0, 0.0, 170.0, 30.0, 0.0, 0.0, 0.0)    # turn to an absolute angle across the +-180 seam
    _place(sim, 4, 0.0, 0.0, 0.0, 0.0, 20.0, 0.5, 0.3, 0.4, -0.2)     # kick one step
    _place(sim, 5, 0.0, 0.0, 0.0, 0.0, -45.0, 0.6, -0.2, 1.0, 0.8)    # stop the ball
    _place(sim, 6, 0.0, 0.0, 0.0, 0.0, 0.0, 3.0, 0.0, 0.5, 0.0)       # ball out of reach: nothing happens
    _place(sim, 7, 0.0, 0.0, 0.0, 0.0, 0.0, 0.5, 0.0, -2.5, 0.0)      # wanted change needs more than max_power
    act = np.zeros((8, 1, 4), np.float32)
    act[0, 0] = [5, -10.0, 25.0, 1]
    act[1, 0] = [5, -10.0, 25.0, 1]
    act[2, 0] = [6, 3, 0, 0]
    act[3, 0] = [7, -175.0, 0, 0]
    act[4, 0] = [8, 30.0, 10.0, 1.5]
    act[5, 0] = [9, 0, 0, 0]
    act[6, 0] = [8, 30.0, 10.0, 1.5]
    act[7, 0] = [8, 30.0, 0.0, 3.0]
    before = sim.get_state()
    sim.step(act)
    st = sim.get_state()
    deg = lambda y, x: np.degrees(np.arctan2(y, x))  # noqa: E731
    norm = lambda d: (d + 180.0) % 360.0 - 180.0     # noqa: E731
    # 0: a standing player faces the point after the turn
    assert st[0, 4] == pytest.approx(deg(25.0, -10.0), abs=1e-3)
    # 1: n = 1 looks from where the player is after this cycle; inertia is compensated exactly
    assert st[1, 4] == pytest.approx(deg(25.0 - st[1, 1], -10.0 - st[1, 0]), abs=1e-3)
    assert (st[1, 0], st[1, 1]) == pytest.approx((0.1, 0.05))
    # 2: the ball as it will be 3 cycles on, seen from where the player will be 3 cycles on
    f = lambda d, n: (1 - d ** n) / (1 - d)  # noqa: E731
    tx, ty = 10.0 + 1.0 * f(bdecay, 3), 5.0 + 0.5 * f(bdecay, 3)
    mx, my = 0.3 * f(decay, 3), -0.1 * f(decay, 3)
    assert st[2, 4] == pytest.approx(deg(ty - my, tx - mx), abs=1e-3)
    # 3: 170 -> -175 is a turn of +15 through the seam
    assert st[3, 4] == pytest.approx(-175.0, abs=1e-3)
    # 4: the ball leaves with first_speed towards the target (one decay later)
    want = 1.5 * np.array([30.0 - 0.5, 10.0 - 0.3]) / np.hypot(30.0 - 0.5, 10.0 - 0.3)
    assert st[4, 11:13] == pytest.approx(want * bdecay, abs=1e-4)
    assert st[4, 9:11] == pytest.approx(np.array([0.5, 0.3]) + want, abs=1e-4)
    # 5: the ball is stopped
    assert st[5, 11:13] == pytest.approx([0.0, 0.0], abs=1e-5)
    # 6: not kickable: the ball just rolls on
    assert st[6, 11] == pytest.approx(0.5 * bdecay) and int(st[6, 19]) & 4 == 0
    # 7: force mode: full power towards the wanted change, short of the wanted speed
    assert int(st[7, 19]) & 4 and -2.5 * bdecay < st[7, 11] < 3.0 * bdecay and st[7, 11] > before[7, 11]


def test_intercept_meets_a_rolling_ball_sooner_than_chasing_it():
    """S2D_CMD_INTERCEPT aims at where the ball can first be met; S2D_CMD_GOTO at the ball's present position runs
    after it.  Same start, ball rolling across the player's front."""
    cfg = H.make_config(2, "command", scenario=_abi.SCENARIO_SHOOT, auto_reset=0, max_steps=10 ** 6, goto_dist_thr=0.5)
    sim = OL.OracleSim(cfg, "f64")
    sim.reset()
    for i in range(2):
        _place(sim, i, -10.0, 0.0, 0.0, 0.0, 0.0, 0.0, -12.0, 0.0, 2.2)
    reached = [None, None]
    for t in range(60):
        st = sim.get_state()
        act = np.zeros((2, 1, 4), np.float32)
        act[0, 0] = [10, 0, 0, 0]
        act[1, 0] = [4, st[1, 9], st[1, 10], 100.0]
        sim.step(act)
        st = sim.get_state()
        for i in range(2):
            if reached[i] is None and np.hypot(st[i, 9] - st[i, 0], st[i, 10] - st[i, 1]) <= 1.085:
                reached[i] = t
    assert reached[0] is not None and (reached[1] is None or reached[0] < reached[1]), reached


def test_player_types_per_match_equal_one_handle_per_match():
    """s2do_set_player_types_per_match (one assignment per match) against what it must mean: match e plays exactly like a
    one-match handle with env_id_offset = e that was given row e as its per-handle assignment."""
    import ctypes as C
    from test_gpu_fullgame import swarm_policy
    from soccer2d_b200.vec_env import per_match_type_assignment
    n, pps = 5, 11
    p = 2 * pps
    lib = _abi.load()
    cfg = H.make_config(n, "command", scenario=_abi.SCENARIO_FULLGAME, seed=21, half_time_cycles=60)
    types = (_abi.PlayerType * 18)()
    assert lib.s2d_generate_player_types(4, C.byref(cfg.sp), types, 18) == 0
    assign = per_match_type_assignment(4, 0, n, pps, 18)
    assert assign.shape == (n, p) and (assign[:, [0, pps]] == 0).all()
    assert all(len(set(r[1:pps])) == pps - 1 and len(set(r[pps + 1:])) == pps - 1 for r in assign.tolist())  # pt_max = 1
    assert np.array_equal(per_match_type_assignment(4, 2, 3, pps, 18), assign[2:])  # keyed on the global match id
    for build in ("f64", "f32"):
        whole = OL.OracleSim(cfg, build)
        whole.set_player_types(types, 18, assign)
        singles = []
        for e in range(n):
            c1 = H.make_config(1, "command", scenario=_abi.SCENARIO_FULLGAME, seed=21, half_time_cycles=60, env_id_offset=e)
            s1 = OL.OracleSim(c1, build)
            s1.set_player_types(types, 18, assign[e])
            s1.reset()
            singles.append(s1)
        obs = whole.reset().copy()
        assert all(np.array_equal(obs[e], singles[e].obs[0]) for e in range(n))
        rng = np.random.default_rng(8)
        for t in range(150):
            act = swarm_policy(whole.obs, p, rng, random_frac=0.2)
            whole.step(act.reshape(n, -1))
            for e in range(n):
                singles[e].step(act[e].reshape(1, -1))
                assert np.array_equal(whole.obs[e], singles[e].obs[0]) and whole.done[e] == singles[e].done[0]
        assert whole.done.any() or t > 100
    # a row with an unknown type is refused
    bad = assign.copy()
    bad[3, 5] = 18
    with pytest.raises(AssertionError):
        OL.OracleSim(cfg, "f64").set_player_types(types, 18, bad)


def test_fullgame_invariants_with_player_types_and_referee():
    """11 v 11 in the f64 build with rcssserver's heterogeneous player types, swarm play for 400 cycles: the physical
    limits of every player's own type hold, the referee's state stays consistent (dead ball inside the pitch, offside
    marks from one team only and only while play goes on, nobody of the other side within 9.15 m of a dead ball)."""
    import ctypes as C
    from test_gpu_fullgame import swarm_policy  # (pure numpy helper; the module's GPU tests are not collected here)
    n, p = 24, 22
    lib = _abi.load()
    cfg = H.make_config(n, "command", scenario=_abi.SCENARIO_FULLGAME, seed=5, half_time_cycles=150)
    types = (_abi.PlayerType * 18)()
    assert lib.s2d_generate_player_types(9, C.byref(cfg.sp), types, 18) == 0
    rng = np.random.default_rng(3)
    type_of = rng.integers(1, 18, size=p)
    type_of[0] = type_of[11] = 0
    sim = OL.OracleSim(cfg, "f64")
    sim.set_player_types(types, 18, type_of)
    sim.reset()
    decay = np.array([types[t].player_decay for t in type_of])
    emax = np.array([types[t].effort_max for t in type_of])
    emin = np.array([types[t].effort_min for t in type_of])
    k = p * 12
    seen_modes, offside_calls, marks_seen, pauses, prev = set(), 0, 0, 0, None
    for t in range(400):
        act = swarm_policy(sim.obs, p, rng, random_frac=0.2)
        sim.step(act.reshape(n, -1))
        s = sim.get_state_fg()
        P = s[:, :k].reshape(n, p, 12)
        assert np.isfinite(s).all()
        hit = P[:, :, 9] != 0  # (a collision multiplies the velocity by -0.1 afterwards: still below the limit)
        assert (np.hypot(P[:, :, 2], P[:, :, 3]) <= 1.05 * decay[None, :] + 1e-9).all() or hit.any()
        assert (P[:, :, 5] >= 0).all() and (P[:, :, 5] <= 8000.0).all()
        assert (P[:, :, 6] >= emin[None, :] - 1e-9).all() and (P[:, :, 6] <= emax[None, :] + 1e-9).all()
        assert (np.abs(P[:, :, 4]) <= 180.0).all()
        mode, side, marks = s[:, k + 8].astype(int), s[:, k + 9].astype(int), s[:, k + 16].astype(np.int64)
        seen_modes |= set(mode.tolist())
        dead = (mode != 2) & (mode != 1) & (mode != 8)
        assert (np.abs(s[dead, k]) <= 52.5 + 1e-9).all() and (np.abs(s[dead, k + 1]) <= 34.0 + 1e-9).all()
        paused = mode == 8  # AfterGoal: the ball lies in the goal, the clock stands still and nothing moves
        assert (np.abs(s[paused, k]) > 52.5).all() and (s[paused, k + 2:k + 4] == 0).all()
        if prev is not None:
            still = paused & (prev[:, k + 8] == 8)
            assert np.array_equal(s[still][:, :k].reshape(-1, p, 12)[:, :, 0:2], prev[still][:, :k].reshape(-1, p, 12)[:, :, 0:2])
            assert np.array_equal(s[still, k + 6], prev[still, k + 6]) and (s[still, k + 10] == prev[still, k + 10] + 1).all()
            kicked_off = (mode == 3) & (prev[:, k + 8] == 8) & (s[:, k + 7] == prev[:, k + 7])  # ... for 50 cycles, then the kick-off
            assert (prev[kicked_off, k + 10] == 49).all() and (side[kicked_off] != prev[kicked_off, k + 9]).all()
            pauses += int(kicked_off.sum())
        if (mode == 3).any():  # kick-off: everybody in its own half
            ko = mode == 3
            # (put back at -+player_size; a player-player collision may then push by < 0.6)
            assert (P[ko][:, :11, 0] <= 0.61).all() and (P[ko][:, 11:, 0] >= -0.61).all()
        prev = s
        assert (marks[mode != 2] == 0).all()
        left_bits, right_bits = marks & 0x7FF, marks >> 11
        assert ((left_bits == 0) | (right_bits == 0)).all()
        marks_seen += int((marks != 0).sum())
        for i in np.nonzero(dead & (s[:, k + 10] > 0))[0]:  # one cycle after the ruling the other side has been cleared
            others = P[i, :, 11] != side[i]
            d = np.hypot(P[i, others, 0] - s[i, k], P[i, others, 1] - s[i, k + 1])
            assert (d >= 9.15 - 0.61).all()  # (placed on the circle; a player-player collision may then push by < 0.6)
        offside_calls += int(((mode == 5) & (s[:, k + 10] == 0)).sum())
    assert {2, 3} <= seen_modes and marks_seen > 0
    assert 8 in seen_modes and pauses > 0, sorted(seen_modes)  # goals were scored and waited out


@pytest.mark.parametrize("collision_model", [0, 1])
def test_fullgame_c_oracle_equals_an_independent_python_twin(collision_model):
    """The FULLGAME spec (include/soccer2d.h) restated twice: oracle/s2d_oracle.c and tests/fullgame_twin.py (built on
    the Python oracle's building blocks).  They are stepped from the same state and compared one cycle ahead - every
    state of 300 cycles of swarm play with heterogeneous players, body actions, goals, restarts and offside calls."""
    import ctypes as C
    import fullgame_twin as T
    from oracle import soccer2d_oracle as O
    from test_gpu_fullgame import swarm_policy
    n, p, half = 5, 22, 120
    lib = _abi.load()
    cfg = H.make_config(n, "command", scenario=_abi.SCENARIO_FULLGAME, seed=21, half_time_cycles=half, goto_dist_thr=0.5,
                        collision_model=collision_model)
    types = (_abi.PlayerType * 18)()
    assert lib.s2d_generate_player_types(4, C.byref(cfg.sp), types, 18) == 0
    rng = np.random.default_rng(11)
    type_of = rng.integers(1, 18, size=p)
    type_of[0] = type_of[11] = 0
    sim = OL.OracleSim(cfg, "f64")
    sim.set_player_types(types, 18, type_of)
    sim.reset()
    base = O.ServerParam(**{k: getattr(cfg.sp, k) for k in _abi._SP_FIELDS if hasattr(O.ServerParam(), k)})
    sps = []
    for j in range(p):
        q = O.ServerParam(**vars(base))
        for k, v in types[type_of[j]].as_dict().items():
            if hasattr(q, k):
                setattr(q, k, float(v))
        sps.append(q)
    k = p * 12
    scale = np.concatenate([np.tile([52.5, 34.0, 1.05, 1.05, 180.0, 8000.0, 1.0, 1.0, 130600.0, 1, 1, 1], p),
                            [52.5, 34.0, 3.0, 3.0, 1], np.ones(12)])
    modes, calls, goals = set(), 0, 0
    for t in range(300):
        before, extra_before = sim.get_state_fg(), sim.get_extra_fg()
        act = swarm_policy(sim.obs, p, rng, random_frac=0.25)
        sim.step(act.reshape(n, -1))
        after, extra_after = sim.get_state_fg(), sim.get_extra_fg()
        for i in range(n):
            m = T.Match(before[i].tolist(), p)
            m.set_extra(extra_before[i])
            rw, done, res = T.cycle(m, act[i, 0], sps, base, int(cfg.seed), int(cfg.env_id_offset) + i,
                                    float(np.float32(cfg.goto_dist_thr)), half, collision_model)
            assert done == bool(sim.done[i]) and res == int(sim.result[i]), (t, i)
            assert rw == pytest.approx(float(sim.reward[i]), abs=1e-9), (t, i)
            if done:
                continue  # (the C side has already started the next match)
            assert m.extra() == extra_after[i].tolist(), (t, i)
            m.ep_return += rw
            got = np.array(m.vector())
            err = np.abs(got - after[i]) / scale
            err[4:k:12] = np.minimum(err[4:k:12], np.abs(360.0 / 180.0 - err[4:k:12]))  # body direction on the +-180 seam
            assert err.max() < 1e-9, (t, i, int(err.argmax()), got[int(err.argmax())], after[i][int(err.argmax())])
            modes.add(m.mode)
            calls += int(m.mode == 5 and m.timer == 0 and before[i][k + 8] == 2)
            goals += int(after[i][k + 11] + after[i][k + 12] > before[i][k + 11] + before[i][k + 12])
    assert {2, 3} <= modes
    if collision_model == 0:  # (the other model plays a different match from the same seed: the comparison above is the point)
        assert len(modes) >= 3 and goals > 0 and calls > 0, (sorted(modes), calls, goals)  # goals, kick-ins, an offside call


@pytest.mark.parametrize("collision_model", [0, 1])
def test_fullgame_c_oracle_equals_the_twin_on_random_states(collision_model):
    """The same comparison from hand-made states instead of a trajectory: players scattered or piled up around the ball,
    the ball near or beyond every line, every play mode with either side and timer, stale offside marks - one cycle
    with random commands.  Reaches the referee branches a swarm trajectory rarely visits."""
    import fullgame_twin as T
    from oracle import soccer2d_oracle as O
    n, p, half = 64, 22, 10 ** 6
    cfg = H.make_config(n, "command", scenario=_abi.SCENARIO_FULLGAME, seed=2, half_time_cycles=half, goto_dist_thr=0.5, auto_reset=0,
                        collision_model=collision_model)
    sim = OL.OracleSim(cfg, "f64")
    sim.reset()
    base = O.ServerParam(**{k: getattr(cfg.sp, k) for k in _abi._SP_FIELDS if hasattr(O.ServerParam(), k)})
    sps = [base] * p
    rng = np.random.default_rng(5)
    k = p * 12
    scale = np.concatenate([np.tile([52.5, 34.0, 1.05, 1.05, 180.0, 8000.0, 1.0, 1.0, 130600.0, 1, 1, 1], p),
                            [52.5, 34.0, 3.0, 3.0, 1], np.ones(12)])
    seen = set()
    for rnd in range(12):
        st = sim.get_state_fg()
        for i in range(n):
            kind = rng.integers(0, 5)
            bx = rng.choice([rng.uniform(-50, 50), 52.4, -52.4, 52.58, -52.58]) if kind else rng.uniform(-50, 50)
            by = rng.choice([rng.uniform(-30, 30), 33.95, -33.95, 34.08, -34.08, 3.0, -6.9]) if kind else rng.uniform(-30, 30)
            st[i, k:k + 4] = [bx, by, rng.uniform(-2.5, 2.5), rng.uniform(-2.5, 2.5)]
            P = st[i, :k].reshape(p, 12)
            P[:, 0] = rng.uniform(-52, 52, p)
            P[:, 1] = rng.uniform(-33, 33, p)
            crowd = rng.integers(0, 7)
            P[:crowd, 0] = bx + rng.uniform(-0.5, 0.5, crowd)  # a pile-up on the ball
            P[:crowd, 1] = by + rng.uniform(-0.5, 0.5, crowd)
            P[:, 2:4] = rng.uniform(-0.4, 0.4, (p, 2))
            P[:, 4] = rng.uniform(-180, 180, p)
            P[:, 5] = rng.uniform(0, 8000, p)
            P[:, 6] = rng.uniform(0.6, 1.0, p)
            P[:, 7] = rng.uniform(0.5, 1.0, p)
            mode = int(rng.choice([2, 2, 2, 3, 4, 5, 6, 7, 8]))
            if rng.uniform() < 0.15:  # the ball in front of a goalkeeper, inside its penalty area
                side = int(rng.integers(0, 2))
                P[11 * side, 0:2] = [(-1) ** (side + 1) * rng.uniform(40, 50), rng.uniform(-15, 15)]
                P[11 * side, 4] = rng.uniform(-180, 180)
                th = np.radians(P[11 * side, 4]) + rng.uniform(-0.3, 0.3)
                d = rng.uniform(0.2, 1.4)
                st[i, k:k + 2] = [P[11 * side, 0] + d * np.cos(th), P[11 * side, 1] + d * np.sin(th)]
            st[i, k + 5] = rnd  # step_number
            st[i, k + 8:k + 11] = [mode, int(rng.integers(1, 3)) if mode != 2 else 0, int(rng.choice([0, 5, 98, 99]))]
            st[i, k + 13] = int(rng.integers(0, 3))  # last touch
            st[i, k + 15] = 0
            marks = int(rng.integers(0, 1 << 11)) << (11 * int(rng.integers(0, 2))) if mode == 2 and rng.uniform() < 0.5 else 0
            st[i, k + 16] = marks
        sim.set_state_fg(st)
        extra = np.zeros((n, p + 2))
        extra[:, :p] = np.where(rng.uniform(size=(n, p)) < 0.1, rng.integers(1, 11, (n, p)), 0)  # some lie on the ground
        extra[:, p:] = np.where(rng.uniform(size=(n, 2)) < 0.2, rng.integers(1, 6, (n, 2)), 0)
        sim.set_extra_fg(extra)
        before = sim.get_state_fg()
        act = H.random_commands(rng, n * p).reshape(n, 1, p, 4)
        act[:, 0, :, 0] = np.where(rng.uniform(size=(n, p)) < 0.3, 3, act[:, 0, :, 0])  # more kicks
        act[:, 0, :, 0] = np.where(rng.uniform(size=(n, p)) < 0.1, 11, act[:, 0, :, 0])  # tackles into the pile-up
        act[:, 0, 0, 0] = np.where(rng.uniform(size=n) < 0.5, 12, act[:, 0, 0, 0])     # the keepers try to catch
        act[:, 0, 11, 0] = np.where(rng.uniform(size=n) < 0.5, 12, act[:, 0, 11, 0])
        sim.step(act.reshape(n, -1))
        after, extra_after = sim.get_state_fg(), sim.get_extra_fg()
        for i in range(n):
            m = T.Match(before[i].tolist(), p)
            m.set_extra(extra[i])
            rw, done, res = T.cycle(m, act[i, 0], sps, base, int(cfg.seed), i, 0.5, half, collision_model)
            assert m.extra() == extra_after[i].tolist(), (rnd, i)
            assert not done and rw == pytest.approx(float(sim.reward[i]), abs=1e-9), (rnd, i)
            m.ep_return += rw
            err = np.abs(np.array(m.vector()) - after[i]) / scale
            err[4:k:12] = np.minimum(err[4:k:12], np.abs(2.0 - err[4:k:12]))
            j = int(err.argmax())
            assert err.max() < 1e-9, (rnd, i, j, m.vector()[j], after[i][j], int(before[i][k + 8]))
            seen.add((int(before[i][k + 8]), m.mode))
    modes_after = {b for _, b in seen}
    # play on, kick-off (after the goal pause), kick-in, free kick (offside / catch), corner, goal kick, after goal
    assert {2, 4, 5, 6, 7, 8} <= modes_after, sorted(seen)
    assert (2, 8) in seen and (8, 8) in seen, sorted(seen)  # a goal stops the clock, and it stays stopped


def test_oracle_config_mirror_matches_the_product_struct_and_defaults():
    """tests/oracle_lib.OracleConfig (what bench.py --impl reference fills without touching the product) has the layout
    of soccer2d_b200._abi.Config, and s2do_default_config restates s2d_default_config for every scenario."""
    import ctypes as C

    from soccer2d_b200 import _abi
    assert C.sizeof(OL.OracleConfig) == C.sizeof(_abi.Config)
    for (na, ta), (nb, tb) in zip(OL.OracleConfig._fields_, _abi.Config._fields_):
        assert na == nb and C.sizeof(ta) == C.sizeof(tb)
        assert getattr(OL.OracleConfig, na).offset == getattr(_abi.Config, nb).offset
    assert [f for f, _ in OL.OracleServerParam._fields_] == [f for f, _ in _abi.ServerParam._fields_]
    lib = _abi.load()
    for scenario in (_abi.SCENARIO_REACHBALL, _abi.SCENARIO_SHOOT, _abi.SCENARIO_FULLGAME):
        mine = OL.default_config(1, scenario)
        theirs = _abi.Config()
        assert lib.s2d_default_config(C.byref(theirs), scenario) == 0
        assert bytes(mine) == bytes(theirs)


def test_oracle_envs_are_compact():
    """one-player scenarios: a 2^16-env oracle steps the same as before with the compact env records (np players each)"""
    cfg = OL.default_config(3000, action_mode=OL.ACT_DISCRETE, change_ball_velocity=1, seed=5)
    a, b = OL.OracleSim(cfg, "f64"), OL.OracleSim(cfg, "f32")
    a.reset(), b.reset()
    rng = np.random.default_rng(0)
    for _ in range(50):
        act = rng.integers(0, 16, size=(3000, 4)).astype(np.uint8)
        a.step(act, 4), b.step(act, 4)
    assert np.array_equal(a.done, b.done) and np.abs(a.obs - b.obs).max() < 1e-3


@pytest.mark.parametrize("collision_model", [0, 1])
def test_one_player_collision_models_c_equals_python(collision_model):
    """The ball-player contact of the one-player scenarios under both collision models (include/soccer2d.h): C f64 and the
    Python oracle's `collisions` agree from hand-placed overlaps (moving ball, ball at rest, coincident centres, a
    moving player), and under BACKTRACE the player is moved as well."""
    from oracle import soccer2d_oracle as O
    from test_gpu_parity_cases import HAND_PLACED_STATES
    cases = [c for c in HAND_PLACED_STATES[:4]] + [[0, 0, 0.2, 0, 0, 8000, 1, 1, 130600, 1.0, 0, -0.5, 0, 1, 5, 0, 3, 3, 1],      # head-on, both moving
             [0, 0, 0.3, 0.1, 0, 8000, 1, 1, 130600, 0.5, 0.2, 0, 0, 1, 5, 0, 3, 3, 1]]     # the player runs into a resting ball
    n = len(cases)
    cfg = H.make_config(n, "continuous", seed=1, min_distance_to_ball=0.05, max_steps=100000, collision_model=collision_model,
                        sp=dict(dash_power_rate=0.0))
    sim = OL.OracleSim(cfg, "f64")
    sim.reset()
    sp = O.ServerParam().as_f32()
    sp.dash_power_rate = 0.0
    moved_player, hits = False, 0
    for i, c in enumerate(cases):
        sim.set_state(i, np.array(c + [0], dtype=np.float64))
    sim.step(np.zeros((n, 1), np.float32))
    for i, c in enumerate(cases):
        p = O.Player(x=c[0], y=c[1], vx=c[2], vy=c[3], body=c[4], stamina=c[5], effort=c[6], recovery=c[7], capacity=c[8])
        b = O.Ball(x=c[9], y=c[10], vx=c[11], vy=c[12])
        O.obj_inc(p, sp.player_accel_max, sp.player_speed_max, sp.player_decay)
        O.obj_inc(b, sp.ball_accel_max, sp.ball_speed_max, sp.ball_decay)
        before = (p.x, p.y)
        O.collisions(b, [p], sp, collision_model)
        got = sim.get_state(i)
        assert b.collided == p.collided == (int(got[19]) & 3 == 3)
        assert np.allclose([p.x, p.y, p.vx, p.vy], got[0:4], rtol=0, atol=1e-12)
        assert np.allclose([b.x, b.y, b.vx, b.vy], got[9:13], rtol=0, atol=1e-12)
        assert math.hypot(b.x - p.x, b.y - p.y) >= sp.player_size + sp.ball_size
        hits += b.collided
        moved_player |= (p.x, p.y) != before
    assert hits >= 3 and moved_player == (collision_model == 1)
