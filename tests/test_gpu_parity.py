"""GPU parity tests (run on the B200 box with -m gpu).  Everything goes through the C ABI of libsoccer2d.so
(via the Python host) and is checked against the CPU oracle on identical seeded inputs:

  * bit-exact against the oracle's fp32 build (same operation order as the kernels): obs, reward, done,
    result, terminal observations, full state, collision flags, statistics;
  * against the oracle's f64 build (the "truth", rcssserver computes in double): flags bit-exact, floats
    within 1e-5 relative to each quantity's scale over 1 000 cycles (the north-star tolerance);
  * against the golden vectors produced by the reference's own reach_ball_env.py.
"""
import numpy as np
import pytest
import torch

import helpers as H
import oracle_lib as OL
from soccer2d_b200 import Soccer2DVecEnv, _abi
from test_gpu_parity_cases import HAND_PLACED_STATES

pytestmark = pytest.mark.gpu

MODE_KW = {
    "discrete": dict(use_continuous_action=False),
    "continuous": dict(use_continuous_action=True, use_turning=False),
    "turning": dict(use_continuous_action=True, use_turning=True),
}


def make_env(n, mode, **kw):
    args = dict(MODE_KW[mode])
    args.update(kw)
    return Soccer2DVecEnv(n, device="cuda:0", **args)


def gpu_state(env):
    """[N, 20] float64 in the order of oracle_lib.STATE_FIELDS."""
    f, u = env.state_planes()
    f = f.cpu().numpy().astype(np.float64)
    u = u.cpu().numpy().astype(np.int64)
    out = np.zeros((env.num_envs, 20))
    out[:, 0:4] = f[0]
    out[:, 4:8] = f[1]
    out[:, 8] = f[3][:, 2]
    out[:, 9:13] = f[2]
    out[:, 13:15] = f[3][:, 0:2]
    out[:, 15] = f[3][:, 3]
    out[:, 16:20] = u
    return out


def set_gpu_state(env, i, v):
    f, u = env.state_planes()
    v = [float(x) for x in v]
    f[0, i] = torch.tensor(v[0:4], device=f.device)
    f[1, i] = torch.tensor(v[4:8], device=f.device)
    f[2, i] = torch.tensor(v[9:13], device=f.device)
    f[3, i] = torch.tensor([v[13], v[14], v[8], v[15]], device=f.device)
    u[i, :3] = torch.tensor([int(v[16]), int(v[17]), int(v[18])], dtype=torch.int32, device=u.device)


def assert_same_step(env, sim, exact=True, angle_scale=1.0):
    obs, rew = env.obs.cpu().numpy(), env.reward.cpu().numpy()
    done, res = env.done_u8.cpu().numpy(), env.result.cpu().numpy()
    assert np.array_equal(done, sim.done) and np.array_equal(res, sim.result)
    if exact:
        assert np.array_equal(obs, sim.obs)
        assert np.array_equal(rew, sim.reward)
        if env.terminal_obs is not None:
            d = done.astype(bool)
            assert np.array_equal(env.terminal_obs.cpu().numpy()[d], sim.term_obs[d])
    else:
        assert H.obs_close(obs, sim.obs, angle_scale=angle_scale) < H.TOL
        assert np.abs(rew - sim.reward).max() < H.TOL * 100.0
    return int(done.sum())


@pytest.mark.parametrize("mode", ["discrete", "continuous", "turning"])
def test_bit_exact_against_fp32_oracle_1000_cycles(mode):
    n = 1000  # ragged: not a multiple of the warp or the block
    env = make_env(n, mode, seed=7, change_ball_velocity=True, terminal_obs=True)
    sim = OL.OracleSim(env.cfg, "f32")
    assert np.array_equal(env.reset(), sim.reset())
    assert np.array_equal(gpu_state(env), sim.get_state())
    rng = np.random.default_rng(0)
    episodes = 0
    for t in range(1000):
        act = H.random_actions(rng, mode, n)
        env.step_torch(torch.from_numpy(act))
        sim.step(act)
        episodes += assert_same_step(env, sim)
        if t % 100 == 99:
            assert np.array_equal(gpu_state(env), sim.get_state())
    assert episodes > 2 * n
    st, so = env.stats(), sim.stats(_abi.Stats())
    for k in ("episodes", "goals", "outs", "timeouts", "episode_steps", "env_steps"):
        assert st[k] == getattr(so, k), k
    assert st["return_sum"] == pytest.approx(so.return_sum, rel=1e-9)
    assert st["episodes"] == st["goals"] + st["outs"] + st["timeouts"] == episodes
    env.close()


@pytest.mark.parametrize("mode", ["discrete", "continuous", "turning"])
def test_non_default_server_param_kernels_bit_exact(mode):
    """A non-default ServerParam selects the kernels that read their constants from the constant bank (the default
    one uses compile-time constants): 45-degree dash_angle_step, slowness on top, other decays / rates."""
    n = 600
    sp = dict(dash_angle_step=45.0, slowness_on_top_for_left_team=1.25, player_decay=0.5, ball_decay=0.9,
              side_dash_rate=0.5, back_dash_rate=0.6, stamina_inc_max=30.0, player_speed_max=0.8, inertia_moment=3.0)
    env = make_env(n, mode, seed=11, change_ball_velocity=True, terminal_obs=True, max_steps=80, server_param=sp)
    sim = OL.OracleSim(env.cfg, "f32")
    assert np.array_equal(env.reset(), sim.reset())
    rng = np.random.default_rng(0)
    episodes = 0
    for t in range(400):
        act = H.random_actions(rng, mode, n)
        env.step_torch(torch.from_numpy(act))
        sim.step(act)
        episodes += assert_same_step(env, sim)
    assert np.array_equal(gpu_state(env), sim.get_state())
    assert episodes > n
    env.close()


# fp32 thresholds (`dist < 5`, `|x| > 52.5`) can fall on the other side of the f64 truth's: measured over all 2^20 envs x
# 1 000 cycles of the bench workload this happens about 2.5e-6 times per env-step (profiles/parity_flips.json).  After such
# a flip the two runs of that env are different episodes, so the env leaves the comparison; the test bounds the RATE on
# seeds that were not looked at beforehand instead of demanding zero flips on hand-picked ones.
MAX_FLIP_RATE = 2.0e-5


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("mode", ["discrete", "continuous", "turning"])
def test_against_f64_truth_1000_cycles(mode, seed):
    """north star: flags bit-exact, floats within 1e-5 relative over 1 000 cycles against the double oracle - for every
    env until (if ever) an fp32 threshold flips one of its flags; the flip rate is bounded."""
    n = 1024
    env = make_env(n, mode, seed=seed, change_ball_velocity=True)
    sim = OL.OracleSim(env.cfg, "f64")
    assert H.obs_close(env.reset(), sim.reset()) < H.TOL
    rng = np.random.default_rng(100 + seed)
    alive = np.ones(n, bool)
    episodes = compared = 0
    for _ in range(1000):
        act = H.random_actions(rng, mode, n)
        env.step_torch(torch.from_numpy(act))
        sim.step(act)
        done, res = env.done_u8.cpu().numpy(), env.result.cpu().numpy()
        compared += int(alive.sum())
        alive &= (done == sim.done) & (res == sim.result)
        assert H.obs_close(env.obs.cpu().numpy()[alive], sim.obs[alive]) < H.TOL
        assert np.abs(env.reward.cpu().numpy()[alive] - sim.reward[alive]).max() < H.TOL * 100.0
        episodes += int(done[alive].sum())
    assert episodes > n
    flips = int((~alive).sum())
    assert flips <= max(2, MAX_FLIP_RATE * compared), (flips, compared)
    g, o = gpu_state(env)[alive], sim.get_state()[alive]
    assert np.array_equal(g[:, 16:], o[:, 16:])  # step_number, cycle, episode, collision flags
    assert H.state_err(g, o) < H.TOL
    env.close()


@pytest.mark.parametrize("mode,k", [("discrete", 16), ("discrete", 5), ("continuous", 8), ("continuous", 3), ("turning", 4)])
def test_fused_substeps_equal_oracle_and_single_steps(mode, k):
    n = 777
    kw = dict(seed=3, change_ball_velocity=True, max_steps=40, terminal_obs=True)
    env = make_env(n, mode, substeps=k, **kw)
    one = make_env(n, mode, substeps=1, **kw)
    sim = OL.OracleSim(env.cfg, "f32")
    env.reset_torch()
    one.reset_torch()
    sim.reset()
    rng = np.random.default_rng(9)
    for _ in range(25):
        act = H.random_actions(rng, mode, n, k)
        env.step_torch(torch.from_numpy(act))
        sim.step(act, k)
        assert_same_step(env, sim)
        for j in range(k):
            one.step_torch(torch.from_numpy(np.ascontiguousarray(act[:, j:j + 1])))
        assert torch.equal(env.obs, one.obs) and torch.equal(env.state, one.state)
    assert np.array_equal(gpu_state(env), sim.get_state())
    a, b = env.stats(), one.stats()
    assert a.pop("return_sum") == pytest.approx(b.pop("return_sum"), rel=1e-9)  # double atomics: order varies
    assert a == b
    env.close()
    one.close()


@pytest.mark.parametrize("step", [22.5, 0.0, 1.0])
def test_sincos_memo_hits_and_misses_bit_exact(step):
    """Launches of K >= 4 ReachBall Discrete(n) cycles read the dash's sin / cos from the whole-degree memo
    (s2d_math.cuh: sincos_deg_memo) and fall back to the polynomial for any other angle.  dash_angle_step = 22.5 or 0
    (no snapping) makes half of the 16 directions fractional, so warps mix hits and misses; step 1 with a few bodies
    pushed off the whole degrees does the same from the state's side.  All of it must equal the oracle (which only
    knows the polynomial) and the K = 1 launches (which never read the memo) bit for bit."""
    n, k = 640, 8
    kw = dict(seed=17, change_ball_velocity=True, max_steps=60, terminal_obs=True)
    if step != 1.0:
        kw["server_param"] = dict(dash_angle_step=step)
    env = make_env(n, "discrete", substeps=k, **kw)
    one = make_env(n, "discrete", substeps=1, **kw)
    sim = OL.OracleSim(env.cfg, "f32")
    env.reset_torch()
    one.reset_torch()
    sim.reset()
    if step == 1.0:
        st = sim.get_state()
        for i in range(0, n, 3):
            v = st[i].copy()
            v[4] = float(np.float32(v[4] * 0.5 + 0.37))  # body: off the whole degrees
            set_gpu_state(env, i, v)
            set_gpu_state(one, i, v)
            sim.set_state(i, v)
    rng = np.random.default_rng(4)
    for _ in range(20):
        act = H.random_actions(rng, "discrete", n, k)
        env.step_torch(torch.from_numpy(act))
        sim.step(act, k)
        assert_same_step(env, sim)
        for j in range(k):
            one.step_torch(torch.from_numpy(np.ascontiguousarray(act[:, j:j + 1])))
        assert torch.equal(env.obs, one.obs) and torch.equal(env.state, one.state)
    assert np.array_equal(gpu_state(env), sim.get_state())
    env.close()
    one.close()


def test_shards_reproduce_the_global_run():
    """RNG is keyed on the GLOBAL env id: 3 shards (as 3 ranks would hold) == one big env."""
    n, k = 3000, 4
    kw = dict(seed=99, change_ball_velocity=True, max_steps=30, substeps=k)
    whole = make_env(n, "discrete", **kw)
    bounds = [0, 1000, 1900, 3000]
    parts = [make_env(b - a, "discrete", env_id_offset=a, **kw) for a, b in zip(bounds[:-1], bounds[1:])]
    whole.reset_torch()
    for p in parts:
        p.reset_torch()
    rng = np.random.default_rng(2)
    for _ in range(20):
        act = H.random_actions(rng, "discrete", n, k)
        whole.step_torch(torch.from_numpy(act))
        for p, a, b in zip(parts, bounds[:-1], bounds[1:]):
            p.step_torch(torch.from_numpy(act[a:b]))
    assert torch.equal(whole.obs, torch.cat([p.obs for p in parts]))
    assert torch.equal(whole.reward, torch.cat([p.reward for p in parts]))
    assert np.array_equal(gpu_state(whole), np.concatenate([gpu_state(p) for p in parts]))
    tot = {k_: sum(p.stats()[k_] for p in parts) for k_ in ("episodes", "goals", "outs", "timeouts", "episode_steps")}
    assert all(whole.stats()[k_] == v for k_, v in tot.items())


def test_masked_reset_and_no_auto_reset():
    n = 300
    env = make_env(n, "discrete", seed=4, auto_reset=False, max_steps=10, change_ball_velocity=True)
    sim = OL.OracleSim(env.cfg, "f32")
    env.reset_torch()
    sim.reset()
    rng = np.random.default_rng(1)
    finished = np.zeros(n, bool)
    for t in range(40):
        act = H.random_actions(rng, "discrete", n)
        env.step_torch(torch.from_numpy(act))
        sim.step(act)
        assert_same_step(env, sim)
        g = gpu_state(env)
        assert np.array_equal(g, sim.get_state())
        # an env that finished keeps its DONE flag until it is reset (the reference keeps simulating too)
        finished |= env.done.cpu().numpy()
        assert np.array_equal((g[:, 19].astype(int) & _abi.FLAG_DONE) != 0, finished)
        if t % 7 == 6:
            mask = env.done.clone()
            finished &= ~mask.cpu().numpy()
            before = env.obs.clone()
            env.reset_torch(mask)
            sim.reset(mask.cpu().numpy())
            assert np.array_equal(env.obs.cpu().numpy(), sim.obs)
            keep = ~mask
            assert torch.equal(env.obs[keep], before[keep])
            assert np.array_equal(gpu_state(env), sim.get_state())


@pytest.mark.parametrize("n", [1, 31, 33, 257])
def test_ragged_sizes(n):
    env = make_env(n, "continuous", seed=n, terminal_obs=True, max_steps=7)
    sim = OL.OracleSim(env.cfg, "f32")
    canary = env.obs.clone()
    assert np.array_equal(env.reset(), sim.reset())
    rng = np.random.default_rng(n)
    for _ in range(30):
        act = H.random_actions(rng, "continuous", n)
        env.step_torch(torch.from_numpy(act))
        sim.step(act)
        assert_same_step(env, sim)
    assert canary.shape == env.obs.shape


@pytest.mark.parametrize("collision_model", ["midpoint", "backtrace"])
def test_collisions_stamina_and_boundaries_match_oracle(collision_model):
    """Hand-placed states: ball inside the player (moving and at rest, centres coincident), player at the
    pitch edge, exhausted stamina (effort / recovery decay), capacity nearly used up - under both collision models."""
    cases = HAND_PLACED_STATES + [[0, 0, 0.2, 0, 0, 8000, 1, 1, 130600, 1.0, 0, -0.5, 0, 1, 5, 0, 3, 3, 1],      # head-on, both moving
             [0, 0, 0.3, 0.1, 0, 8000, 1, 1, 130600, 0.5, 0.2, 0, 0, 1, 5, 0, 3, 3, 1]]     # the player runs into a resting ball
    n = len(cases)
    env = make_env(n, "continuous", seed=1, min_distance_to_ball=0.05, max_steps=100000, terminal_obs=True,
                   collision_model=collision_model)
    sim = OL.OracleSim(env.cfg, "f32")
    env.reset_torch()
    sim.reset()
    for i, c in enumerate(cases):
        set_gpu_state(env, i, c + [0])
        sim.set_state(i, np.array(c + [0], dtype=np.float64))
    rng = np.random.default_rng(4)
    seen_flags = 0
    for _ in range(60):
        act = H.random_actions(rng, "continuous", n)
        act[6:8] = 0.0  # straight ahead: burns stamina as fast as possible
        env.step_torch(torch.from_numpy(act))
        sim.step(act)
        assert_same_step(env, sim)
        g = gpu_state(env)
        assert np.array_equal(g, sim.get_state())
        seen_flags |= int(np.bitwise_or.reduce(g[:, 19].astype(int)))
    assert seen_flags & _abi.FLAG_BALL_COLLIDED and seen_flags & _abi.FLAG_PLAYER_COLLIDED
    st = env.stats()
    assert st["outs"] >= 2  # the two edge cases left the pitch


def test_golden_obs_and_reward_from_the_reference_code(golden):
    """The reference's own state_to_observation / check_trainer_observation outputs (tests/golden) against
    the GPU.  A step is made the identity on the golden state: the ball starts one velocity behind with
    ball_decay = 1, and dash_power_rate = 0 keeps the player in place."""
    sp = dict(ball_decay=1.0, dash_power_rate=0.0, player_size=0.0, ball_size=0.0)  # sizes 0: no collisions
    # --- observations
    rows = golden["obs"]
    n = len(rows)
    env = make_env(n, "continuous", seed=0, server_param=sp, max_steps=10**6, min_distance_to_ball=0.0,
                   auto_reset=False)
    env.reset_torch()
    for i, r in enumerate(rows):
        bx, by, bvx, bvy, px, py, body = r["state"]
        v = [px, py, 0, 0, body, 8000, 1, 1, 130600, np.float32(bx) - np.float32(bvx), np.float32(by) - np.float32(bvy),
             bvx, bvy, 0, 0, 0, 0, 1, 1, 0]
        set_gpu_state(env, i, v)
    env.step_torch(torch.zeros((n, 1)))
    got = env.obs.cpu().numpy()
    want = np.array([r["obs"] for r in rows])
    assert H.obs_close(got, want) < 2.0e-6
    # --- reward / done / info chains (incl. the +10 on Out quirk and the 201st-step Timeout)
    for chain in golden["reward"]:
        steps = chain["steps"]
        kw = chain["kwargs"]
        env = make_env(1, "continuous", seed=0, server_param=sp, max_steps=kw["max_steps"],
                       min_distance_to_ball=kw["min_distance_to_ball"], auto_reset=False)
        env.reset_torch()
        mem_d, mem_a = 0.0, 0.0
        for s in steps:
            bx, by, _, _, px, py, body = s["state"]
            set_gpu_state(env, 0, [px, py, 0, 0, body, 8000, 1, 1, 130600, bx, by, 0, 0, mem_d, mem_a, 0,
                                   s["step_number"] - 1, 1, 1, 0])
            env.step_torch(torch.zeros((1, 1)))
            assert bool(env.done[0]) == s["done"]
            assert _abi.RESULT_NAMES[int(env.result[0])] == s["result"]
            assert float(env.reward[0]) == pytest.approx(s["reward"], abs=2e-4)  # distances ~100 m in fp32
            mem_d, mem_a = s["mem_dist"], s["mem_ang"]
            snap = env.export_env(0)
            assert snap.mem_distance_to_ball == pytest.approx(mem_d, rel=1e-6, abs=1e-5)
            d = abs(snap.mem_body_ball_angle_diff - mem_a)
            assert min(d, 360 - d) < 2e-4 or s["mem_dist"] < 1e-3
        env.close()


def test_decode_table_matches_reference_golden(golden):
    """Discrete(n) -> Dash direction (reach_ball_env.py:84) for n in {16, 8, 36, 7}: one dash from rest moves
    the player along body + rint(direction)."""
    for n_act in (16, 8, 36, 7):
        rows = [r for r in golden["decode"]["discrete"] if r["n"] == n_act]
        env = make_env(n_act, "discrete", action_space_size=n_act, seed=1, max_steps=10**6, min_distance_to_ball=0.0)
        env.reset_torch()
        for a in range(n_act):
            set_gpu_state(env, a, [0, 0, 0, 0, 0, 8000, 1, 1, 130600, 30, 30, 0, 0, 0, 0, 0, 0, 1, 1, 0])
        env.step_torch(torch.arange(n_act, dtype=torch.uint8).view(-1, 1))
        g = gpu_state(env)
        for a in range(n_act):
            want = np.rint(np.float32(rows[a]["dir"]))  # dash_angle_step = 1 snaps to whole degrees
            got = np.degrees(np.arctan2(g[a, 1], g[a, 0]))
            d = abs(got - want)
            assert min(d, 360 - d) < 1e-3, (n_act, a)
        env.close()


def test_step_host_and_gym_api_match_torch_path():
    from sample_environments.environment_factory import EnvironmentFactory

    n = 513
    a = make_env(n, "discrete", seed=8, change_ball_velocity=True, max_steps=25, terminal_obs=True)
    b = make_env(n, "discrete", seed=8, change_ball_velocity=True, max_steps=25, terminal_obs=True)
    assert np.array_equal(a.reset(), b.reset())
    rng = np.random.default_rng(3)
    dones = 0
    for _ in range(60):
        act = H.random_actions(rng, "discrete", n)
        a.step_torch(torch.from_numpy(act))
        obs, rew, done, infos = b.step(act[:, 0])
        assert np.array_equal(obs, a.obs.cpu().numpy()) and np.array_equal(rew, a.reward.cpu().numpy())
        assert np.array_equal(done, a.done.cpu().numpy())
        for i in np.nonzero(done)[0]:
            assert infos[i]["result"] in ("Goal", "Out", "Timeout")
            assert np.array_equal(infos[i]["terminal_observation"], a.terminal_obs[i].cpu().numpy())
            dones += 1
        assert all(infos[i]["result"] is None for i in np.nonzero(~done)[0])
    assert dones > n

    # single-env gym API (old 4-tuple), against the fp32 oracle with auto_reset off
    kwargs = dict(use_continuous_action=False, action_space_size=16, max_steps=30, change_ball_velocity=True)
    env = EnvironmentFactory().create("ReachBall", None, None, "/tmp", seed=21, **kwargs)
    assert env.action_space.n == 16 and env.observation_space.shape == (10,)
    sim = OL.OracleSim(env._vec.cfg, "f32")
    for episode in range(3):
        obs = env.reset()
        assert obs.shape == (10,) and np.array_equal(obs, sim.reset(np.ones(1, np.uint8))[0])
        done = False
        steps = 0
        while not done:
            act = int(rng.integers(16))
            obs, reward, done, info = env.step(act)
            sim.step(np.array([[act]], np.uint8))
            assert np.array_equal(obs, sim.obs[0]) and reward == float(sim.reward[0]) and done == bool(sim.done[0])
            assert info == {"result": _abi.RESULT_NAMES[int(sim.result[0])]}
            steps += 1
        assert info["result"] in ("Goal", "Out", "Timeout") and steps <= 31
    env.close()
    env.close()  # idempotent, the reference scripts call it twice (dqn_stable_baselines3.py:73,86)


def test_unbound_and_bad_arguments_fail_loudly():
    import ctypes as C
    lib = _abi.load()
    h = C.c_void_p()
    cfg = H.make_config(64)
    assert lib.s2d_create(C.byref(cfg), C.byref(h)) == 0
    assert lib.s2d_step(h, 1, None) == _abi.S2D_ERR_UNBOUND
    assert lib.s2d_reset(h, None, None) == _abi.S2D_ERR_UNBOUND
    bufs = _abi.Buffers()
    assert lib.s2d_bind(h, C.byref(bufs)) == _abi.S2D_ERR_UNBOUND
    assert b"required" in lib.s2d_last_error(h)
    assert lib.s2d_destroy(h) == 0
    env = make_env(64, "discrete")
    assert lib.s2d_step(env.handle, 0, None) == _abi.S2D_ERR_INVALID
    cfg = H.make_config(64, device=99)
    assert lib.s2d_create(C.byref(cfg), C.byref(h)) == _abi.S2D_ERR_INVALID


def test_full_size_properties_1m_envs():
    """BASELINE configs[1] size (2^20 envs, K = 16): size-independent properties instead of the oracle."""
    n, k = 1 << 20, 16
    kw = dict(seed=0, change_ball_velocity=True, substeps=k)
    a = make_env(n, "discrete", **kw)
    b = make_env(n, "discrete", substeps=1, seed=0, change_ball_velocity=True)
    a.reset_torch()
    b.reset_torch()
    assert torch.equal(a.obs, b.obs)
    g = torch.Generator(device="cuda").manual_seed(0)
    total_done = 0
    for _ in range(14):  # 224 cycles: every env passes its 200-step timeout at least once
        act = torch.randint(0, 16, (n, k), dtype=torch.uint8, device="cuda", generator=g)
        a.step_torch(act)
        for j in range(k):
            b.step_torch(act[:, j:j + 1])
        total_done += int(a.done.sum())
    # determinism + fusion: 16 fused substeps == 16 single-step launches, bit for bit, on all 2^20 envs
    assert torch.equal(a.state, b.state) and torch.equal(a.obs, b.obs)
    f, u = a.state_planes()
    sp = a.cfg.sp
    speed = torch.hypot(f[0, :, 2], f[0, :, 3])
    assert float(speed.max()) <= sp.player_speed_max * sp.player_decay * (1 + 1e-6)
    assert float(torch.hypot(f[2, :, 2], f[2, :, 3]).max()) <= 3.0
    assert float(f[1, :, 1].min()) >= 0.0 and float(f[1, :, 1].max()) <= sp.stamina_max
    assert float(f[1, :, 2].min()) >= sp.effort_min and float(f[1, :, 2].max()) <= sp.effort_max
    assert float(f[1, :, 0].abs().max()) <= 180.0
    assert int(u[:, 0].max()) <= 200 and int(u[:, 0].min()) >= 0          # step_number
    assert int(u[:, 1].min()) >= 224                                       # cycle counts the reset cycles too
    assert bool(torch.isfinite(a.obs).all()) and float(a.obs[:, 0:2].abs().max()) <= 1.0
    st = a.stats()
    assert st["episodes"] == st["goals"] + st["outs"] + st["timeouts"] >= n
    assert st["env_steps"] == n * k * 14 and st == b.stats() | {"return_sum": st["return_sum"]}
    assert st["return_sum"] == pytest.approx(b.stats()["return_sum"], rel=1e-9)
    assert int((u[:, 2] - 1).sum()) == st["episodes"]                      # episode counters add up
    assert total_done > 0


# ---- command actions (proto PlayerAction vocabulary) and the SHOOT scenario (BASELINE configs[2]) -------------

def test_command_mode_reachball_bit_exact():
    n = 700
    env = Soccer2DVecEnv(n, device="cuda:0", seed=9, use_command_action=True, goto_dist_thr=0.4, terminal_obs=True,
                         change_ball_velocity=True, max_steps=120, min_distance_to_ball=0.3)
    sim = OL.OracleSim(env.cfg, "f32")
    assert np.array_equal(env.reset(), sim.reset())
    rng = np.random.default_rng(0)
    flags = 0
    for t in range(500):
        act = H.random_commands(rng, n)
        if t % 2:
            act = np.where(rng.uniform(size=(n, 1, 1)) < 0.7, H.chase_and_shoot(sim.obs), act).astype(np.float32)
        env.step_torch(torch.from_numpy(act))
        sim.step(act)
        assert_same_step(env, sim)
        if t % 50 == 49:
            g = gpu_state(env)
            assert np.array_equal(g, sim.get_state())
            flags |= int(np.bitwise_or.reduce(g[:, 19].astype(int)))
    assert env.stats()["episodes"] > n


@pytest.mark.parametrize("mode,k", [("discrete", 1), ("discrete", 16), ("command", 1), ("command", 4)])
def test_shoot_bit_exact_against_fp32_oracle(mode, k):
    n = 900
    env = Soccer2DVecEnv(n, scenario="shoot", device="cuda:0", seed=21, substeps=k, terminal_obs=True, max_steps=150,
                         use_command_action=mode == "command")
    sim = OL.OracleSim(env.cfg, "f32")
    assert env.action_space.shape == ((4,) if mode == "command" else ()) and env.obs.shape == (n, 10)
    assert np.array_equal(env.reset(), sim.reset())
    rng = np.random.default_rng(1)
    for t in range(640 // k):
        if mode == "discrete":
            act = rng.integers(0, 24, size=(n, k)).astype(np.uint8)
        else:
            act = np.repeat(H.chase_and_shoot(sim.obs, rng, kick_prob=0.9), k, axis=1)
            rnd = H.random_commands(rng, n, k)
            act = np.where(rng.uniform(size=(n, k, 1)) < 0.1, rnd, act).astype(np.float32)
        env.step_torch(torch.from_numpy(act))
        sim.step(act, k)
        assert_same_step(env, sim)
    assert np.array_equal(gpu_state(env), sim.get_state())
    st, so = env.stats(), sim.stats(_abi.Stats())
    assert (st["episodes"], st["goals"], st["outs"], st["timeouts"]) == (so.episodes, so.goals, so.outs, so.timeouts)
    assert st["episodes"] > 0
    if mode == "command":
        assert st["goals"] > 50 and st["outs"] > 0


def test_shoot_against_f64_truth():
    """Against the double oracle, cycle by cycle: flags bit-exact, floats within 1e-5.
    Dribbling is chaotic - every kick multiplies a position error by about the ball's travel (x16) - so over many
    kicks NO fixed tolerance can hold between an fp32 and an f64 run.  The double oracle is therefore re-synchronised
    to the GPU state before every cycle and the two are compared one cycle ahead, for every state the policy visits
    (kicks, collisions, goals, resets included)."""
    n = 128
    env = Soccer2DVecEnv(n, scenario="shoot", device="cuda:0", seed=77, use_command_action=True, max_steps=150)
    sim = OL.OracleSim(env.cfg, "f64")
    assert H.obs_close(env.reset(), sim.reset()) < H.TOL
    goals = kicks = 0
    for t in range(400):
        g = gpu_state(env)
        for i in range(n):
            sim.set_state(i, g[i])
        act = H.chase_and_shoot(env.obs.cpu().numpy())
        env.step_torch(torch.from_numpy(act))
        sim.step(act)
        # next to the ball the direction to it is ill-conditioned (position error / distance): scale those columns
        assert_same_step(env, sim, exact=False, angle_scale=0.02)
        goals += int((sim.result == 1).sum())
        g2, o2 = gpu_state(env), sim.get_state()
        kicks += int(((g2[:, 19].astype(int) & _abi.FLAG_KICKED) != 0).sum())
        assert np.array_equal(g2[:, 16:], o2[:, 16:])  # step_number, cycle, episode, collision / kick flags
        assert H.state_err(g2, o2) < H.TOL
    assert goals > 20 and kicks > 200


def test_shoot_goal_line_and_kickable_edge_cases():
    line = 52.5 + 0.085
    base = [0, 0, 0, 0, 0, 8000, 1, 1, 130600]
    cases = [
        base + [line - 0.5, 6.9, 1.0, 0.1, 50, 1, 0, 3, 3, 1],
        base + [line - 0.5, 7.0, 1.0, 0.2, 50, 1, 0, 3, 3, 1],
        base + [line - 0.5, 0.0, 0.4, 0.0, 50, 1, 0, 3, 3, 1],
        base + [-line + 0.5, 0.0, -1.0, 0.0, 50, 104, 0, 3, 3, 1],
        base + [10, 33.9, 0.0, 0.5, 40, 45, 0, 3, 3, 1],
        [20, 10, 0, 0, 30, 8000, 1, 1, 130600, 20 + 1.08, 10, 0, 0, 1.08, 34, 0, 3, 3, 1],
        [20, 10, 0, 0, 30, 8000, 1, 1, 130600, 20 + 1.09, 10, 0, 0, 1.09, 34, 0, 3, 3, 1],
        [51, 0, 0, 0, 0, 8000, 1, 1, 130600, 51.5, 0, 0, 0, 0.5, 1, 0, 3, 3, 1],
    ]
    n = len(cases)
    env = Soccer2DVecEnv(n, scenario="shoot", device="cuda:0", seed=2, use_command_action=True, max_steps=50)
    sim = OL.OracleSim(env.cfg, "f32")
    env.reset_torch()
    sim.reset()
    for i, c in enumerate(cases):
        set_gpu_state(env, i, c + [0])
        sim.set_state(i, np.array(c + [0], dtype=np.float64))
    act = np.zeros((n, 1, 4), np.float32)
    act[5:, 0, :3] = [3, 100, 0]
    env.step_torch(torch.from_numpy(act))
    sim.step(act)
    assert_same_step(env, sim)
    assert env.result.cpu().tolist()[:5] == [1, 2, 0, 2, 2]  # Goal, Out (outside the post), -, Out (own goal), Out (side)
    g = gpu_state(env)
    assert np.array_equal(g, sim.get_state())
    assert int(g[5, 19]) & _abi.FLAG_KICKED and not int(g[6, 19]) & _abi.FLAG_KICKED
    snap = env.export_env(5)
    assert snap.players[0].kicked == 1 and snap.num_players == 1


def test_shoot_gym_api_and_factory():
    from sample_environments.environment_factory import EnvironmentFactory
    env = EnvironmentFactory().create("Shoot", None, None, "/tmp", seed=3, max_steps=40)
    assert env.action_space.n == 24 and env.observation_space.shape == (10,)
    sim = OL.OracleSim(env._vec.cfg, "f32")
    rng = np.random.default_rng(0)
    obs = env.reset()
    assert np.array_equal(obs, sim.reset(np.ones(1, np.uint8))[0])
    done = False
    while not done:
        a = int(rng.integers(24))
        obs, reward, done, info = env.step(a)
        sim.step(np.array([[a]], np.uint8))
        assert np.array_equal(obs, sim.obs[0]) and reward == float(sim.reward[0])
    assert info["result"] in ("Goal", "Out", "Timeout")
    env.close()
    vec = EnvironmentFactory().create_vec("shoot", 64, device="cuda:0", use_command_action=True)
    assert vec.actions.shape == (64, 1, 4)
    vec.close()


def test_pipelined_host_steps_equal_synchronous_ones():
    """s2d_bind_pipeline / s2d_submit_host / s2d_wait_host: two steps in flight, results identical to step_host."""
    n, k = 5000, 4
    kw = dict(device="cuda:0", seed=12, substeps=k, use_continuous_action=False, change_ball_velocity=True, max_steps=30)
    a, b = Soccer2DVecEnv(n, **kw), Soccer2DVecEnv(n, **kw)
    assert np.array_equal(a.reset(), b.reset())
    rng = np.random.default_rng(0)
    acts = [torch.from_numpy(H.random_actions(rng, "discrete", n, k)).pin_memory() for _ in range(12)]
    want = []
    for act in acts:
        o, r, d, res = a.step_host(act)
        want.append((o.copy(), r.copy(), d.copy(), res.copy()))
    tickets = []
    got = []
    for i, act in enumerate(acts):
        tickets.append(b.submit_host(act))
        if i >= 1:
            o, r, d, res = b.wait_host(tickets[i - 1])
            got.append((o.copy(), r.copy(), d.copy(), res.copy()))
    o, r, d, res = b.wait_host(tickets[-1])
    got.append((o.copy(), r.copy(), d.copy(), res.copy()))
    for w, g in zip(want, got):
        for x, y in zip(w, g):
            assert np.array_equal(x, y)
    torch.cuda.synchronize()
    assert torch.equal(a.state, b.state) and a.stats() == b.stats() | {"return_sum": a.stats()["return_sum"]}


@pytest.mark.parametrize("scenario,mode,k", [("reachball", "discrete", 1), ("reachball", "discrete", 16), ("reachball", "turning", 1),
                                             ("shoot", "command", 1)])
def test_noise_on_bit_exact_and_reproducible(scenario, mode, k):
    """player_rand / ball_rand / kick_rand noise from Philox keyed on (seed, env, cycle, agent): the GPU equals the fp32
    oracle bit for bit, K fused cycles equal K single ones, and a shard reproduces its slice of the global run."""
    n = 640
    kw = dict(seed=31, noise=True, terminal_obs=True, max_steps=100, substeps=k)
    if scenario == "shoot":
        env = Soccer2DVecEnv(n, scenario="shoot", device="cuda:0", use_command_action=True, **kw)
        shard = Soccer2DVecEnv(200, scenario="shoot", device="cuda:0", use_command_action=True, env_id_offset=300, **kw)
    else:
        env = make_env(n, mode, change_ball_velocity=True, **kw)
        shard = make_env(200, mode, change_ball_velocity=True, env_id_offset=300, **kw)
    sim = OL.OracleSim(env.cfg, "f32")
    assert np.array_equal(env.reset(), sim.reset())
    shard.reset_torch()
    rng = np.random.default_rng(0)
    for t in range(320 // k):
        if mode == "command":
            act = H.chase_and_shoot(sim.obs, rng, kick_prob=0.9)
            act = np.where(rng.uniform(size=(n, 1, 1)) < 0.15, H.random_commands(rng, n), act).astype(np.float32)
        else:
            act = H.random_actions(rng, mode, n, k)
        env.step_torch(torch.from_numpy(act))
        shard.step_torch(torch.from_numpy(np.ascontiguousarray(act[300:500])))
        sim.step(act, k)
        assert_same_step(env, sim)
    assert np.array_equal(gpu_state(env), sim.get_state())
    assert torch.equal(env.obs[300:500], shard.obs) and np.array_equal(gpu_state(env)[300:500], gpu_state(shard))
    quiet = OL.OracleSim(H.make_config(n, mode, scenario=env.cfg.scenario, seed=31, change_ball_velocity=1, max_steps=100), "f32")
    quiet.reset()
    assert not np.array_equal(quiet.obs, sim.obs)


@pytest.mark.parametrize("scenario,n", [("reachball", 33), ("reachball", 257), ("shoot", 130), ("fullgame", 7)])
def test_kernels_stay_inside_their_buffers(scenario, n):
    """compute-sanitizer is not available on this pool, so: every buffer is bound as the middle of a larger tensor
    filled with a sentinel, ragged sizes, K > 1, terminal obs on - the guard bytes on both sides must survive."""
    import ctypes as C
    lib = _abi.load()
    env = Soccer2DVecEnv(n, scenario=scenario, device="cuda:0", seed=1, substeps=3, terminal_obs=True,
                         **({"use_continuous_action": False, "max_steps": 4, "change_ball_velocity": True} if scenario == "reachball"
                            else {"max_steps": 4} if scenario == "shoot" else {"half_time_cycles": 3}))
    guard = 4096
    big = {}

    def guarded(t):
        nbytes = t.numel() * t.element_size()
        buf = torch.full((nbytes + 2 * guard,), 0xA5, dtype=torch.uint8, device="cuda")
        big[len(big)] = (buf, nbytes)
        buf[guard:guard + nbytes] = 0
        return buf[guard:guard + nbytes]

    views = {k: guarded(getattr(env, k)) for k in ("state", "obs", "reward", "done_u8", "result", "terminal_obs", "stats_buf")}
    actions = guarded(env.actions)
    if env.actions.dtype == torch.uint8:
        actions.copy_(torch.randint(0, 16, (actions.numel(),), dtype=torch.uint8, device="cuda"))
    else:
        a = torch.rand(env.actions.shape, device="cuda") * 100 - 50
        a[..., 0] = torch.randint(0, 5, a.shape[:-1], device="cuda").float()
        actions.copy_(a.view(-1).view(torch.uint8))
    b = _abi.Buffers(state=views["state"].data_ptr(), actions=actions.data_ptr(), obs=views["obs"].data_ptr(),
                     reward=views["reward"].data_ptr(), done=views["done_u8"].data_ptr(), result=views["result"].data_ptr(),
                     terminal_obs=views["terminal_obs"].data_ptr(), stats=views["stats_buf"].data_ptr())
    torch.cuda.synchronize()
    _abi.check(lib.s2d_bind(env.handle, C.byref(b)), env.handle)
    stream = torch.cuda.current_stream().cuda_stream
    _abi.check(lib.s2d_reset(env.handle, None, stream), env.handle)
    for _ in range(6):
        _abi.check(lib.s2d_step(env.handle, 3, stream), env.handle)
    mask = torch.ones(n, dtype=torch.uint8, device="cuda")
    _abi.check(lib.s2d_reset(env.handle, mask.data_ptr(), stream), env.handle)
    torch.cuda.synchronize()
    for buf, nbytes in big.values():
        assert bool((buf[:guard] == 0xA5).all()) and bool((buf[guard + nbytes:] == 0xA5).all())
    assert float(views["obs"].view(torch.float32).abs().sum()) > 0  # the kernels did write
    env.close()


def test_checkpoint_resume_is_exact():
    """state_dict / load_state_dict: the SoA state (incl. RNG counters = episode / cycle) is the whole checkpoint."""
    n, k = 2000, 4
    env = make_env(n, "discrete", seed=5, substeps=k, noise=True, change_ball_velocity=True, max_steps=20)
    env.reset_torch()
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = [torch.randint(0, 16, (n, k), dtype=torch.uint8, device="cuda", generator=g) for _ in range(20)]
    for a in acts[:10]:
        env.step_torch(a)
    ckpt = env.state_dict()
    outs = []
    for a in acts[10:]:
        env.step_torch(a)
        outs.append((env.obs.clone(), env.reward.clone(), env.done.clone()))
    final_stats = env.stats()
    env.load_state_dict(ckpt)
    for a, (o, r, d) in zip(acts[10:], outs):
        env.step_torch(a)
        assert torch.equal(env.obs, o) and torch.equal(env.reward, r) and torch.equal(env.done, d)
    st = env.stats()
    assert {k_: st[k_] for k_ in ("episodes", "goals", "outs", "timeouts")} == {k_: final_stats[k_] for k_ in ("episodes", "goals", "outs", "timeouts")}
