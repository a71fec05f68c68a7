"""GPU: the FULLGAME scenario (BASELINE configs[3], up to 11 v 11, one warp per match) against the CPU oracle:
bit for bit against its fp32 build, cycle by cycle against its f64 build."""
import numpy as np
import pytest
import torch

import helpers as H
import oracle_lib as OL
from soccer2d_b200 import Soccer2DVecEnv, _abi

pytestmark = pytest.mark.gpu


def swarm_policy(obs, np_players, rng=None, random_frac=0.0):
    """every player runs to the ball (Body_GoToPoint) and kicks it towards the opponents' goal when within 1 m"""
    obs = np.asarray(obs, np.float64)
    n = obs.shape[0]
    pps = np_players // 2
    a = np.zeros((n, 1, np_players, 4), np.float32)
    bx, by = obs[:, 0] * 52.5, obs[:, 1] * 34.0
    P = obs[:, 4:4 + 5 * np_players].reshape(n, np_players, 5)
    px, py, body = P[:, :, 0] * 52.5, P[:, :, 1] * 34.0, P[:, :, 4] * 180.0
    d = np.hypot(bx[:, None] - px, by[:, None] - py)
    gx = np.where(np.arange(np_players) < pps, 52.5, -52.5)[None, :]
    gd = np.degrees(np.arctan2(0.0 - by[:, None], gx - bx[:, None])) - body
    gd = (gd + 180.0) % 360.0 - 180.0
    near = d < 1.0
    a[:, 0, :, 0] = np.where(near, 3, 4)
    a[:, 0, :, 1] = np.where(near, 100.0, bx[:, None])
    a[:, 0, :, 2] = np.where(near, gd, by[:, None])
    a[:, 0, :, 3] = np.where(near, 0.0, 100.0)
    if rng is not None and random_frac > 0:
        rnd = H.random_commands(rng, n * np_players).reshape(n, 1, np_players, 4)
        a = np.where(rng.uniform(size=(n, 1, np_players, 1)) < random_frac, rnd, a).astype(np.float32)
    return a


def gpu_state_fg(env):
    """[N, np*12 + 17] float64 in the layout of s2do_get_state_fg"""
    pl = {k: v.cpu().numpy() for k, v in env.fullgame_planes().items()}
    n, p = env.num_envs, env.num_players
    pps = p // 2
    out = np.zeros((n, p * 12 + 17))
    players = out[:, :p * 12].reshape(n, p, 12)
    players[:, :, 0:4] = pl["pa"]
    players[:, :, 4:8] = pl["pb"]
    players[:, :, 8] = pl["pc"]
    ei, ej = pl["ei"].astype(np.int64) & 0xFFFFFFFF, pl["ej"].astype(np.int64) & 0xFFFFFFFF
    for j in range(p):
        players[:, j, 9] = (ej[:, 2] >> j) & 1
        players[:, j, 10] = (ej[:, 3] >> j) & 1
        players[:, j, 11] = 1 if j < pps else 2
    k = p * 12
    out[:, k:k + 4] = pl["ball"]
    out[:, k + 4] = (ei[:, 3] >> 20) & 1
    out[:, k + 5] = pl["ei"][:, 0]
    out[:, k + 6] = ei[:, 1]
    out[:, k + 7] = ei[:, 2]
    out[:, k + 8] = ei[:, 3] & 0xff
    out[:, k + 9] = (ei[:, 3] >> 8) & 3
    out[:, k + 10] = (ei[:, 3] >> 12) & 0xff
    out[:, k + 11] = ej[:, 0]
    out[:, k + 12] = ej[:, 1]
    out[:, k + 13] = (ei[:, 3] >> 10) & 3
    out[:, k + 14] = pl["ef"][:, 0]
    out[:, k + 15] = (ei[:, 3] >> 21) & 1
    out[:, k + 16] = pl["ef"][:, 2].copy().view(np.uint32)  # offside marks: a bit mask kept in a float slot
    return out


def gpu_extra_fg(env):
    """[N, np + 2]: tackle counters per player, catch bans of the two keepers (layout of s2do_get_extra_fg)"""
    ek = env.fullgame_planes()["ek"].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    n, p = env.num_envs, env.num_players
    out = np.zeros((n, p + 2))
    for j in range(p):
        out[:, j] = (ek[:, j >> 3] >> (4 * (j & 7))) & 15
    out[:, p], out[:, p + 1] = ek[:, 3] & 15, (ek[:, 3] >> 4) & 15
    return out


def set_gpu_extra_fg(env, extra):
    n, p = env.num_envs, env.num_players
    ek = np.zeros((n, 4), np.int64)
    for j in range(p):
        ek[:, j >> 3] |= extra[:, j].astype(np.int64) << (4 * (j & 7))
    ek[:, 3] = extra[:, p].astype(np.int64) | (extra[:, p + 1].astype(np.int64) << 4)
    env.fullgame_planes()["ek"].copy_(torch.from_numpy(ek.astype(np.uint32).view(np.int32)))


def same_state(env, sim):
    assert np.array_equal(gpu_state_fg(env), sim.get_state_fg())
    assert np.array_equal(gpu_extra_fg(env), sim.get_extra_fg())


def same_step(env, sim):
    assert np.array_equal(env.done_u8.cpu().numpy(), sim.done) and np.array_equal(env.result.cpu().numpy(), sim.result)
    assert np.array_equal(env.obs.cpu().numpy(), sim.obs)
    assert np.array_equal(env.reward.cpu().numpy(), sim.reward)
    d = sim.done.astype(bool)
    if env.terminal_obs is not None and d.any():
        assert np.array_equal(env.terminal_obs.cpu().numpy()[d], sim.term_obs[d])


@pytest.mark.parametrize("pps,k,default_sp,collision_model", [(11, 1, True, "midpoint"), (11, 4, True, "midpoint"),
                                                             (3, 1, True, "midpoint"), (11, 1, False, "midpoint"),
                                                             (11, 1, True, "backtrace"), (5, 3, False, "backtrace")])
@pytest.mark.parametrize("lanes", ["1", "2"])
def test_fullgame_bit_exact_against_fp32_oracle(pps, k, default_sp, collision_model, lanes, monkeypatch):
    monkeypatch.setenv("S2D_FG_LANES", lanes)  # thread per match / two lanes per match (read when the handle is created)
    n = 131
    sp = None if default_sp else dict(player_decay=0.45, kickable_margin=0.8, slowness_on_top_for_right_team=1.1,
                                     ball_decay=0.95, kick_power_rate=0.03)
    env = Soccer2DVecEnv(n, scenario="fullgame", device="cuda:0", seed=5, substeps=k, terminal_obs=True,
                         players_per_side=pps, half_time_cycles=120, server_param=sp, collision_model=collision_model)
    p = 2 * pps
    assert env.obs.shape == (n, 120) and env.actions.shape == (n, k, p, 4)
    sim = OL.OracleSim(env.cfg, "f32")
    assert np.array_equal(env.reset(), sim.reset())
    assert np.array_equal(gpu_state_fg(env), sim.get_state_fg())
    rng = np.random.default_rng(0)
    modes = set()
    for t in range(600 // k):
        act = np.repeat(swarm_policy(sim.obs, p, rng, random_frac=0.15), k, axis=1)
        env.step_torch(torch.from_numpy(act))
        sim.step(act.reshape(n, -1), k)
        same_step(env, sim)
        modes |= set(sim.obs[:, 114].astype(int).tolist())
        if t % 25 == 24:
            same_state(env, sim)
    same_state(env, sim)
    st, so = env.stats(), sim.stats(_abi.Stats())
    for key in ("episodes", "goals", "outs", "timeouts", "episode_steps", "env_steps"):
        assert st[key] == getattr(so, key), key
    assert st["return_sum"] == pytest.approx(so.return_sum, rel=1e-9)
    assert st["episodes"] == 2 * n if k == 1 else st["episodes"] >= n
    assert _abi.RESULT_NAMES and {2, 3} <= modes  # PlayOn and KickOff at least
    total_goals = int(sim.get_state_fg()[:, p * 12 + 11:p * 12 + 13].sum())
    assert total_goals >= 0


def rules_policy(obs, rng):
    """11 v 11 with the whole vocabulary: run to the ball, SmartKick it at the other goal, tackle it off an opponent's
    foot, keepers catch what comes close"""
    obs = np.asarray(obs, np.float64)
    n = obs.shape[0]
    a = np.zeros((n, 1, 22, 4), np.float32)
    bx, by = obs[:, 0] * 52.5, obs[:, 1] * 34.0
    P = obs[:, 4:114].reshape(n, 22, 5)
    px, py, body = P[:, :, 0] * 52.5, P[:, :, 1] * 34.0, P[:, :, 4] * 180.0
    dx, dy = bx[:, None] - px, by[:, None] - py
    d = np.hypot(dx, dy)
    rel = (np.degrees(np.arctan2(dy, dx)) - body + 180.0) % 360.0 - 180.0
    gx = np.where(np.arange(22) < 11, 52.5, -52.5)[None, :] * np.ones((n, 1))
    near = d < 1.0
    a[:, 0, :, 0] = np.where(near, 13, 4)                      # SmartKick | GoToPoint
    a[:, 0, :, 1] = np.where(near, gx, bx[:, None])
    a[:, 0, :, 2] = np.where(near, rng.uniform(-5, 5, (n, 22)), by[:, None])
    a[:, 0, :, 3] = np.where(near, 2.8, 100.0)
    tackle = (d < 1.9) & ~near & (np.abs(rel) < 60) & (rng.uniform(size=(n, 22)) < 0.5)
    a[:, 0, :, 0] = np.where(tackle, 11, a[:, 0, :, 0])
    a[:, 0, :, 1] = np.where(tackle, np.rint(rng.uniform(-90, 90, (n, 22))), a[:, 0, :, 1])
    for kpr in (0, 11):                                           # the keepers stay home and catch
        home = np.stack([np.full(n, -48.0 if kpr == 0 else 48.0), np.clip(by, -6, 6)], axis=1)
        a[:, 0, kpr] = np.stack([np.full(n, 4.0), home[:, 0], home[:, 1], np.full(n, 100.0)], axis=1)
        reach = d[:, kpr] < 1.3
        a[reach, 0, kpr] = np.stack([np.full(reach.sum(), 12.0), np.rint(rel[reach, kpr]), np.zeros(reach.sum()), np.zeros(reach.sum())], axis=1)
    return a


@pytest.mark.parametrize("k", [1, 3])
def test_fullgame_tackles_catches_smart_kicks_and_the_goal_pause_bit_exact(k):
    """round 2 rules (include/soccer2d.h): tackle, goalkeeper catch, Body_SmartKick, the AfterGoal pause and the kick-off
    confinement, in a match where all of them happen - GPU against the fp32 oracle, bit for bit."""
    n = 96
    env = Soccer2DVecEnv(n, scenario="fullgame", device="cuda:0", seed=9, substeps=k, terminal_obs=True, half_time_cycles=300)
    sim = OL.OracleSim(env.cfg, "f32")
    assert np.array_equal(env.reset(), sim.reset())
    rng = np.random.default_rng(2)
    modes, floored, bans = set(), 0, 0
    for t in range(900 // k):
        act = np.repeat(rules_policy(sim.obs, rng), k, axis=1)
        env.step_torch(torch.from_numpy(act))
        sim.step(act.reshape(n, -1), k)
        same_step(env, sim)
        modes |= set(sim.obs[:, 114].astype(int).tolist())
        if t % 10 == 9:
            same_state(env, sim)
            ex = gpu_extra_fg(env)
            floored += int((ex[:, :22] > 0).sum())
            bans += int((ex[:, 22:] > 0).sum())
    same_state(env, sim)
    assert {2, 3, 5, 8} <= modes, sorted(modes)  # play on, kick-off, free kick (catch / offside), after goal
    assert floored > 0 and bans > 0, (floored, bans)
    st, so = env.stats(), sim.stats(_abi.Stats())
    assert st["episodes"] == so.episodes > 0
    env.close()


@pytest.mark.parametrize("pps, k", [(11, 1), (11, 3), (4, 2)])
def test_fullgame_player_types_per_match_bit_exact(pps, k):
    """s2d_set_player_types_per_match: every match has its own assignment of the 18 types (rcssserver hands its types out
    match by match).  Bit-exact against the oracle; a shard with env_id_offset reproduces its part of the global run; a
    match whose row equals the per-handle assignment plays exactly like a handle that has that assignment."""
    n = 150  # (ragged: three blocks of 64 columns, the last one mostly scratch)
    kw = dict(scenario="fullgame", device="cuda:0", seed=5, substeps=k, terminal_obs=True, half_time_cycles=80,
              players_per_side=pps)
    env = Soccer2DVecEnv(n, hetero_seed=7, hetero_per_match=True, **kw)
    np_ = 2 * pps
    assert env.type_of_player.shape == (n, np_) and (env.type_of_player[:, 0] == 0).all() and (env.type_of_player[:, pps] == 0).all()
    assert len({tuple(r) for r in env.type_of_player.tolist()}) == n  # all different
    types = (_abi.PlayerType * 18)()
    for j, t in enumerate(env.player_types):
        for name, v in t.items():
            setattr(types[j], name, v)
    sim = OL.OracleSim(env.cfg, "f32")
    sim.set_player_types(types, 18, env.type_of_player)
    shard = Soccer2DVecEnv(50, env_id_offset=70, hetero_seed=7, hetero_per_match=True, **kw)
    assert np.array_equal(shard.type_of_player, env.type_of_player[70:120])
    one = Soccer2DVecEnv(n, **kw)
    one.set_player_types(env.player_types, env.type_of_player[33])
    assert np.array_equal(env.reset(), sim.reset())
    shard.reset(), one.reset()
    assert np.array_equal(gpu_state_fg(env), sim.get_state_fg())
    rng = np.random.default_rng(2)
    for t in range(-(-170 // k) + 3):  # (past the end of the first match: 2 x 80 cycles)
        act = np.repeat(swarm_policy(sim.obs, np_, rng, random_frac=0.15), k, axis=1)
        env.step_torch(torch.from_numpy(act))
        shard.step_torch(torch.from_numpy(np.ascontiguousarray(act[70:120])))
        one.step_torch(torch.from_numpy(act))
        sim.step(act.reshape(n, -1), k)
        same_step(env, sim)
        assert np.array_equal(shard.obs.cpu().numpy(), env.obs.cpu().numpy()[70:120])
        assert np.array_equal(one.obs.cpu().numpy()[33], env.obs.cpu().numpy()[33])
    assert np.array_equal(gpu_state_fg(env), sim.get_state_fg())
    assert env.stats()["episodes"] == sim.stats(_abi.Stats()).episodes > 0
    for e in (env, shard, one):
        e.close()


@pytest.mark.parametrize("noise", [False, True])
def test_fullgame_heterogeneous_players_bit_exact(noise):
    """rcssserver player types (PlayerType): every player but the goalkeepers plays with one of 17 drawn types; the
    kernel reads the per-player values from the type table, the oracle from per-player copies of ServerParam."""
    n, k = 67, 2
    kw = dict(scenario="fullgame", device="cuda:0", seed=11, substeps=k, terminal_obs=True, half_time_cycles=100, noise=noise)
    env = Soccer2DVecEnv(n, hetero_seed=3, **kw)
    assert len(env.player_types) == 18 and env.type_of_player[0] == 0 and env.type_of_player[11] == 0
    assert len(set(env.type_of_player.tolist())) > 5
    types = (_abi.PlayerType * 18)()
    for j, t in enumerate(env.player_types):
        for name, v in t.items():
            setattr(types[j], name, v)
    sim = OL.OracleSim(env.cfg, "f32")
    sim.set_player_types(types, 18, env.type_of_player)
    plain = Soccer2DVecEnv(n, **kw)
    assert np.array_equal(env.reset(), sim.reset())
    plain.reset()
    st = gpu_state_fg(env)
    assert np.array_equal(st, sim.get_state_fg())
    effort0 = st[:, :22 * 12].reshape(n, 22, 12)[0, :, 6]  # effort starts at the type's effort_max
    assert np.allclose(effort0, [env.player_types[t]["effort_max"] for t in env.type_of_player])
    rng = np.random.default_rng(1)
    for t in range(150):
        act = np.repeat(swarm_policy(sim.obs, 22, rng, random_frac=0.15), k, axis=1)
        a = torch.from_numpy(act)
        env.step_torch(a)
        plain.step_torch(a)
        sim.step(act.reshape(n, -1), k)
        same_step(env, sim)
    assert np.array_equal(gpu_state_fg(env), sim.get_state_fg())
    assert not np.array_equal(env.obs.cpu().numpy(), plain.obs.cpu().numpy())  # the types do change the game
    # back to homogeneous players: identical to a handle that never had types
    env.lib.s2d_set_player_types(env.handle, None, 0, None)
    env.reset()
    plain.state.copy_(env.state)  # the same matches (the two have paused for different goals: their server clocks differ)
    plain.obs.copy_(env.obs)
    for t in range(5):
        act = np.repeat(swarm_policy(plain.obs.cpu().numpy(), 22, rng, random_frac=0.15), k, axis=1)
        a = torch.from_numpy(act)
        env.step_torch(a)
        plain.step_torch(a)
        assert np.array_equal(env.obs.cpu().numpy(), plain.obs.cpu().numpy())


def test_fullgame_referee_and_collision_cases():
    """Hand-placed balls and players: kick-in, corner kick, goal kick, goals at both ends, dead-ball rules (the
    other side's kick is ignored, the awarded side's kick resumes play, drop ball after 100 cycles), two players
    on the same spot, a pile-up around the ball."""
    line = 52.5 + 0.085
    #        ball x, y, vx, vy, play_mode, mode_side, last_touch
    cases = [
        (10.0, 34.0, 0.0, 0.5, 2, 0, 1),      # 0 over the side line, last touched by left -> KickIn right
        (-5.0, -34.0, 0.2, -0.5, 2, 0, 2),    # 1 other side line, last touched by right -> KickIn left
        (line - 0.2, 20.0, 1.0, 0.0, 2, 0, 2),  # 2 over the right goal line, defender (right) touched -> corner for left
        (line - 0.2, -20.0, 1.0, 0.0, 2, 0, 1),  # 3 attacker (left) touched -> goal kick for right
        (line - 0.2, 3.0, 1.0, 0.1, 2, 0, 1),   # 4 into the right goal -> left scores, kick-off for right
        (-line + 0.2, -3.0, -1.0, 0.0, 2, 0, 2),  # 5 into the left goal -> right scores
        (0.0, 0.0, 0.0, 0.0, 4, 1, 0),          # 6 kick-in for left: right player's kick is ignored, left's resumes
        (0.0, 0.0, 0.0, 0.0, 5, 2, 0),          # 7 free kick for right, nobody kicks -> drop ball after 100 cycles
        (20.0, 10.0, 0.0, 0.0, 2, 0, 0),        # 8 two players on the same spot + a third overlapping
        (-20.0, -10.0, 0.1, 0.05, 2, 0, 0),     # 9 pile-up: four players overlapping the (slowly) moving ball
    ]
    n = len(cases)
    env = Soccer2DVecEnv(n, scenario="fullgame", device="cuda:0", seed=9, half_time_cycles=1000, terminal_obs=True)
    sim = OL.OracleSim(env.cfg, "f32")
    env.reset_torch()
    sim.reset()
    pl = env.fullgame_planes()

    def put_ball(i, x, y, vx, vy, mode, side, touch):
        sim.L.s2do_set_ball_fg(sim.h, i, x, y, vx, vy, mode, side, touch)
        pl["ball"][i] = torch.tensor([x, y, vx, vy], device="cuda")
        pl["ei"][i, 3] = mode | (side << 8) | (touch << 10)

    def put_player(i, j, x, y, vx, vy, body):
        sim.L.s2do_set_player_fg(sim.h, i, j, x, y, vx, vy, body)
        pl["pa"][i, j] = torch.tensor([x, y, vx, vy], device="cuda")
        pl["pb"][i, j, 0] = body

    for i, c in enumerate(cases):
        put_ball(i, *c)
    put_player(6, 12, 0.5, 0.0, 0, 0, 180.0)   # right player next to the dead ball
    put_player(6, 3, -0.5, 0.0, 0, 0, 0.0)     # left player next to it
    put_player(8, 1, 20.0, 10.0, 0.3, 0, 0)
    put_player(8, 13, 20.0, 10.0, -0.3, 0, 180.0)
    put_player(8, 2, 20.3, 10.1, 0, 0.1, 90.0)
    for j, (dx, dy) in zip((4, 5, 15, 16), ((0.2, 0.0), (-0.2, 0.1), (0.0, -0.25), (0.1, 0.2))):
        put_player(9, j, -20.0 + dx, -10.0 + dy, 0.1, -0.1, 45.0)
    assert np.array_equal(gpu_state_fg(env), sim.get_state_fg())

    act = np.zeros((n, 1, 22, 4), np.float32)
    act[6, 0, 12, :3] = [3, 100, 0]  # the right player kicks during the left team's kick-in: ignored
    env.step_torch(torch.from_numpy(act))
    sim.step(act.reshape(n, -1))
    same_step(env, sim)
    g = gpu_state_fg(env)
    assert np.array_equal(g, sim.get_state_fg())
    k = 22 * 12
    mode, side, sl, sr = g[:, k + 8], g[:, k + 9], g[:, k + 11], g[:, k + 12]
    assert (mode[0], side[0]) == (_abi_pm("KICK_IN"), 2) and (mode[1], side[1]) == (_abi_pm("KICK_IN"), 1)
    assert (mode[2], side[2]) == (_abi_pm("CORNER_KICK"), 1) and (mode[3], side[3]) == (_abi_pm("GOAL_KICK"), 2)
    assert (mode[4], side[4], sl[4], sr[4]) == (_abi_pm("AFTER_GOAL"), 1, 1, 0)  # AfterGoal_ + the side that scored
    assert (mode[5], side[5], sl[5], sr[5]) == (_abi_pm("AFTER_GOAL"), 2, 0, 1)
    assert mode[6] == _abi_pm("KICK_IN") and not g[6, 12 * 12 + 10]  # still a dead ball, no kick registered
    assert g[8, 1 * 12 + 9] and g[8, 13 * 12 + 9] and g[8, 2 * 12 + 9]  # collided flags
    assert g[9, k + 4] == 1  # ball collided
    assert float(env.reward[4]) > 9.9 and float(env.reward[5]) < -9.9

    act[:] = 0
    act[6, 0, 3, :3] = [3, 60, 0]    # now the left player kicks: play on
    for t in range(110):
        env.step_torch(torch.from_numpy(act))
        sim.step(act.reshape(n, -1))
        same_step(env, sim)
        if t == 0:
            g = gpu_state_fg(env)
            assert g[6, k + 8] == _abi_pm("PLAY_ON") and g[6, 3 * 12 + 10] == 1
            assert g[7, k + 8] == 5
        act[6] = 0
    g = gpu_state_fg(env)
    assert np.array_equal(g, sim.get_state_fg())
    assert g[7, k + 8] == _abi_pm("PLAY_ON")  # dropped after 100 cycles
    # the goals: 50 stopped cycles (the server clock did not advance), then the kick-off for the side that conceded
    assert (g[4, k + 8], g[4, k + 9]) == (_abi_pm("KICK_OFF"), 2) and (g[5, k + 8], g[5, k + 9]) == (_abi_pm("KICK_OFF"), 1)
    assert g[4, k + 6] == g[0, k + 6] - 50 and g[4, k + 10] == 60  # cycle; the kick-off has been waiting for 60 cycles
    snap = env.export_env(4)
    assert (snap.left_score, snap.right_score, snap.num_players) == (1, 0, 22)
    assert snap.players[11].side == 2 and snap.players[11].uniform_number == 1


def test_fullgame_offside():
    """OffsideRef: a pass marks the team-mates beyond the ball, the half-way line and the second-last defender; a marked
    player within 2.5 m of the ball gives the defenders a free kick where it stands.  Onside receiver, a pass from a
    kick-in and a ball won back by the defenders do not."""
    n = 4
    env = Soccer2DVecEnv(n, scenario="fullgame", device="cuda:0", seed=13, half_time_cycles=1000)
    sim = OL.OracleSim(env.cfg, "f32")
    env.reset_torch()
    sim.reset()
    pl = env.fullgame_planes()

    def put_ball(i, x, y, vx, vy, mode, side, touch):
        sim.L.s2do_set_ball_fg(sim.h, i, x, y, vx, vy, mode, side, touch)
        pl["ball"][i] = torch.tensor([x, y, vx, vy], device="cuda")
        pl["ei"][i, 3] = mode | (side << 8) | (touch << 10)

    def put_player(i, j, x, y, body):
        sim.L.s2do_set_player_fg(sim.h, i, j, x, y, 0.0, 0.0, body)
        pl["pa"][i, j] = torch.tensor([x, y, 0.0, 0.0], device="cuda")
        pl["pb"][i, j, 0] = body

    for i in range(n):
        put_ball(i, 10.5, 0.0, 0.0, 0.0, 4 if i == 2 else 2, 1 if i == 2 else 0, 0)  # env 2: kick-in for the left team
        put_player(i, 3, 10.0, 0.0, 0.0)                       # the passer
        put_player(i, 9, 25.0 if i == 1 else 45.0, 5.0, 0.0)   # the receiver: behind / beyond the second-last defender
    put_player(3, 15, 30.0, 20.0, 180.0)                       # env 3: a defender who will win the ball
    k = 22 * 12
    act = np.zeros((n, 1, 22, 4), np.float32)
    act[:, 0, 3, :3] = [3, 100, 0]

    def step():
        env.step_torch(torch.from_numpy(act))
        sim.step(act.reshape(n, -1))
        same_step(env, sim)
        g = gpu_state_fg(env)
        assert np.array_equal(g, sim.get_state_fg())
        return g

    g = step()  # the pass
    assert g[:, k + 8].tolist() == [2, 2, 2, 2]                # play on everywhere (the kick-in was taken)
    assert g[:, k + 16].tolist() == [1 << 9, 0, 0, 1 << 9]     # receiver marked, onside, exempt restart, marked
    act[:] = 0
    # the ball reaches the receiver (envs 0-2) / the defender (env 3), who kicks it on
    for i in range(3):
        put_ball(i, 44.0 if i != 1 else 24.0, 5.0, 0.0, 0.0, 2, 0, 1)
    put_ball(3, 29.5, 20.0, 0.0, 0.0, 2, 0, 1)
    act[3, 0, 15, :3] = [3, 50, 0]
    g = step()
    assert (g[0, k + 8], g[0, k + 9], g[0, k + 16]) == (5, 2, 0)   # offside: free kick for the right team ...
    assert g[0, k:k + 4].tolist() == [45.0, 5.0, 0.0, 0.0]         # ... where the receiver stands
    assert (g[1, k + 8], g[2, k + 8], g[3, k + 8]) == (2, 2, 2)    # no call
    assert g[3, k + 16] == 0 and g[3, 15 * 12 + 10] == 1           # the defenders won the ball: the mark is void
    # the free kick is taken; the left team stays 9.15 m away meanwhile
    for _ in range(3):
        g = step()
    assert g[0, k + 8] == 5 and np.hypot(g[0, 9 * 12] - 45.0, g[0, 9 * 12 + 1] - 5.0) >= 9.15 - 1e-4
    env.close()


def _abi_pm(name):
    return {"BEFORE_KICK_OFF": 0, "TIME_OVER": 1, "PLAY_ON": 2, "KICK_OFF": 3, "KICK_IN": 4, "FREE_KICK": 5,
            "CORNER_KICK": 6, "GOAL_KICK": 7, "AFTER_GOAL": 8}[name]


def test_fullgame_against_f64_truth_cycle_by_cycle():
    """22 players chasing one ball is chaotic (every kick and collision amplifies a rounding difference), so - as for
    Shoot - the double oracle is re-synchronised to the GPU state before every cycle and compared one cycle ahead:
    referee state and flags bit-exact, floats within 1e-5 of each quantity's scale, over 300 cycles of swarm play
    (kicks, pile-ups, goals, restarts)."""
    n, p = 48, 22
    env = Soccer2DVecEnv(n, scenario="fullgame", device="cuda:0", seed=3, half_time_cycles=100)
    sim = OL.OracleSim(env.cfg, "f64")
    env.reset_torch()
    sim.reset()
    scale = np.tile([52.5, 34.0, 1.05, 1.05, 180.0, 8000.0, 1.0, 1.0, 130600.0], p)
    worst = worst_hit = 0.0
    kicks = collisions = 0
    for t in range(300):
        g = gpu_state_fg(env)
        sim.set_state_fg(g)
        act = swarm_policy(env.obs.cpu().numpy(), p)
        env.step_torch(torch.from_numpy(act))
        sim.step(act.reshape(n, -1))
        assert np.array_equal(env.done_u8.cpu().numpy(), sim.done) and np.array_equal(env.result.cpu().numpy(), sim.result)
        g2, o2 = gpu_state_fg(env), sim.get_state_fg()
        P2, Q2 = g2[:, :p * 12].reshape(n, p, 12), o2[:, :p * 12].reshape(n, p, 12)
        assert np.array_equal(P2[:, :, 9:], Q2[:, :, 9:])                      # collided, kicked, side
        assert np.array_equal(g2[:, p * 12 + 4:p * 12 + 14], o2[:, p * 12 + 4:p * 12 + 14])  # ball flag, counters, referee
        d = np.abs(P2[:, :, :9] - Q2[:, :, :9])
        d[:, :, 4] = np.minimum(d[:, :, 4], np.abs(360.0 - d[:, :, 4]))
        d = d / scale.reshape(p, 9)[None]
        db = np.abs(g2[:, p * 12:p * 12 + 4] - o2[:, p * 12:p * 12 + 4]) / [52.5, 34.0, 3.0, 3.0]
        # Collision resolution pushes overlapping objects apart along the line of centres; for nearly coincident
        # centres (a pile-up on the ball) that direction is ill-conditioned, so objects that collided this cycle
        # get a loose bound and everything else the 1e-5 one.
        hit = P2[:, :, 9] != 0
        ball_hit = g2[:, p * 12 + 4] != 0
        worst = max(worst, float(d[~hit].max()), float(db[~ball_hit].max()) if (~ball_hit).any() else 0.0)
        worst_hit = max(worst_hit, float(d[hit].max()) if hit.any() else 0.0, float(db[ball_hit].max()) if ball_hit.any() else 0.0)
        # (the reward is 0.01 x the ball's displacement: the same exemption for a ball that was pushed out of a pile-up)
        dr = np.abs(env.reward.cpu().numpy() - sim.reward)
        assert dr[~ball_hit].max(initial=0.0) < 1e-4 and dr.max() < 0.05
        kicks += int(P2[:, :, 10].sum())
        collisions += int(P2[:, :, 9].sum())
    assert worst < H.TOL, worst
    assert worst_hit < 2e-2, worst_hit
    assert kicks > 100 and collisions > 1000


def test_fullgame_gym_api():
    from sample_environments.environment_factory import EnvironmentFactory
    env = EnvironmentFactory().create("FullGame", None, None, "/tmp", seed=1, players_per_side=5, half_time_cycles=20)
    assert env.action_space.shape == (10, 4) and env.observation_space.shape == (120,)
    obs = env.reset()
    done, steps = False, 0
    while not done:
        act = swarm_policy(obs[None], 10)[0, 0]
        obs, reward, done, info = env.step(act)
        steps += 1
    assert steps == 40 and info["result"] in ("Goal", "Out", "Timeout")
    env.close()


def test_fullgame_with_noise_bit_exact():
    n, p = 96, 22
    env = Soccer2DVecEnv(n, scenario="fullgame", device="cuda:0", seed=6, noise=True, half_time_cycles=100, terminal_obs=True)
    sim = OL.OracleSim(env.cfg, "f32")
    assert np.array_equal(env.reset(), sim.reset())
    rng = np.random.default_rng(3)
    for t in range(300):
        act = swarm_policy(sim.obs, p, rng, random_frac=0.15)
        env.step_torch(torch.from_numpy(act))
        sim.step(act.reshape(n, -1))
        same_step(env, sim)
    assert np.array_equal(gpu_state_fg(env), sim.get_state_fg())
    assert env.stats()["episodes"] == n


def set_gpu_state_fg(env, vec):
    """inverse of gpu_state_fg: write [N, np*12 + 17] states into the planes of the env"""
    pl = env.fullgame_planes()
    n, p = env.num_envs, env.num_players
    P = vec[:, :p * 12].reshape(n, p, 12)
    k = p * 12
    pl["pa"].copy_(torch.from_numpy(P[:, :, 0:4].astype(np.float32)))
    pl["pb"].copy_(torch.from_numpy(P[:, :, 4:8].astype(np.float32)))
    pl["pc"].copy_(torch.from_numpy(P[:, :, 8].astype(np.float32)))
    pl["ball"].copy_(torch.from_numpy(vec[:, k:k + 4].astype(np.float32)))
    ef = np.zeros((n, 4), np.float32)
    ef[:, 0] = vec[:, k + 14]
    ef[:, 2] = vec[:, k + 16].astype(np.uint32).view(np.float32)
    pl["ef"].copy_(torch.from_numpy(ef))
    ei = np.zeros((n, 4), np.int64)
    ei[:, 0], ei[:, 1], ei[:, 2] = vec[:, k + 5], vec[:, k + 6], vec[:, k + 7]
    ei[:, 3] = (vec[:, k + 8].astype(np.int64) | (vec[:, k + 9].astype(np.int64) << 8) | (vec[:, k + 13].astype(np.int64) << 10)
                | (vec[:, k + 10].astype(np.int64) << 12) | (vec[:, k + 4].astype(np.int64) << 20) | (vec[:, k + 15].astype(np.int64) << 21))
    pl["ei"].copy_(torch.from_numpy(ei.astype(np.uint32).view(np.int32)))
    ej = np.zeros((n, 4), np.int64)
    ej[:, 0], ej[:, 1] = vec[:, k + 11], vec[:, k + 12]
    for j in range(p):
        ej[:, 2] |= P[:, j, 9].astype(np.int64) << j
        ej[:, 3] |= P[:, j, 10].astype(np.int64) << j
    pl["ej"].copy_(torch.from_numpy(ej.astype(np.uint32).view(np.int32)))


@pytest.mark.parametrize("lanes", ["1", "2"])
@pytest.mark.parametrize("collision_model", ["midpoint", "backtrace"])
def test_fullgame_random_states_bit_exact(collision_model, lanes, monkeypatch):
    """One cycle from hand-made states (players piled on the ball, the ball near or beyond every line, every play mode,
    stale offside marks) with random commands: reaches the referee branches a trajectory rarely visits.  Both mappings of
    the 11 v 11 kernel: one thread per match (what a full-size shard runs) and two lanes per match (small shards)."""
    monkeypatch.setenv("S2D_FG_LANES", lanes)  # (read when the handle is created)
    n, p = 512, 22
    env = Soccer2DVecEnv(n, scenario="fullgame", device="cuda:0", seed=2, half_time_cycles=10 ** 6, auto_reset=False,
                         collision_model=collision_model)
    sim = OL.OracleSim(env.cfg, "f32")
    env.reset_torch()
    sim.reset()
    rng = np.random.default_rng(5)
    k = p * 12
    modes_after = set()
    for rnd in range(10):
        st = sim.get_state_fg()
        for i in range(n):
            kind = rng.integers(0, 5)
            bx = rng.choice([rng.uniform(-50, 50), 52.4, -52.4, 52.58, -52.58]) if kind else rng.uniform(-50, 50)
            by = rng.choice([rng.uniform(-30, 30), 33.95, -33.95, 34.08, -34.08, 3.0, -6.9]) if kind else rng.uniform(-30, 30)
            st[i, k:k + 4] = np.float32([bx, by, rng.uniform(-2.5, 2.5), rng.uniform(-2.5, 2.5)])
            P = st[i, :k].reshape(p, 12)
            P[:, 0] = rng.uniform(-52, 52, p)
            P[:, 1] = rng.uniform(-33, 33, p)
            crowd = rng.integers(0, 7)
            P[:crowd, 0] = bx + rng.uniform(-0.5, 0.5, crowd)
            P[:crowd, 1] = by + rng.uniform(-0.5, 0.5, crowd)
            P[:, 2:4] = rng.uniform(-0.4, 0.4, (p, 2))
            P[:, 4] = rng.uniform(-180, 180, p)
            P[:, 5] = rng.uniform(0, 8000, p)
            P[:, 6] = rng.uniform(0.6, 1.0, p)
            P[:, 7] = rng.uniform(0.5, 1.0, p)
            P[:, 0:9] = P[:, 0:9].astype(np.float32)
            P[:, 9:11] = 0
            mode = int(rng.choice([2, 2, 2, 3, 4, 5, 6, 7, 8]))
            if rng.uniform() < 0.15:  # the ball in front of a goalkeeper, inside its penalty area
                side = int(rng.integers(0, 2))
                P[11 * side, 0:2] = np.float32([(-1) ** (side + 1) * rng.uniform(40, 50), rng.uniform(-15, 15)])
                P[11 * side, 4] = np.float32(rng.uniform(-180, 180))
                th = np.radians(P[11 * side, 4]) + rng.uniform(-0.3, 0.3)
                d = rng.uniform(0.2, 1.4)
                st[i, k:k + 2] = np.float32([P[11 * side, 0] + d * np.cos(th), P[11 * side, 1] + d * np.sin(th)])
            st[i, k + 4] = 0
            st[i, k + 5] = rnd
            st[i, k + 8:k + 11] = [mode, int(rng.integers(1, 3)) if mode != 2 else 0, int(rng.choice([0, 5, 98, 99]))]
            st[i, k + 13] = int(rng.integers(0, 3))
            st[i, k + 14] = 0.0
            st[i, k + 15] = 0
            st[i, k + 16] = int(rng.integers(0, 1 << 11)) << (11 * int(rng.integers(0, 2))) if mode == 2 and rng.uniform() < 0.5 else 0
        sim.set_state_fg(st)
        set_gpu_state_fg(env, st)
        extra = np.zeros((n, p + 2))
        extra[:, :p] = np.where(rng.uniform(size=(n, p)) < 0.1, rng.integers(1, 11, (n, p)), 0)  # some lie on the ground
        extra[:, p:] = np.where(rng.uniform(size=(n, 2)) < 0.2, rng.integers(1, 6, (n, 2)), 0)
        sim.set_extra_fg(extra)
        set_gpu_extra_fg(env, extra)
        same_state(env, sim)
        act = H.random_commands(rng, n * p).reshape(n, 1, p, 4)
        act[:, 0, :, 0] = np.where(rng.uniform(size=(n, p)) < 0.3, 3, act[:, 0, :, 0])
        act[:, 0, :, 0] = np.where(rng.uniform(size=(n, p)) < 0.1, 11, act[:, 0, :, 0])  # tackles into the pile-up
        act[:, 0, 0, 0] = np.where(rng.uniform(size=n) < 0.5, 12, act[:, 0, 0, 0])     # the keepers try to catch
        act[:, 0, 11, 0] = np.where(rng.uniform(size=n) < 0.5, 12, act[:, 0, 11, 0])
        env.step_torch(torch.from_numpy(act))
        sim.step(act.reshape(n, -1))
        same_step(env, sim)
        same_state(env, sim)
        g = gpu_state_fg(env)
        modes_after |= set(g[:, k + 8].astype(int).tolist())
    assert {2, 4, 5, 6, 7, 8} <= modes_after  # (a goal is followed by the 50-cycle pause now: kick-off is not reached in one cycle)
    env.close()
