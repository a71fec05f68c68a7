"""CPU-only: what the FULLGAME rules added in round 2 DO, stated as behaviour and checked on the C oracle (f64) - tackle
(Player::tackle), goalkeeper catch (Player::goalieCatch), Body_SmartKick staged over 2-3 cycles, the AfterGoal pause and
the kick-off confinement are covered as invariants in tests/test_oracle_c.py.  Spec: include/soccer2d.h.  The GPU is held
to the oracle bit for bit in tests/test_gpu_fullgame.py."""
import numpy as np

import helpers as H  # noqa: F401
import oracle_lib as OL

P, PPS = 22, 11
K = P * 12


def make(n, seed=1):
    cfg = OL.default_config(n, 2, action_mode=OL.ACT_COMMAND, players_per_side=PPS, half_time_cycles=10 ** 6, auto_reset=0, seed=seed)
    sim = OL.OracleSim(cfg, "f64")
    sim.reset()
    return sim


def scene(sim, ball, players, mode=2):
    """every env: the kick-off formation, except `players` {index: (x, y, vx, vy, body)}, ball (x, y, vx, vy), play mode"""
    st = sim.get_state_fg()
    for j, v in players.items():
        st[:, j * 12:j * 12 + 5] = v
    st[:, K:K + 4] = ball
    st[:, K + 8], st[:, K + 9], st[:, K + 10] = mode, 0, 0
    sim.set_state_fg(st)
    sim.set_extra_fg(np.zeros((sim.n, P + 2)))
    return st


def commands(n, per_player):
    a = np.zeros((n, 1, P, 4), np.float32)
    for j, c in per_player.items():
        a[:, 0, j] = c
    return a.reshape(n, -1)


def test_tackle_succeeds_with_the_stated_probability_and_floors_the_tackler():
    n = 4000
    sim = make(n)
    # player 5 at the origin facing +x, the ball 1.6 m ahead and 0.8 m to the side: fail = 0.8^6 + 0.64^6
    scene(sim, (1.6, 0.8, 0, 0), {5: (0, 0, 0, 0, 0)})
    sim.step(commands(n, {5: (11, 0.0, 0, 0)}))
    s = sim.get_state_fg()
    moved = np.hypot(s[:, K + 2], s[:, K + 3]) > 0
    want = 1.0 - (0.8 ** 6 + 0.64 ** 6)
    assert abs(moved.mean() - want) < 4 * np.sqrt(want * (1 - want) / n), (moved.mean(), want)
    # a successful tackle pushes the ball along the body direction (dir 0): 100 * 0.027 * (1 - 0.5 * atan2(0.8, 1.6) / 180)
    eff = 2.7 * (1.0 - 0.5 * np.degrees(np.arctan2(0.8, 1.6)) / 180.0)
    assert np.allclose(s[moved, K + 2], eff * 0.94, atol=1e-6) and np.allclose(s[moved, K + 3], 0.0, atol=1e-9)
    assert (s[moved, 5 * 12 + 10] == 1).all() and (s[~moved, 5 * 12 + 10] == 0).all()  # counts as a kick
    # whether it worked or not, the tackler lies on the ground for 10 cycles: dashes do nothing, then they do again
    assert (sim.get_extra_fg()[:, 5] == 10).all()
    x0 = s[:, 5 * 12].copy()
    for c in range(10):
        sim.step(commands(n, {5: (1, 100.0, 0, 0)}))
        assert np.array_equal(sim.get_state_fg()[:, 5 * 12], x0), c
    sim.step(commands(n, {5: (1, 100.0, 0, 0)}))
    assert (sim.get_state_fg()[:, 5 * 12] > x0 + 0.3).all() and (sim.get_extra_fg()[:, 5] == 0).all()


def test_tackle_fails_from_behind_and_in_dead_ball_modes():
    sim = make(64)
    scene(sim, (-0.5, 0.1, 0, 0), {5: (0, 0, 0, 0, 0)})       # the ball behind the player: tackle_back_dist = 0
    sim.step(commands(64, {5: (11, 0.0, 0, 0)}))
    s = sim.get_state_fg()
    assert (s[:, K + 2] == 0).all() and (sim.get_extra_fg()[:, 5] == 10).all()
    scene(sim, (0.8, 0.0, 0, 0), {5: (0, 0, 0, 0, 0)}, mode=4)  # kick-in: the ball is dead, nobody tackles it away
    sim.step(commands(64, {5: (11, 0.0, 0, 0)}))
    assert (sim.get_state_fg()[:, K + 2] == 0).all()


def test_goalkeeper_catch_gives_a_free_kick_and_is_banned_for_five_cycles():
    n = 8
    sim = make(n)
    # left keeper (player 0) at (-48, 0) facing +x; the ball rolls 0.8 m in front of it; an attacker (player 20) nearby
    scene(sim, (-47.2, 0.2, -0.3, 0.0), {0: (-48, 0, 0, 0, 0), 20: (-46, 1, 0, 0, 180)})
    sim.step(commands(n, {0: (12, 14.0, 0, 0)}))  # catch towards the ball (14 degrees left of the body direction)
    s, ex = sim.get_state_fg(), sim.get_extra_fg()
    assert (s[:, K + 8] == 5).all() and (s[:, K + 9] == 1).all() and (s[:, K + 13] == 1).all()  # FreeKick left, last touch left
    assert np.allclose(s[:, K:K + 2], s[:, 0:2]) and (s[:, K + 2:K + 4] == 0).all()                # the ball in its hands
    assert (ex[:, P] == 5).all()
    d = np.hypot(s[:, 20 * 12] - s[:, K], s[:, 20 * 12 + 1] - s[:, K + 1])
    assert (d >= 9.15 - 1e-9).all()                                                                 # the attacker is sent away
    # a field player cannot catch, nor can a keeper outside its penalty area, nor one that is still banned
    scene(sim, (-47.2, 0.2, 0, 0), {0: (-48, 0, 0, 0, 0), 3: (-47.9, 0.3, 0, 0, 0)})
    sim.step(commands(n, {3: (12, 0.0, 0, 0)}))
    assert (sim.get_state_fg()[:, K + 8] == 2).all()
    scene(sim, (-30.0, 0.2, 0, 0), {0: (-30.8, 0, 0, 0, 0)})
    sim.step(commands(n, {0: (12, 14.0, 0, 0)}))
    assert (sim.get_state_fg()[:, K + 8] == 2).all()
    scene(sim, (-47.2, 0.2, 0, 0), {0: (-48, 0, 0, 0, 0)})
    ban = np.zeros((n, P + 2))
    ban[:, P] = 3
    sim.set_extra_fg(ban)
    sim.step(commands(n, {0: (12, 14.0, 0, 0)}))
    assert (sim.get_state_fg()[:, K + 8] == 2).all() and (sim.get_extra_fg()[:, P] == 2).all()


def test_smart_kick_stages_when_one_kick_is_not_enough():
    """ball at rest beside the player, wanted: 2.9 m/s towards a far target.  One kick yields at most
    100 * 0.027 * rate < 2.7: KickOneStep (force mode) releases a slow ball at once; SmartKick stages the ball in front
    of the player first and releases a faster one a cycle or two later."""
    n = 4
    target = (40.0, 10.0)
    start = {5: (0, 0, 0, 0, np.degrees(np.arctan2(10.0, 40.0)))}  # facing the target
    ball = (0.25, -0.75, 0, 0)                                        # beside / behind: a poor kick rate

    def release_speed(code, cycles):
        sim = make(n)
        scene(sim, ball, start)
        best = 0.0
        for _ in range(cycles):
            sim.step(commands(n, {5: (code, target[0], target[1], 2.9)}))
            s = sim.get_state_fg()
            v = np.array([s[0, K + 2], s[0, K + 3]]) / 0.94  # the velocity the ball left with
            to_target = np.array(target) - s[0, K:K + 2]
            if np.hypot(s[0, K] - s[0, 5 * 12], s[0, K + 1] - s[0, 5 * 12 + 1]) > 1.085:  # out of reach: released
                cos = v @ to_target / (np.linalg.norm(v) * np.linalg.norm(to_target) + 1e-12)
                return float(np.linalg.norm(v)), float(cos)
            best = max(best, float(np.linalg.norm(v)))
        return best, 0.0

    one_speed, one_cos = release_speed(8, 1)
    smart_speed, smart_cos = release_speed(13, 4)
    assert one_speed < 2.2 and smart_speed > one_speed + 0.3 and smart_speed > 2.4, (one_speed, smart_speed)
    assert smart_cos > 0.995, smart_cos
    # when one kick is enough, SmartKick IS KickOneStep
    a, b = make(n), make(n)
    for sim in (a, b):
        scene(sim, (0.5, 0.1, 0, 0), start)
    a.step(commands(n, {5: (8, target[0], target[1], 1.5)}))
    b.step(commands(n, {5: (13, target[0], target[1], 1.5)}))
    assert np.array_equal(a.get_state_fg(), b.get_state_fg())
