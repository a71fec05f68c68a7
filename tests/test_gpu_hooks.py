"""GPU: hook-style scenario classes (the reference's extension API, soccer_2d_env.py:317-354) on top of the GPU cycle.
The test scenario is ReachBall written the reference's way - four hooks over proto-shaped State / PlayerAction /
TrainerAction objects - and must reproduce the double-precision oracle of the same episode."""
import numpy as np
import pytest

import helpers as H  # noqa: F401
from oracle import soccer2d_oracle as O
from soccer2d_b200 import pb2_lite as pb2
from soccer2d_b200.spaces import Box, Discrete
from soccer_2d_env import Soccer2DEnv

pytestmark = pytest.mark.gpu

PLACEMENTS = [(-20.0, 4.0, 135.0, 10.0, -3.0, 1.5, 30.0), (30.0, -20.0, 10.0, -5.0, 12.0, 0.0, 0.0), (0.0, 0.0, 300.0, 6.0, 0.0, 2.0, 180.0)]


class HookReachBall(Soccer2DEnv):
    """ReachBall through the four hooks, as a user of the reference would write it."""

    def __init__(self, **kw):
        super().__init__(None, **kw)
        self.action_space = Discrete(16)
        self.observation_space = Box(low=-1, high=1, shape=(10,), dtype=np.float32)
        self.cfg = O.ReachBallConfig(use_continuous_action=False, max_steps=40)
        self.mem = (0.0, 0.0)
        self.episode = 0

    def action_to_rpc_actions(self, action, player_state):
        self.step_number += 1
        _, power, rel = O.decode_action(self.cfg, int(action), 0.5)
        return pb2.PlayerAction(dash=pb2.Dash(power=power, relative_direction=rel))

    def state_to_observation(self, state):
        wm = state.world_model
        s = getattr(wm, "self")
        return np.array(O.build_obs(wm.ball.position.x, wm.ball.position.y, wm.ball.velocity.x, wm.ball.velocity.y,
                                    s.position.x, s.position.y, s.body_direction))

    def check_trainer_observation(self, state):
        wm = state.world_model
        p = wm.teammates[0]
        done, reward, res, d, a = O.check_trainer(self.cfg, self.mem[0], self.mem[1], self.step_number, wm.ball.position.x,
                                                  wm.ball.position.y, p.position.x, p.position.y, p.body_direction)
        self.mem = (d, a)
        return done, reward, {"result": O.RESULT_NAMES[res]}

    def abs_reset(self):
        obs, trainer_state = self.env_reset()
        self.check_trainer_observation(trainer_state)
        return obs

    def trainer_reset_actions(self):
        self.step_number = 0
        px, py, body, bx, by, speed, d = PLACEMENTS[self.episode % len(PLACEMENTS)]
        self.episode += 1
        vx, vy = O.polar(speed, d)
        return [pb2.TrainerAction(do_move_ball=pb2.DoMoveBall(position=pb2.RpcVector2D(x=bx, y=by), velocity=pb2.RpcVector2D(x=vx, y=vy))),
                pb2.TrainerAction(do_move_player=pb2.DoMovePlayer(our_side=True, uniform_number=1, position=pb2.RpcVector2D(x=px, y=py),
                                                                  body_direction=body)),
                pb2.TrainerAction(do_recover=pb2.DoRecover())]


class PlacedOracle(O.ReachBallOracle):
    def sample_reset(self):
        px, py, body, bx, by, speed, d = PLACEMENTS[self.episode % len(PLACEMENTS)]
        vx, vy = O.polar(speed, d)
        return px, py, body, bx, by, O.f32(vx), O.f32(vy)


def test_hook_defined_scenario_matches_the_oracle():
    env = HookReachBall(device="cuda:0")
    assert env._hook_mode
    ora = PlacedOracle(O.ReachBallConfig(use_continuous_action=False, max_steps=40, sp=O.ServerParam().as_f32()), auto_reset=False)
    rng = np.random.default_rng(0)
    results = []
    for episode in range(3):
        obs = env.reset()
        want = ora.reset()
        assert H.obs_close(obs, want) < H.TOL
        done = False
        while not done:
            a = int(rng.integers(16))
            obs, reward, done, info = env.step(a)
            wo, wr, wd, wres, _ = ora.step(a)
            assert done == wd and info["result"] == O.RESULT_NAMES[wres]
            assert H.obs_close(obs, wo) < H.TOL and reward == pytest.approx(wr, abs=1e-3)
        results.append(info["result"])
    assert all(r in ("Goal", "Out", "Timeout") for r in results)
    env.close()


def test_hook_mode_lowers_every_supported_action():
    class Probe(HookReachBall):
        def action_to_rpc_actions(self, action, player_state):
            return action  # the test passes PlayerAction objects straight through

    env = Probe(device="cuda:0")
    env.reset()
    start = env._vec.export_env(0)
    env.step(pb2.PlayerAction(turn=pb2.Turn(relative_direction=40.0)))
    assert env._vec.export_env(0).players[0].body_direction == pytest.approx(start.players[0].body_direction + 40.0 - 360.0 * (start.players[0].body_direction + 40.0 > 180.0), abs=1e-3)
    env.step(pb2.PlayerAction(body_go_to_point=pb2.Body_GoToPoint(target_point=pb2.RpcVector2D(x=10.0, y=-3.0), distance_threshold=0.5, max_dash_power=100)))
    env.step([pb2.PlayerAction(body_hold_ball=pb2.Body_HoldBall())])
    # put the ball next to the player and kick it
    snap = env._vec.export_env(0)
    p = snap.players[0]
    env._apply_trainer_actions(pb2.TrainerAction(do_move_ball=pb2.DoMoveBall(position=pb2.RpcVector2D(x=p.x + 0.5, y=p.y), velocity=pb2.RpcVector2D(x=0, y=0))))
    env.step(pb2.PlayerAction(kick=pb2.Kick(power=100, relative_direction=0.0)))
    snap = env._vec.export_env(0)
    assert snap.players[0].kicked == 1 and np.hypot(snap.ball_vx, snap.ball_vy) > 1.0
    # the proxy's other body actions: stop the ball, pass it on at 1.2 m per cycle, turn towards it, turn to an angle
    p = snap.players[0]
    env._apply_trainer_actions(pb2.TrainerAction(do_move_ball=pb2.DoMoveBall(position=pb2.RpcVector2D(x=p.x + 0.4, y=p.y + 0.2), velocity=pb2.RpcVector2D(x=0.3, y=-0.2))))
    env._apply_trainer_actions(pb2.TrainerAction(do_move_player=pb2.DoMovePlayer(our_side=True, uniform_number=1, position=pb2.RpcVector2D(x=p.x, y=p.y), body_direction=10.0)))
    env.step(pb2.PlayerAction(body_stop_ball=pb2.Body_StopBall()))
    snap = env._vec.export_env(0)
    assert np.hypot(snap.ball_vx, snap.ball_vy) < 1e-4
    env.step(pb2.PlayerAction(body_kick_one_step=pb2.Body_KickOneStep(target_point=pb2.RpcVector2D(x=snap.ball_x + 30.0, y=snap.ball_y), first_speed=1.2, force_mode=True)))
    snap = env._vec.export_env(0)
    assert (snap.ball_vx, snap.ball_vy) == pytest.approx((1.2 * 0.94, 0.0), abs=1e-4)
    env.step(pb2.PlayerAction(body_turn_to_ball=pb2.Body_TurnToBall(cycle=1)))
    snap = env._vec.export_env(0)
    p = snap.players[0]
    # (cycle = 1: faces where the ball is after this cycle, from where the player is after this cycle = now)
    assert p.body_direction == pytest.approx(np.degrees(np.arctan2(snap.ball_y - p.y, snap.ball_x - p.x)), abs=0.05)
    env.step(pb2.PlayerAction(body_turn_to_angle=pb2.Body_TurnToAngle(angle=-120.0)))
    assert env._vec.export_env(0).players[0].body_direction == pytest.approx(-120.0, abs=1e-3)
    env.step(pb2.PlayerAction(body_turn_to_point=pb2.Body_TurnToPoint(target_point=pb2.RpcVector2D(x=0.0, y=0.0), cycle=1)))
    p = env._vec.export_env(0).players[0]
    assert p.body_direction == pytest.approx(np.degrees(np.arctan2(0.0 - p.y, 0.0 - p.x)), abs=0.05)
    # intercept: the ball rolls away; the player gets it back under control within a few cycles
    snap = env._vec.export_env(0)
    p = snap.players[0]
    env._apply_trainer_actions(pb2.TrainerAction(do_move_ball=pb2.DoMoveBall(position=pb2.RpcVector2D(x=p.x + 3.0, y=p.y + 2.0), velocity=pb2.RpcVector2D(x=0.8, y=0.3))))
    for _ in range(25):
        env.step(pb2.PlayerAction(body_intercept=pb2.Body_Intercept(save_recovery=False)))
        snap = env._vec.export_env(0)
        p = snap.players[0]
        if np.hypot(snap.ball_x - p.x, snap.ball_y - p.y) <= 1.085:
            break
    else:
        raise AssertionError("the ball was never intercepted")
    env.close()
