#!/usr/bin/env python
"""Generate the known-answer fixtures for the ReachBall env contract by EXECUTING THE REFERENCE'S OWN
CODE (sample_environments/reach_ball_env.py) in this container.

The reference cannot travel to the GPU box, so the outputs are committed as
tests/golden/reach_ball_contract.json and this script is committed next to them.

How the reference is run (nothing is copied from it):
  * /root/reference is put on sys.path, so `service_pb2`, `soccer_2d_env`, `server`,
    `sample_environments.reach_ball_env` are the reference's own modules.
  * `gym` and `pyrusgeom` are not installed here: tests/golden/_shims provides stand-ins (gym.Env /
    spaces.Box / spaces.Discrete; Vector2D / AngleDeg restated).  The geometry helper is therefore a
    restatement - everything else (decode, obs packing, reward/done/info, reset sampling) is the
    reference's code, executed verbatim on real `pb2.State` messages.
  * Soccer2DEnv.__init__ (soccer_2d_env.py:30-95) spawns the gRPC server / rcssserver / proxy, which
    are not available offline; it is replaced by a no-op so that ReachBallEnv.__init__
    (reach_ball_env.py:18-51) still runs its own kwargs parsing and space construction.

Usage:  python tests/golden/make_golden.py            (writes the json next to this file)
"""
import json
import logging
import math
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("S2D_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REF)

import numpy as np  # noqa: E402

import service_pb2 as pb2  # noqa: E402  (reference's generated module)
import soccer_2d_env  # noqa: E402

soccer_2d_env.Soccer2DEnv.__init__ = lambda self, *a, **k: None  # no processes offline
from sample_environments.reach_ball_env import ReachBallEnv  # noqa: E402
from sample_environments.environment_factory import EnvironmentFactory  # noqa: E402

_log = logging.getLogger("golden")
_log.addHandler(logging.NullHandler())
_log.setLevel(logging.CRITICAL)
_log.propagate = False


def make_env(**kw):
    return EnvironmentFactory().create("ReachBall", render_mode=None, logger=_log, log_dir="/tmp", **kw)


def player_state(bx, by, bvx, bvy, px, py, body):
    s = pb2.State()
    wm = s.world_model
    wm.ball.position.x, wm.ball.position.y = bx, by
    wm.ball.velocity.x, wm.ball.velocity.y = bvx, bvy
    wm.self.position.x, wm.self.position.y = px, py
    wm.self.body_direction = body
    return s


def trainer_state(bx, by, px, py, body):
    s = pb2.State()
    wm = s.world_model
    wm.ball.position.x, wm.ball.position.y = bx, by
    tm = wm.teammates.add()
    tm.position.x, tm.position.y = px, py
    tm.body_direction = body
    return s


def action_fields(pa):
    which = pa.WhichOneof("action")
    if which == "dash":
        return {"type": "dash", "power": pa.dash.power, "dir": pa.dash.relative_direction}
    if which == "turn":
        return {"type": "turn", "power": 0.0, "dir": pa.turn.relative_direction}
    raise AssertionError(which)


def gen_decode():
    out = {"discrete": [], "continuous": [], "turning": []}
    for n in (16, 8, 36, 7):
        env = make_env(use_continuous_action=False, action_space_size=n)
        for a in range(n):
            for wrap in (int, np.int64, lambda v: np.array(v), lambda v: np.array([v])):
                f = action_fields(env.action_to_rpc_actions(wrap(a), None))
            out["discrete"].append({"n": n, "a": a, **f})
        assert env.step_number == 4 * n  # :55 increments per call
    env = make_env(use_continuous_action=True, use_turning=False)
    rng = np.random.RandomState(7)
    vals = [-1.0, -0.5, 0.0, 0.25, 1.0] + list(rng.uniform(-1, 1, 60).astype(np.float32))
    for v in vals:
        a = np.array([v], dtype=np.float32)
        f = action_fields(env.action_to_rpc_actions(a, None))
        out["continuous"].append({"a": float(a[0]), **f})
    env = make_env(use_continuous_action=True, use_turning=True)
    rng = np.random.RandomState(11)
    real_rand = np.random.rand
    try:
        for i in range(80):
            a = rng.uniform(-1.3, 1.3, 4).astype(np.float32)  # exercises the clip at :64
            u = float(rng.uniform())
            np.random.rand = lambda u=u: u  # the U(0,1) draw at :71, made reproducible
            f = action_fields(env.action_to_rpc_actions(a, None))
            out["turning"].append({"a": [float(x) for x in a], "u": u, **f})
    finally:
        np.random.rand = real_rand
    return out


def rand_state(rng):
    bx, by = rng.uniform(-55, 55), rng.uniform(-36, 36)
    px, py = rng.uniform(-55, 55), rng.uniform(-36, 36)
    sp, d = rng.uniform(0, 3), rng.uniform(-180, 180)
    if rng.uniform() < 0.2:
        sp = 0.0
    if rng.uniform() < 0.15:  # near the player: exercises the Goal branch
        px, py = bx + rng.uniform(-6, 6), by + rng.uniform(-6, 6)
    body = rng.uniform(-180, 180)
    f32 = lambda v: float(np.float32(v))  # noqa: E731  (inputs recorded as the f32 the proto holds)
    return [f32(bx), f32(by), f32(sp * math.cos(math.radians(d))), f32(sp * math.sin(math.radians(d))),
            f32(px), f32(py), f32(body)]


def gen_obs():
    env = make_env()
    rng = np.random.RandomState(3)
    states = [[10, -3, 0.5, 0, -20, 4, 135], [-50, 30, -1.2, 2.1, 50, -30, -179], [0, 0, 0, 0, 3, 4, 0],
              [0, 0, 0, 0, 53, 0, 90], [20, 20, 0, 0, 20.5, 20.5, -45], [5, 5, 0, 0, 5, 5, 180],
              [0, 0, -1, 0, 10, 0, -180], [0, 0, 0, -2, -10, 0, 180]]
    states = [[float(np.float32(v)) for v in s] for s in states] + [rand_state(rng) for _ in range(250)]
    out = []
    for s in states:
        obs = env.state_to_observation(player_state(*s))
        assert obs.dtype == np.float64 and obs.shape == (10,)
        out.append({"state": s, "obs": [float(v) for v in obs]})
    return out


def gen_reward():
    rng = np.random.RandomState(5)
    chains = []
    # chain 0 = SURVEY Appendix B; the rest are random walks incl. all three endings
    fixed = [[10, -3, 0.5, 0, -20, 4, 135], [-50, 30, -1.2, 2.1, 50, -30, -179], [0, 0, 0, 0, 3, 4, 0],
             [0, 0, 0, 0, 53, 0, 90], [20, 20, 0, 0, 20.5, 20.5, -45]]
    specs = [({"min_distance_to_ball": 5.0, "max_steps": 200}, fixed, [1, 2, 3, 4, 5])]
    specs.append(({"min_distance_to_ball": 5.0, "max_steps": 200},
                  [fixed[0], [52.9, 0, 0, 0, 53, 0, 0]], [201, 201]))
    for c in range(12):
        kw = {"min_distance_to_ball": float(rng.choice([5.0, 1.0, 0.5, 8.0])), "max_steps": int(rng.choice([200, 20, 5]))}
        n = 40
        sts = [rand_state(rng) for _ in range(n)]
        start = int(rng.choice([1, kw["max_steps"] - 10 if kw["max_steps"] > 10 else 1]))
        specs.append((kw, sts, list(range(start, start + n))))
    for kw, sts, step_numbers in specs:
        env = make_env(**kw)
        steps = []
        for s, k in zip(sts, step_numbers):
            s = [float(np.float32(v)) for v in s]
            env.step_number = k
            done, reward, info = env.check_trainer_observation(trainer_state(s[0], s[1], s[4], s[5], s[6]))
            steps.append({"state": s, "step_number": k, "done": bool(done), "reward": float(reward),
                          "result": info["result"], "mem_dist": float(env.distance_to_ball),
                          "mem_ang": float(env.body_ball_angle_diff)})
        chains.append({"kwargs": kw, "steps": steps})
    return chains


def gen_reset():
    """Distribution facts of trainer_reset_actions / get_ball_velocity (reach_ball_env.py:170-218).
    The reference draws from Python's Mersenne Twister; the new simulator uses Philox, so what is
    pinned is the DISTRIBUTION (supports, moments, acceptance region), not the stream."""
    out = {}
    for name, kw in (("default", {}), ("dqn_script", {"change_ball_position": True, "change_ball_velocity": True}),
                     ("fixed_ball", {"change_ball_position": False, "ball_position_x": 10, "ball_position_y": -5,
                                     "ball_speed": 1.5, "ball_direction": 30})):
        env = make_env(**kw)
        random.seed(1234)
        rows = []
        for _ in range(20000):
            env.step_number = 17
            acts = env.trainer_reset_actions()
            assert env.step_number == 0 and len(acts) == 3
            mb, mp, rc = acts
            assert mb.WhichOneof("action") == "do_move_ball" and mp.WhichOneof("action") == "do_move_player"
            assert rc.WhichOneof("action") == "do_recover"
            assert mp.do_move_player.our_side is True and mp.do_move_player.uniform_number == 1
            rows.append([mp.do_move_player.position.x, mp.do_move_player.position.y, mp.do_move_player.body_direction,
                         mb.do_move_ball.position.x, mb.do_move_ball.position.y,
                         mb.do_move_ball.velocity.x, mb.do_move_ball.velocity.y])
        r = np.array(rows, dtype=np.float64)
        speed = np.hypot(r[:, 5], r[:, 6])
        travel = speed * (1.0 - 0.96 ** env.max_steps) / (1.0 - 0.96)
        d = np.arctan2(r[:, 6], r[:, 5])
        tx, ty = r[:, 3] + travel * np.cos(d), r[:, 4] + travel * np.sin(d)
        out[name] = {
            "kwargs": kw, "n": len(rows),
            "min": [float(v) for v in r.min(0)], "max": [float(v) for v in r.max(0)],
            "mean": [float(v) for v in r.mean(0)], "std": [float(v) for v in r.std(0)],
            "all_int_pos_body": bool(np.all(r[:, :5] == np.round(r[:, :5]))),
            "speed_mean": float(speed.mean()), "speed_max": float(speed.max()),
            "speed_hist_0_3_12bins": [int(v) for v in np.histogram(speed, bins=12, range=(0, 3))[0]],
            "target_max_abs_x": float(np.abs(tx).max()), "target_max_abs_y": float(np.abs(ty).max()),
            "first_rows": rows[:5],
        }
    return out


def gen_spaces():
    out = []
    for kw in ({}, {"use_continuous_action": False}, {"use_continuous_action": False, "action_space_size": 8},
               {"use_continuous_action": True, "use_turning": True}):
        env = make_env(**kw)
        a, o = env.action_space, env.observation_space
        out.append({"kwargs": kw, "action_space": {"type": type(a).__name__, "n": getattr(a, "n", None),
                                                   "shape": list(a.shape), "dtype": str(a.dtype)},
                    "observation_space": {"shape": list(o.shape), "dtype": str(o.dtype),
                                          "low": float(o.low.min()), "high": float(o.high.max())},
                    "defaults": {k: getattr(env, k) for k in
                                 ("change_ball_position", "change_ball_velocity", "ball_position_x", "ball_position_y",
                                  "ball_speed", "ball_direction", "min_distance_to_ball", "max_steps",
                                  "use_continuous_action", "action_space_size", "use_turning")}})
    try:
        EnvironmentFactory().create("ReachCenter", None, _log, "/tmp")
        unknown = "no error"
    except ValueError as e:
        unknown = f"ValueError: {e}"
    return {"envs": out, "unknown_env": unknown, "metadata": soccer_2d_env.Soccer2DEnv.metadata}


def main():
    data = {"_generated_by": "tests/golden/make_golden.py (executes /root/reference sample_environments/reach_ball_env.py)",
            "_geometry": "pyrusgeom restated in tests/golden/_shims (not installed offline)",
            "decode": gen_decode(), "obs": gen_obs(), "reward": gen_reward(), "reset": gen_reset(),
            "spaces": gen_spaces()}
    path = os.path.join(HERE, "reach_ball_contract.json")
    with open(path, "w") as f:
        json.dump(data, f, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
