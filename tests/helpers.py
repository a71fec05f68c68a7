"""Shared helpers for the parity tests."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gym-soccer-2d-env_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from soccer2d_b200 import _abi  # noqa: E402

MODES = {"discrete": _abi.ACT_DISCRETE, "continuous": _abi.ACT_CONTINUOUS, "turning": _abi.ACT_TURNING}
# obs columns that are angles / 180 or / 360: -1 and +1 (or -0.5 / +0.5) are the same direction
OBS_ANGLE_PERIOD = {0: 2.0, 1: 2.0, 7: 1.0}
# relative tolerance of the north star (1e-5) against the natural scale of each quantity
TOL = 1.0e-5


MODES["command"] = _abi.ACT_COMMAND


def make_config(num_envs, mode="discrete", scenario=_abi.SCENARIO_REACHBALL, **kw):
    """S2DConfig with the reference defaults (s2d_default_config) and overrides; kwargs named like the
    struct fields, plus `sp={...}` for ServerParam overrides."""
    lib = _abi.load()
    cfg = _abi.Config()
    assert lib.s2d_default_config(C.byref(cfg), scenario) == 0
    cfg.num_envs = num_envs
    cfg.action_mode = MODES[mode] if isinstance(mode, str) else mode
    sp = kw.pop("sp", {})
    for k, v in kw.items():
        assert hasattr(cfg, k), k
        setattr(cfg, k, v)
    for k, v in sp.items():
        assert hasattr(cfg.sp, k), k
        setattr(cfg.sp, k, v)
    return cfg


def random_actions(rng, mode, n, k=1, n_actions=16):
    mode = MODES[mode] if isinstance(mode, str) else mode
    if mode == _abi.ACT_DISCRETE:
        return rng.integers(0, n_actions, size=(n, k)).astype(np.uint8)
    if mode == _abi.ACT_CONTINUOUS:
        return rng.uniform(-1, 1, size=(n, k)).astype(np.float32)
    if mode == _abi.ACT_COMMAND:
        return random_commands(rng, n, k)
    return rng.uniform(-1.25, 1.25, size=(n, k, 4)).astype(np.float32)


def random_commands(rng, n, k=1, body_actions=True):
    """[n, k, 4] float32 {cmd, a, b, c}: a mix of none / dash / turn / kick / go-to-point and (body_actions) the
    proxy's turn-to-point / -ball / -angle, kick-one-step, stop-ball and intercept, with out-of-range arguments now and then
    (the clamps are part of the contract)."""
    a = np.zeros((n, k, 4), np.float32)
    menu = [0, 1, 1, 1, 2, 3, 3, 4, 4] + ([5, 6, 7, 8, 8, 9, 10, 10, 11, 12, 13, 13, 14] if body_actions else [])
    cmd = rng.choice(menu, size=(n, k))
    a[..., 0] = cmd
    dash, turn, kick, goto = cmd == 1, cmd == 2, cmd == 3, cmd == 4
    a[..., 1] = np.where(dash | kick, rng.uniform(-30, 130, (n, k)), a[..., 1])
    a[..., 2] = np.where(dash | kick, rng.uniform(-200, 200, (n, k)), a[..., 2])
    a[..., 1] = np.where(turn, rng.uniform(-200, 200, (n, k)), a[..., 1])
    a[..., 1] = np.where(goto, rng.uniform(-55, 55, (n, k)), a[..., 1])
    a[..., 2] = np.where(goto, rng.uniform(-36, 36, (n, k)), a[..., 2])
    a[..., 3] = np.where(goto, rng.uniform(20, 120, (n, k)), a[..., 3])
    to_point, to_ball, to_angle, one_step = cmd == 5, cmd == 6, cmd == 7, (cmd == 8) | (cmd == 13)  # 13: smart kick
    tackle_or_catch = (cmd == 11) | (cmd == 12)  # (14 does not exist: no command)
    a[..., 1] = np.where(tackle_or_catch, rng.uniform(-200, 200, (n, k)), a[..., 1])
    a[..., 1] = np.where(to_point | one_step, rng.uniform(-55, 55, (n, k)), a[..., 1])
    a[..., 2] = np.where(to_point | one_step, rng.uniform(-36, 36, (n, k)), a[..., 2])
    a[..., 3] = np.where(to_point, rng.integers(-2, 70, (n, k)), a[..., 3])
    a[..., 3] = np.where(one_step, rng.uniform(-0.5, 3.5, (n, k)), a[..., 3])
    a[..., 1] = np.where(to_ball, rng.integers(0, 12, (n, k)), a[..., 1])
    a[..., 1] = np.where(to_angle, rng.uniform(-400, 400, (n, k)), a[..., 1])
    return a


def dash_and_shoot(obs):
    """Like chase_and_shoot, but moves with explicit Dash(100, direction to the ball) commands: no turn-or-dash
    decision inside the simulator, so an fp32 and an f64 run cannot split on that threshold."""
    a = chase_and_shoot(obs)
    obs = np.asarray(obs, np.float64)
    far = a[:, 0, 0] == 4
    rel = obs[:, 0] * 180.0  # body-to-ball angle
    a[far, 0, 0] = 1
    a[far, 0, 1] = 100.0
    a[far, 0, 2] = np.rint(rel[far])
    a[far, 0, 3] = 0.0
    return a


def chase_and_shoot(obs, rng=None, kick_prob=1.0):
    """Scripted 1v0 policy from a [n, 10] observation (ReachBall layout): go to the ball, kick it towards the
    centre of the right goal when it is close.  Returns [n, 1, 4] float32 commands."""
    obs = np.asarray(obs, np.float64)
    n = obs.shape[0]
    px, py, bx, by = obs[:, 2] * 52.5, obs[:, 3] * 34.0, obs[:, 4] * 52.5, obs[:, 5] * 34.0
    body = obs[:, 1] * 180.0
    dist = np.hypot(bx - px, by - py)
    a = np.zeros((n, 1, 4), np.float32)
    near = dist < 1.0
    goal_dir = np.degrees(np.arctan2(0.0 - by, 52.5 - bx)) - body
    goal_dir = (goal_dir + 180.0) % 360.0 - 180.0
    a[:, 0, 0] = np.where(near, 3, 4)
    a[:, 0, 1] = np.where(near, 100.0, bx)
    a[:, 0, 2] = np.where(near, goal_dir, by)
    a[:, 0, 3] = np.where(near, 0.0, 100.0)
    if rng is not None and kick_prob < 1.0:
        skip = near & (rng.uniform(size=n) > kick_prob)
        a[skip, 0, 0] = 0
    return a


def obs_close(a, b, tol=TOL, angle_scale=1.0):
    """max |a - b| over the columns, angles compared on the circle.  Obs columns are already normalised to O(1).
    `angle_scale` < 1 down-weights the angle columns: the direction of a short vector (player next to the ball, slow
    ball) is ill-conditioned - an absolute position error eps moves it by eps / length."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = np.abs(a - b)
    for col, period in OBS_ANGLE_PERIOD.items():
        d[..., col] = np.minimum(d[..., col], np.abs(period - d[..., col])) * angle_scale
    return d.max() if d.size else 0.0


# natural scales for the 16 float state fields of oracle_lib.STATE_FIELDS (positions: pitch half length,
# velocities: speed max, body / mem_ang: 180 deg, stamina: stamina_max, capacity: stamina_capacity, ...)
STATE_SCALE = np.array([52.5, 34.0, 1.05, 1.05, 180.0, 8000.0, 1.0, 1.0, 130600.0, 52.5, 34.0, 3.0, 3.0,
                        100.0, 180.0, 100.0])


def state_err(a, b):
    """max over envs of |a-b| / scale for the 16 float fields (angles on the circle)."""
    a = np.asarray(a, np.float64)[..., :16]
    b = np.asarray(b, np.float64)[..., :16]
    d = np.abs(a - b)
    for col in (4, 14):
        d[..., col] = np.minimum(d[..., col], np.abs(360.0 - d[..., col]))
    return (d / STATE_SCALE).max()
