"""FullGameEnv - a whole match, up to 11 v 11 (BASELINE configs[3]) behind the gym API: every step takes one
proto-style command {cmd, a, b, c} per player (shape [2 * players_per_side, 4]: Dash / Turn / Kick / Body_GoToPoint)
and returns the 120-value global observation (ball, 22 x {x, y, vx, vy, body}, play mode, side, scores, time), the
left team's reward, done at time over and info['result'] in {'Goal' (left wins), 'Out' (right wins), 'Timeout' (draw)}.

Not part of the reference (one player, referee off); spec in include/soccer2d.h, kernels in csrc/s2d_fullgame.cuh.
"""
from __future__ import annotations

import numpy as np

from soccer_2d_env import Soccer2DEnv
from soccer2d_b200.vec_env import FULLGAME_DEFAULTS


class FullGameEnv(Soccer2DEnv):
    scenario = "fullgame"

    def __init__(self, render_mode=None, logger=None, log_dir=None, *, device="cuda", seed: int = 0,
                 server_param: dict | None = None, noise: bool = False, **kwargs):
        known = {k: kwargs[k] for k in FULLGAME_DEFAULTS if k in kwargs}
        super().__init__(render_mode, logger=logger, log_dir=log_dir, device=device, seed=seed,
                         server_param=server_param, noise=noise, **known)
        for k, v in dict(FULLGAME_DEFAULTS, **known).items():
            setattr(self, k, v)

    def _shape_action(self, action):
        a = np.asarray(action, dtype=np.float32)
        want = (self._vec.num_players, 4)
        if a.shape != want:
            raise ValueError(f"expected one command per player, shape {want}, got {a.shape}")
        return a.reshape((1, 1) + want)
