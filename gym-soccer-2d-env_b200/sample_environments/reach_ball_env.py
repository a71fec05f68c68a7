"""ReachBallEnv - the reference's only concrete scenario (sample_environments/reach_ball_env.py:17-218):
one player has to get within `min_distance_to_ball` of the ball.

Same constructor (`render_mode, logger, log_dir, **kwargs` with the 11 kwargs of :26-36 and their
defaults), same spaces (:39-48), same step/reset results.  Action decode (:53-85), observation (:87-111),
reward/done/info (:113-161) and the reset distribution (:170-218) are implemented in
csrc/s2d_scenarios.cuh and run on the GPU; the extra keyword-only arguments `device`, `seed` and
`server_param` select where and how the episode is simulated.
"""
from __future__ import annotations

from soccer_2d_env import Soccer2DEnv
from soccer2d_b200.vec_env import REACHBALL_DEFAULTS


class ReachBallEnv(Soccer2DEnv):
    scenario = "reachball"

    def __init__(self, render_mode=None, logger=None, log_dir=None, *, device="cuda", seed: int = 0,
                 server_param: dict | None = None, use_command_action: bool = False, noise: bool = False,
                 **kwargs):
        # the reference reads kwargs with .get() and ignores unknown keys (:26-36)
        known = {k: kwargs[k] for k in REACHBALL_DEFAULTS if k in kwargs}
        super().__init__(render_mode, logger=logger, log_dir=log_dir, device=device, seed=seed,
                         server_param=server_param, use_command_action=use_command_action, noise=noise, **known)
        for k, v in dict(REACHBALL_DEFAULTS, **known).items():
            setattr(self, k, v)
