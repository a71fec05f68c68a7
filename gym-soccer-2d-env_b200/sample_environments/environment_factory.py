"""EnvironmentFactory - name -> env (sample_environments/environment_factory.py:14-28): case-insensitive
"reachball", anything else raises ValueError with the reference's message.  `create_vec` is the new
vectorised entry point (N lockstep episodes on one GPU)."""
from __future__ import annotations

from sample_environments.reach_ball_env import ReachBallEnv
from sample_environments.full_game_env import FullGameEnv
from sample_environments.shoot_env import ShootEnv
from soccer_2d_env import Soccer2DEnv
from soccer2d_b200.vec_env import Soccer2DVecEnv


class EnvironmentFactory:
    def create(self, env_name: str, render_mode: str, logger, log_dir: str, **kwargs) -> Soccer2DEnv:
        if env_name.lower() == "reachball":
            return ReachBallEnv(render_mode=render_mode, logger=logger, log_dir=log_dir, **kwargs)
        if env_name.lower() == "shoot":  # new scenario, not in the reference
            return ShootEnv(render_mode=render_mode, logger=logger, log_dir=log_dir, **kwargs)
        if env_name.lower() == "fullgame":  # new scenario, not in the reference
            return FullGameEnv(render_mode=render_mode, logger=logger, log_dir=log_dir, **kwargs)
        raise ValueError(f"Environment {env_name} not found.")

    def create_vec(self, env_name: str, num_envs: int, **kwargs) -> Soccer2DVecEnv:
        if env_name.lower() in ("reachball", "shoot", "fullgame"):
            return Soccer2DVecEnv(num_envs, scenario=env_name.lower(), **kwargs)
        raise ValueError(f"Environment {env_name} not found.")
