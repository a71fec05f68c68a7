"""Soccer2DEnv - the gym.Env base class of the reference (soccer_2d_env.py:23-446), same public surface,
new internals.

The reference's constructor spawns a gRPC server process, an rcssserver and a proxy, and every `step`
blocks on four multiprocessing queues (soccer_2d_env.py:226-269).  Here the episode lives in GPU memory
inside a one-env `Soccer2DVecEnv` and `step` is one C-ABI call (s2d_step): one fused kernel (decode + server
cycle + reward/done + observation) that reads the action from and writes the results to pinned host memory.

Kept: constructor signature (`render_mode, run_grpc_server, run_rcssserver, run_trainer_player, logger,
log_dir`; the three run_* flags are accepted and ignored - there is nothing to spawn), `metadata`,
`action_space` / `observation_space`, `reset() -> obs`, `step(a) -> (obs, reward, done, info)` (old-gym
4-tuple, soccer_2d_env.py:269), `render`, idempotent `close`, and the names of the four scenario hooks
(soccer_2d_env.py:317-354).  The hooks are not called per step: scenario logic is fused into the kernel.
Dropped by construction: queues, cycle-desync repair (`_wait_for_agents`, :141-177), per-step logging.
"""
from __future__ import annotations

import logging

import numpy as np

from soccer2d_b200 import _abi
from soccer2d_b200.spaces import Box, Discrete, Env
from soccer2d_b200.vec_env import Soccer2DVecEnv


class Soccer2DEnv(Env):
    metadata = {"render.modes": ["human"]}  # soccer_2d_env.py:28
    scenario = "reachball"

    def __init__(self, render_mode: str = None, run_grpc_server: bool = True, run_rcssserver: bool = True,
                 run_trainer_player: bool = True, logger: logging.Logger = None, log_dir: str = None,
                 *, device="cuda", seed: int = 0, server_param: dict | None = None, use_command_action: bool = False,
                 noise: bool = False, **scenario_kwargs):
        self.log_dir = log_dir
        self.logger = logger
        if self.logger is None:
            self.logger = logging.getLogger(type(self).__name__)
            self.logger.addHandler(logging.NullHandler())
        self.render_mode = render_mode
        self.run_grpc_server = run_grpc_server      # accepted for signature compatibility; nothing is spawned
        self.run_rcssserver = run_rcssserver
        self.run_trainer_player = run_trainer_player
        # defaults of the reference base class (soccer_2d_env.py:58-59); scenario classes replace them
        self.action_space = Discrete(4)
        self.observation_space = Box(low=-1, high=1, shape=(2,), dtype=np.float32)
        self._vec = Soccer2DVecEnv(1, scenario=self.scenario, device=device, seed=seed, substeps=1, auto_reset=False,
                                   server_param=server_param, use_command_action=use_command_action,
                                   noise=noise, host_mapped_io=True, **scenario_kwargs)
        self.action_space = self._vec.action_space
        self.observation_space = self._vec.observation_space
        self.step_number = 0
        self.logger.info("Soccer2DEnv ready: 1 episode on %s (no rcssserver / proxy / gRPC processes)", self._vec.device)

    # ---- gym API ------------------------------------------------------------------------------------
    def reset(self) -> np.ndarray:
        return self.abs_reset()

    def abs_reset(self) -> np.ndarray:
        player_observation, _trainer_state = self.env_reset()
        return player_observation

    def env_reset(self) -> tuple:
        """Placement, recover, one idle server cycle, reward priming - all inside s2d_reset
        (soccer_2d_env.py:179-206 + reach_ball_env.py:163-197).  Returns (obs, snapshot of the env)."""
        obs = self._vec.reset()
        self.step_number = 0
        return obs[0].copy(), self._vec.export_env(0)

    def step(self, action) -> tuple:
        a = self._shape_action(action)
        obs, reward, done, result = self._vec.step_host(a)
        self.step_number += 1
        return obs[0].copy(), float(reward[0]), bool(done[0]), {"result": _abi.RESULT_NAMES[int(result[0])]}

    def render(self, mode="human"):
        return None

    def close(self):
        vec = getattr(self, "_vec", None)
        if vec is not None:
            vec.close()

    # ---- helpers ------------------------------------------------------------------------------------
    def _shape_action(self, action):
        mode = self._vec.action_mode
        if mode == _abi.ACT_DISCRETE:
            a = int(np.asarray(action).reshape(-1)[0])
            n = self._vec.cfg.action_space_size
            if not 0 <= a < n:
                raise ValueError(f"discrete action {a} outside Discrete({n})")
            return np.array([[a]], dtype=np.uint8)
        a = np.asarray(action, dtype=np.float32).reshape(-1)
        want = 4 if mode in (_abi.ACT_TURNING, _abi.ACT_COMMAND) else 1
        if a.size != want:
            raise ValueError(f"expected an action with {want} value(s), got shape {np.asarray(action).shape}")
        return a.reshape((1, 1, 4) if want == 4 else (1, 1))

    @property
    def distance_to_ball(self) -> float:
        return float(self._vec.export_env(0).mem_distance_to_ball)

    @property
    def body_ball_angle_diff(self) -> float:
        return float(self._vec.export_env(0).mem_body_ball_angle_diff)

    # ---- the reference's scenario hooks (soccer_2d_env.py:317-354): names kept for subclasses that
    # ---- introspect them; the GPU path does not call them -------------------------------------------
    def action_to_rpc_actions(self, action, player_state=None):
        raise NotImplementedError("action decode is fused into the step kernel (csrc/s2d_scenarios.cuh)")

    def state_to_observation(self, state=None):
        raise NotImplementedError("observation build is fused into the step kernel (csrc/s2d_scenarios.cuh)")

    def check_trainer_observation(self, observation=None):
        raise NotImplementedError("reward/done is fused into the step kernel (csrc/s2d_scenarios.cuh)")

    def trainer_reset_actions(self):
        raise NotImplementedError("reset placement is drawn inside the reset kernel (csrc/s2d_scenarios.cuh)")
