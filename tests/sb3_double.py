"""A test double for the part of Stable-Baselines3 2.4 that drives a VecEnv during off-policy training.

SB3 is not installed in this image, so the reference's caller (`DQN("MlpPolicy", env).learn(..., callback=
InfoCollectorCallback())`, dqn_stable_baselines3.py:36-41, and the DDPG twin, ddpg_stable_baselines3.py:18-37) cannot run
here.  This module restates - from SB3's documented behaviour, in its own words - every call those scripts make on the
environment object, so that `Soccer2DVecEnv` is exercised through the same contract:

  VecEnvWrapper / VecMonitor   wrapping (`num_envs`, spaces, `env_is_wrapped(Monitor)`, `get_attr("render_mode")`),
                               `reset()`, `step_async` + `step_wait`, `infos[:]` copied into a list, `infos[i].copy()`
                               for finished episodes and an "episode" record {r, l, t} added to it
  collect_rollouts             `env.step(actions)` each iteration, `callback.update_locals(locals())` + `on_step()`,
                               `_update_info_buffer` (reads `info.get("episode")`, `info.get("is_success")` of EVERY info),
                               `_store_transition` (replaces next_obs[i] by `infos[i]["terminal_observation"]` when done[i]),
                               ReplayBuffer.add (reads `info.get("TimeLimit.truncated", False)` of EVERY info)
  InfoCollectorCallback        reads `self.locals["infos"]`, keeps the infos with a non-empty `info["result"]`
                               (utils/info_collector_callback.py:17-28)

Test infrastructure only."""
import time

import numpy as np


class Monitor:  # only ever used as the class argument of env_is_wrapped
    pass


class VecEnvWrapperDouble:
    def __init__(self, venv):
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space = venv.observation_space
        self.action_space = venv.action_space
        # SB3's VecEnv.__init__ asks the env for its render mode
        modes = venv.get_attr("render_mode")
        assert len(modes) == venv.num_envs and all(m == modes[0] for m in modes[:8])
        self.render_mode = modes[0]

    def step_async(self, actions):
        self.venv.step_async(actions)

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def reset(self):
        return self.venv.reset()

    def close(self):
        return self.venv.close()

    def get_attr(self, name, indices=None):
        return self.venv.get_attr(name, indices)

    def set_attr(self, name, value, indices=None):
        return self.venv.set_attr(name, value, indices)

    def env_method(self, name, *args, indices=None, **kwargs):
        return self.venv.env_method(name, *args, indices=indices, **kwargs)

    def env_is_wrapped(self, wrapper_class, indices=None):
        return self.venv.env_is_wrapped(wrapper_class, indices=indices)

    def seed(self, seed=None):
        return self.venv.seed(seed)


class VecMonitorDouble(VecEnvWrapperDouble):
    """episode return / length bookkeeping the way VecMonitor does it"""

    def __init__(self, venv):
        super().__init__(venv)
        try:
            already = venv.env_is_wrapped(Monitor)[0]
        except AttributeError:
            already = False
        assert already is False
        self.episode_count = 0
        self.t_start = time.time()
        self.episode_returns = None
        self.episode_lengths = None

    def reset(self):
        obs = self.venv.reset()
        self.episode_returns = np.zeros(self.num_envs, dtype=np.float32)
        self.episode_lengths = np.zeros(self.num_envs, dtype=np.int32)
        return obs

    def step_wait(self):
        obs, rewards, dones, infos = self.venv.step_wait()
        self.episode_returns += rewards
        self.episode_lengths += 1
        new_infos = list(infos[:])
        for i in range(len(dones)):
            if dones[i]:
                info = infos[i].copy()
                info["episode"] = {"r": self.episode_returns[i], "l": self.episode_lengths[i],
                                   "t": round(time.time() - self.t_start, 6)}
                self.episode_count += 1
                self.episode_returns[i] = 0
                self.episode_lengths[i] = 0
                new_infos[i] = info
        return obs, rewards, dones, new_infos


class ReplayBufferDouble:
    def __init__(self, size, num_envs, obs_shape, action_shape, action_dtype):
        self.size, self.pos, self.full = max(size // num_envs, 1), 0, False
        self.observations = np.zeros((self.size, num_envs) + obs_shape, np.float32)
        self.next_observations = np.zeros((self.size, num_envs) + obs_shape, np.float32)
        self.actions = np.zeros((self.size, num_envs) + action_shape, action_dtype)
        self.rewards = np.zeros((self.size, num_envs), np.float32)
        self.dones = np.zeros((self.size, num_envs), np.float32)
        self.timeouts = np.zeros((self.size, num_envs), np.float32)

    def add(self, obs, next_obs, action, reward, done, infos):
        self.observations[self.pos] = obs
        self.next_observations[self.pos] = next_obs
        self.actions[self.pos] = np.asarray(action).reshape(self.actions.shape[1:])
        self.rewards[self.pos] = reward
        self.dones[self.pos] = done
        self.timeouts[self.pos] = np.array([info.get("TimeLimit.truncated", False) for info in infos])
        self.pos += 1
        if self.pos == self.size:
            self.full, self.pos = True, 0


class CallbackDouble:
    """the BaseCallback members collect_rollouts touches, around the reference's InfoCollectorCallback logic"""

    def __init__(self):
        self.locals, self.infos, self.n_calls = {}, [], 0

    def update_locals(self, locals_):
        self.locals.update(locals_)

    def on_step(self):
        self.n_calls += 1
        infos = self.locals.get("infos")
        if infos is not None:
            for info in infos:
                if info["result"] and len(info["result"]) > 0:
                    self.infos.append(info)
        return True


class OffPolicyDouble:
    """the env-facing half of OffPolicyAlgorithm: `_last_obs`, `collect_rollouts`, `_update_info_buffer`,
    `_store_transition`"""

    def __init__(self, env, policy, buffer_size=1 << 18):
        self.env = env if isinstance(env, VecEnvWrapperDouble) else VecMonitorDouble(env)
        self.policy = policy
        shape = getattr(env.action_space, "shape", ()) or ()
        discrete = hasattr(env.action_space, "n")
        self.replay_buffer = ReplayBufferDouble(buffer_size, env.num_envs, env.observation_space.shape,
                                                (1,) if discrete else tuple(shape), np.int64 if discrete else np.float32)
        self.num_timesteps = 0
        self.episode_num = 0
        self.ep_info_buffer = []
        self._last_obs = None

    def _update_info_buffer(self, infos, dones=None):
        for info in infos:
            ep, ok = info.get("episode"), info.get("is_success")
            if ep is not None:
                self.ep_info_buffer.append(ep)
            assert ok is None

    def _store_transition(self, buffer_action, new_obs, reward, dones, infos):
        next_obs = new_obs.copy()
        for i, done in enumerate(dones):
            if done and infos[i].get("terminal_observation") is not None:
                next_obs[i] = infos[i]["terminal_observation"]
        self.replay_buffer.add(self._last_obs, next_obs, buffer_action, reward, dones, infos)
        self._last_obs = new_obs

    def collect_rollouts(self, callback, n_steps):
        env = self.env
        if self._last_obs is None:
            self._last_obs = env.reset()
            assert self._last_obs.shape == (env.num_envs,) + env.observation_space.shape
        for _ in range(n_steps):
            actions = self.policy(self._last_obs)
            new_obs, rewards, dones, infos = env.step(actions)
            self.num_timesteps += env.num_envs
            callback.update_locals(locals())
            if not callback.on_step():
                return False
            self._update_info_buffer(infos, dones)
            self._store_transition(actions, new_obs, rewards, dones, infos)
            self.episode_num += int(np.count_nonzero(dones))
        return True
