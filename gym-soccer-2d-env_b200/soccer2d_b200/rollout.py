"""Device-resident DQN rollout on Soccer2DVecEnv - the caller side of the step path (BASELINE configs[4]).

The reference trains with Stable-Baselines3 (`DQN("MlpPolicy", env)`, dqn_stable_baselines3.py:36-41): one env,
numpy observations, a numpy replay buffer and Python `infos` dicts per step.  With 10^4..10^6 lockstep envs that
plumbing is the bottleneck, so this module keeps the whole loop on the GPU:

  obs (the tensor the step kernel writes) -> Q-network -> epsilon-greedy argmax written straight into the env's
  action tensor -> s2d_step -> (reward, done, terminal obs) -> ring replay buffer in HBM -> TD update.

No host copies, no per-env Python objects; episode outcomes (Goal / Out / Timeout, what InfoCollectorCallback
tallies, utils/info_collector_callback.py:37-53) come from the kernel's statistics counters.
SB3 is not installed in this image; the network shape (64-64 ReLU MLP), the Huber TD loss, the target network
and the epsilon schedule follow SB3's DQN defaults.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import torch
from torch import nn


class QNetwork(nn.Module):
    """SB3 DQN "MlpPolicy": obs -> 64 -> 64 -> n_actions, ReLU."""

    def __init__(self, obs_dim: int, n_actions: int, hidden: int = 64):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(obs_dim, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(),
                                 nn.Linear(hidden, n_actions))

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        return self.net(obs)


class DeviceReplayBuffer:
    """Ring buffer of transitions in GPU memory; whole batches of N envs are appended per step."""

    def __init__(self, capacity: int, obs_dim: int, device):
        self.capacity, self.device = int(capacity), device
        self.obs = torch.empty((capacity, obs_dim), dtype=torch.float32, device=device)
        self.next_obs = torch.empty((capacity, obs_dim), dtype=torch.float32, device=device)
        self.action = torch.empty(capacity, dtype=torch.uint8, device=device)
        self.reward = torch.empty(capacity, dtype=torch.float32, device=device)
        self.done = torch.empty(capacity, dtype=torch.bool, device=device)
        self.pos, self.size = 0, 0

    def add_batch(self, obs, action, reward, next_obs, done):
        n = obs.shape[0]
        if n >= self.capacity:  # keep the newest `capacity` rows
            sl = slice(n - self.capacity, n)
            self.obs.copy_(obs[sl]); self.next_obs.copy_(next_obs[sl]); self.action.copy_(action[sl])
            self.reward.copy_(reward[sl]); self.done.copy_(done[sl])
            self.pos, self.size = 0, self.capacity
            return
        first = min(n, self.capacity - self.pos)
        for dst, src in ((self.obs, obs), (self.next_obs, next_obs), (self.action, action), (self.reward, reward),
                         (self.done, done)):
            dst[self.pos:self.pos + first].copy_(src[:first])
            if first < n:
                dst[:n - first].copy_(src[first:])
        self.pos = (self.pos + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def sample(self, batch: int, generator=None):
        idx = torch.randint(0, self.size, (batch,), device=self.device, generator=generator)
        return self.obs[idx], self.action[idx].long(), self.reward[idx], self.next_obs[idx], self.done[idx]


@dataclass
class DQNConfig:
    gamma: float = 0.99
    lr: float = 1e-4                 # SB3 DQN default
    batch_size: int = 4096
    buffer_size: int = 1 << 21
    learning_starts: int = 1 << 16   # transitions
    train_every: int = 1             # env steps (of all N envs) between gradient steps
    gradient_steps: int = 1
    target_update_every: int = 250   # gradient steps
    eps_start: float = 1.0
    eps_end: float = 0.05
    eps_fraction: float = 0.3        # of total steps (SB3: exploration_fraction... 0.1; larger for short runs)
    seed: int = 0
    log: list = field(default_factory=list)


class DeviceDQN:
    """DQN whose rollout never leaves the GPU.  `env` must be a Discrete-action Soccer2DVecEnv with substeps=1,
    auto_reset=True and terminal_obs=True."""

    def __init__(self, env, cfg: DQNConfig | None = None):
        assert env.substeps == 1 and env.auto_reset and env.terminal_obs is not None
        assert env.actions.dtype == torch.uint8, "Discrete action space required"
        self.env, self.cfg, self.device = env, cfg or DQNConfig(), env.device
        torch.manual_seed(self.cfg.seed)
        self.n_actions = env.action_space.n
        self.q = QNetwork(env.obs_dim, self.n_actions).to(self.device)
        self.q_target = QNetwork(env.obs_dim, self.n_actions).to(self.device)
        self.q_target.load_state_dict(self.q.state_dict())
        self.opt = torch.optim.Adam(self.q.parameters(), lr=self.cfg.lr)
        self.buffer = DeviceReplayBuffer(self.cfg.buffer_size, env.obs_dim, self.device)
        self.gen = torch.Generator(device=self.device).manual_seed(self.cfg.seed)
        self.env_steps = 0       # lockstep steps taken (each = num_envs transitions)
        self.grad_steps = 0
        self._obs = env.reset_torch().clone()

    # -- acting ----------------------------------------------------------------------------------------
    @torch.no_grad()
    def act(self, obs: torch.Tensor, epsilon: float) -> torch.Tensor:
        greedy = self.q(obs).argmax(dim=1)
        if epsilon <= 0.0:
            return greedy.to(torch.uint8)
        n = obs.shape[0]
        explore = torch.rand(n, device=self.device, generator=self.gen) < epsilon
        rand_a = torch.randint(0, self.n_actions, (n,), device=self.device, generator=self.gen)
        return torch.where(explore, rand_a, greedy).to(torch.uint8)

    @torch.no_grad()
    def rollout_step(self, epsilon: float, store: bool = True):
        """One lockstep step of all envs; the action tensor is written in place (zero copy)."""
        env = self.env
        a = self.act(self._obs, epsilon)
        env.actions.view(-1).copy_(a)
        obs, reward, done, _ = env.step_torch()
        if store:
            next_obs = torch.where(done.unsqueeze(1), env.terminal_obs, obs)  # finished envs already show their new episode
            self.buffer.add_batch(self._obs, a, reward, next_obs, done)
        self._obs.copy_(obs)
        self.env_steps += 1

    # -- learning --------------------------------------------------------------------------------------
    def train_step(self):
        c = self.cfg
        obs, act, rew, nxt, done = self.buffer.sample(c.batch_size, self.gen)
        with torch.no_grad():
            target = rew + c.gamma * (~done).float() * self.q_target(nxt).max(dim=1).values
        qsa = self.q(obs).gather(1, act.unsqueeze(1)).squeeze(1)
        loss = nn.functional.smooth_l1_loss(qsa, target)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        nn.utils.clip_grad_norm_(self.q.parameters(), 10.0)  # SB3 max_grad_norm
        self.opt.step()
        self.grad_steps += 1
        if self.grad_steps % c.target_update_every == 0:
            self.q_target.load_state_dict(self.q.state_dict())
        return loss

    def epsilon(self, total_steps: int) -> float:
        c = self.cfg
        frac = min(1.0, self.env_steps / max(1.0, c.eps_fraction * total_steps))
        return c.eps_start + frac * (c.eps_end - c.eps_start)

    def learn(self, total_steps: int, report_every: int = 0):
        """`total_steps` lockstep steps (x num_envs transitions).  Returns the list of report dicts: per interval
        the outcome counts of the finished episodes, like InfoCollectorCallback's per-100-episodes print."""
        c, env = self.cfg, self.env
        last = env.stats()
        t0 = time.perf_counter()
        for step in range(1, total_steps + 1):
            self.rollout_step(self.epsilon(total_steps))
            if self.buffer.size >= c.learning_starts and step % c.train_every == 0:
                for _ in range(c.gradient_steps):
                    self.train_step()
            if report_every and step % report_every == 0:
                now = env.stats()
                d = {k: now[k] - last[k] for k in ("episodes", "goals", "outs", "timeouts", "episode_steps", "return_sum")}
                last = now
                ep = max(1, d["episodes"])
                rep = {"step": step, "transitions": step * env.num_envs, "epsilon": round(self.epsilon(total_steps), 3),
                       "episodes": d["episodes"], "goal_rate": d["goals"] / ep, "out_rate": d["outs"] / ep,
                       "timeout_rate": d["timeouts"] / ep, "mean_return": d["return_sum"] / ep,
                       "mean_length": d["episode_steps"] / ep, "wall_s": round(time.perf_counter() - t0, 2)}
                c.log.append(rep)
        return c.log

    def learn_fused(self, total_steps: int, k: int = 8, report_every: int = 0):
        """The same training loop with the rollout collected by the fused kernel (s2d_rollout_mlp_collect): one launch
        plays `k` epsilon-greedy cycles with the current Q-network inside the step kernel and writes the k x N
        transitions straight into time-major tensors, which are appended to the replay buffer in one go; then
        k / train_every x gradient_steps gradient steps, as `learn` would have made over those cycles.  TF32
        Q-values decide the actions (see Soccer2DVecEnv.rollout_mlp)."""
        c, env = self.cfg, self.env
        n, dev = env.num_envs, self.device
        traj = {"obs": torch.empty((k + 1, n, env.obs_dim), device=dev), "actions": torch.empty((k, n), dtype=torch.uint8, device=dev),
                "reward": torch.empty((k, n), device=dev), "done": torch.empty((k, n), dtype=torch.uint8, device=dev)}
        last, t0 = env.stats(), time.perf_counter()
        next_report = report_every
        for step in range(k, total_steps + 1, k):
            env.rollout_mlp(mlp_layers(self.q), k, self.epsilon(total_steps), traj=traj)
            self.buffer.add_batch(traj["obs"][:k].reshape(k * n, -1), traj["actions"].reshape(-1), traj["reward"].reshape(-1),
                                  traj["obs"][1:].reshape(k * n, -1), traj["done"].reshape(-1).bool())
            self.env_steps += k
            if self.buffer.size >= c.learning_starts:
                for _ in range(max(1, k // c.train_every) * c.gradient_steps):
                    self.train_step()
            if report_every and step >= next_report:
                next_report += report_every
                now = env.stats()
                d = {kk: now[kk] - last[kk] for kk in ("episodes", "goals", "outs", "timeouts", "episode_steps", "return_sum")}
                last = now
                ep = max(1, d["episodes"])
                c.log.append({"step": step, "transitions": step * n, "epsilon": round(self.epsilon(total_steps), 3),
                              "episodes": d["episodes"], "goal_rate": d["goals"] / ep, "out_rate": d["outs"] / ep,
                              "timeout_rate": d["timeouts"] / ep, "mean_return": d["return_sum"] / ep,
                              "mean_length": d["episode_steps"] / ep, "wall_s": round(time.perf_counter() - t0, 2)})
        self._obs.copy_(env.obs)  # what the agent sees next, for rollout_step / evaluate
        return c.log

    @torch.no_grad()
    def evaluate(self, steps: int, fused: bool = False) -> dict:
        """Greedy rollout (the reference's `test()`, dqn_ddpg_stable_baselines3.py:56-75): outcome rates.
        fused: run it with Soccer2DVecEnv.rollout_mlp (TF32 Q-values: the action can differ on near-ties)."""
        before = self.env.stats()
        if fused and self.env.action_mode == 0 and self.env.cfg.action_space_size <= (24 if self.env.scenario == "shoot" else 16):
            # the Q-network inside the step kernel: 16 cycles per launch, nothing leaves the SM in between
            layers = mlp_layers(self.q)
            for lo in range(0, steps, 16):
                self.env.rollout_mlp(layers, min(16, steps - lo))
            self._obs.copy_(self.env.obs)
        else:
            for _ in range(steps):
                self.rollout_step(0.0, store=False)
        after = self.env.stats()
        d = {k: after[k] - before[k] for k in ("episodes", "goals", "outs", "timeouts", "return_sum")}
        ep = max(1, d["episodes"])
        return {"episodes": d["episodes"], "goal_rate": d["goals"] / ep, "out_rate": d["outs"] / ep,
                "timeout_rate": d["timeouts"] / ep, "mean_return": d["return_sum"] / ep}


@torch.no_grad()
def measure_rollout(env, qnet: nn.Module, steps: int, warmup: int = 5, use_graph: bool = True) -> float:
    """env-steps/s of the closed loop obs -> Q-net -> argmax -> step (no learning), timed with CUDA events.
    use_graph: the loop body (policy kernels + the step kernel, launched through the C ABI on the capturing stream)
    is captured once into a CUDA graph and replayed - the launch-bound case of a few thousand envs."""
    def one():
        a = qnet(env.obs).argmax(dim=1).to(torch.uint8)
        env.actions.view(-1).copy_(a)
        env.step_torch()

    run = one
    if use_graph:
        side = torch.cuda.Stream(device=env.device)
        side.wait_stream(torch.cuda.current_stream(env.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                one()
        torch.cuda.current_stream(env.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        cycle = env.state_planes()[1][0, 1]  # server cycle of env 0 (a view of the state the kernel updates)
        before = int(cycle)
        with torch.cuda.graph(graph):
            one()
        assert int(cycle) == before, "capture must not execute"
        graph.replay()
        torch.cuda.synchronize()
        assert int(cycle) in (before + 1, before + 2), "the captured step kernel did not run"  # +2: an auto-reset cycle
        run = graph.replay  # (the handle's host-side env_steps counter does not see replays)
    for _ in range(warmup):
        run()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        run()
    e.record()
    torch.cuda.synchronize()
    return env.num_envs * steps / (s.elapsed_time(e) * 1e-3)


def mlp_layers(qnet: nn.Module) -> list:
    """[(weight, bias)] of the three Linear layers of a QNetwork, as Soccer2DVecEnv.rollout_mlp takes them"""
    lin = [m for m in qnet.modules() if isinstance(m, nn.Linear)]
    assert len(lin) == 3, "a 64-64 MLP has three Linear layers"
    return [(m.weight.detach().contiguous(), m.bias.detach().contiguous()) for m in lin]


@torch.no_grad()
def measure_fused_rollout(env, qnet: nn.Module, launches: int, k: int | None = None, warmup: int = 5) -> float:
    """env-steps/s of the same closed loop with the Q-network INSIDE the step kernel (s2d_rollout_mlp: `k` cycles of
    observe -> Q -> argmax -> step per launch, TF32 tensor-core MMAs), timed with CUDA events."""
    k = env.substeps if k is None else k
    layers = mlp_layers(qnet)
    for _ in range(warmup):
        env.rollout_mlp(layers, k)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(launches):
        env.rollout_mlp(layers, k)
    e.record()
    torch.cuda.synchronize()
    return env.num_envs * k * launches / (s.elapsed_time(e) * 1e-3)


# ---------------------------------------------------------------------------------------------------------
# DDPG (continuous actions) - the counterpart of the reference's ddpg_stable_baselines3.py
# ---------------------------------------------------------------------------------------------------------

class Actor(nn.Module):
    """obs -> action in [-1, 1]^A (tanh), 64-64 ReLU"""

    def __init__(self, obs_dim: int, act_dim: int, hidden: int = 64):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(obs_dim, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(),
                                 nn.Linear(hidden, act_dim), nn.Tanh())

    def forward(self, obs):
        return self.net(obs)


class Critic(nn.Module):
    def __init__(self, obs_dim: int, act_dim: int, hidden: int = 64):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(obs_dim + act_dim, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(),
                                 nn.Linear(hidden, 1))

    def forward(self, obs, act):
        return self.net(torch.cat([obs, act], dim=1)).squeeze(1)


@dataclass
class DDPGConfig:
    gamma: float = 0.99
    lr: float = 1e-3                 # SB3 DDPG default
    tau: float = 0.005
    batch_size: int = 4096
    buffer_size: int = 1 << 21
    learning_starts: int = 1 << 16
    action_noise: float = 0.2        # std of the Gaussian exploration noise (SB3: user-supplied NormalActionNoise)
    seed: int = 0
    log: list = field(default_factory=list)


class DeviceDDPG:
    """DDPG with the rollout, the replay buffer and the learner on the GPU.  `env`: a ReachBall Soccer2DVecEnv with a
    Box action space - Box(-1, 1, (1,)) (use_continuous_action=True: ddpg_stable_baselines3.py) or Box(-1, 1, (4,))
    (use_turning=True: dqn_ddpg_stable_baselines3.py) - with substeps=1, auto_reset and terminal_obs."""

    def __init__(self, env, cfg: DDPGConfig | None = None):
        assert env.substeps == 1 and env.auto_reset and env.terminal_obs is not None
        assert env.actions.dtype == torch.float32 and env.actions.dim() in (2, 3), "a Box action space is required"
        self.env, self.cfg, self.device = env, cfg or DDPGConfig(), env.device
        torch.manual_seed(self.cfg.seed)
        od, ad = env.obs_dim, (1 if env.actions.dim() == 2 else env.actions.shape[-1])
        self.act_dim = ad
        self.actor, self.actor_t = Actor(od, ad).to(self.device), Actor(od, ad).to(self.device)
        self.critic, self.critic_t = Critic(od, ad).to(self.device), Critic(od, ad).to(self.device)
        self.actor_t.load_state_dict(self.actor.state_dict())
        self.critic_t.load_state_dict(self.critic.state_dict())
        self.opt_a = torch.optim.Adam(self.actor.parameters(), lr=self.cfg.lr)
        self.opt_c = torch.optim.Adam(self.critic.parameters(), lr=self.cfg.lr)
        n = self.cfg.buffer_size
        dev = self.device
        self.b_obs = torch.empty((n, od), device=dev)
        self.b_next = torch.empty((n, od), device=dev)
        self.b_act = torch.empty((n, ad), device=dev)
        self.b_rew = torch.empty(n, device=dev)
        self.b_done = torch.empty(n, dtype=torch.bool, device=dev)
        self.pos = self.size = 0
        self.gen = torch.Generator(device=dev).manual_seed(self.cfg.seed)
        self.env_steps = 0
        self._obs = env.reset_torch().clone()

    def _store(self, obs, act, rew, nxt, done):
        n, cap = obs.shape[0], self.cfg.buffer_size
        idx = (torch.arange(n, device=self.device) + self.pos) % cap
        self.b_obs[idx], self.b_act[idx], self.b_rew[idx], self.b_next[idx], self.b_done[idx] = obs, act, rew, nxt, done
        self.pos = (self.pos + n) % cap
        self.size = min(cap, self.size + n)

    @torch.no_grad()
    def rollout_step(self, noise_std: float, store: bool = True, random: bool = False):
        env = self.env
        if random:
            a = torch.rand((env.num_envs, self.act_dim), device=self.device, generator=self.gen) * 2 - 1
        else:
            a = self.actor(self._obs)
            if noise_std > 0:
                a = (a + noise_std * torch.randn(a.shape, device=self.device, generator=self.gen)).clamp_(-1, 1)
        env.actions.view(env.num_envs, self.act_dim).copy_(a)  # [N, 1(, 4)] float32: written in place, zero copy
        obs, reward, done, _ = env.step_torch()
        if store:
            nxt = torch.where(done.unsqueeze(1), env.terminal_obs, obs)
            self._store(self._obs, a, reward, nxt, done)
        self._obs.copy_(obs)
        self.env_steps += 1

    def train_step(self):
        c = self.cfg
        idx = torch.randint(0, self.size, (c.batch_size,), device=self.device, generator=self.gen)
        obs, act, rew, nxt, done = self.b_obs[idx], self.b_act[idx], self.b_rew[idx], self.b_next[idx], self.b_done[idx]
        with torch.no_grad():
            target = rew + c.gamma * (~done).float() * self.critic_t(nxt, self.actor_t(nxt))
        loss_c = nn.functional.mse_loss(self.critic(obs, act), target)
        self.opt_c.zero_grad(set_to_none=True)
        loss_c.backward()
        self.opt_c.step()
        loss_a = -self.critic(obs, self.actor(obs)).mean()
        self.opt_a.zero_grad(set_to_none=True)
        loss_a.backward()
        self.opt_a.step()
        with torch.no_grad():
            for net, tgt in ((self.actor, self.actor_t), (self.critic, self.critic_t)):
                for p, pt in zip(net.parameters(), tgt.parameters()):
                    pt.lerp_(p, c.tau)

    def learn(self, total_steps: int, report_every: int = 0):
        c, env = self.cfg, self.env
        last, t0 = env.stats(), time.perf_counter()
        for step in range(1, total_steps + 1):
            warm = self.size < c.learning_starts
            self.rollout_step(c.action_noise, random=warm)
            if not warm:
                self.train_step()
            if report_every and step % report_every == 0:
                now = env.stats()
                d = {k: now[k] - last[k] for k in ("episodes", "goals", "outs", "timeouts", "return_sum")}
                last = now
                ep = max(1, d["episodes"])
                c.log.append({"step": step, "transitions": step * env.num_envs, "episodes": d["episodes"],
                              "goal_rate": d["goals"] / ep, "out_rate": d["outs"] / ep, "timeout_rate": d["timeouts"] / ep,
                              "mean_return": d["return_sum"] / ep, "wall_s": round(time.perf_counter() - t0, 2)})
        return c.log

    def learn_fused(self, total_steps: int, k: int = 8, report_every: int = 0):
        """`learn` with the rollout collected by the fused kernel (Soccer2DVecEnv.rollout_actor): one launch plays `k`
        cycles with the current actor inside the step kernel (uniform exploration noise of the same variance as the
        Gaussian of `learn`: half-width sqrt(3) * action_noise; uniformly random actions while the buffer warms up) and
        writes the k x N transitions out time-major; then k gradient steps."""
        c, env = self.cfg, self.env
        n, dev, ad = env.num_envs, self.device, self.act_dim
        traj = {"obs": torch.empty((k + 1, n, env.obs_dim), device=dev), "actions": torch.empty((k, n, ad), device=dev),
                "reward": torch.empty((k, n), device=dev), "done": torch.empty((k, n), dtype=torch.uint8, device=dev)}
        last, t0 = env.stats(), time.perf_counter()
        next_report = report_every
        for step in range(k, total_steps + 1, k):
            warm = self.size < c.learning_starts
            env.rollout_actor(mlp_layers(self.actor), k, 1.0 if warm else min(1.0, 3 ** 0.5 * c.action_noise), traj=traj)
            self._store(traj["obs"][:k].reshape(k * n, -1), traj["actions"].reshape(k * n, ad), traj["reward"].reshape(-1),
                        traj["obs"][1:].reshape(k * n, -1), traj["done"].reshape(-1).bool())
            self.env_steps += k
            if not warm:
                for _ in range(k):
                    self.train_step()
            if report_every and step >= next_report:
                next_report += report_every
                now = env.stats()
                d = {kk: now[kk] - last[kk] for kk in ("episodes", "goals", "outs", "timeouts", "return_sum")}
                last = now
                ep = max(1, d["episodes"])
                c.log.append({"step": step, "transitions": step * n, "episodes": d["episodes"],
                              "goal_rate": d["goals"] / ep, "out_rate": d["outs"] / ep, "timeout_rate": d["timeouts"] / ep,
                              "mean_return": d["return_sum"] / ep, "wall_s": round(time.perf_counter() - t0, 2)})
        self._obs.copy_(env.obs)
        return c.log

    @torch.no_grad()
    def evaluate(self, steps: int, fused: bool = False) -> dict:
        before = self.env.stats()
        if fused:
            layers = mlp_layers(self.actor)
            for lo in range(0, steps, 16):
                self.env.rollout_actor(layers, min(16, steps - lo))
            self._obs.copy_(self.env.obs)
            steps = 0
        for _ in range(steps):
            self.rollout_step(0.0, store=False)
        after = self.env.stats()
        d = {k: after[k] - before[k] for k in ("episodes", "goals", "outs", "timeouts", "return_sum")}
        ep = max(1, d["episodes"])
        return {"episodes": d["episodes"], "goal_rate": d["goals"] / ep, "out_rate": d["outs"] / ep,
                "timeout_rate": d["timeouts"] / ep, "mean_return": d["return_sum"] / ep}
