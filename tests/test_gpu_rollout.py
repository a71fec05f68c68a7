"""GPU: the device-resident DQN rollout (BASELINE configs[4]) on top of the step path."""
import numpy as np
import pytest
import torch

import helpers as H
import oracle_lib as OL
from soccer2d_b200 import Soccer2DVecEnv
from soccer2d_b200.rollout import DeviceDQN, DeviceReplayBuffer, DQNConfig, QNetwork, measure_rollout

pytestmark = pytest.mark.gpu
KW = dict(use_continuous_action=False, action_space_size=16, change_ball_velocity=True)


def test_replay_buffer_ring():
    buf = DeviceReplayBuffer(10, 3, "cuda:0")
    mk = lambda lo, n: (torch.arange(lo, lo + n, device="cuda").float().unsqueeze(1).repeat(1, 3),)  # noqa: E731
    for lo, n in ((0, 4), (4, 4), (8, 4)):
        o = mk(lo, n)[0]
        buf.add_batch(o, torch.zeros(n, dtype=torch.uint8, device="cuda"), o[:, 0], o + 0.5, o[:, 0] > 100)
    assert buf.size == 10 and buf.pos == 2
    assert sorted(buf.obs[:, 0].tolist()) == [2, 3, 4, 5, 6, 7, 8, 9, 10, 11]
    assert torch.equal(buf.next_obs, buf.obs + 0.5) and torch.equal(buf.reward, buf.obs[:, 0])
    o = mk(100, 25)[0]
    buf.add_batch(o, torch.zeros(25, dtype=torch.uint8, device="cuda"), o[:, 0], o, o[:, 0] > 0)
    assert sorted(buf.obs[:, 0].tolist()) == list(range(115, 125))


def test_rollout_transitions_match_the_oracle():
    """What the rollout stores (obs, action, reward, next_obs incl. terminal observations, done) is exactly what
    the fp32 oracle produces for the same actions."""
    n = 512
    env = Soccer2DVecEnv(n, device="cuda:0", seed=5, terminal_obs=True, max_steps=20, **KW)
    agent = DeviceDQN(env, DQNConfig(buffer_size=n * 64, seed=1))
    sim = OL.OracleSim(env.cfg, "f32")
    prev = sim.reset().copy()
    assert np.array_equal(agent._obs.cpu().numpy(), prev)
    for t in range(40):
        agent.rollout_step(epsilon=0.5)
        sl = slice(t * n, (t + 1) * n)
        act = agent.buffer.action[sl].cpu().numpy()
        sim.step(act.reshape(n, 1))
        d = sim.done.astype(bool)
        want_next = np.where(d[:, None], sim.term_obs, sim.obs)
        assert np.array_equal(agent.buffer.obs[sl].cpu().numpy(), prev)
        assert np.array_equal(agent.buffer.reward[sl].cpu().numpy(), sim.reward)
        assert np.array_equal(agent.buffer.done[sl].cpu().numpy(), d)
        assert np.array_equal(agent.buffer.next_obs[sl].cpu().numpy(), want_next)
        prev = sim.obs.copy()
    assert agent.buffer.size == 40 * n and int(agent.buffer.done.sum()) > n


def test_dqn_learns_to_reach_the_ball():
    """Short training run: the greedy policy must reach the ball far more often than the random policy."""
    env = Soccer2DVecEnv(4096, device="cuda:0", seed=0, terminal_obs=True, **KW)
    agent = DeviceDQN(env, DQNConfig(seed=0, learning_starts=1 << 15, lr=5e-4, eps_fraction=0.5))
    random_policy = agent.evaluate(0)
    for _ in range(250):
        agent.rollout_step(1.0, store=False)
    st = env.stats()
    random_goal_rate = st["goals"] / max(1, st["episodes"])
    agent.learn(1500)
    trained = agent.evaluate(400)
    assert trained["episodes"] > 4096 and random_policy["episodes"] == 0
    assert trained["goal_rate"] > max(0.6, 3 * random_goal_rate), (trained, random_goal_rate)


def test_measure_rollout_runs():
    env = Soccer2DVecEnv(1 << 16, device="cuda:0", seed=0, **KW)
    env.reset_torch()
    q = QNetwork(10, 16).to("cuda:0")
    rate = measure_rollout(env, q, steps=20)
    assert rate > 1e7


def test_ddpg_learns_to_reach_the_ball():
    """continuous Dash direction (ddpg_stable_baselines3.py's configuration): the trained actor beats the random policy"""
    from soccer2d_b200.rollout import DDPGConfig, DeviceDDPG
    kw = dict(change_ball_position=False, use_continuous_action=True, use_turning=False)
    env = Soccer2DVecEnv(4096, device="cuda:0", seed=0, terminal_obs=True, **kw)
    agent = DeviceDDPG(env, DDPGConfig(seed=0, learning_starts=1 << 15))
    for _ in range(250):
        agent.rollout_step(0.0, store=False, random=True)
    st = env.stats()
    random_goal_rate = st["goals"] / max(1, st["episodes"])
    agent.learn(1500)
    trained = agent.evaluate(400)
    assert trained["episodes"] > 4096
    assert trained["goal_rate"] > max(0.6, 2 * random_goal_rate), (trained, random_goal_rate)
