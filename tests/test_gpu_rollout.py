"""GPU: the device-resident DQN rollout (BASELINE configs[4]) on top of the step path."""
import numpy as np
import pytest
import torch

import helpers as H
import oracle_lib as OL
from soccer2d_b200 import Soccer2DVecEnv, _abi
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
    # the same greedy evaluation with the Q-network inside the step kernel (TF32): the same policy, the same outcome rates
    fused = agent.evaluate(400, fused=True)
    assert fused["episodes"] > 4096 and abs(fused["goal_rate"] - trained["goal_rate"]) < 0.05, (fused, trained)


def test_measure_rollout_runs():
    from soccer2d_b200.rollout import measure_fused_rollout
    env = Soccer2DVecEnv(1 << 16, device="cuda:0", seed=0, **KW)
    env.reset_torch()
    q = QNetwork(10, 16).to("cuda:0")
    rate = measure_rollout(env, q, steps=20)
    assert rate > 1e7
    assert measure_fused_rollout(env, q, launches=10, k=16) > 5 * rate


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


# ---- the Q-network inside the kernel (s2d_rollout_mlp) ---------------------------------------------------------------

def _layers(q):
    lin = [m for m in q.net if isinstance(m, torch.nn.Linear)]
    return [(l.weight.detach().contiguous(), l.bias.detach().contiguous()) for l in lin]


@pytest.mark.parametrize("precision", ["tf32", "tf32_mma_sync"])
@pytest.mark.parametrize("scenario,n_actions,n", [("reachball", 16, 4096 + 37), ("reachball", 12, 1000), ("shoot", 24, 2000 + 5)])
def test_fused_policy_rollout_matches_policy_then_step(scenario, n_actions, n, precision):
    """K cycles of observe -> Q -> argmax -> step in one launch.  The env half is checked bit for bit (replaying the
    recorded actions through the ordinary step kernel gives the same state), the policy half against torch fp32:
    Q-values within TF32 accuracy, and the greedy action equal except where fp32 itself sees a near-tie.
    precision: "tf32" = tcgen05.mma with the accumulators in tensor memory (default), "tf32_mma_sync" = warp-level mma.sync."""
    from soccer2d_b200.rollout import QNetwork
    torch.manual_seed(0)
    k = 5
    kw = dict(device="cuda:0", seed=4, scenario=scenario, action_space_size=n_actions, change_ball_velocity=True)
    if scenario == "reachball":
        kw["use_continuous_action"] = False
    width = 24 if scenario == "shoot" else 16
    fused = Soccer2DVecEnv(n, substeps=k, **kw)
    plain = Soccer2DVecEnv(n, substeps=1, **kw)
    qnet = QNetwork(10, n_actions).cuda()
    with torch.no_grad():  # weights of a trained net are O(1): make the test's larger than the default init
        for p in qnet.parameters():
            p.mul_(3.0)
    layers = _layers(qnet)
    fused.reset_torch()
    plain.reset_torch()
    actions = torch.zeros((n, k), dtype=torch.uint8, device="cuda")
    q_seen = torch.zeros((n, width), dtype=torch.float32, device="cuda")
    agree = total = 0
    worst_q = 0.0
    for launch in range(50):  # 250 cycles: episodes end and restart inside the launches
        fused.rollout_mlp(layers, k, 0.0, actions, q_seen, precision=precision)
        for j in range(k):
            with torch.no_grad():
                q32 = qnet(plain.obs)
            a32 = q32.argmax(dim=1)
            taken = actions[:, j].long()
            agree += int((a32 == taken).sum())
            total += n
            # where the kernel chose differently, fp32 itself must see (almost) a tie between the two actions
            gap = (q32.gather(1, a32[:, None]) - q32.gather(1, taken[:, None])).squeeze(1)
            scale = q32.abs().max(dim=1).values + 1.0
            assert float((gap / scale).max()) < 5e-3
            if j == k - 1:
                worst_q = max(worst_q, float(((q_seen[:, :n_actions] - q32).abs().max(dim=1).values / scale).max()))
                if n_actions < width:
                    assert float(q_seen[:, n_actions:].max()) < -1e38
            plain.step_torch(actions[:, j:j + 1].contiguous())
        assert torch.equal(fused.state, plain.state) and torch.equal(fused.obs, plain.obs)
        assert torch.equal(fused.done_u8, plain.done_u8) or k > 1  # (done of a fused launch = any cycle ended)
    assert worst_q < 3e-3, worst_q
    assert agree / total > 0.985, agree / total
    sf, sp = fused.stats(), plain.stats()
    assert sf["episodes"] == sp["episodes"] > 0 and sf["env_steps"] == sp["env_steps"] == n * k * 50
    fused.close()
    plain.close()


def test_fused_policy_rollout_epsilon_and_errors():
    from soccer2d_b200.rollout import QNetwork
    n, k = 8192, 8
    env = Soccer2DVecEnv(n, device="cuda:0", seed=1, substeps=k, use_continuous_action=False, action_space_size=16)
    qnet = QNetwork(10, 16).cuda()
    layers = _layers(qnet)
    env.reset_torch()
    greedy = torch.zeros((n, k), dtype=torch.uint8, device="cuda")
    env.rollout_mlp(layers, k, 0.0, greedy)
    env.reset_torch()  # (a new episode number: other start states, so compare distributions, not actions)
    explore = torch.zeros((n, k), dtype=torch.uint8, device="cuda")
    env.rollout_mlp(layers, k, 1.0, explore)
    counts = torch.bincount(explore.flatten().long(), minlength=16).float()
    assert float(counts.min()) > 0.8 * n * k / 16 and float(counts.max()) < 1.2 * n * k / 16  # uniform over 16 actions
    assert torch.bincount(greedy.flatten().long(), minlength=16).float().max() > 2.0 * n * k / 16  # a policy is not uniform
    with pytest.raises(ValueError):
        env.rollout_mlp(layers[:2] + [(layers[2][0][:8].contiguous(), layers[2][1][:8].contiguous())], k)
    cont = Soccer2DVecEnv(64, device="cuda:0", use_continuous_action=True)
    with pytest.raises(_abi.Soccer2DError):
        cont.rollout_mlp([(torch.zeros(64, 10, device="cuda"), torch.zeros(64, device="cuda")),
                          (torch.zeros(64, 64, device="cuda"), torch.zeros(64, device="cuda")),
                          (torch.zeros(16, 64, device="cuda"), torch.zeros(16, device="cuda"))], 1)
    env.close()
    cont.close()


def test_ddpg_drives_the_turning_action_space():
    """Box(4) = [turn_prob, turn_angle, dash_prob, dash_angle] (dqn_ddpg_stable_baselines3.py's configuration): the agent
    fills the [N, 1, 4] action tensor in place and learns on it (a short run: it must at least beat the random policy)."""
    from soccer2d_b200.rollout import DDPGConfig, DeviceDDPG
    kw = dict(change_ball_position=False, use_continuous_action=True, use_turning=True)
    env = Soccer2DVecEnv(2048, device="cuda:0", seed=0, terminal_obs=True, **kw)
    assert tuple(env.actions.shape) == (2048, 1, 4)
    agent = DeviceDDPG(env, DDPGConfig(seed=0, learning_starts=1 << 14))
    assert agent.act_dim == 4
    for _ in range(250):
        agent.rollout_step(0.0, store=False, random=True)
    st = env.stats()
    random_goal_rate = st["goals"] / max(1, st["episodes"])
    agent.learn(1200)
    trained = agent.evaluate(300)
    assert trained["episodes"] > 1000 and trained["goal_rate"] > random_goal_rate + 0.1, (trained, random_goal_rate)
    env.close()


def test_fused_rollout_collects_the_transitions_a_replay_buffer_needs():
    """s2d_rollout_mlp_collect: obs[k] -> action[k] -> reward[k], done[k], obs[k + 1], time-major.  Replaying the recorded
    actions through the ordinary step kernel, one cycle per launch, reproduces every recorded tensor bit for bit."""
    from soccer2d_b200.rollout import mlp_layers
    torch.manual_seed(1)
    n, k = 3000 + 11, 6
    kw = dict(device="cuda:0", seed=8, use_continuous_action=False, action_space_size=16, change_ball_velocity=True, max_steps=40)
    fused = Soccer2DVecEnv(n, substeps=k, **kw)
    plain = Soccer2DVecEnv(n, substeps=1, **kw)
    qnet = QNetwork(10, 16).cuda()
    fused.reset_torch()
    plain.reset_torch()
    traj = {"obs": torch.zeros((k + 1, n, 10), device="cuda"), "actions": torch.zeros((k, n), dtype=torch.uint8, device="cuda"),
            "reward": torch.zeros((k, n), device="cuda"), "done": torch.zeros((k, n), dtype=torch.uint8, device="cuda")}
    ended = 0
    for launch in range(12):
        first_obs = plain.obs.clone()
        fused.rollout_mlp(mlp_layers(qnet), k, 0.2, traj=traj)
        assert torch.equal(traj["obs"][0], first_obs)
        for j in range(k):
            plain.step_torch(traj["actions"][j].unsqueeze(1).contiguous())
            assert torch.equal(traj["obs"][j + 1], plain.obs)
            assert torch.equal(traj["reward"][j], plain.reward) and torch.equal(traj["done"][j], plain.done_u8)
        assert torch.equal(fused.state, plain.state) and torch.equal(fused.obs, plain.obs)
        assert torch.allclose(traj["reward"].sum(dim=0), fused.reward, rtol=1e-5, atol=1e-5)
        ended += int(traj["done"].sum())
    assert ended == fused.stats()["episodes"] > n
    fused.close()
    plain.close()


def test_dqn_learns_from_fused_collection():
    env = Soccer2DVecEnv(4096, device="cuda:0", seed=0, terminal_obs=True, **KW)
    agent = DeviceDQN(env, DQNConfig(seed=0, learning_starts=1 << 15, lr=5e-4, eps_fraction=0.5))
    log = agent.learn_fused(1504, k=8, report_every=752)
    assert len(log) == 2 and log[-1]["transitions"] == 1504 * 4096
    trained = agent.evaluate(400, fused=True)
    assert trained["episodes"] > 4096 and trained["goal_rate"] > 0.6, trained
    agent.rollout_step(0.0, store=False)  # the unfused path carries on from the fused one's observation
    env.close()


@pytest.mark.parametrize("precision", ["tf32", "tf32_mma_sync"])
@pytest.mark.parametrize("turning", [False, True])
def test_fused_actor_rollout_matches_actor_then_step(turning, precision):
    """s2d_rollout_actor_collect: the DDPG actor inside the step kernel (tcgen05, and the warp-level mma.sync kernel),
    Box(1) and Box(4) action spaces.  The recorded
    actions equal torch's fp32 actor on the recorded observations to TF32 accuracy; replaying them through the ordinary
    step kernel reproduces every recorded observation, reward and done flag bit for bit."""
    from soccer2d_b200.rollout import Actor, mlp_layers
    torch.manual_seed(2)
    n, k = 2048 + 9, 5
    ad = 4 if turning else 1
    kw = dict(device="cuda:0", seed=6, use_continuous_action=True, use_turning=turning, change_ball_velocity=True, max_steps=50)
    fused = Soccer2DVecEnv(n, substeps=k, **kw)
    plain = Soccer2DVecEnv(n, substeps=1, **kw)
    actor = Actor(10, ad).cuda()
    with torch.no_grad():
        for p in actor.parameters():
            p.mul_(2.0)
    fused.reset_torch()
    plain.reset_torch()
    traj = {"obs": torch.zeros((k + 1, n, 10), device="cuda"), "actions": torch.zeros((k, n, ad), device="cuda"),
            "reward": torch.zeros((k, n), device="cuda"), "done": torch.zeros((k, n), dtype=torch.uint8, device="cuda")}
    worst = 0.0
    for launch in range(20):
        fused.rollout_actor(mlp_layers(actor), k, 0.0, traj=traj, precision=precision)
        assert torch.equal(traj["obs"][0], plain.obs)
        for j in range(k):
            with torch.no_grad():
                want = actor(plain.obs)
            worst = max(worst, float((traj["actions"][j] - want).abs().max()))
            plain.step_torch(traj["actions"][j].reshape(plain.actions.shape).contiguous())
            assert torch.equal(traj["obs"][j + 1], plain.obs)
            assert torch.equal(traj["reward"][j], plain.reward) and torch.equal(traj["done"][j], plain.done_u8)
        assert torch.equal(fused.state, plain.state)
    assert worst < 5e-3, worst
    assert fused.stats()["episodes"] == plain.stats()["episodes"] > 0
    # exploration noise: uniform of the given half-width around the deterministic action, clipped
    det = torch.zeros((k, n, ad), device="cuda")
    noisy = torch.zeros((k, n, ad), device="cuda")
    state = fused.state.clone()
    fused.rollout_actor(mlp_layers(actor), k, 0.0, traj={"actions": det}, precision=precision)
    fused.state.copy_(state)
    fused.rollout_actor(mlp_layers(actor), k, 0.25, traj={"actions": noisy}, precision=precision)
    d = (noisy[0] - det[0])  # first cycle: the same state, so the same deterministic action
    inside = det[0].abs() < 0.7
    assert float(d[inside].abs().max()) <= 0.25 + 1e-6 and float(d[inside].std()) == pytest.approx(0.25 / 3 ** 0.5, rel=0.05)
    assert float(noisy.abs().max()) <= 1.0
    fused.close()
    plain.close()


def test_ddpg_learns_from_fused_collection():
    from soccer2d_b200.rollout import DDPGConfig, DeviceDDPG
    kw = dict(change_ball_position=False, use_continuous_action=True, use_turning=False)
    env = Soccer2DVecEnv(4096, device="cuda:0", seed=0, terminal_obs=True, **kw)
    agent = DeviceDDPG(env, DDPGConfig(seed=0, learning_starts=1 << 15))
    agent.learn_fused(1200, k=8)
    trained = agent.evaluate(320, fused=True)
    assert trained["episodes"] > 4096 and trained["goal_rate"] > 0.6, trained
    agent.rollout_step(0.0, store=False)
    env.close()


@pytest.mark.parametrize("n", [33, 1000])
def test_fused_rollouts_stay_inside_their_buffers(n):
    """(compute-sanitizer is not available on this pool) every output of the fused kernels is the middle of a larger
    tensor filled with a sentinel; ragged sizes; the guard bytes on both sides must survive."""
    from soccer2d_b200.rollout import Actor, mlp_layers
    guard, k = 4096, 3
    big = []

    def guarded(shape, dtype):
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        buf = torch.full((nbytes + 2 * guard,), 0xA5, dtype=torch.uint8, device="cuda")
        big.append((buf, nbytes))
        return buf[guard:guard + nbytes].view(dtype).view(shape)

    env = Soccer2DVecEnv(n, device="cuda:0", seed=1, substeps=k, use_continuous_action=False, action_space_size=16, max_steps=5)
    env.reset_torch()
    q = QNetwork(10, 16).cuda()
    traj = {"obs": guarded((k + 1, n, 10), torch.float32), "actions": guarded((k, n), torch.uint8),
            "reward": guarded((k, n), torch.float32), "done": guarded((k, n), torch.uint8)}
    acts, qv = guarded((n, k), torch.uint8), guarded((n, 16), torch.float32)
    for _ in range(4):
        env.rollout_mlp(mlp_layers(q), k, 0.3, traj=traj)
        env.rollout_mlp(mlp_layers(q), k, 0.3, actions_out=acts, q_out=qv)
    cont = Soccer2DVecEnv(n, device="cuda:0", seed=1, substeps=k, use_continuous_action=True, use_turning=True, max_steps=5)
    cont.reset_torch()
    actor = Actor(10, 4).cuda()
    ctraj = {"obs": guarded((k + 1, n, 10), torch.float32), "actions": guarded((k, n, 4), torch.float32),
             "reward": guarded((k, n), torch.float32), "done": guarded((k, n), torch.uint8)}
    for _ in range(4):
        cont.rollout_actor(mlp_layers(actor), k, 0.1, traj=ctraj)
    torch.cuda.synchronize()
    for buf, nbytes in big:
        assert bool((buf[:guard] == 0xA5).all()) and bool((buf[guard + nbytes:] == 0xA5).all())
    assert float(traj["obs"].abs().sum()) > 0 and float(ctraj["actions"].abs().sum()) > 0 and int(traj["done"].sum()) > 0
    env.close()
    cont.close()


def test_fused_policy_rollout_bf16_option():
    """precision="bf16": the env half is still bit-exact; Q-values agree with torch fp32 to bf16 accuracy and the greedy
    action mostly; it is refused for actors and noisy handles."""
    from soccer2d_b200.rollout import Actor, mlp_layers
    torch.manual_seed(0)
    n, k = 4096 + 37, 5
    kw = dict(device="cuda:0", seed=4, use_continuous_action=False, action_space_size=16, change_ball_velocity=True)
    fused, plain = Soccer2DVecEnv(n, substeps=k, **kw), Soccer2DVecEnv(n, substeps=1, **kw)
    qnet = QNetwork(10, 16).cuda()
    with torch.no_grad():
        for p in qnet.parameters():
            p.mul_(3.0)
    fused.reset_torch()
    plain.reset_torch()
    actions = torch.zeros((n, k), dtype=torch.uint8, device="cuda")
    q_seen = torch.zeros((n, 16), device="cuda")
    agree = total = 0
    worst = 0.0
    for launch in range(20):
        fused.rollout_mlp(mlp_layers(qnet), k, 0.0, actions, q_seen, precision="bf16")
        for j in range(k):
            with torch.no_grad():
                q32 = qnet(plain.obs)
            agree += int((q32.argmax(dim=1) == actions[:, j].long()).sum())
            total += n
            if j == k - 1:
                scale = q32.abs().max(dim=1).values + 1.0
                worst = max(worst, float(((q_seen - q32).abs().max(dim=1).values / scale).max()))
            plain.step_torch(actions[:, j:j + 1].contiguous())
        assert torch.equal(fused.state, plain.state) and torch.equal(fused.obs, plain.obs)
    assert worst < 4e-2 and agree / total > 0.9, (worst, agree / total)
    noisy = Soccer2DVecEnv(64, device="cuda:0", noise=True, use_continuous_action=False)
    with pytest.raises(_abi.Soccer2DError):
        noisy.rollout_mlp(mlp_layers(qnet), 1, precision="bf16")
    for e in (fused, plain, noisy):
        e.close()


def test_info_collector_reads_results_from_the_device():
    from utils.info_collector_callback import InfoCollectorCallback
    env = Soccer2DVecEnv(2048, device="cuda:0", seed=3, max_steps=30, **KW)
    env.reset_torch()
    cb = InfoCollectorCallback()
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(80):
        env.step_torch(torch.randint(0, 16, (2048, 1), dtype=torch.uint8, device="cuda", generator=g))
        cb.collect(env)
    st = env.stats()
    got = [i["result"] for i in cb.infos]
    assert (got.count("Goal"), got.count("Out"), got.count("Timeout")) == (st["goals"], st["outs"], st["timeouts"])
    assert len(got) == st["episodes"] > 2048
    import logging
    goal, out, timeout = cb.update_results_dict(logging.getLogger("x")).values()
    assert len(goal) == (len(got) + 99) // 100 and abs(goal[0] + out[0] + timeout[0] - 100.0) < 1e-9
    env.close()
