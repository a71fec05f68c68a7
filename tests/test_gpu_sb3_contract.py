"""Soccer2DVecEnv driven through Stable-Baselines3's VecEnv contract (tests/sb3_double.py restates the calls SB3 2.4
makes: VecMonitor wrapping, OffPolicyAlgorithm.collect_rollouts, ReplayBuffer.add, and the reference's
InfoCollectorCallback._on_step, utils/info_collector_callback.py:17-28) in the loop shape of
dqn_stable_baselines3.py:36-55, at N in {1, 4 096, 65 536}; everything the env returns is checked against the oracle."""
import time

import numpy as np
import pytest
import torch

import oracle_lib as OL
import sb3_double as SB3
from soccer2d_b200 import Soccer2DVecEnv
from soccer2d_b200.vec_env import LazyInfos

pytestmark = pytest.mark.gpu

KW = dict(device="cuda:0", use_continuous_action=False, action_space_size=16, change_ball_position=True,
          change_ball_velocity=True, min_distance_to_ball=5.0, max_steps=25, terminal_obs=True)


@pytest.mark.parametrize("n,steps", [(1, 300), (4096, 80), (65536, 40)])
def test_collect_rollouts_shape_of_the_dqn_script(n, steps):
    env = Soccer2DVecEnv(n, seed=7, **KW)
    sim = OL.OracleSim(env.cfg, "f32")
    rng = np.random.default_rng(n)
    trace = []

    def policy(obs):  # stands in for DQN's epsilon-greedy sample: an int64 action per env, shape [N]
        assert obs.shape == (n, 10) and obs.dtype == np.float32
        a = rng.integers(0, 16, size=n)
        trace.append(a)
        return a

    algo = SB3.OffPolicyDouble(env, policy, buffer_size=n * steps)
    cb = SB3.CallbackDouble()
    first = None
    ref_obs = sim.reset().copy()
    episodes = 0
    for t in range(steps):
        assert algo.collect_rollouts(cb, 1)
        if first is None:
            first = algo.replay_buffer.observations[0].copy()
            assert np.array_equal(first, ref_obs)  # reset() -> [N, 10] float32, the oracle's reset
        sim.step(trace[-1].astype(np.uint8).reshape(n, 1))
        pos = t % algo.replay_buffer.size
        done = sim.done.astype(bool)
        # what SB3 stores: next_obs = the terminal observation where the episode ended, else the new observation
        want_next = np.where(done[:, None], sim.term_obs, sim.obs)
        assert np.array_equal(algo.replay_buffer.next_observations[pos], want_next)
        assert np.array_equal(algo.replay_buffer.rewards[pos], sim.reward)
        assert np.array_equal(algo.replay_buffer.dones[pos].astype(bool), done)
        assert np.array_equal(algo._last_obs, sim.obs)  # after auto-reset: first observation of the next episode
        episodes += int(done.sum())
    assert cb.n_calls == steps and algo.num_timesteps == n * steps
    # the reference callback kept exactly the finished episodes, each with its result (and, via VecMonitor, its record)
    assert len(cb.infos) == episodes == algo.episode_num == len(algo.ep_info_buffer) == env.stats()["episodes"]
    assert episodes > 0
    st = env.stats()
    tally = {k: sum(1 for i in cb.infos if i["result"] == k) for k in ("Goal", "Out", "Timeout")}
    assert tally == {"Goal": st["goals"], "Out": st["outs"], "Timeout": st["timeouts"]}
    for info in cb.infos[:200]:
        assert info["terminal_observation"].shape == (10,) and 1 <= info["episode"]["l"] <= KW["max_steps"] + 1
    assert abs(sum(float(e["r"]) for e in algo.ep_info_buffer) - st["return_sum"]) < 1e-3 * max(1.0, abs(st["return_sum"]))
    assert sum(int(e["l"]) for e in algo.ep_info_buffer) == st["episode_steps"]
    env.close()


def test_infos_are_lazy_and_shared():
    n = 4096
    env = Soccer2DVecEnv(n, seed=3, **KW)
    env.reset()
    rng = np.random.default_rng(0)
    seen = 0
    for _ in range(60):
        env.step_async(rng.integers(0, 16, size=n))
        obs, rew, done, infos = env.step_wait()
        assert isinstance(infos, LazyInfos) and len(infos) == n and infos._cache is None  # nothing built yet
        idx = np.flatnonzero(done)
        if idx.size:
            i = int(idx[0])
            assert infos[i]["result"] in ("Goal", "Out", "Timeout") and infos[i]["terminal_observation"].shape == (10,)
            assert len(infos._cache) == idx.size
            lst = infos[:]
            assert isinstance(lst, list) and len(lst) == n and lst[i] is infos[i]
            running = int(np.flatnonzero(~done)[0])
            assert infos[running] is infos[-1] or done[-1]
            with pytest.raises(TypeError):
                infos[running]["episode"] = {}        # shared and read-only ...
            mine = infos[running].copy()              # ... VecMonitor copies before it writes
            mine["episode"] = {}
            assert infos[running] == {"result": None}
            seen += idx.size
    assert seen > 0
    # the arrays of a step stay valid for host_ring - 1 further steps, and stale infos refuse to answer
    env.step_async(rng.integers(0, 16, size=n))
    obs0, rew0, done0, infos0 = env.step_wait()
    keep = obs0.copy()
    for _ in range(env.host_ring - 1):
        env.step_async(rng.integers(0, 16, size=n))
        env.step_wait()
    assert np.array_equal(obs0, keep)
    env.step_async(rng.integers(0, 16, size=n))
    env.step_wait()
    with pytest.raises(RuntimeError):
        infos0[0]
    env.step_async(rng.integers(0, 16, size=n))
    o2, _, _, _ = env.step_wait(copy=True)  # private copies on request
    assert o2.flags.owndata
    env.close()


def test_wrapper_surface():
    env = Soccer2DVecEnv(8, seed=1, **KW)
    assert env.get_attr("render_mode") == [None] * 8 and env.get_attr("num_envs", indices=[1, 3]) == [8, 8]
    assert env.env_is_wrapped(SB3.Monitor) == [False] * 8
    env.set_attr("render_mode", "human")
    assert env.get_attr("render_mode", indices=0) == ["human"]
    with pytest.raises(AttributeError):
        env.set_attr("max_steps", 10)           # a kernel constant
    assert env.env_method("stats", indices=[0, 1])[0]["episodes"] == 0
    with pytest.raises(AttributeError):
        env.env_method("no_such_method")
    assert env.seed(5) == [None] * 8
    mon = SB3.VecMonitorDouble(env)
    assert mon.reset().shape == (8, 10)
    env.close()


def test_step_wait_costs_less_than_twice_step_host():
    """at 65 536 envs the SB3 face (step_async / step_wait, infos included) must stay within 2x of the
    bare host-buffer call."""
    n = 65536
    env = Soccer2DVecEnv(n, seed=5, **dict(KW, max_steps=200))
    env.reset()
    rng = np.random.default_rng(1)
    acts = [torch.from_numpy(rng.integers(0, 16, size=(n, 1)).astype(np.uint8)).pin_memory() for _ in range(4)]

    def bench(fn, reps=60):
        for i in range(10):
            fn(acts[i % 4])
        ts = []
        for i in range(reps):
            t0 = time.perf_counter()
            fn(acts[i % 4])
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts))

    def wait(a):
        env.step_async(a)
        return env.step_wait()

    t_host = bench(env.step_host)
    t_wait = bench(wait)
    print(f"step_host {t_host * 1e6:.0f} us, step_async+step_wait {t_wait * 1e6:.0f} us ({t_wait / t_host:.2f}x) at {n} envs")
    assert t_wait < 2.0 * t_host
    env.close()
