#!/usr/bin/env python
"""Alternating train / test rounds on ReachBall with the lockstep GPU simulator - the counterpart of the reference's
dqn_ddpg_stable_baselines3.py: its env kwargs (:19-31: fixed ball at the centre, `use_turning=True`, i.e. the Box(4)
action [turn_prob, turn_angle, dash_prob, dash_angle]) with DDPG, or Discrete(16) with DQN when
--discrete is given (:36-48), `test()` after every training round (:56-75) and a results table (:77-86).

    python gym-soccer-2d-env_b200/dqn_ddpg_b200.py --rounds 5 --train-steps 600
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch  # noqa: E402

from soccer2d_b200 import Soccer2DVecEnv  # noqa: E402
from soccer2d_b200.rollout import DDPGConfig, DeviceDDPG, DeviceDQN, DQNConfig  # noqa: E402

# dqn_ddpg_stable_baselines3.py:19-31
KWARGS = dict(change_ball_position=False, change_ball_velocity=False, ball_position_x=0, ball_position_y=0, ball_speed=0,
              ball_direction=0, min_distance_to_ball=5.0, max_steps=200, use_continuous_action=True, action_space_size=16,
              use_turning=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--train-steps", type=int, default=600, help="lockstep cycles per training round")
    ap.add_argument("--test-steps", type=int, default=300, help="lockstep cycles per test")
    ap.add_argument("--discrete", action="store_true", help="Discrete(16) actions and DQN instead of Box(4) and DDPG")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--save-dir", default=None, help="where to write model_<round>_<goal rate>.pt (as the reference saves models)")
    args = ap.parse_args()
    kw = dict(KWARGS)
    if args.discrete:
        kw.update(use_continuous_action=False, use_turning=False)
    env = Soccer2DVecEnv(args.envs, device=args.device, seed=args.seed, terminal_obs=True, **kw)
    warm = min(1 << 16, 16 * args.envs)
    agent = (DeviceDQN(env, DQNConfig(seed=args.seed, learning_starts=warm)) if args.discrete
             else DeviceDDPG(env, DDPGConfig(seed=args.seed, learning_starts=warm)))
    results = [agent.evaluate(args.test_steps)]  # the untrained policy, as the reference tests before training
    print(json.dumps({"round": 0, "test": results[-1]}), flush=True)
    for rnd in range(1, args.rounds + 1):
        log = agent.learn(args.train_steps, report_every=args.train_steps)
        results.append(agent.evaluate(args.test_steps))
        print(json.dumps({"round": rnd, "train": log[-1] if log else None, "test": results[-1]}), flush=True)
        if args.save_dir:
            os.makedirs(args.save_dir, exist_ok=True)
            net = agent.q if args.discrete else agent.actor
            torch.save(net.state_dict(), os.path.join(args.save_dir, f"model_{rnd}_{results[-1]['goal_rate']:.2f}.pt"))
    print("round  goal   out    timeout")
    for i, r in enumerate(results):
        print(f"{i:5d}  {r['goal_rate']:.3f}  {r['out_rate']:.3f}  {r['timeout_rate']:.3f}")
    env.close()


if __name__ == "__main__":
    main()
