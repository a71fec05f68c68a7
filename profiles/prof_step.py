"""Profiling driver (run under ncu on the GPU box): a few launches of the step kernel in both regimes.
usage: python profiles/prof_step.py [k16|k1|both] [launches]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-soccer-2d-env_b200"))
import torch  # noqa: E402

from soccer2d_b200 import Soccer2DVecEnv  # noqa: E402

KW = dict(use_continuous_action=False, action_space_size=16, change_ball_position=True, change_ball_velocity=True)
which = sys.argv[1] if len(sys.argv) > 1 else "both"
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 4
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 13  # 13 x 16 = 208 cycles: resets are in steady state
for name, n, k in (("k16", 1 << 20, 16), ("k1", 1 << 23, 1)):
    if which not in (name, "both"):
        continue
    env = Soccer2DVecEnv(n, device="cuda:0", seed=0, substeps=k, **KW)
    g = torch.Generator(device="cuda").manual_seed(0)
    pool = [torch.randint(0, 16, (n, k), dtype=torch.uint8, device="cuda", generator=g) for _ in range(2)]
    env.reset_torch()
    for i in range((warm if k > 1 else 5) + launches):
        env.bind_actions(pool[i % 2])
        env.step_torch()
    torch.cuda.synchronize()
    print(name, env.stats())
    env.close()
