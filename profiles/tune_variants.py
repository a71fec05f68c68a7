"""Timing driver (GPU box): the less common kernel variants - rcssserver noise on, a non-default ServerParam (constants
from the constant bank instead of immediates), heterogeneous players.  python profiles/tune_variants.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-soccer-2d-env_b200"))
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import torch  # noqa: E402

from soccer2d_b200 import Soccer2DVecEnv  # noqa: E402
from tune_scenarios import commands, time_env  # noqa: E402

KW = dict(use_continuous_action=False, action_space_size=16, change_ball_position=True, change_ball_velocity=True)
g = torch.Generator(device="cuda").manual_seed(0)
for label, extra in (("default", {}), ("noise on", {"noise": True}), ("runtime ServerParam", {"server_param": {"player_decay": 0.41}})):
    for n, k in ((1 << 20, 16), (1 << 23, 1)):
        env = Soccer2DVecEnv(n, device="cuda:0", seed=0, substeps=k, **KW, **extra)
        pool = [torch.randint(0, 16, (n, k), dtype=torch.uint8, device="cuda", generator=g) for _ in range(2)]
        ms = time_env(env, pool, 14 if k > 1 else 5, 20)
        print(f"reachball {label}: K={k} {n} envs  {ms:.4f} ms  {n * k / ms / 1e6:.1f} G env-steps/s")
        env.close()
n = 1 << 18
for label, extra in (("default", {}), ("heterogeneous players", {"hetero_seed": 1}), ("noise on", {"noise": True})):
    for k in (1, 16):
        env = Soccer2DVecEnv(n, scenario="fullgame", device="cuda:0", seed=0, substeps=k, **extra)
        pool = [commands((n, k, 22)) for _ in range(2)]
        ms = time_env(env, pool, 5, 15)
        print(f"fullgame {label}: K={k} {n} matches  {ms:.4f} ms  {n * k / ms / 1e6:.2f} G match-cycles/s")
        env.close()
