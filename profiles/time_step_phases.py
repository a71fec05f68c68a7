"""How the K = 16 ReachBall launch time moves with the age of the episodes: per-launch CUDA-event times over 200
launches (means of blocks of 20), for the bench configuration and for episodes that (almost) never end."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-soccer-2d-env_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from soccer2d_b200 import Soccer2DVecEnv  # noqa: E402

n, k = 1 << 20, 16
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(1234)
pool = [torch.randint(0, 16, (n, k), dtype=torch.uint8, device=dev, generator=gen) for _ in range(4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, kw in (("bench configuration", bench.SCENARIO_KW),
                 ("episodes never end by goal or time-out", dict(bench.SCENARIO_KW, min_distance_to_ball=0.0, max_steps=1 << 30))):
    env = Soccer2DVecEnv(n, device=dev, seed=0, substeps=k, **kw)
    env.reset_torch()
    bench.time_launches(env, pool, 3, flush)
    ms = bench.time_launches(env, pool, 200, flush)
    st = env.stats()
    print(name, " ".join(f"{sum(ms[i:i + 20]) / 20 * 1e3:.1f}" for i in range(0, 200, 20)), "us per launch (blocks of 20);",
          "episodes ended per env-step: %.5f" % (st["episodes"] / max(1, st["env_steps"])))
    env.close()
