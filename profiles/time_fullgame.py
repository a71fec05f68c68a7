"""Per-launch CUDA-event times of the 11v11 step kernel (K = 1, 2^18 matches) over the first cycles of a match, with and
without an L2 flush in front of every launch; `swarm` makes every player chase the ball (collisions every cycle)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-soccer-2d-env_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from soccer2d_b200 import Soccer2DVecEnv  # noqa: E402

n, k = int(os.environ.get("FG_MATCHES", 1 << 18)), int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(99)
env = Soccer2DVecEnv(n, scenario="fullgame", device=dev, seed=0, substeps=k)
pool = [bench.commands(torch, gen, dev, (n, k, 22)) for _ in range(int(os.environ.get("FG_POOL", 2)))]
if os.environ.get("FG_CMD"):  # every player the same kind of command (0 none, 1 dash, 2 turn, 3 kick, 4 go-to-point): no mixed warps
    for p in pool:
        p[..., 0] = float(os.environ["FG_CMD"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, fl in (("flush", flush), ("noflush", None)):
    env.reset_torch()
    ms = bench.time_launches(env, pool, steps, fl)
    print(name, "K", k, " ".join(f"{x * 1e3:.0f}" for x in ms), "us;  mean %.1f us" % (sum(ms) / len(ms) * 1e3))
    st = env.stats()
    pl = env.fullgame_planes()
    ej = pl["ej"]
    print("   matches with a collision in the last cycle: %.3f, with a kick: %.3f; play modes: %s" % (
        float((ej[:, 2] != 0).float().mean()), float((ej[:, 3] != 0).float().mean()),
        torch.bincount((pl["ei"][:, 3] & 0xff).long(), minlength=9).tolist()))
