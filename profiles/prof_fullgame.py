"""Profiling driver (under ncu on the GPU box): a few launches of the 11v11 step kernel, K=1 and K=16."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-soccer-2d-env_b200"))
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import torch  # noqa: E402

from soccer2d_b200 import Soccer2DVecEnv  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)


def commands(shape):
    a = torch.zeros(shape + (4,), device="cuda")
    cmd = torch.randint(0, 5, shape, device="cuda", generator=g)
    a[..., 0] = cmd.float()
    a[..., 1] = torch.where(cmd == 4, torch.rand(shape, device="cuda", generator=g) * 100 - 50, torch.rand(shape, device="cuda", generator=g) * 100)
    a[..., 2] = torch.where(cmd == 4, torch.rand(shape, device="cuda", generator=g) * 60 - 30, torch.rand(shape, device="cuda", generator=g) * 360 - 180)
    a[..., 3] = 100.0
    return a


launches = int(sys.argv[1]) if len(sys.argv) > 1 else 5  # capture with ncu -s <launch index> -c 1
for k in ((1, 16) if launches == 5 else (1,)):
    n = int(os.environ.get("FG_MATCHES", 1 << 18))
    env = Soccer2DVecEnv(n, scenario="fullgame", device="cuda:0", seed=0, substeps=k)
    pool = [commands((n, k, 22)) for _ in range(2)]
    env.reset_torch()
    for i in range(launches):
        env.bind_actions(pool[i % 2])
        env.step_torch()
    torch.cuda.synchronize()
    env.close()
