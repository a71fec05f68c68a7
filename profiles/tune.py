"""Timing driver for kernel-tuning experiments (run on the GPU box):
   S2D_LIB=build/exp/lib_x.so python profiles/tune.py   -> ms per launch for K=16 @ 2^20 envs and K=1 @ 2^23 envs"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-soccer-2d-env_b200"))
import torch  # noqa: E402

from soccer2d_b200 import Soccer2DVecEnv  # noqa: E402

CONT = os.environ.get("S2D_TUNE_CONTINUOUS", "0") == "1"  # Box(1) actions (the dash direction) instead of Discrete(16)
KW = dict(use_continuous_action=CONT, action_space_size=16, change_ball_position=True,
          change_ball_velocity=os.environ.get("S2D_TUNE_STILL", "0") != "1")  # S2D_TUNE_STILL=1: the ball rests (reference default)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = []
for name, n, k, warm, steps in (("k16", 1 << 20, 16, int(os.environ.get("S2D_TUNE_WARM", 14)), 30), ("k1", 1 << 23, 1, 10, 30)):  # S2D_TUNE_WARM=150: steady state of the episode ends
    env = Soccer2DVecEnv(n, device="cuda:0", seed=0, substeps=k, **KW)
    g = torch.Generator(device="cuda").manual_seed(0)
    if CONT:
        pool = [torch.rand((n, k), dtype=torch.float32, device="cuda", generator=g) * 2 - 1 for _ in range(2)]
    else:
        pool = [torch.randint(0, 16, (n, k), dtype=torch.uint8, device="cuda", generator=g) for _ in range(2)]
    env.reset_torch()
    for i in range(warm):
        env.bind_actions(pool[i % 2])
        env.step_torch()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i, (a, b) in enumerate(ev):
        env.bind_actions(pool[i % 2])
        flush.zero_()
        a.record()
        env.step_torch()
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    med = ms[len(ms) // 2]
    bytes_per_env = 160 + k + 46
    out.append(f"{name}: median {med:.4f} ms  min {ms[0]:.4f}  {n * k / med / 1e6:.1f} G env-steps/s  {n * bytes_per_env / med / 1e6:.0f} GB/s")
    env.close()
print(os.environ.get("S2D_LIB", "default"), " | ".join(out))
