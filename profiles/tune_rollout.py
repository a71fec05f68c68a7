"""Timing driver (GPU box): closed-loop rollout with the Q-network inside the kernel (s2d_rollout_mlp) against the
torch-policy loop (policy launch + step launch per cycle, CUDA graph).  python profiles/tune_rollout.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-soccer-2d-env_b200"))
import torch  # noqa: E402

from soccer2d_b200 import Soccer2DVecEnv  # noqa: E402
from soccer2d_b200.rollout import QNetwork, measure_rollout  # noqa: E402

KW = dict(use_continuous_action=False, action_space_size=16, change_ball_position=True, change_ball_velocity=True)
torch.manual_seed(0)
qnet = QNetwork(10, 16).cuda()
layers = [(m.weight.detach().contiguous(), m.bias.detach().contiguous()) for m in qnet.net if isinstance(m, torch.nn.Linear)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for n, k, prec in ((1 << 20, 16, "tf32"), (1 << 20, 16, "tf32_mma_sync"), (1 << 20, 1, "tf32"), (1 << 20, 1, "tf32_mma_sync"),
                   (1 << 16, 16, "tf32"), (1 << 16, 16, "tf32_mma_sync"), (1 << 20, 16, "bf16")):
    env = Soccer2DVecEnv(n, device="cuda:0", seed=0, substeps=k, **KW)
    env.reset_torch()
    for _ in range(14 if k > 1 else 5):
        env.rollout_mlp(layers, k, precision=prec)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in ev:
        flush.zero_()
        a.record()
        env.rollout_mlp(layers, k, precision=prec)
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    med = ms[len(ms) // 2]
    print(f"fused policy+step ({prec}): {n} envs K={k}: median {med:.4f} ms  min {ms[0]:.4f}  {n * k / med / 1e6:.2f} G env-steps/s  stats {env.stats()['episodes']}")
    env.close()
from soccer2d_b200.rollout import Actor, mlp_layers  # noqa: E402
for turning, aprec in ((False, "tf32"), (True, "tf32"), (False, "tf32_mma_sync"), (True, "tf32_mma_sync")):
    n, k = 1 << 20, 16
    env = Soccer2DVecEnv(n, device="cuda:0", seed=0, substeps=k, use_continuous_action=True, use_turning=turning,
                         change_ball_position=True, change_ball_velocity=True)
    actor = Actor(10, 4 if turning else 1).cuda()
    env.reset_torch()
    for _ in range(14):
        env.rollout_actor(mlp_layers(actor), k, precision=aprec)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in ev:
        flush.zero_()
        a.record()
        env.rollout_actor(mlp_layers(actor), k, precision=aprec)
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    print(f"fused actor+step ({'Box(4) turning' if turning else 'Box(1)'}, {aprec}): {n} envs K={k}: median {ms[len(ms) // 2]:.4f} ms  "
          f"{n * k / ms[len(ms) // 2] / 1e6:.2f} G env-steps/s")
    env.close()
env = Soccer2DVecEnv(1 << 20, device="cuda:0", seed=0, substeps=1, **KW)
print("torch policy + step (CUDA graph), 2^20 envs:", f"{measure_rollout(env, qnet, steps=50, use_graph=True) / 1e9:.3f} G env-steps/s")
