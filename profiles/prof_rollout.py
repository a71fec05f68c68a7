"""Profiling driver (under ncu on the GPU box): a few launches of the fused policy + step kernel, K = 16, 2^20 envs.
   python profiles/prof_rollout.py [tf32|tf32_mma_sync|bf16|actor|actor_mma_sync]   (actor: the DDPG actor, Box(4))"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-soccer-2d-env_b200"))
import torch  # noqa: E402

from soccer2d_b200 import Soccer2DVecEnv  # noqa: E402
from soccer2d_b200.rollout import QNetwork  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
torch.manual_seed(0)
if mode.startswith("actor"):
    from soccer2d_b200.rollout import Actor, mlp_layers  # noqa: E402
    env = Soccer2DVecEnv(1 << 20, device="cuda:0", seed=0, substeps=16, use_continuous_action=True, use_turning=True,
                         change_ball_position=True, change_ball_velocity=True)
    actor = Actor(10, 4).cuda()
    env.reset_torch()
    for _ in range(4):
        env.rollout_actor(mlp_layers(actor), 16, precision="tf32_mma_sync" if mode.endswith("mma_sync") else "tf32")
    torch.cuda.synchronize()
    sys.exit(0)
qnet = QNetwork(10, 16).cuda()
layers = [(m.weight.detach().contiguous(), m.bias.detach().contiguous()) for m in qnet.net if isinstance(m, torch.nn.Linear)]
env = Soccer2DVecEnv(1 << 20, device="cuda:0", seed=0, substeps=16, use_continuous_action=False, action_space_size=16,
                     change_ball_position=True, change_ball_velocity=True)
env.reset_torch()
for _ in range(4):
    env.rollout_mlp(layers, 16, precision=mode)
torch.cuda.synchronize()
