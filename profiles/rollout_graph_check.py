import sys, torch
sys.path.insert(0,'gym-soccer-2d-env_b200')
from soccer2d_b200 import Soccer2DVecEnv
from soccer2d_b200.rollout import QNetwork, measure_rollout
KW=dict(use_continuous_action=False, action_space_size=16, change_ball_velocity=True)
for n in (4096, 65536, 1<<20):
    for g in (False, True):
        env=Soccer2DVecEnv(n, device="cuda:0", seed=0, **KW); env.reset_torch()
        torch.manual_seed(0); q=QNetwork(10,16).to("cuda:0")
        print(n, g, f"{measure_rollout(env,q,50,use_graph=g):.3e}")
        env.close()
