"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck): ragged sizes, all action
modes, both scenarios, fullgame with fewer than 11 a side, masked reset, pipelined host stepping."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-soccer-2d-env_b200"))
import torch  # noqa: E402

from soccer2d_b200 import Soccer2DVecEnv  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
for n in (1, 33, 257):
    for kw in (dict(use_continuous_action=False), dict(use_continuous_action=True), dict(use_continuous_action=True, use_turning=True),
               dict(use_command_action=True), dict(use_continuous_action=False, noise=True),
               dict(use_continuous_action=False, server_param=dict(player_decay=0.5))):
        for k in (1, 3, 16):
            env = Soccer2DVecEnv(n, device="cuda:0", substeps=k, terminal_obs=True, max_steps=5, change_ball_velocity=True, **kw)
            env.reset_torch()
            for _ in range(4):
                a = env.actions
                if a.dtype == torch.uint8:
                    a.copy_(torch.randint(0, 16, a.shape, dtype=torch.uint8, device="cuda", generator=g))
                else:
                    a.copy_(torch.rand(a.shape, device="cuda", generator=g) * 2 - 1)
                    if kw.get("use_command_action"):
                        a[..., 0] = torch.randint(0, 5, a.shape[:-1], device="cuda", generator=g).float()
                env.step_torch()
            env.reset_torch(env.done)
            env.stats()
            env.close()
    for scen, extra in (("shoot", {}), ("shoot", dict(use_command_action=True)), ("fullgame", dict(players_per_side=11, half_time_cycles=4)),
                        ("fullgame", dict(players_per_side=3, half_time_cycles=4, noise=True))):
        env = Soccer2DVecEnv(n, scenario=scen, device="cuda:0", substeps=2, terminal_obs=True, **extra)
        env.reset_torch()
        for _ in range(5):
            a = env.actions
            if a.dtype == torch.uint8:
                a.copy_(torch.randint(0, 24, a.shape, dtype=torch.uint8, device="cuda", generator=g))
            else:
                a.copy_(torch.rand(a.shape, device="cuda", generator=g) * 100 - 50)
                a[..., 0] = torch.randint(0, 5, a.shape[:-1], device="cuda", generator=g).float()
            env.step_torch()
        env.export_env(0)
        env.close()
env = Soccer2DVecEnv(1000, device="cuda:0", substeps=4, use_continuous_action=False)
env.reset_torch()
acts = [torch.randint(0, 16, (1000, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
t = env.submit_host(acts[0])
for i in range(1, 6):
    t2 = env.submit_host(acts[i % 2])
    env.wait_host(t)
    t = t2
env.wait_host(t)
env.step_host(acts[0])
torch.cuda.synchronize()
env.close()
print("sanitize driver finished")
