"""Timing driver (GPU box): Shoot and FullGame step kernels, ms per launch and env-steps/s.
   python profiles/tune_scenarios.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-soccer-2d-env_b200"))
import torch  # noqa: E402

from soccer2d_b200 import Soccer2DVecEnv  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda").manual_seed(0)


def commands(shape):
    """random {cmd, a, b, c}: dash / turn / kick / go-to-point mix"""
    a = torch.zeros(shape + (4,), device="cuda")
    cmd = torch.randint(0, 5, shape, device="cuda", generator=g)
    a[..., 0] = cmd.float()
    a[..., 1] = torch.where(cmd == 4, torch.rand(shape, device="cuda", generator=g) * 100 - 50, torch.rand(shape, device="cuda", generator=g) * 100)
    a[..., 2] = torch.where(cmd == 4, torch.rand(shape, device="cuda", generator=g) * 60 - 30, torch.rand(shape, device="cuda", generator=g) * 360 - 180)
    a[..., 3] = 100.0
    return a


def time_env(env, pool, warm, steps):
    env.reset_torch()
    for i in range(warm):
        env.bind_actions(pool[i % len(pool)])
        env.step_torch()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i, (a, b) in enumerate(ev):
        env.bind_actions(pool[i % len(pool)])
        flush.zero_()
        a.record()
        env.step_torch()
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return ms[len(ms) // 2]


def main():
    rows = []
    for name, n, k, mode in (("shoot discrete K=16", 1 << 22, 16, "discrete"), ("shoot discrete K=1", 1 << 22, 1, "discrete"),
                             ("shoot command K=1", 1 << 22, 1, "command")):
        env = Soccer2DVecEnv(n, scenario="shoot", device="cuda:0", seed=0, substeps=k, use_command_action=mode == "command")
        if mode == "discrete":
            pool = [torch.randint(0, 24, (n, k), dtype=torch.uint8, device="cuda", generator=g) for _ in range(2)]
            abytes = k
        else:
            pool = [commands((n, k)) for _ in range(2)]
            abytes = 16 * k
        med = time_env(env, pool, 14 if k > 1 else 6, 20)
        per_env = 160 + abytes + 46
        rows.append(f"{name}: {n} envs  {med:.4f} ms  {n * k / med / 1e6:.1f} G env-steps/s  {n * per_env / med / 1e6:.0f} GB/s algorithmic")
        env.close()
        del env, pool
        torch.cuda.empty_cache()

    for name, n, k in (("fullgame 11v11 K=1", 1 << 18, 1), ("fullgame 11v11 K=16", 1 << 18, 16)):
        env = Soccer2DVecEnv(n, scenario="fullgame", device="cuda:0", seed=0, substeps=k, half_time_cycles=3000)
        pool = [commands((n, k, 22)) for _ in range(2)]
        med = time_env(env, pool, 6, 20)
        per_env = 2 * (22 * 36 + 64) + 22 * 16 * k + 480 + 6
        rows.append(f"{name}: {n} envs  {med:.4f} ms  {n * k / med / 1e6:.2f} G env-steps/s ({22 * n * k / med / 1e6:.1f} G agent-steps/s)  "
                    f"{n * per_env / med / 1e6:.0f} GB/s algorithmic  stats {env.stats()['episodes']}")
        env.close()
        del env, pool
        torch.cuda.empty_cache()
    print("\n".join(rows))


if __name__ == "__main__":
    main()
