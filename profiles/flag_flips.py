#!/usr/bin/env python
"""fp32-vs-f64 flag-flip rate of the ReachBall step path, over ALL seeds (no selection).

The kernels compute in IEEE binary32 (the reference's proto scalars are float32); the north-star truth is the f64
oracle.  A threshold test (`dist < min_distance_to_ball`, `|x| > 52.5`, `step_number > max_steps` is integer) can land
within fp32 rounding of its threshold, and then `done` differs between the two for that env-step - after which the two
runs of that env are different episodes.  This script measures how often: it steps the oracle's fp32 build (bit-identical
to the CUDA kernels, tests/test_gpu_parity.py) and its f64 build side by side on the bench workload (2^20 envs,
Discrete(16), 1 000 cycles, the same action stream), and records for every env the first cycle at which done / result
differ.  rate = diverged envs / env-steps compared before divergence.  CPU only (test infrastructure).

    python profiles/flag_flips.py [--envs N] [--cycles C] [--seed S] -> profiles/parity_flips.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as OL  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--cycles", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--k", type=int, default=8, help="cycles per oracle call (flags are compared per cycle: K = 1 calls)")
    args = ap.parse_args()
    n = args.envs
    cfg = OL.default_config(n, 0, action_mode=OL.ACT_DISCRETE, seed=args.seed, change_ball_velocity=1, change_ball_position=1,
                            max_steps=200, min_distance_to_ball=5.0, action_space_size=16)
    threads = len(os.sched_getaffinity(0))
    truth, spec = OL.OracleSim(cfg, "f64"), OL.OracleSim(cfg, "f32")
    for kind in ("f64", "f32"):
        OL.lib(kind).s2do_set_threads(threads)
    truth.reset()
    spec.reset()
    rng = np.random.default_rng(args.seed)
    alive = np.ones(n, bool)
    compared = 0
    first = np.full(n, -1, np.int32)
    max_obs_err = 0.0
    t0 = time.time()
    for c in range(args.cycles):
        act = rng.integers(0, 16, size=(n, 1)).astype(np.uint8)
        _, _, dt, rt = truth.step(act, 1)
        _, _, ds, rs = spec.step(act, 1)
        compared += int(alive.sum())
        diff = alive & ((dt != ds) | (rt != rs))
        first[diff] = c
        alive &= ~diff
        if c % 50 == 49:
            err = np.abs(truth.obs[alive] - spec.obs[alive].astype(np.float64))
            for col, period in ((0, 2.0), (1, 2.0), (7, 1.0)):
                err[:, col] = np.minimum(err[:, col], np.abs(period - err[:, col]))
            max_obs_err = max(max_obs_err, float(err.max()))
    flips = int((~alive).sum())
    out = {"flag_flip_rate_vs_f64": flips / compared, "flips": flips, "env_steps_compared": compared, "envs": n,
           "cycles": args.cycles, "seeds": "all (Philox seed %d, every env of the batch, no selection)" % args.seed,
           "max_obs_err_of_undiverged_envs": max_obs_err, "tolerance": 1e-5,
           "what": "done/result of the fp32 spec (== the CUDA kernels bit for bit) vs the f64 oracle, first divergence per env; "
                   "inherent to fp32 thresholds, not a defect of the kernels",
           "how": "python profiles/flag_flips.py (oracle f32 build vs f64 build, CPU)", "seconds": round(time.time() - t0, 1)}
    with open(os.path.join(ROOT, "profiles", "parity_flips.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
