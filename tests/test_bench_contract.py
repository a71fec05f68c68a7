"""CPU-only: the parts of bench.py's contract that need no GPU - the reference arm prints one JSON line with the
required keys, ranks other than 0 stay silent, the B200 arm refuses to run without a GPU, and the committed ncu
summaries are found under the keys bench.py looks up."""
import json
import os
import subprocess
import sys

import helpers as H

BENCH = os.path.join(H.ROOT, "bench.py")
REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(args, env=None):
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env={**os.environ, **(env or {})},
                          timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d) and d["impl"] == "reference" and d["metric"] == "env_steps_per_sec"
    assert d["value"] > 1e5 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # under torchrun only rank 0 runs and prints the reference arm
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_is_the_same_workload_and_never_touches_the_product():
    """the CPU arm must step the B200 arm's workload (2^20 envs x 16 cycles per step, same `config` dict) and
    its process must neither import the product package nor map libsoccer2d.so."""
    code = (
        "import json, runpy, sys, io, contextlib\n"
        f"sys.argv = [{BENCH!r}, '--impl', 'reference', '--steps', '1', '--warmup', '0']\n"
        "buf = io.StringIO()\n"
        "with contextlib.redirect_stdout(buf):\n"
        f"    runpy.run_path({BENCH!r}, run_name='__main__')\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert 'libsoccer2d' not in maps, 'the reference arm mapped the product library'\n"
        "assert not any(m.startswith('soccer2d_b200') for m in sys.modules), 'the reference arm imported the product'\n"
        "assert 'liboracle_f64' in maps\n"
        "print(buf.getvalue())\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    sys.path.insert(0, H.ROOT)
    import bench
    assert d["config"] == bench.workload_config(1, 1 << 20, 16)  # what the B200 arm prints, key for key
    assert d["config"]["envs_per_gpu"] == 1 << 20 and d["config"]["substeps"] == 16
    assert f"{1 << 20} envs x 16 cycles per step" in d["cpu_baseline"]["sample"]
    assert abs(d["value"] - (1 << 24) / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]  # one step = 2^24 env-steps


def test_b200_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return  # (on the GPU box the arm runs; covered by the bench itself)
    r = _run(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_committed_ncu_summaries_match_the_keys_bench_reads():
    sys.path.insert(0, H.ROOT)
    import bench
    assert bench.load_traffic("k16", 1 << 20) > 1e8 and bench.load_traffic("k1", 1 << 23) > 1e9
    assert 200 < bench.load_traffic("k16", 1 << 20, "warp_instructions_per_env_step") < 400
    assert 50 < bench.load_traffic("k16", 1 << 20, "issue_active_pct") <= 100
    assert bench.load_traffic("fullgame_k1", 1 << 18) > 1e8 and bench.load_traffic("k16", 12345) is None
    peak, src = bench.load_peaks()
    assert 5000 < peak < 9000 and ("measured" in src or "fallback" in src)
