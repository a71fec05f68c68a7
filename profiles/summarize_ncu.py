"""Turns ncu captures (gpurun_out/*.ncu-rep, `ncu --set full --clock-control none --import-source on`) into the two
summaries kept under profiles/: one CSV row per captured launch with the metrics the README quotes, and traffic.json
(DRAM bytes per launch, read by bench.py for `roofline.traffic`).
usage: python profiles/summarize_ncu.py OUT.csv TRAFFIC.json NAME:UNITS=path.ncu-rep [...]
       NAME = the key bench.py looks up (k16_envs_1048576, ...), UNITS = env-steps (or match-cycles) one launch computes"""
import csv
import json
import os
import subprocess
import sys

METRICS = """launch__grid_size launch__block_size launch__registers_per_thread gpu__time_duration.sum dram__bytes_read.sum
dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__thread_inst_executed_per_inst_executed.ratio smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio""".split()


def rows_of(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        yield dict(zip(head, r)), dict(zip(head, units))


def main():
    out, traffic_out, caps = sys.argv[1], sys.argv[2], [a.split("=", 1) for a in sys.argv[3:]]
    table, unit_row = [], None
    summary = {"_source": "ncu --set full --clock-control none --import-source on, one launch per capture (profiles/refresh.sh); "
                          "rows in " + os.path.basename(out)}
    for name_units, path in caps:
        name, units = name_units.split(":")
        units = int(units)
        for r, u in rows_of(path):
            unit_row = unit_row or [""] * 2 + [u.get(m, "") for m in METRICS]
            table.append([name, r["Kernel Name"]] + [r.get(m, "") for m in METRICS])
            to_bytes = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(r["dram__bytes_read.sum"]) * to_bytes[u["dram__bytes_read.sum"]]
            wr = float(r["dram__bytes_write.sum"]) * to_bytes[u["dram__bytes_write.sum"]]
            summary[name] = {
                "kernel": r["Kernel Name"], "grid": int(r["launch__grid_size"]), "dram_read_bytes": int(rd),
                "dram_write_bytes": int(wr), "traffic_bytes": int(rd + wr),
                "warp_instructions": int(float(r["smsp__inst_executed.sum"])),
                "warp_instructions_per_env_step": round(float(r["smsp__inst_executed.sum"]) / (units / 32.0)
                                                        if "fullgame" not in name else float(r["smsp__inst_executed.sum"]) / units),
                "issue_active_pct": round(float(r["smsp__issue_active.avg.pct_of_peak_sustained_active"]), 1),
                "dram_throughput_pct": round(float(r["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]), 1),
                "duration_us_under_ncu": float(r["gpu__time_duration.sum"]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}[u["gpu__time_duration.sum"]],
            }
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["capture", "Kernel Name"] + METRICS)
        w.writerow(unit_row)
        w.writerows(table)
    with open(traffic_out, "w") as f:
        json.dump(summary, f, indent=1)
        f.write("\n")
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
