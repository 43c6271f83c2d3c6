#!/usr/bin/env python
"""Summarise an ncu report (read here, where ncu is installed but no GPU is): one CSV row per profiled launch with the
metrics DESIGN.md quotes, and (with --traffic) the per-kernel DRAM bytes per launch as the JSON that bench.py puts
next to the algorithmic bytes (profiles/r2_traffic.json).

  python tools/ncu_summary.py gpurun_out/r2_hot_full.ncu-rep [more.ncu-rep ...] --csv profiles/r2_ncu_full_hot_kernels.csv \
         --traffic profiles/r2_traffic.json --note "..."
"""
import argparse
import csv
import io
import json
import subprocess

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
# kernel-name fragment -> the name bench.py uses
NAMES = {"fwd_fused_kernel": "embed_fwd", "fwd_miss_kernel": "embed_miss", "bwd_plan": "bwd_plan",
         "bwd_sgd_apply_kernel": "bwd_sgd", "interact_fwd": "interact_fwd", "interact_bwd": "interact_bwd",
         "gemm3x_tf32_kernel": "mlp_gemm"}
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    return hdr, units, rd[2:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--csv", required=True)
    ap.add_argument("--traffic")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    lines, traffic = [], {}
    for rep in a.reports:
        hdr, units, rows = rows_of(rep)
        col = {h: i for i, h in enumerate(hdr)}
        keep = [m for m in METRICS if m in col]
        if not lines:
            lines.append(["# " + a.note])
            lines.append(["report", "ID", "Kernel Name", "Block Size", "Grid Size"] + keep)
            lines.append(["", "", "", "", ""] + [units[col[m]] for m in keep])
        for r in rows:
            name = r[col["Kernel Name"]]
            lines.append([rep.split("/")[-1], r[col["ID"]], name[:90], r[col["Block Size"]], r[col["Grid Size"]]] +
                         [r[col[m]] for m in keep])
            for frag, nm in NAMES.items():
                if frag in name and "dram__bytes_read.sum" in col:
                    b = sum(float(r[col[m]].replace(",", "")) * TO_BYTES.get(units[col[m]], 1.0)
                            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                    t = traffic.setdefault(nm, {"bytes": 0.0, "launches": 0})
                    t["bytes"] += b
                    t["launches"] += 1
                    break
    with open(a.csv, "w", newline="") as f:
        csv.writer(f).writerows(lines)
    if a.traffic:
        json.dump({"source": f"{a.csv} ({a.note})",
                   "kernels": {k: {"dram_bytes_per_launch": v["bytes"] / v["launches"], "launches": v["launches"]}
                               for k, v in traffic.items()}}, open(a.traffic, "w"), indent=1)
    print(f"{len(lines) - 3} launches -> {a.csv}" + (f", traffic of {sorted(traffic)} -> {a.traffic}" if a.traffic else ""))


if __name__ == "__main__":
    main()
