#!/usr/bin/env python
"""Turns an ncu report (captured on the GPU box, read here without a GPU) into the text summary committed under profiles/
and into profiles/ncu_traffic.json, from which bench.py takes `roofline.traffic`.

  python profiles/summarise.py gpurun_out/X.ncu-rep profiles/ncu_rNN_all_kernels_summary.txt --batch 8192 --width 752 --height 480
"""
import argparse
import csv
import io
import json
import os
import subprocess

METRICS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_not_selected",
]


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def to_us(value, unit):
    v = float(value.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--batch", type=int, required=True)
    ap.add_argument("--width", type=int, default=752)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--header", default="")
    ap.add_argument("--traffic-json", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json"))
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(names)}
    lines = [l for l in a.header.split("\\n") if l]
    traffic = {}
    for r in data:
        kname = r[col["Kernel Name"]]
        lines.append(f"Kernel Name: {kname}")
        for m in METRICS:
            if m in col:
                lines.append(f"{m}: {r[col[m]]} {units[col[m]]}")
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        us = to_us(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
        lines.append(f"derived: dram read+write = {(rd + wr) / 1e6:.3f} MB in {us:.1f} us = {(rd + wr) / us / 1e3:.1f} GB/s (under ncu: cold cache, serialised)")
        if "scan_kernel" in kname:
            alg = a.batch * a.width * a.height
            lines.append(f"derived: algorithmic bytes = {a.batch} x {a.width} x {a.height} = {alg} B; dram traffic / algorithmic = {(rd + wr) / alg:.4f}")
            traffic = {"kernel": "scan_kernel", "batch": a.batch, "width": a.width, "height": a.height, "dram_bytes_read": rd, "dram_bytes_write": wr,
                       "dram_bytes": rd + wr, "algorithmic_bytes": alg, "source": os.path.basename(a.out)}
        lines.append("")
    with open(a.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    if traffic:
        with open(a.traffic_json, "w") as f:
            json.dump(traffic, f, indent=1)
    print(f"wrote {a.out}" + (f" and {a.traffic_json}" if traffic else ""))


if __name__ == "__main__":
    main()
