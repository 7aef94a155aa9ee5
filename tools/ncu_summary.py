#!/usr/bin/env python
"""Exports the raw page and per-line table of the four wave-kernel captures of a tag
(gpurun_out/prof_<tag>_<kernel>.ncu-rep, written by tools/ncu_capture.sh) into profiles/ and
prints / stores their headline counters (profiles/<tag>_kernels_summary.json,
profiles/kernel_traffic.json).   usage: tools/ncu_summary.py <tag>"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
KERNELS = ("walk", "shade", "finish", "prepass")


def num(d, key):
    for n in d:
        if n == key or n.startswith(key):
            try:
                return float(d[n].replace(",", ""))
            except ValueError:
                return None
    return None


out = {}
for k in KERNELS:
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}_{k}.ncu-rep")
    raw = os.path.join(ROOT, "profiles", f"{tag}_{k}_ncu_raw.csv")
    with open(raw, "w") as fh:
        subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=fh, stderr=subprocess.DEVNULL, check=True)
    with open(os.path.join(ROOT, "profiles", f"{tag}_{k}_lines.txt"), "w") as fh:
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, f"wave_{k}_kernel", "60"],
                       stdout=fh, stderr=subprocess.STDOUT)
    rows = list(csv.reader(open(raw)))
    d, u = dict(zip(rows[0], rows[-1])), dict(zip(rows[0], rows[1]))
    assert u.get("gpu__time_duration.sum") in ("us", "usecond"), u.get("gpu__time_duration.sum")
    assert u.get("dram__bytes_read.sum") == "Mbyte" and u.get("dram__bytes_write.sum") == "Mbyte"
    stalls = {n.split("issue_stalled_")[1].replace(".sum", ""): float(v.replace(",", ""))
              for n, v in d.items()
              if n.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in n
              and v.replace(",", "").replace(".", "").isdigit()}
    total = sum(stalls.values()) or 1.0
    out[k] = {
        "duration_us": num(d, "gpu__time_duration.sum"),
        "warp_inst": num(d, "smsp__inst_executed.sum"),
        "issue_active_pct": num(d, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        "lsu_wavefronts_pct": num(d, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "dram_pct": num(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "occupancy_pct": num(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "regs": num(d, "launch__registers_per_thread"), "grid": num(d, "launch__grid_size"),
        "l1_hit": num(d, "l1tex__t_sector_hit_rate.pct"), "l2_hit": num(d, "lts__t_sector_hit_rate.pct"),
        "dram_read_mb": num(d, "dram__bytes_read.sum"), "dram_write_mb": num(d, "dram__bytes_write.sum"),
        "stalls_pct": {a: round(100 * b / total, 1) for a, b in sorted(stalls.items(), key=lambda x: -x[1])[:6]},
    }
    print(k, out[k])
json.dump(out, open(os.path.join(ROOT, "profiles", f"{tag}_kernels_summary.json"), "w"), indent=1)
traffic = {
    "source": f"profiles/{tag}_*_ncu_raw.csv (tools/ncu_capture.sh: ncu --set full --clock-control none; one "
              "steady-state wave of one slot group, 4 Mi slots)",
    "note": "DRAM bytes of one launch; the scene (3 KB for house.toml) stays in L1, so this is the wavefront's "
            "path state, ray / hit records, queues and environment lookups",
    "kernels": {f"wave_{k}_kernel": {
        "dram_bytes_per_launch": int((v["dram_read_mb"] + v["dram_write_mb"]) * 1e6),
        "dram_bytes_read": int(v["dram_read_mb"] * 1e6), "dram_bytes_write": int(v["dram_write_mb"] * 1e6),
        "launch_duration_us_under_ncu": v["duration_us"]} for k, v in out.items()}}
json.dump(traffic, open(os.path.join(ROOT, "profiles", "kernel_traffic.json"), "w"), indent=2)
