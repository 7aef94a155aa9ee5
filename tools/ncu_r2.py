#!/usr/bin/env python
"""Headline counters of one `ncu --set full` capture, from the .ncu-rep or its exported raw page.

usage: tools/ncu_r2.py <capture.ncu-rep | raw.csv> [label]   -> prints one JSON object
Counters: duration, warp instructions, issue utilisation, active threads per instruction, branch
uniformity, instruction-cache behaviour (per-SM sm__icc hit rate, GPC-level gcc instruction request
rate), L1 / L2 hit rates, DRAM bytes, occupancy, registers, the top stall reasons.
"""
import csv
import io
import json
import subprocess
import sys


def load(path):
    if path.endswith(".ncu-rep"):
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        text = open(path).read()
    rows = list(csv.reader(io.StringIO(text)))
    return dict(zip(rows[0], rows[-1])), dict(zip(rows[0], rows[1]))


def num(d, key):
    v = d.get(key, "")
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def summary(path, label=None):
    d, u = load(path)
    stalls = {n.split("issue_stalled_")[1]: float(v.replace(",", "")) for n, v in d.items()
              if n.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in n
              and v.replace(",", "").replace(".", "").isdigit()}
    total = sum(stalls.values()) or 1.0
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    dram = sum((num(d, k) or 0.0) * scale.get(u.get(k, "byte"), 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    t = num(d, "gpu__time_duration.sum")
    t_ms = t * {"ms": 1.0, "us": 1e-3, "usecond": 1e-3, "ns": 1e-6, "nsecond": 1e-6, "msecond": 1.0, "s": 1e3}.get(u.get("gpu__time_duration.sum"), 1.0)
    return {
        "label": label or path, "kernel": d.get("Kernel Name"), "duration_ms": t_ms,
        "warp_instructions": num(d, "smsp__inst_executed.sum"),
        "issue_active_pct": num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "active_threads_per_instruction": num(d, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "branch_targets_uniform_pct": num(d, "smsp__sass_average_branch_targets_threads_uniform.pct"),
        "icc_hit_rate_pct": num(d, "sm__icc_request_hit_rate.pct"),
        "gcc_instruction_requests_pct_of_peak": num(d, "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed"),
        "l1_hit_pct": num(d, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": num(d, "lts__t_sector_hit_rate.pct"),
        "dram_bytes": dram, "lsu_wavefronts_pct": num(d, "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
        "warps_active_pct": num(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers": num(d, "launch__registers_per_thread"), "grid": num(d, "launch__grid_size"),
        "stalls_pct": {a: round(100 * b / total, 1) for a, b in sorted(stalls.items(), key=lambda x: -x[1])[:6]},
    }


if __name__ == "__main__":
    print(json.dumps(summary(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None), indent=1))
