#!/usr/bin/env python
"""Whole-step figures from one ncu pass over EVERY launch of one render:
    RT_WAVE_GRAPH=0 QB_NO_WARMUP=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,\
smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
--log-file gpurun_out/ncu_<tag>_step.csv python tools/quick_bench.py 256 1
(the host-paced loop runs the same kernels as the graph loop; ncu serialises them, so durations are
cold-cache and only their shares mean anything — instruction and byte counts are exact).
usage: tools/ncu_step.py <csv> <config letter> <tag>  -> profiles/<tag>_step_by_kernel.json and the
config's entry of profiles/measured.json (read by bench.py)."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, config, tag = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 8]
hdr = next(i for i, r in enumerate(rows) if "Metric Name" in r)
col = {n: rows[hdr].index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3,
         "ms": 1.0, "msecond": 1.0, "inst": 1.0, "": 1.0}
per = {}
for r in rows[hdr + 1:]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").strip()
    name = name.split("::")[-1]
    metric, unit = r[col["Metric Name"]], r[col["Metric Unit"]]
    v = float(r[col["Metric Value"]].replace(",", "")) * SCALE.get(unit, 1.0)
    k = per.setdefault(name, {"launches": set()})
    k["launches"].add(r[col["ID"]])
    k[metric] = k.get(metric, 0.0) + v
out = {}
for name, k in per.items():
    out[name] = {"launches": len(k["launches"]), "duration_ms_serialised": k.get("gpu__time_duration.sum", 0.0),
                 "warp_instructions": k.get("smsp__inst_executed.sum", 0.0),
                 "thread_instructions": k.get("smsp__thread_inst_executed.sum", 0.0),
                 "dram_bytes": k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)}
    o = out[name]
    o["active_threads_per_instruction"] = o["thread_instructions"] / o["warp_instructions"] if o["warp_instructions"] else None
total = {m: sum(o[m] for o in out.values()) for m in ("duration_ms_serialised", "warp_instructions", "thread_instructions", "dram_bytes")}
for o in out.values():
    o["share_of_serialised_time"] = o["duration_ms_serialised"] / total["duration_ms_serialised"]
result = {"source": f"profiles/{tag}_step_by_kernel.json (one ncu pass over every launch of one step; tools/ncu_step.py)",
          "totals": total, "by_kernel": out}
json.dump(result, open(os.path.join(ROOT, "profiles", f"{tag}_step_by_kernel.json"), "w"), indent=1)
mpath = os.path.join(ROOT, "profiles", "measured.json")
measured = json.load(open(mpath)) if os.path.exists(mpath) else {}
measured[config] = {"source": result["source"], "dram_bytes_per_step": total["dram_bytes"],
                    "issue": {"warp_instructions_per_step": total["warp_instructions"],
                              "thread_instructions_per_step": total["thread_instructions"]}}
json.dump(measured, open(mpath, "w"), indent=1)
for name, o in sorted(out.items(), key=lambda kv: -kv[1]["duration_ms_serialised"]):
    print(f"{name:28} x{o['launches']:5d} {o['duration_ms_serialised']:9.2f} ms ({100 * o['share_of_serialised_time']:5.1f} %) "
          f"warp inst {o['warp_instructions'] / 1e9:8.2f} G lanes {o['active_threads_per_instruction'] or 0:5.2f} dram {o['dram_bytes'] / 1e9:8.2f} GB")
print("total", {k: f"{v:.4g}" for k, v in total.items()})
