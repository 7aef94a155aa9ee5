#!/usr/bin/env python
"""Splits the SASS of the persistent kernel (rt_local.cuh) at its BAR.SYNC instructions and
prints, per region between two barriers: warp instructions, average active threads, stall
samples (and how many of them sit on the closing barrier).
usage: tools/ncu_phases.py <report.ncu-rep>"""
import csv
import subprocess
import sys

rep = sys.argv[1]
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(text.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
def num(r, name):
    try:
        return float(r[col[name]].replace(",", ""))
    except (ValueError, IndexError):
        return 0.0
regions, cur = [], {"inst": 0.0, "thr": 0.0, "samples": 0.0, "bar_samples": 0.0, "n": 0, "first": None, "calls": 0}
tot_inst = tot_s = 0.0
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]]
    inst, thr, smp = num(r, "Instructions Executed"), num(r, "Thread Instructions Executed"), num(r, "# Samples")
    tot_inst += inst
    tot_s += smp
    cur["inst"] += inst
    cur["thr"] += thr
    cur["samples"] += smp
    cur["n"] += 1
    if cur["first"] is None:
        cur["first"] = r[col["Address"]]
    if "BAR.SYNC" in src:
        cur["bar_samples"] = smp
        cur["bar_exec"] = inst
        regions.append(cur)
        cur = {"inst": 0.0, "thr": 0.0, "samples": 0.0, "bar_samples": 0.0, "n": 0, "first": None, "calls": 0}
regions.append(cur)
print(f"total warp-inst {tot_inst:,.0f}, samples {tot_s:,.0f}")
print(" region  sass-inst   warp-inst%  thr/inst  samples%  of which on the closing barrier%")
for i, g in enumerate(regions):
    if g["inst"] == 0:
        continue
    print(f"{i:7d} {g['n']:10d} {100 * g['inst'] / tot_inst:11.2f} {g['thr'] / max(g['inst'], 1):9.2f} "
          f"{100 * g['samples'] / tot_s:9.2f} {100 * g['bar_samples'] / tot_s:9.2f}")
