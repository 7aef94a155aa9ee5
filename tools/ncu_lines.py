#!/usr/bin/env python
"""Joins an ncu SASS source page with nvdisasm line info and aggregates per source line.

usage: tools/ncu_lines.py <report.ncu-rep> <kernel mangled-name substring> [top N]
Prints, per CUDA source line: warp instructions, thread instructions, SIMT efficiency,
stall samples — the view used for profiles/*.md.
"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "rsoderh-raytracing_b200", "librt_backend.so")


def sass_lines(kernel_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
    text = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    out, cur, active = [], None, False
    for line in text.splitlines():
        m = re.match(r"^\.text\.(\S+):", line)
        if m:
            active = kernel_substr in m.group(1)
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return out


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    csv_text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                              text=True).stdout
    rows = list(csv.reader(io.StringIO(csv_text)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    lines = sass_lines(kernel)
    if len(lines) != len(data):
        print(f"warning: {len(lines)} SASS instructions from nvdisasm vs {len(data)} in the report",
              file=sys.stderr)
    agg = {}
    src_cache = {}
    tot = [0, 0, 0]
    for (addr, text, loc), r in zip(lines, data):
        wi = int(float(r[col["Instructions Executed"]] or 0))
        ti = int(float(r[col["Thread Instructions Executed"]] or 0))
        sm = int(float(r[col["# Samples"]] or 0))
        a = agg.setdefault(loc, [0, 0, 0, {}])
        a[0] += wi; a[1] += ti; a[2] += sm
        tot[0] += wi; tot[1] += ti; tot[2] += sm
        for k in col:
            if k.startswith("stall_") and "Not Issued" not in k:
                v = int(float(r[col[k]] or 0))
                if v:
                    a[3][k] = a[3].get(k, 0) + v
    print(f"total warp-inst {tot[0]:,} thread-inst {tot[1]:,} avg threads {tot[1] / max(tot[0], 1):.2f} samples {tot[2]:,}")
    print(f"{'file:line':28} {'warp-inst%':>10} {'thr/inst':>8} {'samples%':>8}  top stalls / source")
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
        f, ln = loc if loc else ("?", 0)
        path = os.path.join(ROOT, "rsoderh-raytracing_b200", "csrc", f)
        if path not in src_cache:
            src_cache[path] = open(path).read().splitlines() if os.path.exists(path) else []
        src = src_cache[path][ln - 1].strip() if 0 < ln <= len(src_cache[path]) else ""
        stalls = ", ".join(f"{k[6:]}:{v * 100 // max(a[2], 1)}%" for k, v in
                           sorted(a[3].items(), key=lambda kv: -kv[1])[:3])
        print(f"{f + ':' + str(ln):28} {a[0] * 100 / max(tot[0], 1):10.2f} {a[1] / max(a[0], 1):8.2f} "
              f"{a[2] * 100 / max(tot[2], 1):8.2f}  [{stalls}] {src[:70]}")


if __name__ == "__main__":
    main()
