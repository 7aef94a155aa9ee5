#!/usr/bin/env python
"""SASS footprint of every kernel of librt_backend.so (cuobjdump -sass; 16 bytes per instruction on
sm_100a): the instruction-cache argument behind the wavefront split needs the numbers.
usage: tools/sass_sizes.py [lib] > profiles/<tag>_sass_sizes.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rsoderh-raytracing_b200", "librt_backend.so")
text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
sizes, name = {}, None
for line in text.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        sizes[name] = 0
    elif name and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
        sizes[name] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(sizes), capture_output=True, text=True).stdout.splitlines()
rows = sorted(zip(demangled, sizes.values()), key=lambda r: -r[1])
print(f"{'kernel':90} {'instructions':>12} {'KB':>8}")
for n, c in rows:
    n = re.sub(r"\(.*", "", n)
    print(f"{n[:90]:90} {c:12d} {c * 16 / 1024:8.1f}")
