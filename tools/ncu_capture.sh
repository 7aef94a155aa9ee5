#!/bin/bash
# Full ncu captures of one STEADY-STATE launch of each wavefront kernel (every slot of the
# group populated), house.toml 1080p. Run on a GPU box:  tools/ncu_capture.sh <tag> [spp]
# 1. a duration-only pass lists the wave_walk launches; the launch picked is the one in the
#    middle of the longest-running half (the start-up and drain waves are shorter);
# 2. the same launch index of each of the four kernels is captured with --set full.
# NCU_CMD / NCU_KERNELS override the profiled command and the kernels captured.
# Writes gpurun_out/prof_<tag>_<kernel>.ncu-rep and gpurun_out/ncu_<tag>_pick.txt.
set -u
cd "$(dirname "$0")/.."
tag=${1:-cap}; spp=${2:-32}
mkdir -p gpurun_out
cmd=${NCU_CMD:-"python tools/quick_bench.py $spp 1"}
kernels=${NCU_KERNELS:-"walk shade finish prepass"}
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:wave_walk --csv \
    --log-file gpurun_out/ncu_${tag}_list.csv $cmd > gpurun_out/ncu_${tag}_list.log 2>&1
skip=$(python - "$tag" <<'PY'
import csv, sys
tag = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/ncu_{tag}_list.csv")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Metric Value" in r)
col = rows[hdr].index("Metric Value")
durs = [float(r[col].replace(",", "")) for r in rows[hdr + 1:]]
order = sorted(range(len(durs)), key=lambda i: -durs[i])
top = sorted(order[: max(1, len(durs) // 2)])
pick = top[len(top) // 2]
open(f"gpurun_out/ncu_{tag}_pick.txt", "w").write(
    f"{len(durs)} wave_walk launches; picked index {pick} ({durs[pick]:.0f} ns); "
    f"median of the longer half {sorted(durs[i] for i in top)[len(top) // 2]:.0f} ns; max {max(durs):.0f} ns\n")
print(pick)
PY
)
cat gpurun_out/ncu_${tag}_pick.txt
for k in $kernels; do
  ncu --set full --clock-control none --import-source on -k regex:wave_$k --launch-skip $skip -c 1 \
      -o gpurun_out/prof_${tag}_$k -f $cmd > gpurun_out/ncu_${tag}_$k.log 2>&1
  tail -1 gpurun_out/ncu_${tag}_$k.log
done
