#!/bin/bash
# Runs tools/quick_bench.py for every build/variants/*.so (plus the in-tree library).
# usage: tools/run_variants.sh [spp] [reps] [scene] [w] [h]
cd "$(dirname "$0")/.."
python tools/quick_bench.py "$@"
for so in build/variants/*.so; do
  RT_BACKEND_LIB=$PWD/$so timeout 120 python tools/quick_bench.py "$@" || echo "$so FAILED"
done
