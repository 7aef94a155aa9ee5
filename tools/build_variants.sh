#!/bin/bash
# Builds tuning variants of librt_backend.so into build/variants/<name>.so.
# usage: tools/build_variants.sh name1:"-DFLAG=1 -DX=2" name2:"..." ...
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
pids=()
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  ( $NVCC $flags -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -prec-div=true \
      -prec-sqrt=true -ftz=false -Xcompiler -fPIC -Iinclude -Xptxas -v -shared -o build/variants/$name.so \
      rsoderh-raytracing_b200/csrc/rt_backend.cu -ldl 2> build/variants/$name.log \
    && grep -A2 "render_fast" build/variants/$name.log | grep -E "Used" | sed "s/^/$name: /" ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
