#!/bin/bash
# Builds a tuning variant of the backend: tools/build_variant.sh NAME -DRT_LOCAL_THREADS=256 ...
# -> build/variants/librt_backend_NAME.so (select it with RT_BACKEND_LIB=<path>)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
/usr/local/cuda/bin/nvcc "$@" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
    --fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC -Iinclude -Xptxas -v \
    -shared -o build/variants/librt_backend_$name.so rsoderh-raytracing_b200/csrc/rt_backend.cu -ldl -lpthread \
    2> build/variants/ptxas_$name.log || { cat build/variants/ptxas_$name.log; exit 1; }
grep -A2 "local_wave" build/variants/ptxas_$name.log | grep -E "spill|Used" | sed "s/^/$name: /"
