# Build recipe for the B200 path-tracing backend.
#   make host      -> rsoderh-raytracing_b200/librt_host.so     (C++17, no GPU)
#   make backend   -> rsoderh-raytracing_b200/librt_backend.so  (nvcc, sm_100a)
#   make cli       -> rsoderh-raytracing_b200/rt_render         (headless renderer)
#   make oracle    -> oracle/liboracle.so                        (test infrastructure only)
PKG      := rsoderh-raytracing_b200
CXX      ?= g++
NVCC     ?= /usr/local/cuda/bin/nvcc
CUDA_HOME ?= /usr/local/cuda
CXXFLAGS := -O2 -std=c++17 -fPIC -Wall -Wextra -ffp-contract=off -fno-fast-math -Iinclude
NVFLAGS  := $(EXTRA_NVFLAGS) -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
            --fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
            -Xcompiler -fPIC -Iinclude -Xptxas -v
NCCL_INC ?= /usr/include
NCCL_LIB ?= /usr/lib/x86_64-linux-gnu

HOST_SRC := $(PKG)/host/toml_lite.cpp $(PKG)/host/obj.cpp $(PKG)/host/scene.cpp \
            $(PKG)/host/bvh.cpp $(PKG)/host/env.cpp $(PKG)/host/camera.cpp \
            $(PKG)/host/synthetic.cpp $(PKG)/host/host_api.cpp
HOST_HDR := $(PKG)/host/scene.hpp $(PKG)/host/toml_lite.hpp include/rt_host.h include/rt_types.h

.PHONY: all host backend cli oracle clean
all: host backend cli oracle

host: $(PKG)/librt_host.so
$(PKG)/librt_host.so: $(HOST_SRC) $(HOST_HDR)
	$(CXX) $(CXXFLAGS) -shared -o $@ $(HOST_SRC) -lz

backend: $(PKG)/librt_backend.so
$(PKG)/librt_backend.so: $(PKG)/csrc/rt_backend.cu $(wildcard $(PKG)/csrc/*.cuh) include/rt_backend.h include/rt_types.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(PKG)/csrc/rt_backend.cu -ldl -lpthread 2> $(PKG)/csrc/ptxas.log || (cat $(PKG)/csrc/ptxas.log; false)
	@grep -E "registers|spill|error" $(PKG)/csrc/ptxas.log | head -40 || true

cli: $(PKG)/rt_render
$(PKG)/rt_render: $(PKG)/host/cli_main.cpp $(PKG)/librt_host.so $(PKG)/librt_backend.so
	$(CXX) $(CXXFLAGS) -o $@ $(PKG)/host/cli_main.cpp -L$(PKG) -lrt_host -lrt_backend \
	    -Wl,-rpath,'$$ORIGIN' -lpthread

oracle: oracle/liboracle.so
oracle/liboracle.so: oracle/shader_oracle.cpp
	$(CXX) -O2 -std=c++17 -fPIC -Wall -Wextra -ffp-contract=off -fno-fast-math -shared \
	    -o $@ oracle/shader_oracle.cpp -lpthread

clean:
	rm -f $(PKG)/librt_host.so $(PKG)/librt_backend.so $(PKG)/rt_render oracle/liboracle.so

# Host build of the production traversal for the CPU test suite (tests/test_fast_emu.py).
.PHONY: emu
emu: tests/emu/libfast_emu.so
tests/emu/libfast_emu.so: tests/emu/fast_emu.cu $(wildcard $(PKG)/csrc/*.cuh) $(PKG)/csrc/rt_fast_host.hpp include/rt_types.h
	$(NVCC) -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a --fmad=false -prec-div=true \
	    -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -Iinclude \
	    -shared -o $@ tests/emu/fast_emu.cu
