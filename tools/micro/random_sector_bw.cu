// Microbenchmark: how much DRAM bandwidth do scattered reads of 16 / 32 / 64 / 128-byte pieces
// reach on this GPU? (The wavefront's path state and environment lookups are such accesses.)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o random_sector_bw random_sector_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t pcg(uint32_t s) {
    s = s * 747796405u + 2891336453u;
    uint32_t w = ((s >> ((s >> 28) + 4u)) ^ s) * 277803737u;
    return (w >> 22) ^ w;
}

// Each thread reads `PIECE` bytes (as float4s) at `ITER` random piece-aligned places, UNROLL in flight.
template <int PIECE, int UNROLL>
__global__ void gather(const float4* __restrict__ buf, uint64_t pieces, int iters, float* sink) {
    uint32_t s = pcg(blockIdx.x * blockDim.x + threadIdx.x + 12345u);
    float acc = 0.f;
    for (int it = 0; it < iters; it += UNROLL) {
        float4 v[UNROLL][PIECE / 16];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            s = pcg(s);
            const uint64_t piece = ((uint64_t)s * pieces) >> 32;
            const float4* q = buf + piece * (PIECE / 16);
#pragma unroll
            for (int k = 0; k < PIECE / 16; k++) v[u][k] = __ldg(q + k);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
#pragma unroll
            for (int k = 0; k < PIECE / 16; k++) acc += v[u][k].x + v[u][k].w;
    }
    if (acc == 123.456f) *sink = acc;
}

template <int PIECE, int UNROLL>
void run(const float4* buf, size_t bytes, float* sink) {
    const uint64_t pieces = bytes / PIECE;
    const int threads = 256, blocks = 148 * 8, iters = 256;
    gather<PIECE, UNROLL><<<blocks, threads>>>(buf, pieces, 32, sink);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    gather<PIECE, UNROLL><<<blocks, threads>>>(buf, pieces, iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double useful = (double)blocks * threads * iters * PIECE;
    printf("piece %3d B, %d in flight per thread: %8.1f GB/s useful, %7.2f G pieces/s (%.3f ms)\n", PIECE, UNROLL,
           useful / ms / 1e6, useful / PIECE / ms / 1e6, ms);
}

int main() {
    const size_t max_bytes = (size_t)8 << 30;
    float4* buf;
    float* sink;
    cudaMalloc(&buf, max_bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, max_bytes);
    // footprints: beyond L2 (126 MB) but small, medium, and far beyond the TLB reach
    for (size_t mb : {32, 64, 128, 256, 512, 2048, 8192}) {
        const size_t bytes = mb << 20;
        printf("-- footprint %zu MB\n", mb);
        run<16, 4>(buf, bytes, sink);
        run<32, 4>(buf, bytes, sink);
        run<64, 4>(buf, bytes, sink);
        run<128, 4>(buf, bytes, sink);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
