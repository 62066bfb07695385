// Micro-benchmark: how fast can the warps of ONE SM drain TMEM with tcgen05.ld.32x32b.x32 (the accumulator read of every
// GEMM epilogue / per-pixel chain stage)?  Prints cycles per warp-level LDTM.x32 (4 KB) and bytes per clock per SM for
// 4, 8 and 12 warps, with one and with four loads in flight per warp.  The per-pixel chains drain 192 (attention program) and
// 384 (shot program) accumulator columns of 128 lanes per tile: this number is their TMEM floor.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I noisediff_b200/csrc tools/ubench_tmem.cu -o /tmp/ubench_tmem
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "common.cuh"

using namespace ndiff;

template <int DEPTH>
__global__ void __launch_bounds__(384, 1) drain_kernel(int iters, long long* cycles, unsigned* sink) {
    __shared__ uint32_t tmem_base_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&tmem_base_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = tmem_base_slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    unsigned acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t v[DEPTH][32];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) tmem_ld32(base + d * 32, v[d]);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) acc ^= v[d][0] ^ v[d][31];
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_base_slot); }
}

template <int DEPTH>
void run(int warps, int ctas) {
    long long* cyc; unsigned* sink;
    cudaMalloc(&cyc, sizeof(long long) * ctas);
    cudaMalloc(&sink, 4);
    const int iters = 4096;
    drain_kernel<DEPTH><<<ctas, warps * 32>>>(64, cyc, sink);          // warm-up
    drain_kernel<DEPTH><<<ctas, warps * 32>>>(iters, cyc, sink);
    cudaError_t err = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per_ld = static_cast<double>(h) / (static_cast<double>(iters) * DEPTH * warps);
    std::printf("{\"warps\": %d, \"ctas\": %d, \"loads_in_flight_per_warp\": %d, \"cycles\": %lld, \"cycles_per_warp_ldtm_x32\": %.2f, "
                "\"tmem_read_bytes_per_clk_per_sm\": %.1f, \"err\": \"%s\"}\n",
                warps, ctas, DEPTH, h, per_ld * warps, 4096.0 / per_ld, cudaGetErrorString(err));
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    for (int ctas : {1, 148})
        for (int warps : {4, 8, 12}) {
            run<1>(warps, ctas);
            run<4>(warps, ctas);
        }
    return 0;
}
