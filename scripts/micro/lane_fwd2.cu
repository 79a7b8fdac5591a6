// Micro-benchmark (B200): the forward pass of k_viterbi_lanes exactly as the kernel runs it (vl_step5_emit, short or general
// branch-error form, decision words stored as they are produced, symbols prefetched one iteration ahead), alone: no group
// set-up, no traceback.  128 threads per CTA, 1..4 CTAs per SM = 1..4 warps per sub-partition, 128 or 240 registers.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../sdrplusplus-dab-radio-plugin_b200/csrc -o lane_fwd2 lane_fwd2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "viterbi_lane_core.h"
#define STEPS 1545

template <bool M128, int CTAS>
__global__ void __launch_bounds__(128, CTAS) k(const uint32_t* __restrict__ sym, uint2* __restrict__ scratch, unsigned long long* out, long long* cyc, const VlConst kc) {
    const uint32_t lane = threadIdx.x & 31u, slot = blockIdx.x * 4u + (threadIdx.x >> 5);
    const uint32_t* __restrict__ srow = sym + lane;
    uint2* __restrict__ dec = scratch + size_t(slot) * (STEPS + 16) * 32u + lane;
    VlState S;
    vl_reset(S);
    uint32_t final_rel = 0;
    uint32_t w[VL_UNROLL];
    const long long t0 = clock64();
#pragma unroll
    for (int q = 0; q < VL_UNROLL; q++) w[q] = __ldg(srow + size_t(q) * 32u);
#pragma unroll 1
    for (uint32_t t = 0; t < STEPS; t += VL_UNROLL) {
        uint32_t wn[VL_UNROLL];
#pragma unroll
        for (int q = 0; q < VL_UNROLL; q++) wn[q] = __ldg(srow + size_t(t + VL_UNROLL + q) * 32u);
        uint2* __restrict__ drow = dec + size_t(t) * 32u;
        vl_step5_emit<M128>(S, w, t, STEPS, [&](const int q, const uint32_t d0, const uint32_t d1) { drow[q * 32] = make_uint2(d0, d1); }, final_rel, kc);
#pragma unroll
        for (int q = 0; q < VL_UNROLL; q++) w[q] = wn[q];
    }
    const long long t1 = clock64();
    out[blockIdx.x * 128 + threadIdx.x] = vl_final_error(S, final_rel);
    if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(cyc), (unsigned long long)(t1 - t0));
}

template <bool M128, int CTAS> void run(const char* name, const uint32_t* sym, uint2* scratch, unsigned long long* out, long long* cyc) {
    const VlConst kc = {0xFFFFFFFFu, 2u, 4u, 16u, 256u, 0x10000u};
    for (int per_sm = 1; per_sm <= CTAS; per_sm++) {
        cudaMemset(cyc, 0, 8);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<M128, CTAS><<<148 * per_sm, 128>>>(sym, scratch, out, cyc, kc);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-40s warps/SMSP=%d  %7.1f cycles per step per warp, %6.1f per SMSP-step, kernel %.3f ms (%s)\n", name, per_sm, double(h) / STEPS,
               double(h) / STEPS / per_sm, ms, cudaGetErrorString(e));
    }
}
int main() {
    uint32_t* sym; uint2* scratch; unsigned long long* out; long long* cyc;
    const size_t rows = STEPS + 16;
    cudaMalloc(&sym, rows * 32 * 4); cudaMalloc(&scratch, size_t(148) * 16 * rows * 32 * 8); cudaMalloc(&out, 148 * 4 * 128 * 8); cudaMalloc(&cyc, 8);
    uint32_t* h = new uint32_t[rows * 32];
    uint32_t r = 1;
    for (size_t i = 0; i < rows * 32; i++) { r = r * 1664525u + 1013904223u; h[i] = (i & 1) ? (r & 0x3F003F00u) : (r & 0x003F003Fu); }   // rate 1/2: two of four symbols punctured
    cudaMemcpy(sym, h, rows * 32 * 4, cudaMemcpyHostToDevice);
    run<false, 4>("short form, 128 registers", sym, scratch, out, cyc);
    run<true, 4>("general form, 128 registers", sym, scratch, out, cyc);
    run<false, 1>("short form, up to 255 registers", sym, scratch, out, cyc);
    return 0;
}
