// Micro-benchmark (B200): cycles per trellis step of the lane-per-trellis forward pass (viterbi_lane_core.h), per warp,
// with 1 / 2 / 4 warps per SM sub-partition, and with parts of the step removed to see what each costs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../sdrplusplus-dab-radio-plugin_b200/csrc -o lane_fwd lane_fwd.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "viterbi_lane_core.h"
#define GROUPS 400   // 5-step groups

template <int V>
__global__ void __launch_bounds__(512, 1) k(const uint32_t* __restrict__ sym, uint2* __restrict__ dec, unsigned long long* out, long long* cyc, const VlConst kc) {
    VlState S;
    vl_reset(S);
    uint64_t fe = 0; uint32_t frel = 0;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t* s = sym + (threadIdx.x & 31);
    uint2* d = dec + size_t(tid);
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (uint32_t g = 0; g < GROUPS; g++) {
        uint32_t w[5], dd[10];
#pragma unroll
        for (int q = 0; q < 5; q++) w[q] = __ldg(s + (g * 5 + q) * 32);
        if (V == 0) vl_step5(S, w, g * 5, GROUPS * 5, dd, frel, kc);
        else {
            uint32_t Ea[8], Eia[8], Eb[8], Eib[8];
            vl_branch<0, true>(w[0], Ea, Eia, kc);
            vl_acs<0, true>(S.R, Ea, Eia, S.CL, dd[0], dd[1], kc); vl_branch<1, true>(w[1], Eb, Eib, kc);
            vl_acs<1, true>(S.R, Eb, Eib, S.CL, dd[2], dd[3], kc); vl_branch<2, true>(w[2], Ea, Eia, kc);
            vl_acs<2, true>(S.R, Ea, Eia, S.CL, dd[4], dd[5], kc); vl_branch<3, true>(w[3], Eb, Eib, kc);
            vl_acs<3, true>(S.R, Eb, Eib, S.CL, dd[6], dd[7], kc); vl_branch<4, true>(w[4], Ea, Eia, kc);
            vl_acs<4, true>(S.R, Ea, Eia, S.CL, dd[8], dd[9], kc);
            vl_repack(S.R);
            if (V == 1) {
                const uint32_t delta = (S.R[0] & 0xFFFFu) - VL_ORIGIN;
                const uint32_t d2 = delta * 0x10001u;
#pragma unroll
                for (int i = 0; i < 32; i++) S.R[i] -= d2;
                vl_set_off(S, S.off + int32_t(delta));
            }
        }
        if (V != 3) {
#pragma unroll
            for (int q = 0; q < 5; q++) d[size_t(g * 5 + q) * stride] = make_uint2(dd[2 * q], dd[2 * q + 1]);
        } else fe += dd[0] ^ dd[3] ^ dd[4] ^ dd[7] ^ dd[9] ^ dd[1] ^ dd[2] ^ dd[5] ^ dd[6] ^ dd[8];
    }
    const long long t1 = clock64();
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) x ^= S.R[i];
    out[tid] = fe + x + vl_final_error(S, frel);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int V> void run(const char* name, const uint32_t* sym, uint2* dec, unsigned long long* out, long long* cyc) {
    const VlConst kc = {0xFFFFFFFFu, 2u, 4u, 16u, 256u, 0x10000u};
    for (int warps = 1; warps <= 4; warps *= 2) {
        k<V><<<148, warps * 128>>>(sym, dec, out, cyc, kc);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-44s warps/SMSP=%d  %.1f cycles per trellis step per warp, %.1f per SMSP-step  (%s)\n", name, warps, double(h) / (GROUPS * 5),
               double(h) / (GROUPS * 5) / warps, cudaGetErrorString(e));
    }
}
int main() {
    uint32_t* sym; uint2* dec; unsigned long long* out; long long* cyc;
    cudaMalloc(&sym, GROUPS * 5 * 32 * 4); cudaMalloc(&dec, size_t(GROUPS) * 5 * 148 * 512 * 8); cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 8);
    uint32_t* h = new uint32_t[GROUPS * 5 * 32];
    uint32_t r = 1;
    for (int i = 0; i < GROUPS * 5 * 32; i++) { r = r * 1664525u + 1013904223u; h[i] = r & 0x3F3F3F3Fu; }   // moderate symbols: frequent renormalisations
    cudaMemcpy(sym, h, GROUPS * 5 * 32 * 4, cudaMemcpyHostToDevice);
    run<0>("full step (events + own renorm + stores)", sym, dec, out, cyc);
    run<1>("no reference-renormalisation check", sym, dec, out, cyc);
    run<2>("no renormalisation at all", sym, dec, out, cyc);
    run<3>("no renormalisation, no decision stores", sym, dec, out, cyc);
    return 0;
}
