// Micro-benchmark (B200): issue rate per SM sub-partition of the instructions the lane Viterbi is made of.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2000
template <int OP>
__global__ void k(uint32_t* out, uint32_t seed, uint32_t m1, long long* cyc) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = seed * (i + 1) + threadIdx.x;
    const uint32_t c = seed ^ 0x7fff7fff;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (OP == 0) r[i] = __viaddmin_u16x2(r[i], c, r[(i + 1) & 15]);
            if (OP == 1) r[i] = __vminu2(r[i], r[(i + 5) & 15] ^ c);
            if (OP == 2) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r[i]) : "r"(r[i]), "r"(m1), "r"(c));
            if (OP == 3) asm volatile("vabsdiff4.u32.s32.s32.add %0, %1, %2, %3;" : "=r"(r[i]) : "r"(r[i]), "r"(c), "r"(0u));
            if (OP == 4) r[i] = __byte_perm(r[i], c, 0x5410 + (r[(i + 1) & 15] & 1));
            if (OP == 5) r[i] = (r[i] + c) ^ r[(i + 3) & 15];     // IADD3 + LOP3
            if (OP == 6) { r[i] = __viaddmin_u16x2(r[i], c, r[(i + 1) & 15]); asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r[(i + 8) & 15]) : "r"(r[(i + 8) & 15]), "r"(m1), "r"(c)); }
            if (OP == 7) r[i] = __vimin3_u16x2(r[i], c, r[(i + 1) & 15]);
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name, int per_iter) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    for (int warps = 1; warps <= 8; warps *= 2) {   // warps per SM sub-partition
        k<OP><<<148, warps * 128>>>(out, 12345u, 0xFFFFFFFFu, cyc);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-28s warps/SMSP=%d  %.3f warp-instr/clk/SMSP\n", name, warps, double(ITERS) * per_iter * warps / double(h));
    }
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("VIADDMNMX.U16x2", 16); run<1>("VIMNMX.U16x2 (+LOP3)", 32); run<2>("IMAD", 16); run<3>("VABSDIFF4", 16);
    run<4>("PRMT (+LOP3,IADD)", 48); run<5>("IADD3+LOP3", 32); run<6>("VIADDMNMX + IMAD pair", 32); run<7>("VIMNMX3.U16x2", 16);
    return 0;
}
