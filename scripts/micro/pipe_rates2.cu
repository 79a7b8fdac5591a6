// Micro-benchmark (B200): which pipe takes IDP4A / LOP3 / SHF / IADD3, measured by pairing each with VIADDMNMX.U16x2 (ALU pipe,
// one warp instruction every other clock): a pair that reaches ~0.8-1.0 warp-instr/clk/SMSP runs on another pipe, ~0.5 on the same.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates2 pipe_rates2.cu && ./pipe_rates2
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
        for (int i = 0; i < 8; i++) {
            const int j = i + 8;
            if (OP == 0) { r[i] = __dp4a(int(r[i]), int(c), int(r[(i + 1) & 7])); r[j] = __dp4a(int(r[j]), int(c), int(r[8 + ((i + 1) & 7)])); }
            if (OP == 1) { r[i] = __viaddmin_u16x2(r[i], c, r[(i + 1) & 7]); r[j] = __dp4a(int(r[j]), int(c), int(r[8 + ((i + 1) & 7)])); }
            if (OP == 2) { r[i] = __viaddmin_u16x2(r[i], c, r[(i + 1) & 7]); asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r[j]) : "r"(r[j]), "r"(c), "r"(r[8 + ((i + 1) & 7)])); }
            if (OP == 3) { r[i] = __viaddmin_u16x2(r[i], c, r[(i + 1) & 7]); asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(r[j]) : "r"(r[j]), "r"(r[8 + ((i + 1) & 7)]), "r"(m1)); }
            if (OP == 4) { r[i] = __viaddmin_u16x2(r[i], c, r[(i + 1) & 7]); asm volatile("vabsdiff4.u32.s32.s32.add %0, %1, %2, %3;" : "=r"(r[j]) : "r"(r[j]), "r"(c), "r"(r[8 + ((i + 1) & 7)])); }
            if (OP == 5) { r[i] = __viaddmin_u16x2(r[i], c, r[(i + 1) & 7]); asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r[j]) : "r"(r[j]), "r"(m1), "r"(r[8 + ((i + 1) & 7)])); }
            if (OP == 6) { asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r[i]) : "r"(r[i]), "r"(c), "r"(r[(i + 1) & 7])); asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r[j]) : "r"(r[j]), "r"(c), "r"(r[8 + ((i + 1) & 7)])); }
            if (OP == 7) { r[i] = __viaddmin_u16x2(r[i], c, r[(i + 1) & 7]); asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(r[j]) : "r"(r[j]), "r"(c), "r"(r[8 + ((i + 1) & 7)])); }
            if (OP == 8) { r[i] = __viaddmin_u16x2(r[i], c, r[(i + 1) & 7]); asm volatile("add.u32 %0, %1, %2;" : "=r"(r[j]) : "r"(r[j]), "r"(r[8 + ((i + 1) & 7)])); }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    for (int warps = 2; warps <= 8; warps *= 2) {
        k<OP><<<148, warps * 128>>>(out, 12345u, 0xFFFFFFFFu, cyc);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-34s warps/SMSP=%d  %.3f warp-instr/clk/SMSP\n", name, warps, double(ITERS) * 16 * warps / double(h));
    }
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("IDP4A + IDP4A"); run<1>("VIADDMNMX + IDP4A"); run<2>("VIADDMNMX + LOP3"); run<3>("VIADDMNMX + SHF"); run<4>("VIADDMNMX + VABSDIFF4");
    run<5>("VIADDMNMX + IMAD"); run<6>("LOP3 + LOP3"); run<7>("VIADDMNMX + PRMT"); run<8>("VIADDMNMX + IADD");
    return 0;
}
