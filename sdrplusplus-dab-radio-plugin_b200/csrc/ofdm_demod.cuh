// ofdm_demod.cuh -- frame demodulation kernel: PLL -> cyclic-prefix phase -> FFT -> DQPSK -> frequency
// de-interleave -> int8 soft bits, one pass over the raw IQ of a completed frame.
//
// Reference (paths relative to /root/reference/vendor/DAB-Radio/src/ofdm):
//   PipelineThread                                    ofdm_demodulator.cpp:650-766
//   apply_pll_avx (phase arithmetic reproduced)       dsp/apply_pll.cpp:82-116
//   CalculateCyclicPhaseError / complex_conj_mul_sum  ofdm_demodulator.cpp:768-777, dsp/complex_conj_mul_sum.cpp:65-99
//   CalculateFFT (FFTW3f, unnormalised forward DFT)   ofdm_demodulator.cpp:891-894
//   CalculateDQPSK                                    ofdm_demodulator.cpp:842-865
//   CalculateViterbiBits / convert_to_viterbi_bit     ofdm_demodulator.cpp:867-889, 58-72
//
// Structure (B200): one CTA of N/8 threads per (stream, chunk of consecutive symbols).
//   * The raw IQ of symbol l+1 is staged into shared memory by a TMA bulk copy (cp.async.bulk + mbarrier,
//     double buffered) while symbol l is computed, so global-load latency never sits on the critical path.
//   * Every thread derotates its 8 FFT inputs straight from the staged bytes into registers (u8 -> f32
//     conversion fused), plus the <= 2 cyclic-prefix samples it needs for the CP correlation.
//   * FFT: decimation in frequency, radices {4|2|1} x 8 x 8 x 8; the first pass runs in registers, the middle
//     passes through shared memory as float2 (64-bit accesses, padded layouts that are conflict free for both
//     the writer and the reader of each buffer), the last radix-8 pass reads 8 contiguous points with four
//     128-bit loads and leaves the spectrum IN REGISTERS: thread t always owns bins kk(t) + (N/8)*k4.
//   * DQPSK therefore needs no spectrum buffer: X_{l-1} is simply the register copy kept from the previous
//     iteration.  Soft bits are scattered (frequency de-interleaver) into a 2K-byte shared row and leave with
//     16-byte coalesced stores.
//   * Frame layout in the ring (common.cuh: frame layout): the FIC symbols in natural order, every CIF of the MSC as 16 PLANES
//     (plane r = the soft bits whose index inside the CIF is r modulo 16).  The time de-interleaver takes bit i of a logical
//     frame from the CIF that is 15 - bitrev4(i mod 16) CIFs old (cif_deinterleaver.cpp:8-11, 62-68), so with planes every
//     Viterbi job reads 16 contiguous runs instead of every 16th byte of 16 rows.  A data symbol is a whole number of 16-bit
//     groups, so it contributes one contiguous run of 2K/16 bytes to each plane: the shared row is filled plane-major and the
//     stores stay coalesced.
#pragma once
#include "ofdm.cuh"

#define DEMOD_PADA(a) ((a) + (((a) >> 6) << 3))
#define DEMOD_L3(a) (10 * ((a) >> 3) + ((a) & 7) + (((a) >> 6) << 3))

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared; completion is signalled on the mbarrier as transaction bytes
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

template <int N, int FMT> struct DemodCfg {
    static constexpr int T = N / 8;                                 // threads per CTA = FFT points / 8
    static constexpr int BPS = (FMT == DABGPU_IQ_U8) ? 2 : 8;       // bytes per IQ sample
    static constexpr int CP = N * 63 / 256;
    static constexpr int TSYM = N + CP;
    static constexpr int K = 3 * N / 4;
    static constexpr int R1 = (N == 2048 || N == 256) ? 4 : (N == 1024 ? 2 : 1);
    static constexpr int NA = N + N / 8;                            // DEMOD_PADA extent
    static constexpr int NB = 11 * N / 8;                           // DEMOD_L3 extent
    static constexpr int FIC_SYMS = (N == 256) ? 8 : 3;               // dab_parameters.h:26-90
    static constexpr int SYMS_PER_CIF = 55296 / (2 * K);
    static constexpr int RUN = 2 * K / 16;                            // bytes one data symbol adds to every plane of its CIF
    static constexpr int CH = (RUN % 16 == 0) ? 16 : 8;               // store granularity (mode III: runs of 24 bytes)
    static constexpr int STAGE = ((TSYM * BPS + 15 + 15) / 16) * 16;   // one symbol + alignment slack
    static constexpr int WARPS = (T + 31) / 32;
    static constexpr size_t SMEM = size_t(NA + NB) * 8 + size_t(2 * K) + 2 * size_t(STAGE) + 16 /*mbarriers*/ + size_t(WARPS) * 8 + 16;
};

// apply_pll_avx phase arithmetic (dsp/apply_pll.cpp:95-108) kept operation by operation in float32 -- the
// quantisation of t at |t| ~ 1e3 cycles dominates the reference's own phase noise -- while sin/cos of the wrapped
// phase come from the SFU (sin.approx, |err| < 5e-7) instead of the 6-term Chebyshev polynomial.
//   nf = float(n & ~3), kf = float(n & 3) * f, kf25 = kf + 0.25f
// round to nearest even by the 1.5 * 2^23 trick (two FADDs on the FMA pipe instead of FRND on the quarter-rate XU pipe);
// identical to rintf for |t| < 2^22, and |t| stays below a few thousand cycles here
__device__ __forceinline__ float pll_rint(const float t) { return __fsub_rn(__fadd_rn(t, 12582912.0f), 12582912.0f); }

__device__ __forceinline__ float2 pll_rotate_sfu(const float2 v, const float nf, const float kf, const float kf25, const float f, const float dt0) {
    const float base = __fmaf_rn(nf, f, dt0);
    float ts = __fadd_rn(base, kf);
    float tc = __fadd_rn(base, kf25);
    ts = __fsub_rn(ts, pll_rint(ts));
    tc = __fsub_rn(tc, pll_rint(tc));
    const float sn = __sinf(ts * 6.283185307179586f), cs = __sinf(tc * 6.283185307179586f);
    return make_float2(__fmaf_rn(cs, v.x, -__fmul_rn(sn, v.y)), __fmaf_rn(cs, v.y, __fmul_rn(sn, v.x)));
}

template <int FMT> __device__ __forceinline__ float2 demod_raw_to_c32(const uint8_t* p) {
    if (FMT == DABGPU_IQ_U8) {
        // (u8 - 127.5) * (1/127.5) as the reference computes it (app_iq_readers.h:19-27), with the u8 -> f32 conversion done
        // by placing the byte in the mantissa of 2^15 (PRMT, ALU pipe) instead of I2F (XU pipe): 32768 + b - 32895.5 is exact
        const uint32_t v = *reinterpret_cast<const uint16_t*>(p);
        const float fx = __uint_as_float(__byte_perm(v, 0x47000000u, 0x7404));
        const float fy = __uint_as_float(__byte_perm(v, 0x47000000u, 0x7414));
        return make_float2(__fmul_rn(__fsub_rn(fx, 32895.5f), 1.0f / 127.5f), __fmul_rn(__fsub_rn(fy, 32895.5f), 1.0f / 127.5f));
    }
    return *reinterpret_cast<const float2*>(p);
}

__device__ __forceinline__ float2 cmul_tw(const float2 v, const float2* __restrict__ tw, const int idx) { return cmulf(v, __ldg(tw + idx)); }
template <int N, int FMT>
__global__ void __launch_bounds__(N / 8, (N == 2048) ? 4 : (N == 1024 ? 6 : 8))
k_ofdm_demod2(const OfdmDev D, const int first_stream, const int sym_per_chunk) {
    using C = DemodCfg<N, FMT>;
    constexpr int T = C::T, CP = C::CP, TSYM = C::TSYM, K = C::K, BPS = C::BPS;
    // the per-symbol single-thread chores go to different warps, so that no warp is late for the CTA barriers by both of them
    constexpr int T_PREFETCH = (T > 32) ? 32 : 0, T_PHASE = (T > 64) ? 64 : 0;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float2* sA = reinterpret_cast<float2*>(smem_raw);
    float2* sB = sA + C::NA;
    uint8_t* s_out = reinterpret_cast<uint8_t*>(sB + C::NB);
    uint8_t* s_stage = s_out + 2 * K;                         // 2 x STAGE, 16-byte aligned (NA, NB, 2K multiples of 16 bytes)
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stage + 2 * C::STAGE);
    float2* s_part = reinterpret_cast<float2*>(s_bar + 2);

    const int s = first_stream + blockIdx.y;
    const OfdmStream st = D.st[s];
    if (st.pending != 1) return;
    const OfdmGeom& g = D.g;
    const int tid = threadIdx.x;
    const int l0 = blockIdx.x * sym_per_chunk;
    if (l0 >= g.L - 1) return;
    const int l1 = min(l0 + sym_per_chunk, g.L - 1);
    const float f = st.frame_f;
    const float2* head = D.head + size_t(s) * (g.Tsym + g.CP);
    const uint32_t slot = D.frames_written[s] & D.slot_mask;
    int8_t* out_frame = D.frames + (size_t(s) * (D.slot_mask + 1u) + slot) * g.frame_bits;

    // staged (TMA) path: the frame is contiguous in the ring and the ring geometry keeps 16-byte blocks inside the stream
    const bool staged = (st.head_contig != 0) && (D.tma_ok != 0);
    const uint8_t* ring_s = D.ring + size_t(s) * D.ring_stride * BPS;
    const unsigned long long frame_abs0 = st.frame_ring_base - st.n_head;   // abs ring index of frame sample 0 (staged path only)
    const bool wraps = (D.ring_mask != ~0ull);

    if (staged && tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue_prefetch = [&](const int l, const int buf) {   // one thread
        const unsigned long long a = frame_abs0 + (unsigned long long)(l) * TSYM;
        const unsigned long long pos = a & D.ring_mask;
        const unsigned long long byte0 = pos * BPS;
        const unsigned long long start16 = byte0 & ~15ull;
        uint8_t* dst = s_stage + buf * C::STAGE;
        if (!wraps || pos + TSYM <= D.ring_mask + 1ull) {
            const uint32_t bytes = uint32_t(((byte0 + TSYM * BPS + 15ull) & ~15ull) - start16);
            mbar_expect_tx(&s_bar[buf], bytes);
            bulk_g2s(dst, ring_s + start16, bytes, &s_bar[buf]);
        } else {
            const unsigned long long ring_bytes = (D.ring_mask + 1ull) * BPS;
            const uint32_t b1 = uint32_t(ring_bytes - start16);
            const uint32_t rest = uint32_t(byte0 + TSYM * BPS - ring_bytes);
            const uint32_t b2 = (rest + 15u) & ~15u;
            mbar_expect_tx(&s_bar[buf], b1 + b2);
            bulk_g2s(dst, ring_s + start16, b1, &s_bar[buf]);
            bulk_g2s(dst + b1, ring_s, b2, &s_bar[buf]);
        }
    };
    if (staged && tid == 0) issue_prefetch(l0, 0);

    // per-thread constants.  The thread owns bins kk + (N/8)*k4: k4 = 1,2,5,6,7 always carry data, k4 = 4 never does, and
    // exactly one of k4 = 0 (kk != 0) / k4 = 3 (kk == 0: carrier +K/2, while bin 0 is the empty DC carrier) does, so every
    // thread has six DQPSK products per symbol and the loop below needs no per-bin branch.  op[] = output positions.
    uint32_t op[6];
    bool dc_thread;
    {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(D.obin) + tid);
        dc_thread = (v.x & 0xFFFFu) == 0xFFFFu;
        op[0] = dc_thread ? (v.y >> 16) : (v.x & 0xFFFFu);
        op[1] = v.x >> 16; op[2] = v.y & 0xFFFFu; op[3] = v.z >> 16; op[4] = v.w & 0xFFFFu; op[5] = v.w >> 16;
        // low half: position in a natural-order row (FIC symbols); high half: position in a plane-major row (MSC symbols), where
        // bit b of the symbol sits at (b mod 16) * RUN + b / 16.  The second bit of a carrier is K positions later, K = 0 mod 16.
#pragma unroll
        for (int k = 0; k < 6; k++) op[k] |= ((op[k] & 15u) * uint32_t(C::RUN) + (op[k] >> 4)) << 16;
    }
    float nf[8];
#pragma unroll
    for (int m = 0; m < 8; m++) nf[m] = float((CP + tid + m * T) & ~3);
    const float k3 = float((CP + tid) & 3);   // (CP + tid + m*T) & 3 is the same for every m (T multiple of 4)
    // cyclic-prefix partners of the FFT inputs m = 6, 7: c = (tid + m*T) - (N - CP)
    const int c6 = tid + 6 * T - (N - CP), c7 = tid + 7 * T - (N - CP);
    const float nf6 = float(c6 & ~3), nf7 = float(c7 & ~3);
    const float kc6 = float(c6 & 3), kc7 = float(c7 & 3);

    // the root of every pass's twiddles for this thread (tw[i] = exp(-2 pi j i / N)); the powers are formed where they are used
    const float2 w_p1a = __ldg(D.tw + tid), w_p1b = __ldg(D.tw + tid + T);
    const float2 w_p2 = __ldg(D.tw + (tid & 63) * (N / 512)), w_p3 = __ldg(D.tw + (tid & 7) * (N / 64));

    float2 prev[8];
#pragma unroll
    for (int k = 0; k < 8; k++) prev[k] = make_float2(0.0f, 0.0f);

    for (int l = l0; l <= l1; l++) {
        const int it = l - l0, buf = it & 1;
        const bool own = (l < l1) || (l == g.L - 1);   // this CTA accounts for the symbol's cyclic-prefix phase
        const float dt0 = __fmul_rn(float(l * TSYM), f);
        const float kf = __fmul_rn(k3, f), kf25 = __fadd_rn(kf, 0.25f);
        float2 x[8];
        float2 acc = make_float2(0.0f, 0.0f);
        if (staged) {
            if (tid == T_PREFETCH && l < l1) issue_prefetch(l + 1, buf ^ 1);   // buffer buf^1 was last read two barriers ago
            mbar_wait(&s_bar[buf], uint32_t(it >> 1) & 1u);
            const unsigned long long a = frame_abs0 + (unsigned long long)(l) * TSYM;
            const uint32_t shift = uint32_t(((a & D.ring_mask) * BPS) & 15ull);
            const uint8_t* raw = s_stage + buf * C::STAGE + shift;
#pragma unroll
            for (int m = 0; m < 8; m++)
                x[m] = pll_rotate_sfu(demod_raw_to_c32<FMT>(raw + (CP + tid + m * T) * BPS), nf[m], kf, kf25, f, dt0);
            if (own) {
                // sum over n < CP of sym[N+n] * conj(sym[n])   (ofdm_demodulator.cpp:768-777)
                if (c6 >= 0) {
                    const float kk = __fmul_rn(kc6, f);
                    const float2 c = pll_rotate_sfu(demod_raw_to_c32<FMT>(raw + c6 * BPS), nf6, kk, __fadd_rn(kk, 0.25f), f, dt0);
                    acc.x += __fmaf_rn(x[6].y, c.y, x[6].x * c.x);
                    acc.y += __fmaf_rn(x[6].y, c.x, -(x[6].x * c.y));
                }
                {
                    const float kk = __fmul_rn(kc7, f);
                    const float2 c = pll_rotate_sfu(demod_raw_to_c32<FMT>(raw + c7 * BPS), nf7, kk, __fadd_rn(kk, 0.25f), f, dt0);
                    acc.x += __fmaf_rn(x[7].y, c.y, x[7].x * c.x);
                    acc.y += __fmaf_rn(x[7].y, c.x, -(x[7].x * c.y));
                }
            }
        } else {
            // generic path (first frame after acquisition, or a caller buffer that TMA cannot address): direct loads
            const uint32_t sym_base = uint32_t(l) * uint32_t(TSYM);
#pragma unroll
            for (int m = 0; m < 8; m++)
                x[m] = pll_rotate_sfu(ofdm_frame_sample(D, s, st, sym_base + CP + tid + m * T, head), nf[m], kf, kf25, f, dt0);
            if (own) {
                if (c6 >= 0) {
                    const float kk = __fmul_rn(kc6, f);
                    const float2 c = pll_rotate_sfu(ofdm_frame_sample(D, s, st, sym_base + c6, head), nf6, kk, __fadd_rn(kk, 0.25f), f, dt0);
                    acc.x += __fmaf_rn(x[6].y, c.y, x[6].x * c.x);
                    acc.y += __fmaf_rn(x[6].y, c.x, -(x[6].x * c.y));
                }
                {
                    const float kk = __fmul_rn(kc7, f);
                    const float2 c = pll_rotate_sfu(ofdm_frame_sample(D, s, st, sym_base + c7, head), nf7, kk, __fadd_rn(kk, 0.25f), f, dt0);
                    acc.x += __fmaf_rn(x[7].y, c.y, x[7].x * c.x);
                    acc.y += __fmaf_rn(x[7].y, c.x, -(x[7].x * c.y));
                }
            }
        }
        if (own) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc.x += __shfl_xor_sync(FULL_MASK, acc.x, o);
                acc.y += __shfl_xor_sync(FULL_MASK, acc.y, o);
            }
            if ((tid & 31) == 0) s_part[tid >> 5] = acc;   // read after the next barriers; rewritten one iteration later
        }

        // ---- FFT pass 1 in registers -----------------------------------------------------------------
        if (C::R1 == 4) {
            dft4<false>(x[0], x[2], x[4], x[6]);
            dft4<false>(x[1], x[3], x[5], x[7]);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int r = tid + h * T;
                const float2 w1 = h ? w_p1b : w_p1a, w2 = cmulf(w1, w1), w3 = cmulf(w2, w1);
                sA[DEMOD_PADA(r)] = x[h];
                sA[DEMOD_PADA(N / 4 + r)] = cmulf(x[2 + h], w1);
                sA[DEMOD_PADA(2 * (N / 4) + r)] = cmulf(x[4 + h], w2);
                sA[DEMOD_PADA(3 * (N / 4) + r)] = cmulf(x[6 + h], w3);
            }
        } else if (C::R1 == 2) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int r = tid + c * T;
                const float2 a = x[c], b = x[c + 4];
                sA[DEMOD_PADA(r)] = make_float2(a.x + b.x, a.y + b.y);
                sA[DEMOD_PADA(N / 2 + r)] = cmul_tw(make_float2(a.x - b.x, a.y - b.y), D.tw, r);
            }
        } else {
            dft8<false>(x);
            float2 wp[8];
            tw_powers7(w_p1a, wp);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                float2 v = x[k];
                if (k > 0) v = cmulf(v, wp[k]);
                sA[DEMOD_PADA(k * (N / 8) + tid)] = v;
            }
        }
        __syncthreads();   // (1)
        // ---- radix-8 pass over blocks of 512 (in place) ------------------------------------------------
        if (N >= 1024) {
            const int q = tid >> 6, r = tid & 63;
            const int base = q * 512 + r;
            float2 y[8];
            float2 wp[8];
            tw_powers7(w_p2, wp);
#pragma unroll
            for (int j = 0; j < 8; j++) y[j] = sA[DEMOD_PADA(base + j * 64)];
            dft8<false>(y);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                float2 v = y[k];
                if (k > 0) v = cmulf(v, wp[k]);
                sA[DEMOD_PADA(base + k * 64)] = v;
            }
            // (2) a block of 512 points belongs to the 64 threads with the same q, in this pass and in the next one: only those
            // two warps meet (named barrier 1 + q), not the CTA
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        }
        // ---- radix-8 pass over blocks of 64: A -> B (layout L3) ----------------------------------------
        {
            const int q = tid >> 3, r = tid & 7;
            const int base = q * 64 + r;
            float2 y[8];
            float2 wp[8];
            tw_powers7(w_p3, wp);
#pragma unroll
            for (int j = 0; j < 8; j++) y[j] = sA[DEMOD_PADA(base + j * 8)];
            dft8<false>(y);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                float2 v = y[k];
                if (k > 0) v = cmulf(v, wp[k]);
                sB[DEMOD_L3(base + k * 8)] = v;
            }
        }
        // (3) a block of 64 points is written by eight consecutive threads and its eight contiguous runs are read back by the same
        // eight: sB is private to the warp
        __syncwarp();
        // ---- last radix-8 pass: 8 contiguous points, spectrum stays in registers -----------------------
        {
            const float4* p = reinterpret_cast<const float4*>(sB + (10 * tid + ((tid >> 3) << 3)));
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float4 v = p[j];
                x[2 * j] = make_float2(v.x, v.y);
                x[2 * j + 1] = make_float2(v.z, v.w);
            }
            dft8<false>(x);
        }
        if (own && tid == T_PHASE) {
            float2 t = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int w = 0; w < C::WARPS; w++) { t.x += s_part[w].x; t.y += s_part[w].y; }
            D.phase_err[size_t(s) * g.L + l] = atan2f(t.y, t.x);
        }
        if (l > l0) {
            // DQPSK X_{l-1} * conj(X_l) on the owned bins, L-infinity normalise, truncate to int8
            // (ofdm_demodulator.cpp:842-889: bit0 = trunc(-127*re/A), bit1 = trunc(+127*im/A)).  The scale carries a
            // 1e-6 guard so that the dominant component truncates to exactly +-127 with the approximate reciprocal.
            const float2 pa[6] = {dc_thread ? prev[3] : prev[0], prev[1], prev[2], prev[5], prev[6], prev[7]};
            const float2 xb[6] = {dc_thread ? x[3] : x[0], x[1], x[2], x[5], x[6], x[7]};
            const bool msc = (l - 1) >= C::FIC_SYMS;
            const uint32_t sel = msc ? 0x4432u : 0x4410u;            // which half of op[]
            const uint32_t second = msc ? uint32_t(K / 16) : uint32_t(K);
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const float2 a = pa[k], b = xb[k];
                const float vr = a.x * b.x + a.y * b.y;
                const float vi = a.y * b.x - a.x * b.y;
                const float sc = __fdividef(127.00012f, fmaxf(fabsf(vr), fabsf(vi)));
                const uint32_t pos = __byte_perm(op[k], 0u, sel);
                s_out[pos] = uint8_t(int(vr * -sc));
                s_out[pos + second] = uint8_t(int(vi * sc));
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) prev[k] = x[k];
        __syncthreads();   // (4) s_out complete; sA/sB/s_part free for the next symbol
        if (l > l0) {
            const int m = l - 1;
            if (m < C::FIC_SYMS) {
                const uint4* src = reinterpret_cast<const uint4*>(s_out);
                uint4* dst = reinterpret_cast<uint4*>(out_frame + size_t(m) * 2u * K);
                for (int i = tid; i < 2 * K / 16; i += T) dst[i] = src[i];
            } else {
                // 16 runs of RUN bytes, one per plane of the symbol's CIF
                const int cif = (m - C::FIC_SYMS) / C::SYMS_PER_CIF, q = (m - C::FIC_SYMS) - cif * C::SYMS_PER_CIF;
                int8_t* base = out_frame + size_t(C::FIC_SYMS) * 2u * K + size_t(cif) * 55296u + size_t(q) * C::RUN;
                constexpr int PER_RUN = C::RUN / C::CH;
                for (int i = tid; i < 2 * K / C::CH; i += T) {
                    const int run = i / PER_RUN, j = i - run * PER_RUN;
                    if (C::CH == 16) *reinterpret_cast<uint4*>(base + run * 3456 + j * 16) = reinterpret_cast<const uint4*>(s_out)[i];
                    else *reinterpret_cast<uint2*>(base + run * 3456 + j * 8) = reinterpret_cast<const uint2*>(s_out)[i];
                }
            }
            // s_out is rewritten only after barrier (3) of the next iteration
        }
    }
}
