// ofdm.cuh -- batched OFDM demodulator: per-stream control kernel (sync state machine) + wide
// frame demodulation kernel (PLL -> cyclic-prefix phase -> FFT -> DQPSK -> frequency de-interleave
// -> int8 soft bits in one pass over the IQ).
//
// Reference (paths relative to /root/reference/vendor/DAB-Radio/src/ofdm):
//   OFDM_Demod::Process and its five states            ofdm_demodulator.cpp:235-577
//   CoordinatorThread / PipelineThread                 ofdm_demodulator.cpp:581-766
//   CalculateCyclicPhaseError / fine frequency update  ofdm_demodulator.cpp:768-840
//   CalculateDQPSK / CalculateViterbiBits              ofdm_demodulator.cpp:842-889, 58-72
//   apply_pll_avx + chebyshev sine                     dsp/apply_pll.cpp:82-116, dsp/chebyshev_sine.h:13-105
//   complex_conj_mul_sum                               dsp/complex_conj_mul_sum.cpp:65-99
//   CircularBuffer / ReconstructionBuffer / frame buf  circular_buffer.h, reconstruction_buffer.h, ofdm_frame_buffer.h
//   tables                                             dab_prs_ref.cpp:24-194, dab_mapper_ref.cpp:10-50
//
// Ordering is the canonical serialised one: when a frame completes, the control
// kernel stops consuming samples of that stream, the demod kernel processes the frame, and the next
// control pass applies the fine-frequency update before the next frame's PRS is looked at.
//
// Layout: the raw IQ of every stream stays where the host (or the caller's device buffer) put it, in a
// ring addressed by absolute sample index; frames are never copied, the demod kernel reads the ring
// directly (2 B/sample for u8 input) and writes only the 1 B/bit soft decisions.  The first symbol of a
// frame comes from a small per-stream "head" buffer because the reference takes it from its correlation
// buffer (which, while acquiring, is not always contiguous with the stream: circular_buffer.h quirk).
#pragma once
#include "common.cuh"

enum { OST_FINDING_NULL = 0, OST_READING_NULL_PRS = 1, OST_COARSE = 2, OST_FINE_TIME = 3, OST_READING_SYMBOLS = 4 };

#define OFDM_PAD(a) ((a) + (((a) >> 5) << 2))
#define OFDM_L1_BATCH 512

struct OfdmStream {
    int32_t state, frames_read, frames_desync, coarse_found;
    float coarse, fine;
    int32_t fine_time_offset, null_start_found, null_end_found;
    float l1_avg;
    uint32_t ring_index, ring_len, corr_len;
    uint32_t frame_fill, n_head;
    int32_t pending;          // 1 = a complete frame waits for / is being processed by k_ofdm_demod
    float frame_f;            // coarse+fine sampled when the frame was handed over (ofdm_demodulator.cpp:672)
    int32_t block_size;
    int32_t produced_last;    // frames emitted during the last process/advance call
    int32_t pad0;
    unsigned long long consumed, push_end, block_end, frame_ring_base;
};

struct OfdmGeom {
    int L, Tsym, Tnull, CP, N, K, frame_bits, frame_samples, sym_per_chunk, n_chunks;
};

struct OfdmDev {
    OfdmGeom g;
    dabgpu_ofdm_config cfg;
    int iq_format;
    // input ring
    const uint8_t* ring;              // base of stream 0
    unsigned long long ring_stride;   // samples between streams
    unsigned long long ring_mask;     // capacity-1 when the capacity is a power of two, else all ones (no wrap)
    // per-stream buffers
    OfdmStream* st;
    float2* null_ring;   // [stream][Tnull]
    float2* corr;        // [stream][Tnull+Tsym]
    float2* head;        // [stream][Tsym+CP]
    float* phase_err;    // [stream][L]
    // tables
    const float2* tw;            // exp(-2*pi*i*n/N)
    const float2* prs_fft_conj;  // conj(PRS spectrum), natural order
    const float2* prs_time_ref;  // conj(IFFT(relative phase of PRS)), natural order
    const uint16_t* dpos;        // padded smem position of natural index k after the in-place DIF FFT
    const uint16_t* outpos;      // padded smem position of the carrier that feeds output bit i (frequency de-interleaver)
    // outputs
    int8_t* frames;              // [stream][slot][frame_bits]
    uint32_t* frames_written;    // [stream]
    dabgpu_frame_info* frame_info;   // [stream][slot]
    uint32_t slot_mask;
    unsigned long long* counters;
};

__device__ __forceinline__ float2 cmulf(const float2 a, const float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__device__ __forceinline__ float2 ofdm_ring_sample(const OfdmDev& D, const int s, const unsigned long long abs_idx) {
    const unsigned long long off = (unsigned long long)s * D.ring_stride + (abs_idx & D.ring_mask);
    if (D.iq_format == DABGPU_IQ_U8) {
        // QuantisedIQToFloatIQ<uint8_t>: (u8 - 127.5) * (1/127.5)   (app_iq_readers.h:23-43, 72-87)
        const uchar2 v = __ldg(reinterpret_cast<const uchar2*>(D.ring) + off);
        return make_float2(__fmul_rn(__fsub_rn(float(v.x), 127.5f), 1.0f / 127.5f), __fmul_rn(__fsub_rn(float(v.y), 127.5f), 1.0f / 127.5f));
    }
    return __ldg(reinterpret_cast<const float2*>(D.ring) + off);
}

// sin(2*pi*x) on [-0.5, 0.5], same polynomial and (AVX+FMA build) evaluation order as chebyshev_sine.h:79-105
__device__ __forceinline__ float cheb_sin(const float x) {
    const float z = __fmul_rn(x, x);
    float b = 3.20396066f;
    b = __fmaf_rn(b, z, -14.07150173f);
    b = __fmaf_rn(b, z, 38.50016403f);
    b = __fmaf_rn(b, z, -67.07687378f);
    b = __fmaf_rn(b, z, 64.83583069f);
    b = __fmaf_rn(b, z, -25.13274193f);
    return __fmul_rn(__fmul_rn(x, b), __fadd_rn(z, -0.25f));
}

// apply_pll_avx (dsp/apply_pll.cpp:82-116): y = x * exp(j*2*pi*(dt0 + n*f)), float32 phase arithmetic reproduced
// operation by operation: base = fma(float(n & ~3), f, dt0); t = base + (float(n&3)*f [+0.25]); t -= rint(t)
__device__ __forceinline__ float2 pll_rotate(const float2 v, const int n, const float f, const float dt0) {
    const float base = __fmaf_rn(float(n & ~3), f, dt0);
    const float kf = __fmul_rn(float(n & 3), f);
    float ts = __fadd_rn(base, kf);
    float tc = __fadd_rn(base, __fadd_rn(kf, 0.25f));
    ts = __fsub_rn(ts, rintf(ts));
    tc = __fsub_rn(tc, rintf(tc));
    const float sn = cheb_sin(ts), cs = cheb_sin(tc);
    return make_float2(__fmaf_rn(cs, v.x, -__fmul_rn(sn, v.y)), __fmaf_rn(cs, v.y, __fmul_rn(sn, v.x)));
}

// ---------------------------------------------------------------------------------------------
// CTA-wide in-place decimation-in-frequency FFT in shared memory, N/8 threads, radix 8 (+ final 4 or 2).
// Stage-1 inputs come in registers: x[j] = input[tid + j*N/8].  Output is digit reversed; D.dpos maps a
// natural index to its (padded) position.  Arrays are split re/im and padded (OFDM_PAD) so that the
// strided stage accesses hit distinct banks.
// ---------------------------------------------------------------------------------------------
template <bool INV> __device__ __forceinline__ float2 mul_neg_i(const float2 t) {   // t * (-i) forward, t * (+i) inverse
    return INV ? make_float2(-t.y, t.x) : make_float2(t.y, -t.x);
}

template <bool INV> __device__ __forceinline__ void dft4(float2& d0, float2& d1, float2& d2, float2& d3) {
    const float2 s02 = make_float2(d0.x + d2.x, d0.y + d2.y), m02 = make_float2(d0.x - d2.x, d0.y - d2.y);
    const float2 s13 = make_float2(d1.x + d3.x, d1.y + d3.y), m13 = mul_neg_i<INV>(make_float2(d1.x - d3.x, d1.y - d3.y));
    d0 = make_float2(s02.x + s13.x, s02.y + s13.y);
    d2 = make_float2(s02.x - s13.x, s02.y - s13.y);
    d1 = make_float2(m02.x + m13.x, m02.y + m13.y);
    d3 = make_float2(m02.x - m13.x, m02.y - m13.y);
}

template <bool INV> __device__ __forceinline__ void dft8(float2 (&a)[8]) {
    const float h = 0.70710678118654752440f;
    float2 b0 = make_float2(a[0].x + a[4].x, a[0].y + a[4].y), c0 = make_float2(a[0].x - a[4].x, a[0].y - a[4].y);
    float2 b1 = make_float2(a[1].x + a[5].x, a[1].y + a[5].y), c1 = make_float2(a[1].x - a[5].x, a[1].y - a[5].y);
    float2 b2 = make_float2(a[2].x + a[6].x, a[2].y + a[6].y), c2 = make_float2(a[2].x - a[6].x, a[2].y - a[6].y);
    float2 b3 = make_float2(a[3].x + a[7].x, a[3].y + a[7].y), c3 = make_float2(a[3].x - a[7].x, a[3].y - a[7].y);
    // c_j *= w^j, w = exp(-+ 2*pi*i/8)
    if (!INV) {
        c1 = make_float2((c1.x + c1.y) * h, (c1.y - c1.x) * h);
        c2 = make_float2(c2.y, -c2.x);
        c3 = make_float2((c3.y - c3.x) * h, -(c3.x + c3.y) * h);
    } else {
        c1 = make_float2((c1.x - c1.y) * h, (c1.y + c1.x) * h);
        c2 = make_float2(-c2.y, c2.x);
        c3 = make_float2(-(c3.x + c3.y) * h, (c3.x - c3.y) * h);
    }
    dft4<INV>(b0, b1, b2, b3);
    dft4<INV>(c0, c1, c2, c3);
    a[0] = b0; a[2] = b1; a[4] = b2; a[6] = b3;
    a[1] = c0; a[3] = c1; a[5] = c2; a[7] = c3;
}

template <bool INV> __device__ __forceinline__ float2 ld_tw(const float2* __restrict__ tw, const int idx) {
    float2 w = __ldg(tw + idx);
    if (INV) w.y = -w.y;
    return w;
}

template <int N, bool INV>
__device__ __forceinline__ void fft_cta(float2 (&x)[8], float* __restrict__ re, float* __restrict__ im, const float2* __restrict__ tw, const int tid) {
    constexpr int NT = N / 8;
    dft8<INV>(x);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        float2 v = x[k];
        if (k > 0) v = cmulf(v, ld_tw<INV>(tw, tid * k));
        const int a = OFDM_PAD(tid + k * NT);
        re[a] = v.x; im[a] = v.y;
    }
#pragma unroll
    for (int M = N / 8; M >= 8; M /= 8) {
        __syncthreads();
        const int m8 = M / 8;
        const int q = tid / m8, r = tid - q * m8;
        const int base = q * M + r;
        float2 y[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { const int a = OFDM_PAD(base + j * m8); y[j] = make_float2(re[a], im[a]); }
        dft8<INV>(y);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float2 v = y[k];
            if (k > 0 && M > 8) v = cmulf(v, ld_tw<INV>(tw, r * k * (N / M)));
            const int a = OFDM_PAD(base + k * m8);
            re[a] = v.x; im[a] = v.y;
        }
    }
    constexpr int RF = (N == 2048 || N == 256) ? 4 : (N == 1024 ? 2 : 1);
    if (RF == 4) {
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int g = tid + h * NT;
            float2 d[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { const int a = OFDM_PAD(4 * g + j); d[j] = make_float2(re[a], im[a]); }
            dft4<INV>(d[0], d[1], d[2], d[3]);
#pragma unroll
            for (int j = 0; j < 4; j++) { const int a = OFDM_PAD(4 * g + j); re[a] = d[j].x; im[a] = d[j].y; }
        }
    } else if (RF == 2) {
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 4; h++) {
            const int p = tid + h * NT;
            const int a0 = OFDM_PAD(2 * p), a1 = OFDM_PAD(2 * p + 1);
            const float2 u = make_float2(re[a0], im[a0]), v = make_float2(re[a1], im[a1]);
            re[a0] = u.x + v.x; im[a0] = u.y + v.y;
            re[a1] = u.x - v.x; im[a1] = u.y - v.y;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Frame demodulation kernel.  grid = (chunks, streams), block = N/8 threads.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ofdm_frame_sample(const OfdmDev& D, const int s, const OfdmStream& st, const uint32_t k,
                                                    const float2* __restrict__ head) {
    if (k < st.n_head) return __ldg(head + k);
    return ofdm_ring_sample(D, s, st.frame_ring_base + (k - st.n_head));
}

template <int N>
__global__ void __launch_bounds__(N / 8, (N == 2048) ? 3 : 4)
k_ofdm_demod(const OfdmDev D, const int first_stream) {
    constexpr int NT = N / 8;
    constexpr int NP = N + N / 8;      // padded length
    __shared__ float s_re[2][NP];
    __shared__ float s_im[2][NP];
    __shared__ float2 s_cp[N / 4];
    __shared__ float2 s_part[NT / 32 > 0 ? NT / 32 : 1];

    const int s = first_stream + blockIdx.y;
    const OfdmStream st = D.st[s];
    if (st.pending != 1) return;
    const OfdmGeom& g = D.g;
    const int tid = threadIdx.x;
    const int l0 = blockIdx.x * g.sym_per_chunk;
    if (l0 >= g.L - 1) return;
    const int l1 = min(l0 + g.sym_per_chunk, g.L - 1);
    const float f = st.frame_f;
    const float2* head = D.head + size_t(s) * (g.Tsym + g.CP);
    const uint32_t slot = D.frames_written[s] & D.slot_mask;
    int8_t* out_frame = D.frames + (size_t(s) * (D.slot_mask + 1u) + slot) * g.frame_bits;
    const int CP = g.CP, K = g.K;

    for (int l = l0; l <= l1; l++) {
        float* re = s_re[l & 1];
        float* im = s_im[l & 1];
        const bool own = (l < l1) || (l == g.L - 1);   // this CTA accounts for the symbol's cyclic-prefix phase
        const float dt0 = __fmul_rn(float(l * g.Tsym), f);
        const uint32_t sym_base = uint32_t(l) * uint32_t(g.Tsym);
        float2 x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int n = CP + tid + j * NT;
            x[j] = pll_rotate(ofdm_frame_sample(D, s, st, sym_base + n, head), n, f, dt0);
        }
        __syncthreads();   // previous iteration's DQPSK reads of buffer (l&1) and of s_cp are complete
        if (own) {
#pragma unroll
            for (int m = 0; m < 2; m++) {
                const int n = tid + m * NT;
                if (n < CP) s_cp[n] = pll_rotate(ofdm_frame_sample(D, s, st, sym_base + n, head), n, f, dt0);
            }
            __syncthreads();
            // sum over n < CP of sym[N+n] * conj(sym[n])   (ofdm_demodulator.cpp:768-777)
            float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int b = tid + j * NT;
                if (b >= N - CP) {
                    const float2 c = s_cp[b - (N - CP)];
                    acc.x += __fmaf_rn(x[j].y, c.y, x[j].x * c.x);
                    acc.y += __fmaf_rn(x[j].y, c.x, -(x[j].x * c.y));
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc.x += __shfl_xor_sync(FULL_MASK, acc.x, o);
                acc.y += __shfl_xor_sync(FULL_MASK, acc.y, o);
            }
            if ((tid & 31) == 0) s_part[tid >> 5] = acc;
        }
        fft_cta<N, false>(x, re, im, D.tw, tid);
        __syncthreads();
        if (own && tid == 0) {
            float2 t = make_float2(0.0f, 0.0f);
            for (int w = 0; w < NT / 32; w++) { t.x += s_part[w].x; t.y += s_part[w].y; }
            D.phase_err[size_t(s) * g.L + l] = atan2f(t.y, t.x);
        }
        if (l > l0) {
            // DQPSK X_{l-1} * conj(X_l), frequency de-interleave, L-infinity normalise, truncate to int8
            // (ofdm_demodulator.cpp:842-889; bit0 = trunc(-127*re/A), bit1 = trunc(+127*im/A))
            const float* pre = s_re[(l - 1) & 1];
            const float* pim = s_im[(l - 1) & 1];
            int8_t* row = out_frame + size_t(l - 1) * 2u * K;
            for (int u = tid; u < K / 4; u += NT) {
                const ushort4 p4 = __ldg(reinterpret_cast<const ushort4*>(D.outpos) + u);
                const unsigned short pp[4] = {p4.x, p4.y, p4.z, p4.w};
                uint32_t w0 = 0, w1 = 0;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int p = pp[e];
                    const float ar = pre[p], ai = pim[p], br = re[p], bi = im[p];
                    const float vr = ar * br + ai * bi;
                    const float vi = ai * br - ar * bi;
                    const float A = fmaxf(fabsf(vr), fabsf(vi));
                    const int q0 = int(__fmul_rn(__fdiv_rn(vr, A), -127.0f));
                    const int q1 = int(__fmul_rn(__fdiv_rn(vi, A), 127.0f));
                    w0 |= uint32_t(uint8_t(q0)) << (8 * e);
                    w1 |= uint32_t(uint8_t(q1)) << (8 * e);
                }
                *reinterpret_cast<uint32_t*>(row + 4 * u) = w0;
                *reinterpret_cast<uint32_t*>(row + K + 4 * u) = w1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Control kernel: one CTA (N/8 threads) per stream walks the reference state machine over the
// newly available samples; heavy per-frame work is left to k_ofdm_demod.
// ---------------------------------------------------------------------------------------------
struct CtlShared {
    OfdmStream st;
    float l1[OFDM_L1_BATCH];
    float red_v[32];
    int red_i[32];
    float red_s[32];
    int flag;
    int nb_read;
};

__device__ __forceinline__ void ctl_update_fine(OfdmStream& st, const int N, const float delta) {   // ofdm_demodulator.cpp:829-840
    const float wrap = 0.5f * (1.0f / float(N)) * 1.01f;
    st.fine = fmodf(st.fine + delta, wrap);
}

__device__ __forceinline__ void ctl_reset(OfdmStream& st) {   // OFDM_Demod::Reset, ofdm_demodulator.cpp:277-289
    st.state = OST_FINDING_NULL;
    st.corr_len = 0;
    st.frames_desync++;
    st.coarse_found = 0;
    st.coarse = 0.0f;
    st.fine = 0.0f;
    st.fine_time_offset = 0;
}

// mean of |re|+|im| over `count` samples starting at abs index a0, computed by one warp (CalculateL1Average :922-932)
__device__ __forceinline__ float ctl_l1_window(const OfdmDev& D, const int s, const unsigned long long a0, const int count, const int lane) {
    float acc = 0.0f;
    for (int i = lane; i < count; i += 32) {
        const float2 v = ofdm_ring_sample(D, s, a0 + i);
        acc += fabsf(v.x) + fabsf(v.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, o);
    return acc / float(count);
}

template <int N>
__global__ void __launch_bounds__(N / 8)
k_ofdm_ctl(const OfdmDev D, const int first_stream, const int n_new_samples, const int block_size, const int is_first) {
    constexpr int NT = N / 8;
    constexpr int NP = N + N / 8;
    constexpr int NW = NT / 32;
    __shared__ float a_re[NP], a_im[NP], b_re[NP], b_im[NP];
    __shared__ CtlShared sh;
    const int s = first_stream + blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const OfdmGeom& g = D.g;
    OfdmStream& st = sh.st;
    if (tid == 0) {
        st = D.st[s];
        if (is_first) {
            st.push_end += (unsigned long long)n_new_samples;
            st.block_size = block_size > 0 ? block_size : n_new_samples;
            st.produced_last = 0;
        }
        if (st.pending == 1) {
            // coordinator (ofdm_demodulator.cpp:608-635): average CP phase -> fine frequency, then emit the frame
            const float* pe = D.phase_err + size_t(s) * g.L;
            float total = 0.0f;
            for (int l = 0; l < g.L; l++) total += pe[l];
            const float avg = total / float(g.L);
            const float err = (1.0f / float(N)) * avg / (2.0f * 3.14159265358979323846f);
            ctl_update_fine(st, N, -D.cfg.fine_freq_update_beta * err);
            st.frames_read++;
            const uint32_t w = D.frames_written[s];
            dabgpu_frame_info fi;
            fi.freq_coarse_offset = st.coarse; fi.freq_fine_offset = st.fine; fi.fine_time_offset = st.fine_time_offset; fi.frame_index = int(w);
            D.frame_info[size_t(s) * (D.slot_mask + 1u) + (w & D.slot_mask)] = fi;
            D.frames_written[s] = w + 1u;
            st.produced_last++;
            st.pending = 0;
            atomicAdd(&D.counters[0], 1ull);   // CNT_FRAMES_DEMOD
        }
    }
    float2* null_ring = D.null_ring + size_t(s) * g.Tnull;
    float2* corr = D.corr + size_t(s) * (g.Tnull + g.Tsym);
    float2* head = D.head + size_t(s) * (g.Tsym + g.CP);
    const int K1 = D.cfg.signal_l1_nb_samples;

    for (;;) {
        __syncthreads();
        if (st.pending) break;
        if (st.consumed == st.block_end) {
            if (st.consumed == st.push_end) break;
            // start of the next Process() call: UpdateSignalAverage (ofdm_demodulator.cpp:934-950)
            const unsigned long long bstart = st.consumed;
            const unsigned long long remaining = st.push_end - bstart;
            const int blen = int(remaining < (unsigned long long)st.block_size ? remaining : (unsigned long long)st.block_size);
            __syncthreads();
            if (blen >= K1) {
                const int M = blen - K1, Lstep = K1 * D.cfg.signal_l1_nb_decimate;
                const int n_win = (M + Lstep - 1) / Lstep;
                for (int w0 = 0; w0 < n_win; w0 += OFDM_L1_BATCH) {
                    const int nb = min(OFDM_L1_BATCH, n_win - w0);
                    for (int w = warp; w < nb; w += NW) {
                        const float v = ctl_l1_window(D, s, bstart + (unsigned long long)(w0 + w) * Lstep, K1, lane);
                        if (lane == 0) sh.l1[w] = v;
                    }
                    __syncthreads();
                    if (tid == 0) {
                        const float beta = D.cfg.signal_l1_update_beta;
                        float a = st.l1_avg;
                        for (int w = 0; w < nb; w++) a = beta * a + (1.0f - beta) * sh.l1[w];
                        st.l1_avg = a;
                    }
                    __syncthreads();
                }
            }
            if (tid == 0) st.block_end = bstart + (unsigned long long)blen;
            __syncthreads();
        }
        const int rem = int(st.block_end - st.consumed);
        const int state = st.state;

        if (state == OST_FINDING_NULL) {
            // FindNullPowerDip (ofdm_demodulator.cpp:291-347)
            const int M = rem - K1;
            const int n_win = M > 0 ? (M + K1 - 1) / K1 : 0;
            const float t0 = st.l1_avg * D.cfg.null_thresh_start, t1 = st.l1_avg * D.cfg.null_thresh_end;
            if (tid == 0) { sh.nb_read = rem; sh.flag = 0; }
            __syncthreads();
            for (int w0 = 0; w0 < n_win; w0 += OFDM_L1_BATCH) {
                const int nb = min(OFDM_L1_BATCH, n_win - w0);
                for (int w = warp; w < nb; w += NW) {
                    const float v = ctl_l1_window(D, s, st.consumed + (unsigned long long)(w0 + w) * K1, K1, lane);
                    if (lane == 0) sh.l1[w] = v;
                }
                __syncthreads();
                if (tid == 0) {
                    for (int w = 0; w < nb; w++) {
                        if (st.null_start_found) {
                            if (sh.l1[w] > t1) { st.null_end_found = 1; sh.nb_read = (w0 + w) * K1 + K1; sh.flag = 1; break; }
                        } else if (sh.l1[w] < t0) {
                            st.null_start_found = 1;
                        }
                    }
                }
                __syncthreads();
                if (sh.flag) break;
            }
            const int nb_read = sh.nb_read;
            // CircularBuffer::ConsumeBuffer(read_all = true): only the last Tnull samples survive
            const int cap = g.Tnull;
            const int first = nb_read > cap ? nb_read - cap : 0;
            for (int i = first + tid; i < nb_read; i += NT)
                null_ring[(st.ring_index + uint32_t(i)) % uint32_t(cap)] = ofdm_ring_sample(D, s, st.consumed + i);
            __syncthreads();
            if (tid == 0) {
                st.ring_index = (st.ring_index + uint32_t(nb_read)) % uint32_t(cap);
                st.ring_len = min(st.ring_len + uint32_t(nb_read), uint32_t(cap));
                st.consumed += (unsigned long long)nb_read;
            }
            __syncthreads();
            if (st.null_end_found) {
                const uint32_t Lr = st.ring_len, start = st.ring_index;
                for (uint32_t i = tid; i < Lr; i += NT) corr[i] = null_ring[(i + start) % uint32_t(cap)];
                __syncthreads();
                if (tid == 0) {
                    st.null_start_found = 0; st.null_end_found = 0;
                    st.corr_len = Lr; st.ring_len = 0;
                    st.state = OST_READING_NULL_PRS;
                }
            }
        } else if (state == OST_READING_NULL_PRS) {
            const int cap = g.Tnull + g.Tsym;
            const int need = cap - int(st.corr_len);
            const int take = rem < need ? rem : need;
            for (int i = tid; i < take; i += NT) corr[st.corr_len + i] = ofdm_ring_sample(D, s, st.consumed + i);
            __syncthreads();
            if (tid == 0) {
                st.corr_len += uint32_t(take);
                st.consumed += (unsigned long long)take;
                if (int(st.corr_len) == cap) st.state = OST_COARSE;
            }
        } else if (state == OST_COARSE) {
            // RunCoarseFreqSync (ofdm_demodulator.cpp:360-471)
            if (!D.cfg.is_coarse_freq_correction) {
                if (tid == 0) { st.coarse = 0.0f; st.state = OST_FINE_TIME; }
                continue;
            }
            const float2* prs = corr + g.Tnull;
            float2 x[8];
#pragma unroll
            for (int j = 0; j < 8; j++) x[j] = prs[tid + j * NT];
            fft_cta<N, false>(x, a_re, a_im, D.tw, tid);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 8; j++) {   // CalculateRelativePhase: conj(X[i]) * X[i+1]
                const int i = tid + j * NT;
                float2 z = make_float2(0.0f, 0.0f);
                if (i < N - 1) {
                    const int p0 = D.dpos[i], p1 = D.dpos[i + 1];
                    const float2 u = make_float2(a_re[p0], a_im[p0]), v = make_float2(a_re[p1], a_im[p1]);
                    z = make_float2(u.x * v.x + u.y * v.y, u.x * v.y - u.y * v.x);
                }
                x[j] = z;
            }
            fft_cta<N, true>(x, b_re, b_im, D.tw, tid);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int i = tid + j * NT;
                const int p = D.dpos[i];
                x[j] = cmulf(make_float2(b_re[p], b_im[p]), __ldg(D.prs_time_ref + i));
            }
            __syncthreads();
            fft_cta<N, false>(x, a_re, a_im, D.tw, tid);
            __syncthreads();
            // CalculateMagnitude: resp[i] = 20*log10(|F[(i + N/2) % N]|), kept in b_re (natural order, unpadded)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int i = tid + j * NT;
                const int p = D.dpos[(i + N / 2) % N];
                b_re[i] = 20.0f * log10f(sqrtf(a_re[p] * a_re[p] + a_im[p] * a_im[p]));
            }
            __syncthreads();
            const int Mh = N / 2;
            int maxoff = int(D.cfg.max_coarse_freq_correction_norm * float(N));
            maxoff = max(0, min(maxoff, Mh));
            // first maximum over i in [-maxoff, maxoff] (fft index i+Mh, skipping index N)
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int i = -maxoff + tid; i <= maxoff; i += NT) {
                const int k = i + Mh;
                if (k == N) continue;
                const float v = b_re[k];
                if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(FULL_MASK, bv, o);
                const int oi = __shfl_xor_sync(FULL_MASK, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { sh.red_v[warp] = bv; sh.red_i[warp] = bi; }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < NW; w++)
                    if (sh.red_v[w] > bv || (sh.red_v[w] == bv && sh.red_i[w] < bi)) { bv = sh.red_v[w]; bi = sh.red_i[w]; }
                // the reference starts from max_value = resp[-maxoff+Mh], index -maxoff and only replaces on '>'
                int max_index = bi;
                if (!(bv > b_re[-maxoff + Mh])) max_index = -maxoff;
                float mag[3]; int pk[3];
                for (int j = 0; j < 3; j++) {
                    int idx = max_index - 1 + j;
                    idx = max(-maxoff, min(idx, maxoff));
                    int k = idx + Mh;
                    if (k >= N) k = N - 1;
                    mag[j] = powf(10.0f, b_re[k] / 20.0f);
                    pk[j] = k - Mh;
                }
                const float sum = mag[0] + mag[1] + mag[2];
                float lerp = 0.0f;
                for (int j = 0; j < 3; j++) lerp += float(pk[j]) * mag[j] / sum;
                const float predicted = -lerp / float(N);
                const float error = predicted - st.coarse;
                const bool large = fabsf(error) > 1.5f / float(N);
                const float beta = (large || !st.coarse_found) ? 1.0f : D.cfg.coarse_freq_slow_beta;
                const float delta = beta * error;
                st.coarse += delta;
                st.coarse_found = 1;
                ctl_update_fine(st, N, -delta);
                st.state = OST_FINE_TIME;
            }
        } else if (state == OST_FINE_TIME) {
            // RunFineTimeSync (ofdm_demodulator.cpp:473-548)
            const float2* prs = corr + g.Tnull;
            const float f = st.coarse + st.fine;
            float2 x[8];
#pragma unroll
            for (int j = 0; j < 8; j++) { const int n = tid + j * NT; x[j] = pll_rotate(prs[n], n, f, 0.0f); }
            fft_cta<N, false>(x, a_re, a_im, D.tw, tid);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int i = tid + j * NT;
                const int p = D.dpos[i];
                x[j] = cmulf(make_float2(a_re[p], a_im[p]), __ldg(D.prs_fft_conj + i));
            }
            fft_cta<N, true>(x, b_re, b_im, D.tw, tid);
            __syncthreads();
            // impulse response in dB, natural order in a_re (unpadded)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int i = tid + j * NT;
                const int p = D.dpos[i];
                a_re[i] = 20.0f * log10f(sqrtf(b_re[p] * b_re[p] + b_im[p] * b_im[p]));
            }
            __syncthreads();
            const float decay = 1.0f - D.cfg.impulse_peak_distance_probability;
            float bv = -INFINITY; int bi = 0x7fffffff; float sum = 0.0f;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int i = tid + j * NT;
                const float v = a_re[i];
                const float nd = float(abs(g.CP - i)) / float(g.Tsym);
                const float wv = (1.0f - decay * nd) * v;
                sum += v;
                if (wv > bv || (wv == bv && i < bi)) { bv = wv; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(FULL_MASK, bv, o);
                const int oi = __shfl_xor_sync(FULL_MASK, bi, o);
                sum += __shfl_xor_sync(FULL_MASK, sum, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { sh.red_v[warp] = bv; sh.red_i[warp] = bi; sh.red_s[warp] = sum; }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < NW; w++) {
                    sum += sh.red_s[w];
                    if (sh.red_v[w] > bv || (sh.red_v[w] == bv && sh.red_i[w] < bi)) { bv = sh.red_v[w]; bi = sh.red_i[w]; }
                }
                // sequential scan semantics: starts from (impulse[0], index 0), replaced only by a strictly larger weighted value
                float max_value = a_re[0]; int max_index = 0;
                if (bv > max_value) { max_value = bv; max_index = bi; }
                const float avg = sum / float(N);
                if ((max_value - avg) < D.cfg.impulse_peak_threshold_db) {
                    ctl_reset(st);
                    sh.flag = 0;
                } else {
                    const int offset = max_index - g.CP;
                    st.n_head = uint32_t(g.Tsym - offset);
                    st.fine_time_offset = offset;
                    sh.flag = 1;
                }
            }
            __syncthreads();
            if (sh.flag) {
                const int offset = st.fine_time_offset;
                const int start = g.Tnull + offset;
                for (uint32_t i = tid; i < st.n_head; i += NT) head[i] = corr[start + i];
                __syncthreads();
                if (tid == 0) {
                    st.frame_fill = st.n_head;
                    st.frame_ring_base = st.consumed;
                    st.corr_len = 0;
                    st.state = OST_READING_SYMBOLS;
                }
            }
        } else {   // OST_READING_SYMBOLS (ofdm_demodulator.cpp:550-577): nothing is copied, the frame stays in the ring
            const int need = g.frame_samples - int(st.frame_fill);
            const int take = rem < need ? rem : need;
            if (take == need) {
                // the NULL symbol at the end of the frame seeds the correlation buffer of the next frame
                const unsigned long long null_abs = st.frame_ring_base + (unsigned long long)(g.L * g.Tsym) - st.n_head;
                for (int i = tid; i < g.Tnull; i += NT) corr[i] = ofdm_ring_sample(D, s, null_abs + i);
            }
            __syncthreads();
            if (tid == 0) {
                st.frame_fill += uint32_t(take);
                st.consumed += (unsigned long long)take;
                if (take == need) {
                    st.corr_len = uint32_t(g.Tnull);
                    st.state = OST_READING_NULL_PRS;
                    st.frame_f = st.coarse + st.fine;
                    st.pending = 1;
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) D.st[s] = st;
}

// Packs the newest frame of every stream that produced one in the last call into a contiguous staging buffer.
__global__ void k_ofdm_gather_latest(const OfdmDev D, const int first_stream, int8_t* __restrict__ stage, uint8_t* __restrict__ produced) {
    const int s = first_stream + blockIdx.y;
    const OfdmStream& st = D.st[s];
    const bool has = st.produced_last > 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) produced[blockIdx.y] = has ? 1 : 0;
    if (!has) return;
    const uint32_t slot = (D.frames_written[s] - 1u) & D.slot_mask;
    const uint4* src = reinterpret_cast<const uint4*>(D.frames + (size_t(s) * (D.slot_mask + 1u) + slot) * D.g.frame_bits);
    uint4* dst = reinterpret_cast<uint4*>(stage + size_t(blockIdx.y) * D.g.frame_bits);
    const int n16 = D.g.frame_bits / 16;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

__global__ void k_ofdm_reset(const OfdmDev D, const int stream, const int n_streams, const int full) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_streams) return;
    const int s = (stream < 0) ? i : stream;
    OfdmStream st = D.st[s];
    if (full) {
        memset(&st, 0, sizeof(st));
    } else {
        ctl_reset(st);
        st.pending = 0;
    }
    D.st[s] = st;
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
struct OfdmState {
    uint64_t launches = 0;
    Profiler* prof = nullptr;
    OfdmDev dev;
    dabgpu_params P;
    int max_streams = 0, frame_slots = 0;
    size_t ring_cap = 0;          // internal ring capacity (samples)
    bool external_ring = false;
    int bps = 2;
    DevBuf d_ring, d_st, d_null_ring, d_corr, d_head, d_phase, d_tw, d_prs_conj, d_prs_time, d_dpos, d_outpos, d_stage, d_produced;
    PinnedBuf h_produced;
    std::vector<unsigned long long> h_written;   // absolute samples written per stream (internal ring)
};

// PRS phases, EN 300 401 clause 14.3.2 tables 23/24 (reference: dab_prs_ref.cpp:24-194).
// per block of 32 carriers: (i << 2) | n ; negative carriers first
static const uint8_t kPrsH[4][32] = {
    {0, 2, 0, 0, 0, 0, 1, 1, 2, 0, 0, 0, 2, 2, 1, 1, 0, 2, 0, 0, 0, 0, 1, 1, 2, 0, 0, 0, 2, 2, 1, 1},
    {0, 3, 2, 3, 0, 1, 3, 0, 2, 1, 2, 3, 2, 3, 3, 0, 0, 3, 2, 3, 0, 1, 3, 0, 2, 1, 2, 3, 2, 3, 3, 0},
    {0, 0, 0, 2, 0, 2, 1, 3, 2, 2, 0, 2, 2, 0, 1, 3, 0, 0, 0, 2, 0, 2, 1, 3, 2, 2, 0, 2, 2, 0, 1, 3},
    {0, 1, 2, 1, 0, 3, 3, 2, 2, 3, 2, 1, 2, 1, 3, 2, 0, 1, 2, 1, 0, 3, 3, 2, 2, 3, 2, 1, 2, 1, 3, 2},
};
static const uint8_t kPrsBlocks1[48] = {1, 6, 8, 13, 3, 6, 10, 15, 2, 5, 10, 15, 1, 6, 11, 15, 2, 6, 10, 13, 1, 7, 9, 14,
                                        3, 13, 9, 5, 2, 14, 9, 4, 2, 14, 11, 7, 0, 14, 9, 7, 3, 15, 11, 4, 3, 12, 9, 5};
static const uint8_t kPrsBlocks2[12] = {2, 7, 10, 14, 1, 6, 8, 6, 2, 13, 8, 7};
static const uint8_t kPrsBlocks3[6] = {2, 7, 8, 14, 10, 6};
static const uint8_t kPrsBlocks4[24] = {0, 5, 9, 14, 2, 6, 8, 15, 3, 5, 11, 14, 0, 13, 8, 6, 0, 13, 10, 6, 2, 13, 11, 4};

static void host_prs_spectrum(int mode, int N, int K, std::vector<float2>& prs) {
    const uint8_t* tab = (mode == 1) ? kPrsBlocks1 : (mode == 2) ? kPrsBlocks2 : (mode == 3) ? kPrsBlocks3 : kPrsBlocks4;
    prs.assign(size_t(N), make_float2(0.0f, 0.0f));
    for (int slot = 0; slot < K; slot++) {
        const int k = (slot < K / 2) ? (slot - K / 2) : (slot - K / 2 + 1);
        const int i = tab[slot / 32] >> 2, n = tab[slot / 32] & 3;
        const float phi = float(M_PI) / 2.0f * float(kPrsH[i][slot % 32] + n);
        prs[size_t(k < 0 ? N + k : k)] = make_float2(cosf(phi), sinf(phi));
    }
}

static void host_dft_double(std::vector<double>& re, std::vector<double>& im, bool inverse) {
    // simple recursive-free radix-2 on doubles for table generation only
    const size_t n = re.size();
    for (size_t i = 1, j = 0; i < n; i++) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = (inverse ? 2.0 : -2.0) * M_PI / double(len);
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; k++) {
                const double wr = cos(ang * double(k)), wi = sin(ang * double(k));
                const double ur = re[i + k], ui = im[i + k];
                const double vr = re[i + k + len / 2] * wr - im[i + k + len / 2] * wi;
                const double vi = re[i + k + len / 2] * wi + im[i + k + len / 2] * wr;
                re[i + k] = ur + vr; im[i + k] = ui + vi;
                re[i + k + len / 2] = ur - vr; im[i + k + len / 2] = ui - vi;
            }
    }
}

static int ofdm_init(OfdmState& O, const dabgpu_config& cfg, const dabgpu_params& P, int frame_slots, int8_t* d_frames,
                     uint32_t* d_frames_written, dabgpu_frame_info* d_frame_info, unsigned long long* d_counters) {
    O.P = P;
    O.max_streams = cfg.max_streams;
    O.frame_slots = frame_slots;
    O.bps = (cfg.iq_format == DABGPU_IQ_U8) ? 2 : 8;
    const int S = cfg.max_streams, N = P.nb_fft, K = P.nb_data_carriers;
    size_t cap = cfg.ring_samples;
    if (cap == 0) {
        cap = 1;
        while (cap < size_t(P.nb_frame_samples) * 2 + size_t(P.nb_null_period + P.nb_symbol_period)) cap <<= 1;
    }
    if (cap & (cap - 1)) return set_error(DABGPU_ERR_INVALID, "ring_samples must be a power of two");
    if (cap < size_t(P.nb_frame_samples) + size_t(P.nb_null_period + P.nb_symbol_period) + 4096)
        return set_error(DABGPU_ERR_INVALID, "ring_samples too small for one frame");
    O.ring_cap = cap;
    int rc;
    if ((rc = O.d_ring.alloc(size_t(S) * cap * O.bps))) return rc;
    if ((rc = O.d_st.alloc(size_t(S) * sizeof(OfdmStream)))) return rc;
    if ((rc = O.d_null_ring.alloc(size_t(S) * P.nb_null_period * sizeof(float2)))) return rc;
    if ((rc = O.d_corr.alloc(size_t(S) * (P.nb_null_period + P.nb_symbol_period) * sizeof(float2)))) return rc;
    if ((rc = O.d_head.alloc(size_t(S) * (P.nb_symbol_period + P.nb_cyclic_prefix) * sizeof(float2)))) return rc;
    if ((rc = O.d_phase.alloc(size_t(S) * P.nb_frame_symbols * sizeof(float)))) return rc;
    if ((rc = O.d_produced.alloc(size_t(S)))) return rc;
    if ((rc = O.h_produced.alloc(size_t(S)))) return rc;
    cudaMemset(O.d_ring.p, 0, O.d_ring.bytes);
    cudaMemset(O.d_st.p, 0, O.d_st.bytes);
    cudaMemset(O.d_null_ring.p, 0, O.d_null_ring.bytes);   // the reference's joint block is zero initialised (joint_allocate.h:21-23)
    cudaMemset(O.d_corr.p, 0, O.d_corr.bytes);
    cudaMemset(O.d_head.p, 0, O.d_head.bytes);
    cudaMemset(O.d_phase.p, 0, O.d_phase.bytes);
    O.h_written.assign(size_t(S), 0ull);

    // tables
    std::vector<float2> tw(static_cast<size_t>(N)), prs;
    for (int n = 0; n < N; n++) {
        const double a = -2.0 * M_PI * double(n) / double(N);
        tw[size_t(n)] = make_float2(float(cos(a)), float(sin(a)));
    }
    host_prs_spectrum(cfg.transmission_mode, N, K, prs);
    std::vector<float2> prs_conj(static_cast<size_t>(N)), prs_time(static_cast<size_t>(N));
    for (int i = 0; i < N; i++) prs_conj[size_t(i)] = make_float2(prs[size_t(i)].x, -prs[size_t(i)].y);
    {
        // conj(IFFT(conj(X[i]) * X[i+1]))   (ofdm_demodulator.cpp:136-140)
        std::vector<double> re(static_cast<size_t>(N), 0.0), im(static_cast<size_t>(N), 0.0);
        for (int i = 0; i < N - 1; i++) {
            const double ar = prs[size_t(i)].x, ai = prs[size_t(i)].y, br = prs[size_t(i + 1)].x, bi = prs[size_t(i + 1)].y;
            re[size_t(i)] = ar * br + ai * bi;
            im[size_t(i)] = ar * bi - ai * br;
        }
        host_dft_double(re, im, true);
        for (int i = 0; i < N; i++) prs_time[size_t(i)] = make_float2(float(re[size_t(i)]), float(-im[size_t(i)]));
    }
    // digit reversal of the in-place DIF (radices 8,8,8,4 / 8,8,8,2 / 8,8,8 / 8,8,4)
    std::vector<int> radices;
    switch (N) {
    case 2048: radices = {8, 8, 8, 4}; break;
    case 1024: radices = {8, 8, 8, 2}; break;
    case 512: radices = {8, 8, 8}; break;
    default: radices = {8, 8, 4}; break;
    }
    std::vector<uint16_t> dpos(static_cast<size_t>(N)), outpos(static_cast<size_t>(K));
    for (int k = 0; k < N; k++) {
        int a = 0, rem = N, kk = k;
        for (int R : radices) { const int d = kk % R; kk /= R; rem /= R; a += d * rem; }
        dpos[size_t(k)] = uint16_t(OFDM_PAD(a));
    }
    {
        // frequency de-interleaver (dab_mapper_ref.cpp:10-50) composed with the carrier -> FFT bin map (ofdm_demodulator.cpp:853-864)
        std::vector<int> cmap;
        int v = 0;
        const int dc = N / 2, lo = dc - K / 2, hi = dc + K / 2;
        for (int i = 0; i < N; i++) {
            if (i > 0) v = (13 * v + N / 4 - 1) % N;
            if (v < lo || v > hi || v == dc) continue;
            cmap.push_back(v < dc ? v - lo : v - lo - 1);
        }
        for (int i = 0; i < K; i++) {
            const int slot = cmap[size_t(i)];
            const int kf = (slot < K / 2) ? (slot - K / 2) : (slot - K / 2 + 1);
            outpos[size_t(i)] = dpos[size_t((N + kf) % N)];
        }
    }
    if ((rc = O.d_tw.alloc(tw.size() * sizeof(float2)))) return rc;
    if ((rc = O.d_prs_conj.alloc(prs_conj.size() * sizeof(float2)))) return rc;
    if ((rc = O.d_prs_time.alloc(prs_time.size() * sizeof(float2)))) return rc;
    if ((rc = O.d_dpos.alloc(dpos.size() * 2))) return rc;
    if ((rc = O.d_outpos.alloc(outpos.size() * 2))) return rc;
    cudaMemcpy(O.d_tw.p, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice);
    cudaMemcpy(O.d_prs_conj.p, prs_conj.data(), prs_conj.size() * sizeof(float2), cudaMemcpyHostToDevice);
    cudaMemcpy(O.d_prs_time.p, prs_time.data(), prs_time.size() * sizeof(float2), cudaMemcpyHostToDevice);
    cudaMemcpy(O.d_dpos.p, dpos.data(), dpos.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(O.d_outpos.p, outpos.data(), outpos.size() * 2, cudaMemcpyHostToDevice);

    OfdmDev& D = O.dev;
    D.g.L = P.nb_frame_symbols; D.g.Tsym = P.nb_symbol_period; D.g.Tnull = P.nb_null_period; D.g.CP = P.nb_cyclic_prefix;
    D.g.N = N; D.g.K = K; D.g.frame_bits = P.nb_frame_bits; D.g.frame_samples = P.nb_frame_samples;
    D.g.sym_per_chunk = (P.nb_frame_symbols == 153) ? 19 : 15;
    D.g.n_chunks = (P.nb_frame_symbols - 1 + D.g.sym_per_chunk - 1) / D.g.sym_per_chunk;
    D.cfg = cfg.ofdm;
    D.iq_format = cfg.iq_format;
    D.ring = O.d_ring.as<uint8_t>();
    D.ring_stride = cap;
    D.ring_mask = cap - 1;
    D.st = O.d_st.as<OfdmStream>();
    D.null_ring = O.d_null_ring.as<float2>();
    D.corr = O.d_corr.as<float2>();
    D.head = O.d_head.as<float2>();
    D.phase_err = O.d_phase.as<float>();
    D.tw = O.d_tw.as<float2>();
    D.prs_fft_conj = O.d_prs_conj.as<float2>();
    D.prs_time_ref = O.d_prs_time.as<float2>();
    D.dpos = O.d_dpos.as<uint16_t>();
    D.outpos = O.d_outpos.as<uint16_t>();
    D.frames = d_frames;
    D.frames_written = d_frames_written;
    D.frame_info = d_frame_info;
    D.slot_mask = uint32_t(frame_slots - 1);
    D.counters = d_counters;
    return DABGPU_OK;
}

static void ofdm_destroy(OfdmState& O) {
    DevBuf* bufs[] = {&O.d_ring, &O.d_st, &O.d_null_ring, &O.d_corr, &O.d_head, &O.d_phase, &O.d_tw, &O.d_prs_conj, &O.d_prs_time,
                      &O.d_dpos, &O.d_outpos, &O.d_stage, &O.d_produced};
    for (DevBuf* b : bufs) b->release();
    O.h_produced.release();
}

static int ofdm_reset(OfdmState& O, int stream, cudaStream_t cs) {
    const int n = (stream < 0) ? O.max_streams : 1;
    k_ofdm_reset<<<(n + 127) / 128, 128, 0, cs>>>(O.dev, stream, n, 0);
    O.launches++;
    CUDA_TRY(cudaGetLastError());
    return DABGPU_OK;
}

template <int N>
static int ofdm_run_t(OfdmState& O, int first, int n, int n_samples, int block_size, cudaStream_t cs) {
    const int max_frames = n_samples / O.P.nb_frame_samples + 2;
    const dim3 dgrid(O.dev.g.n_chunks, n);
    Profiler& pf = *O.prof;
    for (int it = 0; it < max_frames; it++) {
        pf.begin(PROF_OFDM_CTL, cs);
        k_ofdm_ctl<N><<<n, N / 8, 0, cs>>>(O.dev, first, n_samples, block_size, it == 0 ? 1 : 0);
        pf.end(cs);
        pf.begin(PROF_OFDM_DEMOD, cs);
        k_ofdm_demod<N><<<dgrid, N / 8, 0, cs>>>(O.dev, first);
        pf.end(cs);
        O.launches += 2;
    }
    pf.begin(PROF_OFDM_CTL, cs);
    k_ofdm_ctl<N><<<n, N / 8, 0, cs>>>(O.dev, first, n_samples, block_size, 0);
    pf.end(cs);
    O.launches++;
    CUDA_TRY(cudaGetLastError());
    return DABGPU_OK;
}

static int ofdm_run(OfdmState& O, int first, int n, int n_samples, int block_size, cudaStream_t cs) {
    if (n == 0 || n_samples == 0) return DABGPU_OK;
    switch (O.P.nb_fft) {
    case 2048: return ofdm_run_t<2048>(O, first, n, n_samples, block_size, cs);
    case 1024: return ofdm_run_t<1024>(O, first, n, n_samples, block_size, cs);
    case 512: return ofdm_run_t<512>(O, first, n, n_samples, block_size, cs);
    case 256: return ofdm_run_t<256>(O, first, n, n_samples, block_size, cs);
    }
    return set_error(DABGPU_ERR_INVALID, "unsupported FFT size %d", O.P.nb_fft);
}

static int ofdm_process(OfdmState& O, const void* iq_host, size_t stride_bytes, int first, int n, int n_samples, int block_size, cudaStream_t cs) {
    if (O.external_ring) return set_error(DABGPU_ERR_STATE, "a device input buffer is attached: use dabgpu_ofdm_advance");
    // the ring must keep the frame being assembled plus the correlation window: bound the samples per pass
    const size_t max_chunk = O.ring_cap - size_t(O.P.nb_frame_samples) - size_t(O.P.nb_null_period + O.P.nb_symbol_period) - 1024;
    const int bs = block_size > 0 ? block_size : n_samples;
    size_t chunk_max = (max_chunk / size_t(bs)) * size_t(bs);   // keep Process() block boundaries intact
    if (chunk_max == 0) return set_error(DABGPU_ERR_INVALID, "block_size %d does not fit the IQ ring (%zu samples)", bs, O.ring_cap);
    const uint8_t* src = static_cast<const uint8_t*>(iq_host);
    for (size_t done = 0; done < size_t(n_samples);) {
        const size_t len = std::min(chunk_max, size_t(n_samples) - done);
        for (int i = 0; i < n; i++) {
            const int s = first + i;
            unsigned long long w = O.h_written[size_t(s)];
            size_t left = len, off = 0;
            while (left > 0) {
                const size_t pos = size_t(w & (O.ring_cap - 1));
                const size_t run = std::min(left, O.ring_cap - pos);
                CUDA_TRY(cudaMemcpyAsync(O.d_ring.as<uint8_t>() + (size_t(s) * O.ring_cap + pos) * O.bps,
                                         src + size_t(i) * stride_bytes + (done + off) * O.bps, run * O.bps, cudaMemcpyHostToDevice, cs));
                w += run; off += run; left -= run;
            }
            O.h_written[size_t(s)] = w;
        }
        int rc = ofdm_run(O, first, n, int(len), bs, cs);
        if (rc) return rc;
        done += len;
    }
    CUDA_TRY(cudaStreamSynchronize(cs));
    return DABGPU_OK;
}

static int ofdm_attach(OfdmState& O, const void* d_iq, size_t stride_samples, size_t capacity, cudaStream_t cs) {
    CUDA_TRY(cudaStreamSynchronize(cs));
    O.dev.ring = static_cast<const uint8_t*>(d_iq);
    O.dev.ring_stride = stride_samples;
    O.dev.ring_mask = (capacity & (capacity - 1)) == 0 ? (unsigned long long)(capacity - 1) : ~0ull;
    O.external_ring = true;
    k_ofdm_reset<<<(O.max_streams + 127) / 128, 128, 0, cs>>>(O.dev, -1, O.max_streams, 1);
    O.launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(cs));
    return DABGPU_OK;
}

static int ofdm_advance(OfdmState& O, int first, int n, int n_samples, int block_size, cudaStream_t cs) {
    if (!O.external_ring) return set_error(DABGPU_ERR_STATE, "no device input attached: use dabgpu_ofdm_process");
    if (n_samples < 0) return set_error(DABGPU_ERR_INVALID, "negative sample count");
    return ofdm_run(O, first, n, n_samples, block_size > 0 ? block_size : n_samples, cs);
}

static int ofdm_get_status(OfdmState& O, int stream, dabgpu_ofdm_status* out, cudaStream_t cs) {
    CUDA_TRY(cudaStreamSynchronize(cs));
    OfdmStream st;
    CUDA_TRY(cudaMemcpy(&st, O.dev.st + stream, sizeof(st), cudaMemcpyDeviceToHost));
    out->state = st.state;
    out->total_frames_read = st.frames_read;
    out->total_frames_desync = st.frames_desync;
    out->fine_time_offset = st.fine_time_offset;
    out->signal_l1_average = st.l1_avg;
    out->freq_coarse_offset = st.coarse;
    out->freq_fine_offset = st.fine;
    out->frames_queued = 0;
    return DABGPU_OK;
}

static int ofdm_fetch_latest(OfdmState& O, int first, int n, int8_t* frames_host, uint8_t* produced, cudaStream_t cs) {
    int rc;
    const size_t fb = size_t(O.P.nb_frame_bits);
    if ((rc = O.d_stage.alloc(size_t(n) * fb))) return rc;
    const dim3 grid(32, n);
    k_ofdm_gather_latest<<<grid, 256, 0, cs>>>(O.dev, first, O.d_stage.as<int8_t>(), O.d_produced.as<uint8_t>());
    O.launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(frames_host, O.d_stage.p, size_t(n) * fb, cudaMemcpyDeviceToHost, cs));
    CUDA_TRY(cudaMemcpyAsync(O.h_produced.p, O.d_produced.p, size_t(n), cudaMemcpyDeviceToHost, cs));
    CUDA_TRY(cudaStreamSynchronize(cs));
    memcpy(produced, O.h_produced.p, size_t(n));
    return DABGPU_OK;
}
