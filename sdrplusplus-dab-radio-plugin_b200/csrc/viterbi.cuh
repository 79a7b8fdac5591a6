// viterbi.cuh -- K=7 rate-1/4 Viterbi decoder, one warp per trellis.
//
// Replaces (bit-exactly) the decoder the reference selects on an AVX2 host:
//   DAB_Viterbi_Decoder::{reset,update,chainback}   dab/algorithms/dab_viterbi_decoder.cpp:109-181
//   ViterbiDecoder_AVX_u16<7,4>::{update,bfly,renormalise}  VIT/x86/viterbi_decoder_avx_u16.h:47-170
//   ViterbiDecoder_Core::chainback                   VIT/viterbi_decoder_core.h:214-236
// Semantics reproduced: saturating u16 path metrics, branch error = sum |bt - sym| over the 4
// outputs, inverse error = max(1016 - e, 0), decision bit = 1 when the path from state s+32 is
// <= the path from state s (ties => 1), renormalisation by the minimum when metric[0] >= 60455,
// whole-block traceback from state 0.
//
// Mapping: lane l owns butterfly l: inputs old[l] and old[l+32], outputs new[2l] and new[2l+1].
// The two decision bits per lane leave through two warp ballots per step (even-state word, odd-state
// word).  The metric exchange new -> old is two shuffles of the packed (new[2l], new[2l+1]) pair.
// Decisions go to a per-warp-slot scratch area (L2 resident; the kernel is persistent so the footprint
// is resident warps x trellis length), 256 B coalesced per 32 steps.  Traceback walks the scratch
// backwards 32 steps at a time; the survivor state is warp-uniform.
#pragma once
#include "tables.cuh"

enum : uint32_t {
    VJ_DESCRAMBLE = 1u,   // XOR decoded bytes with the energy dispersal PRBS
    VJ_GATHER = 2u,       // punctured symbols are fetched through the 16-CIF time de-interleaver
    VJ_FIB_CRC = 4u,      // output is a FIB group: check the CRC16 of every 32-byte FIB
};

struct __align__(16) VitJobDev {
    const int8_t* src;              // linear: first punctured symbol; gather: soft-bit frame ring of the stream
    uint8_t* out;                   // decoded bytes
    unsigned long long* path_error; // optional
    uint8_t* crc_ok;                // optional (VJ_FIB_CRC): one flag per FIB
    uint32_t seg_step_end[DABGPU_MAX_SEGMENTS];  // cumulative trellis steps at the end of each segment
    uint32_t seg_in_base[DABGPU_MAX_SEGMENTS];   // punctured-symbol index where each segment starts
    uint8_t seg_pi[8];
    uint32_t n_seg;
    uint32_t total_steps;           // 0 => inactive job
    uint32_t n_out_bytes;
    uint32_t flags;
    // gather parameters
    uint32_t newest_cif;            // absolute CIF index of the newest CIF (the one this job is decoding "at")
    uint32_t sub_start_bit;         // first soft bit of the sub-channel inside a CIF
    uint32_t n_fibs;
    uint32_t pad_;
};

struct GatherGeom {
    uint32_t nb_cifs;        // CIFs per transmission frame
    uint32_t frame_bits;     // soft bits per frame
    uint32_t fic_bits;
    uint32_t cif_bits;
    uint32_t slot_mask;      // frame ring depth - 1
};

#define VIT_MAX_ERROR 1016
#define VIT_NONSTART 5080u
#define VIT_RENORM 60455u
// metric[0] grows by at most 1020 per step, so a 32-step chunk that starts below this cannot reach the
// renormalisation threshold inside the chunk: the per-step check can be skipped without changing results.
#define VIT_CHUNK_SAFE (VIT_RENORM - 32u * 1020u)

__device__ __forceinline__ int8_t vit_load_soft(const VitJobDev& J, const GatherGeom& G, uint32_t idx) {
    if (J.flags & VJ_GATHER) {
        const uint32_t age = c_ti_age[idx & 15u];
        const uint32_t cabs = J.newest_cif - age;
        const uint32_t fr = cabs / G.nb_cifs;
        const uint32_t c = cabs - fr * G.nb_cifs;
        const size_t off = size_t(fr & G.slot_mask) * G.frame_bits + G.fic_bits + size_t(c) * G.cif_bits + J.sub_start_bit + idx;
        return __ldg(J.src + off);
    }
    return __ldg(J.src + idx);
}

// Depuncture on the fly: mother-code symbols of trellis step t packed as 4 int8 (punctured => 0).
// Reference: DAB_Viterbi_Decoder::depuncture_symbols, dab_viterbi_decoder.cpp:131-181
__device__ __forceinline__ uint32_t vit_load_step(const VitJobDev& J, const GatherGeom& G, uint32_t t) {
    uint32_t start = 0, pi = J.seg_pi[0], inb = J.seg_in_base[0];
#pragma unroll
    for (int i = 0; i < DABGPU_MAX_SEGMENTS - 1; i++) {
        if (t >= J.seg_step_end[i]) { start = J.seg_step_end[i]; pi = J.seg_pi[i + 1]; inb = J.seg_in_base[i + 1]; }
    }
    const uint32_t u = t - start;
    const uint32_t q = u >> 3, g = u & 7u;
    const uint32_t cnt = (c_pi_cnt[pi] >> (4u * g)) & 0xFu;
    const uint32_t pre = uint32_t(c_pi_pref[pi] >> (8u * g)) & 0xFFu;
    const uint32_t base = inb + q * c_pi_K[pi] + pre;
    uint32_t w = 0;
#pragma unroll
    for (uint32_t r = 0; r < 4; r++) {
        if (r < cnt) w |= uint32_t(uint8_t(vit_load_soft(J, G, base + r))) << (8u * r);
    }
    return w;
}

#define VIT_ACS_STEP(J_)                                                                            \
    {                                                                                               \
        const uint32_t w_ = __shfl_sync(FULL_MASK, word, (J_));                                     \
        uint32_t e_;                                                                                \
        asm("vabsdiff4.u32.s32.s32.add %0, %1, %2, %3;" : "=r"(e_) : "r"(bt), "r"(w_), "r"(0u));    \
        const uint32_t inv_ = uint32_t(max(VIT_MAX_ERROR - int(e_), 0));                            \
        const uint32_t n00 = min(m_lo + e_, 65535u), n10 = min(m_hi + inv_, 65535u);                \
        const uint32_t n01 = min(m_lo + inv_, 65535u), n11 = min(m_hi + e_, 65535u);                \
        const uint32_t be_ = __ballot_sync(FULL_MASK, n10 <= n00);                                  \
        const uint32_t bo_ = __ballot_sync(FULL_MASK, n11 <= n01);                                  \
        if (lane == (J_)) { dec_e = be_; dec_o = bo_; }                                             \
        const uint32_t pk_ = min(n00, n10) | (min(n01, n11) << 16);                                 \
        const uint32_t a_ = __shfl_sync(FULL_MASK, pk_, src_lo);                                    \
        const uint32_t b_ = __shfl_sync(FULL_MASK, pk_, src_hi);                                    \
        m_lo = __byte_perm(a_, 0u, sel);                                                            \
        m_hi = __byte_perm(b_, 0u, sel);                                                            \
    }

#define VIT_RENORM_CHECK()                                                                          \
    if (__ballot_sync(FULL_MASK, m_lo >= VIT_RENORM) & 1u) {                                        \
        const uint32_t mn_ = __reduce_min_sync(FULL_MASK, min(m_lo, m_hi));                         \
        m_lo -= mn_; m_hi -= mn_; acc_err += mn_;                                                   \
    }

__device__ __forceinline__ uint16_t crc16_ccitt_dev(const uint8_t* p, int n) {
    uint32_t crc = 0xFFFFu;
    for (int i = 0; i < n; i++) crc = ((crc << 8) ^ c_crc_ccitt[((crc >> 8) ^ p[i]) & 0xFFu]) & 0xFFFFu;
    return uint16_t(crc ^ 0xFFFFu);
}

__device__ void vit_decode_job(const VitJobDev& J, const GatherGeom& G, uint2* __restrict__ scratch,
                               const uint32_t* __restrict__ prbs_words, const uint32_t lane) {
    const uint32_t N = J.total_steps;
    const uint32_t bt = c_branch[lane];
    const uint32_t src_lo = lane >> 1, src_hi = 16u + (lane >> 1);
    const uint32_t sel = (lane & 1u) ? 0x4432u : 0x4410u;
    // ViterbiDecoder_Core::reset (viterbi_decoder_core.h:202-211), config dab_viterbi_decoder.cpp:31-41
    uint32_t m_lo = (lane == 0) ? 0u : VIT_NONSTART;
    uint32_t m_hi = VIT_NONSTART;
    unsigned long long acc_err = 0;

    for (uint32_t t0 = 0; t0 < N; t0 += 32) {
        const uint32_t t = t0 + lane;
        const uint32_t word = (t < N) ? vit_load_step(J, G, t) : 0u;
        uint32_t dec_e = 0, dec_o = 0;
        const uint32_t m0 = __shfl_sync(FULL_MASK, m_lo, 0);
        const uint32_t n = min(32u, N - t0);
        if (n == 32u && m0 < VIT_CHUNK_SAFE) {
#pragma unroll
            for (int j = 0; j < 32; j++) VIT_ACS_STEP(j)
        } else {
            for (uint32_t j = 0; j < n; j++) {
                VIT_ACS_STEP(j)
                VIT_RENORM_CHECK()
            }
        }
        if (t < N) scratch[t] = make_uint2(dec_e, dec_o);
    }
    if (J.path_error != nullptr && lane == 0) *J.path_error = acc_err + m_lo;
    __syncwarp();

    // Traceback (viterbi_decoder_core.h:214-236): decoded bit j = decision[j+6][state], state walks
    // back from 0; bytes are MSB first.  Words are assembled big-endian: bit 31 of word k = bit 32k.
    const uint32_t nbits = J.n_out_bytes * 8u;
    const uint32_t nwords = (nbits + 31u) >> 5;
    uint32_t state = 0, myword = 0;
    for (int k = int(nwords) - 1; k >= 0; --k) {
        const uint32_t jlo = uint32_t(k) << 5;
        const uint32_t m = min(32u, nbits - jlo);
        uint2 d = make_uint2(0u, 0u);
        if (lane < m) d = scratch[jlo + lane + 6u];
        uint32_t acc = 0;
        if (m == 32u) {
#pragma unroll
            for (int i = 31; i >= 0; --i) {
                const uint32_t w = __shfl_sync(FULL_MASK, (state & 1u) ? d.y : d.x, i);
                const uint32_t bit = (w >> (state >> 1)) & 1u;
                state = (state >> 1) | (bit << 5);
                acc = (acc >> 1) | (bit << 31);
            }
        } else {
            for (int i = int(m) - 1; i >= 0; --i) {
                const uint32_t w = __shfl_sync(FULL_MASK, (state & 1u) ? d.y : d.x, i);
                const uint32_t bit = (w >> (state >> 1)) & 1u;
                state = (state >> 1) | (bit << 5);
                acc = (acc >> 1) | (bit << 31);
            }
        }
        if (lane == (uint32_t(k) & 31u)) myword = acc;
        if ((k & 31) == 0) {
            const uint32_t widx = uint32_t(k) + lane;
            if (widx < nwords) {
                uint32_t v = myword;
                if (J.flags & VJ_DESCRAMBLE) v ^= prbs_words[widx];
                const uint32_t b0 = widx * 4u;
#pragma unroll
                for (uint32_t b = 0; b < 4; b++)
                    if (b0 + b < J.n_out_bytes) J.out[b0 + b] = uint8_t(v >> (24u - 8u * b));
            }
            myword = 0;
        }
    }
    if (J.flags & VJ_FIB_CRC) {
        // FIB = 30 data bytes + CRC16, fic_decoder.cpp:98-116
        __syncwarp();
        if (lane < J.n_fibs) {
            const uint8_t* fib = J.out + 32u * lane;
            const uint16_t rx = uint16_t((uint16_t(fib[30]) << 8) | fib[31]);
            J.crc_ok[lane] = (crc16_ccitt_dev(fib, 30) == rx) ? 1 : 0;
        }
    }
}

#define VIT_WARPS_PER_BLOCK 4

// Persistent kernel: every warp pulls trellises from a global counter until none are left.
__global__ void __launch_bounds__(VIT_WARPS_PER_BLOCK * 32)
k_viterbi(const VitJobDev* __restrict__ jobs, const int n_jobs, int* __restrict__ counter, uint2* __restrict__ scratch,
          const uint32_t scratch_steps, const uint32_t* __restrict__ prbs_words, const GatherGeom G) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t slot = blockIdx.x * VIT_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    uint2* my_scratch = scratch + size_t(slot) * scratch_steps;
    for (;;) {
        int job = 0;
        if (lane == 0) job = atomicAdd(counter, 1);
        job = __shfl_sync(FULL_MASK, job, 0);
        if (job >= n_jobs) break;
        const VitJobDev J = jobs[job];
        if (J.total_steps == 0) continue;
        vit_decode_job(J, G, my_scratch, prbs_words, lane);
        __syncwarp();
    }
}
