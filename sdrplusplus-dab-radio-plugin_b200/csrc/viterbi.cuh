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

// Frame layout in the soft-bit ring: [FIC, natural order, fic_bits][CIF 0][CIF 1]...; every CIF is stored as 16 planes of
// cif_bits/16 bytes, plane r = the soft bits with index r modulo 16 inside the CIF (written that way by k_ofdm_demod2 and
// k_frame_planarize).  Sub-channels start at multiples of 64 bits, so bit j of a sub-channel is byte (start_bit + j) / 16 of plane
// j mod 16 -- and j mod 16 is what selects the CIF in the time de-interleaver.
#define FRAME_PLANES 16u
struct GatherGeom {
    uint32_t cif_shift;      // log2(CIFs per transmission frame): 2, 0, 0, 1 for modes I..IV
    uint32_t nb_cifs;        // CIFs per transmission frame
    uint32_t frame_bits;     // soft bits per frame
    uint32_t fic_bits;
    uint32_t cif_bits;
    uint32_t slot_mask;      // frame ring depth - 1
};
__host__ __device__ __forceinline__ uint32_t geom_plane_stride(const GatherGeom& G) { return G.cif_bits / FRAME_PLANES; }

// Natural-order frames (what On_OFDM_Frame observers and BasicRadio::Process see) <-> the ring layout.  One thread per group of 16
// consecutive soft bits: in the MSC they are one byte of each plane (a warp touches 32 consecutive bytes of every plane), in the
// FIC they stay together.  to_ring: natural -> ring, else ring -> natural.
__device__ __forceinline__ void frame_convert_group(const int8_t* __restrict__ src, int8_t* __restrict__ dst, const uint32_t grp, const GatherGeom& G, const bool to_ring) {
    const uint32_t i0 = grp * 16u;
    if (i0 < G.fic_bits) {
        *reinterpret_cast<uint4*>(dst + i0) = *reinterpret_cast<const uint4*>(src + i0);
        return;
    }
    const uint32_t m = i0 - G.fic_bits, c = m / G.cif_bits, b = m - c * G.cif_bits;
    const uint32_t pbase = G.fic_bits + c * G.cif_bits + (b >> 4), ps = geom_plane_stride(G);
    if (to_ring) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + i0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (uint32_t r = 0; r < 16u; r++) dst[pbase + r * ps] = int8_t((w[r >> 2] >> (8u * (r & 3u))) & 0xFFu);
    } else {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (uint32_t r = 0; r < 16u; r++) w[r >> 2] |= uint32_t(uint8_t(src[pbase + r * ps])) << (8u * (r & 3u));
        *reinterpret_cast<uint4*>(dst + i0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// One frame per blockIdx.y.  nat: natural-order frames, frame y at nat + y * nat_stride.  ring: the frame ring of stream
// first_stream (stream y at + y * stream_stride); the slot is slot_of[first_stream + y] & slot_mask (+ slot_delta), or fixed_slot
// when slot_of is null.
__global__ void k_frame_convert(int8_t* __restrict__ nat, const size_t nat_stride, int8_t* __restrict__ ring, const size_t stream_stride,
                                const uint32_t* __restrict__ slot_of, const int first_stream, const uint32_t slot_delta, const uint32_t fixed_slot,
                                const GatherGeom G, const int to_ring) {
    const uint32_t y = blockIdx.y;
    const uint32_t slot = slot_of ? ((slot_of[first_stream + y] + slot_delta) & G.slot_mask) : fixed_slot;
    int8_t* fr = ring + size_t(y) * stream_stride + size_t(slot) * G.frame_bits;
    int8_t* nt = nat + size_t(y) * nat_stride;
    const uint32_t n_grp = G.frame_bits / 16u;
    for (uint32_t grp = blockIdx.x * blockDim.x + threadIdx.x; grp < n_grp; grp += gridDim.x * blockDim.x) {
        if (to_ring) frame_convert_group(nt, fr, grp, G, true);
        else frame_convert_group(fr, nt, grp, G, false);
    }
}

#define VIT_MAX_ERROR 1016
#define VIT_NONSTART 5080u
#define VIT_RENORM 60455u
// metric[0] grows by at most 1020 per step, so a 32-step chunk that starts below this cannot reach the
// renormalisation threshold inside the chunk: the per-step check can be skipped without changing results.
#define VIT_CHUNK_SAFE (VIT_RENORM - 32u * 1020u)

// Offset (relative to the stream's frame ring) of the plane that holds the soft bits with index r modulo 16 of the logical frame a
// gather job decodes, at the first byte of its sub-channel.  Bit i of the oldest complete logical frame was sent 15 - T[i mod 16]
// CIFs before the newest one, and the interleaver sequence T = {0,8,4,12,...} is the 4-bit reversal of i
// (cif_deinterleaver.cpp:8-11, 62-68).
__device__ __forceinline__ uint32_t vit_plane_offset(const uint32_t newest_cif, const uint32_t sub_start_bit, const GatherGeom& G, const uint32_t r) {
    const uint32_t age = 15u - (__brev(r) >> 28);
    const uint32_t cabs = newest_cif - age;
    const uint32_t fr = cabs >> G.cif_shift;
    const uint32_t c = cabs & (G.nb_cifs - 1u);
    return (fr & G.slot_mask) * G.frame_bits + G.fic_bits + c * G.cif_bits + r * geom_plane_stride(G) + (sub_start_bit >> 4);
}

// Per-trellis gather table: punctured symbol idx of a gather job is J.src[s_rowoff[idx & 15] + (idx >> 4)]; linear jobs (no time
// de-interleaver: FIC, dabgpu_viterbi_decode) read J.src[idx] (all-zero table, shift 0).
__device__ __forceinline__ void vit_fill_rowoff(const VitJobDev& J, const GatherGeom& G, uint32_t* s_rowoff, const uint32_t lane) {
    if (lane < 16u) s_rowoff[lane] = (J.flags & VJ_GATHER) ? vit_plane_offset(J.newest_cif, J.sub_start_bit, G, lane) : 0u;
}

// Depuncture on the fly: mother-code symbols of trellis step t packed as 4 int8 (punctured => 0).
// Reference: DAB_Viterbi_Decoder::depuncture_symbols, dab_viterbi_decoder.cpp:131-181
__device__ __forceinline__ uint32_t vit_load_step(const VitJobDev& J, const uint32_t* s_rowoff, uint32_t t) {
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
    const uint32_t sh = (J.flags & VJ_GATHER) ? 4u : 0u;
    uint32_t w = 0;
#pragma unroll
    for (uint32_t r = 0; r < 4; r++) {
        if (r < cnt) {
            const uint32_t idx = base + r;
            w |= uint32_t(uint8_t(__ldg(J.src + (size_t(s_rowoff[idx & 15u]) + (idx >> sh))))) << (8u * r);
        }
    }
    return w;
}

// The symbols of the 32 steps of a chunk sit in shared memory (one word per step, written by the lane that fetched it)
// and are read back as warp-uniform 128-bit loads, four steps at a time.
#define VIT_WORD(J_) ((((J_) & 3) == 0) ? wq.x : (((J_) & 3) == 1) ? wq.y : (((J_) & 3) == 2) ? wq.z : wq.w)
#define VIT_WORD_FETCH(J_) if (((J_) & 3) == 0) wq = reinterpret_cast<const uint4*>(s_words)[(J_) >> 2];

// ---- exact step (reference arithmetic spelled out: saturating adds, renormalisation check after every step) ----
#define VIT_ACS_STEP(J_)                                                                            \
    {                                                                                               \
        const uint32_t w_ = s_words[(J_)];                                                          \
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

// ---- fast step: both metrics of a lane packed as u16x2 and processed by one VIADD / VIMNMX.U16x2 each.
//   P = (old[l], old[l]), Q = (old[l+32], old[l+32]);  E = (e, 1016-e), E2 = (1016-e, e) built by one IMAD each
//   A = P + E = (old[l]+e, old[l]+inv), B = Q + E2 = (old[l+32]+inv, old[l+32]+e), NEW = min(B, A) = (new[2l], new[2l+1])
//   decision = 1 iff the upper predecessor is <= the lower one (tie => 1): these are the two predicate outputs of
//   VIMNMX.U16x2 (per-halfword B <= A), so the decisions cost no instruction beyond the two ballots.
// Valid only while no sum can reach 65536 and no renormalisation can trigger, which holds when state 0 stays
// <= VIT_FAST_SAFE during the chunk (every metric is within 6*1020 of state 0's: any state is reachable from any
// other in 6 steps) and no soft symbol is -128 (then 1016-e never needs the reference's clamp at 0).  The chunk
// is verified after the fact and redone with an exact step otherwise.
#define VIT_FAST_SAFE 58000u
#define VIT_FAST_STEP(J_)                                                                           \
    {                                                                                               \
        VIT_WORD_FETCH(J_)                                                                          \
        const uint32_t w_ = VIT_WORD(J_);                                                           \
        uint32_t e_;                                                                                \
        asm("vabsdiff4.u32.s32.s32.add %0, %1, %2, %3;" : "=r"(e_) : "r"(bt), "r"(w_), "r"(0u));    \
        const uint32_t E_ = e_ * 0xFFFF0001u + (uint32_t(VIT_MAX_ERROR) << 16);                     \
        const uint32_t E2_ = e_ * 0x0000FFFFu + uint32_t(VIT_MAX_ERROR);                            \
        bool ph_, pl_;                                                                              \
        const uint32_t NEW_ = __vibmin_u16x2(Q + E2_, P + E_, &ph_, &pl_);                          \
        const uint32_t be_ = __ballot_sync(FULL_MASK, pl_);                                         \
        const uint32_t bo_ = __ballot_sync(FULL_MASK, ph_);                                         \
        if (SM) { if (lane == 0) dec[t0 + (J_)] = make_uint2(be_, bo_); }                           \
        else if (lane == (J_)) { dec_e = be_; dec_o = bo_; }                                        \
        const uint32_t a_ = __shfl_sync(FULL_MASK, NEW_, src_lo);                                   \
        const uint32_t b_ = __shfl_sync(FULL_MASK, NEW_, src_hi);                                   \
        P = __byte_perm(a_, 0u, dup);                                                               \
        Q = __byte_perm(b_, 0u, dup);                                                               \
        mx = max(mx, P);                                                                            \
    }

// ---- packed exact step: the same u16x2 arithmetic with the reference's saturation and renormalisation reproduced.
// Metrics are held with a bias of -VIT_PK_BIAS so that a sum can exceed the saturation level without wrapping the
// halfword: ref 65535 <-> stored VIT_PK_CLAMP, and min(sum, VIT_PK_CLAMP) per halfword IS the saturating add of the
// reference.  Needs every metric >= VIT_PK_BIAS at the start of the chunk (guaranteed when state 0 >= VIT_PK_MIN0, by
// the 6-step reachability bound) and no -128 symbol.  The renormalisation test (state 0 >= 60455 after the step) is made
// every step; when it fires the metrics are renormalised exactly as the reference does (subtract the minimum) and the
// rest of the chunk runs on the scalar exact step, because the renormalised metrics fall below the bias.
#define VIT_PK_BIAS 1024u
#define VIT_PK_CLAMP ((65535u - VIT_PK_BIAS) * 0x10001u)
#define VIT_PK_RENORM ((VIT_RENORM - VIT_PK_BIAS) * 0x10001u)
#define VIT_PK_MIN0 (VIT_PK_BIAS + 6u * 1020u)
#define VIT_PK_STEP(J_)                                                                             \
    {                                                                                               \
        VIT_WORD_FETCH(J_)                                                                          \
        const uint32_t w_ = VIT_WORD(J_);                                                           \
        uint32_t e_;                                                                                \
        asm("vabsdiff4.u32.s32.s32.add %0, %1, %2, %3;" : "=r"(e_) : "r"(bt), "r"(w_), "r"(0u));    \
        const uint32_t E_ = e_ * 0xFFFF0001u + (uint32_t(VIT_MAX_ERROR) << 16);                     \
        const uint32_t E2_ = e_ * 0x0000FFFFu + uint32_t(VIT_MAX_ERROR);                            \
        const uint32_t A_ = __vminu2(P + E_, VIT_PK_CLAMP), B_ = __vminu2(Q + E2_, VIT_PK_CLAMP);   \
        bool ph_, pl_;                                                                              \
        const uint32_t NEW_ = __vibmin_u16x2(B_, A_, &ph_, &pl_);                                   \
        const uint32_t be_ = __ballot_sync(FULL_MASK, pl_);                                         \
        const uint32_t bo_ = __ballot_sync(FULL_MASK, ph_);                                         \
        if (lane == (J_)) { dec_e = be_; dec_o = bo_; }                                             \
        const uint32_t a_ = __shfl_sync(FULL_MASK, NEW_, src_lo);                                   \
        const uint32_t b_ = __shfl_sync(FULL_MASK, NEW_, src_hi);                                   \
        P = __byte_perm(a_, 0u, dup);                                                               \
        Q = __byte_perm(b_, 0u, dup);                                                               \
    }

__device__ __forceinline__ uint16_t crc16_ccitt_dev(const uint8_t* p, int n) {
    uint32_t crc = 0xFFFFu;
    for (int i = 0; i < n; i++) crc = ((crc << 8) ^ c_crc_ccitt[((crc >> 8) ^ p[i]) & 0xFFu]) & 0xFFFFu;
    return uint16_t(crc ^ 0xFFFFu);
}

// Traceback of the bits [b_lo, b_hi] (walking down) from `state`; decoded bit j comes from decision word j+6
// (viterbi_decoder_core.h:214-236).  When emit is set, completed 32-bit words (big endian: bit 32k is the MSB of
// word k) are descrambled and stored.  Returns the state reached below b_lo.
__device__ __forceinline__ uint32_t vit_walk(const uint2* __restrict__ dec, const int b_hi, const int b_lo, uint32_t state, const bool emit,
                                             const VitJobDev& J, const uint32_t* __restrict__ prbs_words) {
    uint32_t acc = 0;
    for (int b = b_hi; b >= b_lo; --b) {
        const uint32_t* d = reinterpret_cast<const uint32_t*>(dec + (b + 6));
        const uint32_t w = d[state & 1u];
        const uint32_t bit = (w >> (state >> 1)) & 1u;
        state = (state >> 1) | (bit << 5);
        acc = (acc >> 1) | (bit << 31);
        if (emit && (b & 31) == 0) {
            const uint32_t widx = uint32_t(b) >> 5;
            uint32_t v = acc;
            if (J.flags & VJ_DESCRAMBLE) v ^= prbs_words[widx];
            const uint32_t b0 = widx * 4u;
            if (b0 + 4u <= J.n_out_bytes && ((reinterpret_cast<uintptr_t>(J.out) & 3u) == 0)) {
                *reinterpret_cast<uint32_t*>(J.out + b0) = __byte_perm(v, 0u, 0x0123);
            } else {
#pragma unroll
                for (uint32_t q = 0; q < 4; q++)
                    if (b0 + q < J.n_out_bytes) J.out[b0 + q] = uint8_t(v >> (24u - 8u * q));
            }
            acc = 0;
        }
    }
    return state;
}

#define VIT_OVERLAP 128   // speculative traceback: steps walked before a chunk to let the survivors merge

// SM: the decision words of this trellis live in shared memory (written by lane 0 every step); otherwise in the
// per-warp-slot global scratch (recorded through the lane == step select and stored coalesced per chunk).
// Three step variants, always bit-identical to the reference:
//   fast   (VIT_FAST_STEP)  : packed, no saturation / renormalisation handling, verified after the chunk
//   packed (VIT_PK_STEP)    : packed with the reference's saturation and renormalisation, for the chunks around a
//                             renormalisation (state 0 between VIT_FAST_SAFE and VIT_RENORM)
//   scalar (VIT_ACS_STEP)   : the reference arithmetic spelled out; partial chunks, -128 symbols, and the steps that
//                             follow a renormalisation inside a chunk
template <bool SM>
__device__ void vit_decode_job(const VitJobDev& J, uint2* __restrict__ dec, uint32_t* __restrict__ s_words,
                               const uint32_t* __restrict__ s_rowoff, const uint32_t* __restrict__ prbs_words, const uint32_t lane) {
    const uint32_t N = J.total_steps;
    const uint32_t bt = c_branch[lane];
    const uint32_t src_lo = lane >> 1, src_hi = 16u + (lane >> 1);
    const uint32_t sel = (lane & 1u) ? 0x4432u : 0x4410u;   // one half, zero extended (scalar step)
    const uint32_t dup = (lane & 1u) ? 0x3232u : 0x1010u;   // one half, duplicated (packed steps)
    // ViterbiDecoder_Core::reset (viterbi_decoder_core.h:202-211), config dab_viterbi_decoder.cpp:31-41
    uint32_t P = ((lane == 0) ? 0u : VIT_NONSTART) * 0x10001u;
    uint32_t Q = VIT_NONSTART * 0x10001u;
    unsigned long long acc_err = 0;
    uint32_t growth = 0;   // growth of state 0's metric over the previous chunk: predicts whether the fast step can hold

    uint32_t word = (lane < N) ? vit_load_step(J, s_rowoff, lane) : 0u;
    for (uint32_t t0 = 0; t0 < N; t0 += 32) {
        const uint32_t t = t0 + lane;
        __syncwarp();            // every lane is done with the previous chunk's symbols
        s_words[lane] = word;
        __syncwarp();
        // prefetch the symbols of the next chunk: the gather goes through L2/HBM and must not sit on the ACS chain
        const uint32_t next_word = (t + 32u < N) ? vit_load_step(J, s_rowoff, t + 32u) : 0u;
        uint32_t dec_e = 0, dec_o = 0;
        const uint32_t n = min(32u, N - t0);
        // a -128 symbol makes 1016 - e negative for some branch: needs the clamp of the scalar step
        const uint32_t z = word ^ 0x80808080u;
        const bool has_m128 = __ballot_sync(FULL_MASK, ((z - 0x01010101u) & ~z & 0x80808080u) != 0u) != 0u;
        const uint32_t m0 = __shfl_sync(FULL_MASK, P, 0) & 0xFFFFu;
        bool done = false, stored = false;
        uint32_t jr = 0;         // first step left to the scalar loop
        uint32_t m_lo = 0, m_hi = 0;
        uint4 wq;
        if (n == 32u && !has_m128) {
            const bool pk_ok = m0 >= VIT_PK_MIN0;
            if (!pk_ok || m0 + growth + 256u <= VIT_FAST_SAFE) {
                const uint32_t P0 = P, Q0 = Q;
                uint32_t mx = P;
#pragma unroll
                for (int j = 0; j < 32; j++) VIT_FAST_STEP(j)
                if ((__shfl_sync(FULL_MASK, mx, 0) & 0xFFFFu) <= VIT_FAST_SAFE) { done = true; stored = SM; }
                else { P = P0; Q = Q0; }
            }
            if (!done && pk_ok) {
                P -= VIT_PK_BIAS * 0x10001u; Q -= VIT_PK_BIAS * 0x10001u;
                jr = 32u;
                bool fired = false;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    VIT_PK_STEP(j)
                    if (__any_sync(FULL_MASK, lane == 0u && P >= VIT_PK_RENORM)) { fired = true; jr = uint32_t(j) + 1u; break; }
                }
                if (fired) {
                    // ViterbiDecoder_AVX_u16::renormalise: subtract the minimum over the 64 states
                    const uint32_t mn = __reduce_min_sync(FULL_MASK, min(P & 0xFFFFu, Q & 0xFFFFu));
                    m_lo = (P & 0xFFFFu) - mn; m_hi = (Q & 0xFFFFu) - mn;
                    acc_err += mn + VIT_PK_BIAS;
                } else {
                    P += VIT_PK_BIAS * 0x10001u; Q += VIT_PK_BIAS * 0x10001u;
                    done = true;
                }
            }
        }
        if (!done) {
            if (jr == 0u) { m_lo = P & 0xFFFFu; m_hi = Q & 0xFFFFu; }
            for (uint32_t j = jr; j < n; j++) {
                VIT_ACS_STEP(j)
                VIT_RENORM_CHECK()
            }
            P = m_lo * 0x10001u; Q = m_hi * 0x10001u;
        }
        if (!stored && t < N) dec[t] = make_uint2(dec_e, dec_o);
        const uint32_t m0_end = __shfl_sync(FULL_MASK, P, 0) & 0xFFFFu;
        growth = (m0_end > m0) ? (m0_end - m0) : 0u;
        word = next_word;
    }
    if (J.path_error != nullptr && lane == 0) *J.path_error = acc_err + (P & 0xFFFFu);
    __syncwarp();

    // ---- whole-block traceback from state 0, parallelised over lanes --------------------------------------------
    // Lane c owns the output words [c*W, (c+1)*W).  Only the top lane knows its start state (0); every other lane
    // starts VIT_OVERLAP steps above its range from state 0, which puts it on the surviving path with high
    // probability.  The guess of lane c is then CHECKED against the state the lane above actually reached; any lane
    // whose guess was wrong is walked again from the verified state (top down), so the result is always the exact
    // whole-block traceback of the reference.
    const int nbits = int(J.n_out_bytes) * 8;
    if (nbits > 0) {
        const int nwords = (nbits + 31) >> 5;
        const int W = (nwords + 31) >> 5;
        const int nact = (nwords + W - 1) / W;
        const bool active = int(lane) < nact;
        const int lo_bit = int(lane) * W * 32;
        const int hi_bit = min((int(lane) + 1) * W * 32, nbits);   // exclusive
        const bool top = active && (hi_bit == nbits);
        uint32_t g = 0, h = 0;
        if (active) {
            if (!top) g = vit_walk(dec, min(hi_bit + VIT_OVERLAP, nbits) - 1, hi_bit, 0u, false, J, prbs_words);
            h = vit_walk(dec, hi_bit - 1, lo_bit, g, true, J, prbs_words);
        }
        for (;;) {
            const uint32_t h_above = __shfl_down_sync(FULL_MASK, h, 1);
            const uint32_t bad = __ballot_sync(FULL_MASK, active && !top && g != h_above);
            if (bad == 0u) break;
            const uint32_t fix = 31u - uint32_t(__clz(int(bad)));   // highest wrong lane: everything above it is verified
            if (lane == fix) {
                g = h_above;
                h = vit_walk(dec, hi_bit - 1, lo_bit, g, true, J, prbs_words);
            }
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
#define VIT_SMEM_STEPS 1600   // trellises up to this many steps keep their decisions in shared memory (12.5 KB per warp)
#define VIT_WARP_SMEM (size_t(VIT_SMEM_STEPS) * sizeof(uint2) + 32 * 4 + 16 * 4)   // decisions + chunk symbols + gather table

// Persistent kernel: every warp pulls trellises from a global counter until none are left.
__global__ void __launch_bounds__(VIT_WARPS_PER_BLOCK * 32)
k_viterbi(const VitJobDev* __restrict__ jobs, const int n_jobs, int* __restrict__ counter, uint2* __restrict__ scratch,
          const uint32_t scratch_steps, const uint32_t* __restrict__ prbs_words, const GatherGeom G, const uint32_t* __restrict__ lanes_plan,
          const uint32_t* __restrict__ order) {
    // lanes_plan[0] != 0: this call is decoded by k_viterbi_lanes (viterbi_lanes.cuh), decided on the device by k_vit_plan
    if (lanes_plan != nullptr && lanes_plan[0] != 0u) return;
    // order != nullptr: the active jobs by length class, longest first (k_vit_scatter; VlPlan words 2 and 3 = n_active, oversize)
    const bool ordered = order != nullptr && lanes_plan != nullptr && lanes_plan[3] == 0u;
    const int n_pull = ordered ? int(lanes_plan[2]) : n_jobs;
    extern __shared__ __align__(16) uint8_t s_vit[];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t wib = threadIdx.x >> 5;
    const uint32_t slot = blockIdx.x * VIT_WARPS_PER_BLOCK + wib;
    uint2* my_scratch = scratch + size_t(slot) * scratch_steps;
    uint8_t* base = s_vit + size_t(wib) * VIT_WARP_SMEM;
    uint2* my_smem = reinterpret_cast<uint2*>(base);
    uint32_t* s_words = reinterpret_cast<uint32_t*>(base + size_t(VIT_SMEM_STEPS) * sizeof(uint2));
    uint32_t* s_rowoff = s_words + 32;
    for (;;) {
        int job = 0;
        if (lane == 0) job = atomicAdd(counter, 1);
        job = __shfl_sync(FULL_MASK, job, 0);
        if (job >= n_pull) break;
        const VitJobDev J = jobs[ordered ? order[job] : uint32_t(job)];
        if (J.total_steps == 0) continue;
        __syncwarp();
        vit_fill_rowoff(J, G, s_rowoff, lane);
        __syncwarp();
        if (J.total_steps <= VIT_SMEM_STEPS) vit_decode_job<true>(J, my_smem, s_words, s_rowoff, prbs_words, lane);
        else vit_decode_job<false>(J, my_scratch, s_words, s_rowoff, prbs_words, lane);
        __syncwarp();
    }
}
#define VIT_SMEM_BYTES (size_t(VIT_WARPS_PER_BLOCK) * VIT_WARP_SMEM)
