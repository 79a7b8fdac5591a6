// viterbi_lane_core.h -- K=7 rate-1/4 Viterbi with the whole 64-state trellis held by ONE thread.
//
// This is the arithmetic core of the batch ("lane per trellis") decoder k_viterbi_lanes: 32 independent trellises per
// warp, no shuffles and no ballots.  It is written so that the same code compiles for the device (DPX instructions
// VIADDMNMX.U16x2 / VIMNMX.U16x2) and for the host (plain C++ emulation of those instructions), which is how
// tests/host/lane_core_check.cpp checks the algorithm against the oracle without a GPU.
//
// Bit-exact with the decoder the reference selects on an AVX2 host:
//   ViterbiDecoder_AVX_u16<7,4>::{update,bfly,renormalise}  VIT/x86/viterbi_decoder_avx_u16.h:47-170
//   ViterbiDecoder_Core::chainback                           VIT/viterbi_decoder_core.h:214-236
//   config (max error 1016, non-start 5080, renormalise at 60455)  dab/algorithms/dab_viterbi_decoder.cpp:31-41
//
// Data layout.  The 64 path metrics live in 32 registers, two u16 halves each.  In "layout k" (k = 0..5) register i holds
// the states  lo = insert0(i, k)  and  hi = lo | (1 << k)  (insert0 puts a zero bit at position k of the 5-bit index i).
// For k <= 4 the butterfly inputs old[j], old[j+32] of the pair (j, j^(1<<k)) are R[i] and R[i+16], and the outputs
// (new[2j], new[2j']) and (new[2j+1], new[2j'+1]) are exactly registers 2i and 2i+1 of layout k+1: a trellis step is a
// pure register renaming, the add-compare-select of two butterflies costs six u16x2 instructions.  After five steps the
// metrics are in layout 5 = (s, s+32); 32 byte permutes bring them back to layout 0.  So the code is unrolled by five steps.
//
// Exactness.  The reference keeps u16 metrics with saturating adds and subtracts the minimum whenever metric[0] >= 60455.
// Here a metric is stored as  rel = ref - off  with a per-trellis scalar `off`; every five steps `off` is moved so that
// rel[0] = 8192, which keeps every rel in [2072, 21452] (any state is reachable from any other in 6 steps, so all metrics
// are within 6*1020 of metric[0]) and leaves bit 15 of every half free.
//   * saturation: ref saturates at 65535 <=> rel clamps at CL = 65535 - off.  Only the metric coming from the upper
//     predecessor needs the clamp (B = min(old[j+32] + e', CL)); new = min(old[j] + e, B) and the decision
//     "B <= A" <=> "new == B" are then exactly the reference's saturated results (case analysis in DESIGN.md section 4).
//   * renormalisation: the reference's subtraction of the minimum only changes `off` (and the accumulated path error),
//     never a rel value, so it is a scalar update in a rarely taken branch.
//   * decision bit = 1 iff the path from state j+32 is <= the path from state j (ties => 1), as VPCMPEQW(min, upper) does.
//   * inverse branch error = max(1016 - e, 0) (e can reach 1020 when a soft bit is -128).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VL_HD __host__ __device__ __forceinline__
#else
#define VL_HD inline
#endif

#define VL_MAX_ERROR 1016u
#define VL_NONSTART 5080u
#define VL_RENORM 60455
#define VL_ORIGIN 8192u     // rel value given to state 0 at every own renormalisation
#define VL_UNROLL 5         // trellis steps per layout cycle

// Opaque multipliers (passed as kernel arguments on the device) so that the compiler keeps `a * k + b` as an integer
// multiply-add on the FMA pipe instead of folding it into ALU-pipe adds/shifts: the ACS loop is ALU-pipe bound.
struct VlConst { uint32_t m1, two, x4, x16, x256, x10000; };

// ---- the five instructions the loop is made of -------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define VL_DEVICE_CODE 1
VL_HD uint32_t vl_addmin(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_u16x2(a, b, c); }   // min(a + b, c) per half
VL_HD uint32_t vl_min(uint32_t a, uint32_t b) { return __vminu2(a, b); }
VL_HD uint32_t vl_prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
VL_HD uint32_t vl_absdiff4(uint32_t bt, uint32_t w) {
    uint32_t e;
    asm("vabsdiff4.u32.s32.s32.add %0, %1, %2, %3;" : "=r"(e) : "r"(bt), "r"(w), "r"(0u));
    return e;
}
VL_HD uint32_t vl_mad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
VL_HD uint32_t vl_dot4(uint32_t a, uint32_t b, uint32_t c) { return uint32_t(__dp4a(int(a), int(b), int(c))); }   // c + sum of int8 products (IDP4A, FMA pipe)
#else
VL_HD uint32_t vl_min(uint32_t a, uint32_t b) {
    const uint32_t al = a & 0xFFFFu, bl = b & 0xFFFFu, ah = a >> 16, bh = b >> 16;
    return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
}
VL_HD uint32_t vl_addmin(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t s = (((a & 0xFFFFu) + (b & 0xFFFFu)) & 0xFFFFu) | ((((a >> 16) + (b >> 16)) & 0xFFFFu) << 16);
    return vl_min(s, c);
}
VL_HD uint32_t vl_prmt(uint32_t a, uint32_t b, uint32_t s) {
    const uint64_t v = (uint64_t(b) << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= uint32_t((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
}
VL_HD uint32_t vl_absdiff4(uint32_t bt, uint32_t w) {
    uint32_t e = 0;
    for (int r = 0; r < 4; r++) {
        const int a = int8_t(bt >> (8 * r)), b = int8_t(w >> (8 * r));
        e += uint32_t(a > b ? a - b : b - a);
    }
    return e;
}
VL_HD uint32_t vl_mad(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }
VL_HD uint32_t vl_dot4(uint32_t a, uint32_t b, uint32_t c) {
    int32_t acc = int32_t(c);
    for (int r = 0; r < 4; r++) acc += int32_t(int8_t(a >> (8 * r))) * int32_t(int8_t(b >> (8 * r)));
    return uint32_t(acc);
}
#endif

// ---- compile-time geometry ------------------------------------------------------------------------------------------
VL_HD constexpr uint32_t vl_parity(uint32_t v) { return ((v >> 6) ^ (v >> 5) ^ (v >> 4) ^ (v >> 3) ^ (v >> 2) ^ (v >> 1) ^ v) & 1u; }
// which of the 8 distinct branch words butterfly j uses: bit r = parity((j << 1) & G[r]), G = {109, 79, 83, 109}
// (dab_viterbi_decoder.cpp:18-25, viterbi_branch_table.h:45-55; outputs 0 and 3 share a polynomial)
VL_HD constexpr uint32_t vl_pat(uint32_t j) {
    return vl_parity((j << 1) & 109u) | (vl_parity((j << 1) & 79u) << 1) | (vl_parity((j << 1) & 83u) << 2);
}
// the four expected outputs (+127 / -127 as int8) of pattern p, packed like the received symbol word
VL_HD constexpr uint32_t vl_bt(uint32_t p) {
    return ((p & 1u) ? 0x7Fu : 0x81u) | (((p & 2u) ? 0x7Fu : 0x81u) << 8) | (((p & 4u) ? 0x7Fu : 0x81u) << 16) | (((p & 1u) ? 0x7Fu : 0x81u) << 24);
}
// minus the signs of pattern p's expected symbols, one int8 each: for |w| <= 127 the branch error is
// sum |(+-127) - w| = 508 - sum sign * w = 508 + dot(vl_negsign(p), w)
VL_HD constexpr uint32_t vl_negsign(uint32_t p) {
    return ((p & 1u) ? 0xFFu : 0x01u) | (((p & 2u) ? 0xFFu : 0x01u) << 8) | (((p & 4u) ? 0xFFu : 0x01u) << 16) | (((p & 1u) ? 0xFFu : 0x01u) << 24);
}
VL_HD constexpr uint32_t vl_insert0(uint32_t i, uint32_t k) { return ((i >> k) << (k + 1u)) | (i & ((1u << k) - 1u)); }

struct VlState {
    uint32_t R[32];      // path metrics, layout 0 between calls of vl_step5
    int32_t off;         // ref metric = rel + off
    uint32_t CL;         // saturation level in the rel domain, both halves
    int32_t thr;         // rel[0] >= thr  <=>  the reference renormalises
    uint64_t acc_err;    // what the reference has subtracted so far (DAB_Viterbi_Decoder accumulates it, dab_viterbi_decoder.cpp:109-129)
};

VL_HD void vl_set_off(VlState& S, int32_t off) {
    S.off = off;
    const int32_t cl = 65535 - off;
    S.CL = uint32_t(cl > 32767 ? 32767 : cl) * 0x10001u;
    S.thr = VL_RENORM - off;
}

// ViterbiDecoder_Core::reset (viterbi_decoder_core.h:202-211): start state 0 at error 0, every other state at 5080
VL_HD void vl_reset(VlState& S) {
#pragma unroll
    for (int i = 0; i < 32; i++) S.R[i] = (VL_ORIGIN + VL_NONSTART) * 0x10001u;
    S.R[0] = VL_ORIGIN | ((VL_ORIGIN + VL_NONSTART) << 16);
    S.acc_err = 0;
    vl_set_off(S, -int32_t(VL_ORIGIN));
}

// Branch errors of one trellis step, paired for layout K.  w = the four soft symbols of the step (int8 each, punctured = 0).
// E[p] = (e[p], e[p ^ dK]) and Ei[p] = (1016 - e, clamped at 0) for the two butterflies that share a register.
// Independent of the path metrics: vl_step5 issues it one step ahead so that it overlaps the tail of the previous step.
//
// M128 = false: the caller guarantees that no symbol of the call is -128 (true for everything the OFDM stage produces; the
// de-puncture pass looks).  Then |b - w| + |-b - w| = 254 for b = +-127, so the error of the complementary pattern is exactly
// the inverse error, 1016 - e[p] = e[p ^ 7] with no clamp to apply: Ei[p] = E[p ^ 7], and the eight subtractions, eight clamps
// and eight pairings of the general form (24 of 264 instructions per step) are not needed.  Ei is left untouched.
template <int K, bool M128>
VL_HD void vl_branch(const uint32_t w, uint32_t (&E)[8], uint32_t (&Ei)[8], const VlConst kc) {
    uint32_t e[8], ei[8];
#pragma unroll
    for (int p = 0; p < 8; p++) {
        // short form: the error as a dot product (FMA pipe; VABSDIFF4 shares the ALU pipe with the add-compare-select)
        e[p] = M128 ? vl_absdiff4(vl_bt(uint32_t(p)), w) : vl_dot4(w, vl_negsign(uint32_t(p)), 508u);
        if (M128) {
            const int32_t inv = int32_t(vl_mad(e[p], kc.m1, VL_MAX_ERROR));   // 1016 - e
            ei[p] = uint32_t(inv < 0 ? 0 : inv);
        }
    }
    constexpr uint32_t dk = vl_pat(1u << K);   // pattern difference between the two butterflies sharing a register
#pragma unroll
    for (int p = 0; p < 8; p++) {
        E[p] = vl_mad(e[uint32_t(p) ^ dk], kc.x10000, e[p]);
        if (M128) Ei[p] = vl_mad(ei[uint32_t(p) ^ dk], kc.x10000, ei[p]);
    }
}

// sum of z[i] << i over 16 words whose halves are 0/1, as a depth-4 tree of multiply-adds
VL_HD uint32_t vl_gather16(const uint32_t (&z)[16], const VlConst kc) {
    uint32_t a[8], b[4], c[2];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = vl_mad(z[2 * i + 1], kc.two, z[2 * i]);
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = vl_mad(a[2 * i + 1], kc.x4, a[2 * i]);
#pragma unroll
    for (int i = 0; i < 2; i++) c[i] = vl_mad(b[2 * i + 1], kc.x16, b[2 * i]);
    return vl_mad(c[1], kc.x256, c[0]);
}

// One trellis step from layout K to layout K+1 (add-compare-select of the 32 butterflies).
// d0/d1 receive the decision bits of the even/odd new states: new state n = 2j + (n & 1), i = j with bit K removed,
// position i + 16 * (bit K of j).
template <int K, bool M128>
VL_HD void vl_acs(uint32_t (&R)[32], const uint32_t (&E)[8], const uint32_t (&Ei)[8], const uint32_t CL, uint32_t& d0, uint32_t& d1, const VlConst kc) {
    uint32_t Rn[32], z0[16], z1[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t p = vl_pat(vl_insert0(uint32_t(i), uint32_t(K)));
        const uint32_t M0 = R[i], M1 = R[i + 16];
        const uint32_t ep = E[p], eip = M128 ? Ei[p] : E[p ^ 7u];
        const uint32_t B = vl_addmin(M1, eip, CL);         // upper predecessor, saturating like the reference
        const uint32_t N0 = vl_addmin(M0, ep, B);          // new[2j]
        z0[i] = vl_min(vl_mad(N0, kc.m1, B), 0x00010001u);   // 0 where the upper path won or tied
        const uint32_t D = vl_addmin(M1, ep, CL);
        const uint32_t N1 = vl_addmin(M0, eip, D);         // new[2j+1]
        z1[i] = vl_min(vl_mad(N1, kc.m1, D), 0x00010001u);
        Rn[2 * i] = N0;
        Rn[2 * i + 1] = N1;
    }
#pragma unroll
    for (int i = 0; i < 32; i++) R[i] = Rn[i];
    d0 = ~vl_gather16(z0, kc);
    d1 = ~vl_gather16(z1, kc);
}

// layout 5 (s, s+32) -> layout 0 (2i, 2i+1)
VL_HD void vl_repack(uint32_t (&R)[32]) {
    uint32_t Rn[32];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        Rn[i] = vl_prmt(R[2 * i], R[2 * i + 1], 0x5410u);
        Rn[16 + i] = vl_prmt(R[2 * i], R[2 * i + 1], 0x7632u);
    }
#pragma unroll
    for (int i = 0; i < 32; i++) R[i] = Rn[i];
}

VL_HD uint32_t vl_min_all(const uint32_t (&R)[32]) {
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = vl_min(R[i], R[i + 16]);
#pragma unroll
    for (int s = 8; s >= 1; s >>= 1)
#pragma unroll
        for (int i = 0; i < s; i++) m[i] = vl_min(m[i], m[i + s]);
    const uint32_t lo = m[0] & 0xFFFFu, hi = m[0] >> 16;
    return lo < hi ? lo : hi;
}

// Bookkeeping after the step with index t (0-based): the reference's renormalisation test and the capture of the path
// error of a trellis that ends at step n_steps - 1.  State 0 is the low half of register 0 in every layout.
// final_rel = metric of state 0 after the last step, in the reference's domain; the renormalisations stop with the trellis, so
// the path error chainback returns is S.acc_err + final_rel once the loop is over (vl_final_error).
VL_HD void vl_after_step(VlState& S, const uint32_t t, const uint32_t n_steps, uint32_t& final_rel) {
    const int32_t r0 = int32_t(S.R[0] & 0xFFFFu);
    if (r0 >= S.thr && t < n_steps) {
        // ViterbiDecoder_AVX_u16::renormalise (viterbi_decoder_avx_u16.h:138-170): subtract the minimum over the 64 states
        const int32_t mn = int32_t(vl_min_all(S.R)) + S.off;
        S.acc_err += uint64_t(uint32_t(mn));
        vl_set_off(S, S.off - mn);
    }
    if (t + 1u == n_steps) final_rel = uint32_t(r0 + S.off);
}
VL_HD uint64_t vl_final_error(const VlState& S, const uint32_t final_rel) { return S.acc_err + uint64_t(final_rel); }

// Five trellis steps t0 .. t0+4 starting and ending in layout 0.  emit(k, d0, d1) receives the decision words of step t0+k as
// soon as they exist (the kernel stores them at once: ten live registers less than keeping them to the end of the iteration).
template <bool M128, class Emit>
VL_HD void vl_step5_emit(VlState& S, const uint32_t (&w)[VL_UNROLL], const uint32_t t0, const uint32_t n_steps, Emit&& emit, uint32_t& final_rel,
                         const VlConst kc) {
    uint32_t Ea[8], Eia[8], Eb[8], Eib[8], d0, d1;
    if (!M128) {
#pragma unroll
        for (int p = 0; p < 8; p++) { Eia[p] = 0u; Eib[p] = 0u; }    // never read
    }
    vl_branch<0, M128>(w[0], Ea, Eia, kc);
    vl_acs<0, M128>(S.R, Ea, Eia, S.CL, d0, d1, kc); emit(0, d0, d1); vl_branch<1, M128>(w[1], Eb, Eib, kc); vl_after_step(S, t0 + 0u, n_steps, final_rel);
    vl_acs<1, M128>(S.R, Eb, Eib, S.CL, d0, d1, kc); emit(1, d0, d1); vl_branch<2, M128>(w[2], Ea, Eia, kc); vl_after_step(S, t0 + 1u, n_steps, final_rel);
    vl_acs<2, M128>(S.R, Ea, Eia, S.CL, d0, d1, kc); emit(2, d0, d1); vl_branch<3, M128>(w[3], Eb, Eib, kc); vl_after_step(S, t0 + 2u, n_steps, final_rel);
    vl_acs<3, M128>(S.R, Eb, Eib, S.CL, d0, d1, kc); emit(3, d0, d1); vl_branch<4, M128>(w[4], Ea, Eia, kc); vl_after_step(S, t0 + 3u, n_steps, final_rel);
    vl_acs<4, M128>(S.R, Ea, Eia, S.CL, d0, d1, kc); emit(4, d0, d1); vl_after_step(S, t0 + 4u, n_steps, final_rel);
    vl_repack(S.R);
    // own renormalisation: bring rel[0] back to VL_ORIGIN (delta >= 0: metrics never decrease)
    const uint32_t delta = (S.R[0] & 0xFFFFu) - VL_ORIGIN;
    const uint32_t d2 = delta * 0x10001u;
#pragma unroll
    for (int i = 0; i < 32; i++) S.R[i] -= d2;
    vl_set_off(S, S.off + int32_t(delta));
}

// the same with the decision words returned: dec[2k], dec[2k+1] = step t0+k
template <bool M128 = true>
VL_HD void vl_step5(VlState& S, const uint32_t (&w)[VL_UNROLL], const uint32_t t0, const uint32_t n_steps, uint32_t (&dec)[2 * VL_UNROLL],
                    uint32_t& final_rel, const VlConst kc) {
    vl_step5_emit<M128>(S, w, t0, n_steps, [&](const int k, const uint32_t d0, const uint32_t d1) { dec[2 * k] = d0; dec[2 * k + 1] = d1; }, final_rel, kc);
}

// true if one of the four int8 symbols of a word is -128 (the exact-zero-byte test on w ^ 0x80808080)
VL_HD bool vl_has_m128(const uint32_t w) {
    const uint32_t t = w ^ 0x80808080u;
    return ((t - 0x01010101u) & ~t & 0x80808080u) != 0u;
}

// Traceback: the decision bit of new state n at a step whose index is k modulo 5.
VL_HD uint32_t vl_decision(const uint32_t d0, const uint32_t d1, const uint32_t n, const uint32_t k) {
    const uint32_t word = (n & 1u) ? d1 : d0;
    const uint32_t j = n >> 1;
    const uint32_t pos = ((j >> (k + 1u)) << k) | (j & ((1u << k) - 1u)) | (((j >> k) & 1u) << 4);
    return (word >> pos) & 1u;
}

// The same walk with the survivor state kept in the top six bits of a 32-bit history word h (state = h >> 26; the history is
// also the word of decoded bits under construction, newest bit on top): one traceback step is
//   h' = (h >> 1) | (decision(state) << 31).
// With j = h >> 27 the position above is (j with bit k moved to bit 4) = ((h >> 28) & ~m) | ((h >> 27) & m) | ((h >> (23 + k)) & 16),
// m = 2^k - 1: three shifts and two 3-input logic operations; only bit 0 of the shifted decision word enters the funnel shift.
template <uint32_t K>
VL_HD uint32_t vl_traceback_step(const uint32_t d0, const uint32_t d1, const uint32_t h) {
    const uint32_t word = (h & (1u << 26)) ? d1 : d0;
    constexpr uint32_t m = (1u << K) - 1u;
    const uint32_t pos = (((h >> 28) & ~m) | ((h >> 27) & m) | ((h >> (23u + K)) & 16u)) & 31u;
    const uint32_t x = word >> pos;
#if defined(VL_DEVICE_CODE)
    return __funnelshift_r(h, x, 1u);
#else
    return (h >> 1) | (x << 31);
#endif
}
