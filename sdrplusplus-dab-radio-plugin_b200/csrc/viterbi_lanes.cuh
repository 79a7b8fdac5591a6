// viterbi_lanes.cuh -- batch Viterbi: one LANE per trellis (32 trellises per warp), for calls that carry thousands of
// trellises (hundreds of streams per GPU).  Small calls (one receiver = 76 trellises per frame) keep the warp-per-trellis
// kernel of viterbi.cuh, which is the latency-oriented mapping; both produce the reference's bits exactly.
//
// Replaces the same reference code as viterbi.cuh (VIT/x86/viterbi_decoder_avx_u16.h:47-170, viterbi_decoder_core.h:214-236,
// dab/algorithms/dab_viterbi_decoder.cpp:109-181); the arithmetic is in viterbi_lane_core.h.
//
// Pipeline of one call (all on the context's stream, no host round trip):
//   k_vit_count    histogram of the active trellises by length class (dabgpu_chan_decode takes it in k_chan_build_jobs instead)
//   k_vit_plan     one thread: decides lanes vs warps, orders the classes longest first, assigns symbol rows
//   k_vit_scatter  job indices grouped by class (32 consecutive entries = one warp's trellises)
//   k_vit_prep     time de-interleave + de-puncture into a [step][lane] word matrix per group (one lane per trellis, like the
//                  decoder), so that the decoder reads one coalesced 128-byte row per step
//   k_viterbi_lanes  persistent warps: forward pass (decisions to a per-warp scratch, 256 B coalesced per step), traceback,
//                  energy dispersal, FIB CRC
#pragma once
#include "viterbi.cuh"
#include "viterbi_lane_core.h"

#define VL_BUCKETS 64
#define VL_MAX_STEPS 73728u   // longest trellis the length classes cover (a full CIF at the weakest code is < 50 000 steps)
#define VL_PAD_ROWS 16u       // rounding to the unroll factor + one prefetched iteration

__host__ __device__ inline uint32_t vl_bucket(uint32_t steps) { return steps < 8192u ? (steps >> 8) : 32u + ((steps - 8192u) >> 11); }
// exclusive upper bound of the trellis lengths of a class
__host__ __device__ inline uint32_t vl_bucket_cap(uint32_t b) { return b < 32u ? (b + 1u) * 256u : 8192u + (b - 31u) * 2048u; }
__host__ __device__ inline uint32_t vl_bucket_rows(uint32_t b) { return vl_bucket_cap(b) + VL_PAD_ROWS; }

struct VlPlan {
    uint32_t use_lanes, n_groups, n_active, oversize;   // k_viterbi reads words 0, 2 and 3 through a plain uint32_t pointer
    uint32_t next_group;              // work counter of k_viterbi_lanes
    uint32_t has_m128;                // k_vit_prep found a -128 symbol in the call: the decoder takes the general branch-error form
    uint32_t pad_[2];
    uint32_t count[VL_BUCKETS];       // trellises per class
    uint32_t cursor[VL_BUCKETS];      // scatter cursors
    uint32_t list_base[VL_BUCKETS];   // first entry of the class in the job list
    uint32_t group_base[VL_BUCKETS];  // first group of the class (classes ordered longest first)
    uint32_t row_base[VL_BUCKETS];    // first symbol row of the class
};

// Histogram entry of one trellis per lane (0 steps = none), called by whole warps: the lanes of a class elect a leader that adds
// their number with one atomic.  A call carries two or three classes, one atomic per trellis serialised 78 000 of them on two
// addresses (55 us per 1024-stream step).
__device__ __forceinline__ void vl_count_warp(VlPlan* __restrict__ plan, const uint32_t steps) {
    uint32_t b = 0xFFFFFFFFu;
    if (steps >= VL_MAX_STEPS) plan->oversize = 1u;
    else if (steps != 0u) b = vl_bucket(steps);
    const uint32_t peers = __match_any_sync(FULL_MASK, b);
    if (b != 0xFFFFFFFFu && (threadIdx.x & 31u) == uint32_t(__ffs(int(peers)) - 1)) atomicAdd(&plan->count[b], uint32_t(__popc(peers)));
}

__global__ void k_vit_count(const VitJobDev* __restrict__ jobs, const int n_jobs, VlPlan* __restrict__ plan) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;    // blockDim is a multiple of 32: whole warps reach vl_count_warp
    vl_count_warp(plan, gid < n_jobs ? jobs[gid].total_steps : 0u);
}

// mode: 0 = lanes when at least min_jobs trellises are active, 1 = always, 2 = never
__global__ void k_vit_plan(VlPlan* __restrict__ plan, const int mode, const uint32_t min_jobs, const uint32_t cap_rows, const uint32_t cap_groups) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t groups = 0, rows = 0, list = 0, active = 0;
    for (int b = VL_BUCKETS - 1; b >= 0; --b) {
        const uint32_t c = plan->count[b];
        const uint32_t g = (c + 31u) >> 5;
        plan->group_base[b] = groups;
        plan->row_base[b] = rows;
        plan->list_base[b] = list;
        plan->cursor[b] = 0u;
        groups += g;
        rows += g * vl_bucket_rows(uint32_t(b));
        list += c;
        active += c;
    }
    bool use = (mode == 1) || (mode == 0 && active >= min_jobs);
    if (plan->oversize || rows > cap_rows || groups > cap_groups || active == 0u) use = false;   // the warp kernel takes the call
    plan->n_active = active;
    plan->use_lanes = use ? 1u : 0u;
    plan->n_groups = use ? groups : 0u;
    plan->next_group = 0u;
}

__global__ void k_vit_scatter(const VitJobDev* __restrict__ jobs, const int n_jobs, VlPlan* __restrict__ plan, uint32_t* __restrict__ list) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (plan->oversize) return;     // a trellis beyond the classes: no list (k_viterbi then takes the jobs in index order)
    const uint32_t steps = gid < n_jobs ? jobs[gid].total_steps : 0u;
    const uint32_t b = steps != 0u ? vl_bucket(steps) : 0xFFFFFFFFu;
    // one cursor atomic per class and warp (see vl_count_warp); the lanes of the class take consecutive entries
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t peers = __match_any_sync(FULL_MASK, b);
    const uint32_t leader = uint32_t(__ffs(int(peers)) - 1);
    uint32_t base = 0;
    if (b != 0xFFFFFFFFu && lane == leader) base = atomicAdd(&plan->cursor[b], uint32_t(__popc(peers)));
    base = __shfl_sync(FULL_MASK, base, int(leader));
    if (b != 0xFFFFFFFFu) list[plan->list_base[b] + base + uint32_t(__popc(peers & ((1u << lane) - 1u)))] = uint32_t(gid);
}

// group index -> class, first list entry, number of trellises, first symbol row.  Called by whole warps: every lane tests two
// classes, so the plan is read with one round trip instead of a serial walk over the 64 classes.
__device__ __forceinline__ void vl_locate(const VlPlan* __restrict__ plan, const uint32_t g, uint32_t& list0, uint32_t& n_in, uint32_t& row0) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t hit = 0xFFFFFFFFu;
#pragma unroll
    for (uint32_t h = 0; h < 2u; h++) {
        const uint32_t i = lane + 32u * h;
        const uint32_t c = plan->count[i], gb = plan->group_base[i];
        if (c != 0u && g >= gb && g < gb + ((c + 31u) >> 5)) hit = i;
    }
    const uint32_t b = __reduce_min_sync(FULL_MASK, hit) & (VL_BUCKETS - 1u);   // exactly one class contains g
    const uint32_t gi = g - plan->group_base[b];
    list0 = plan->list_base[b] + gi * 32u;
    n_in = min(32u, plan->count[b] - gi * 32u);
    row0 = plan->row_base[b] + gi * vl_bucket_rows(b);
}

// ---------------------------------------------------------------------------------------------------------------------------
// k_vit_prep: de-puncture (DAB_Viterbi_Decoder::depuncture_symbols, dab_viterbi_decoder.cpp:131-181) of a whole call into the
// [step][lane] symbol matrix the decoder reads.  The punctured symbols of every trellis are contiguous in memory: FIC groups
// and dabgpu_viterbi_decode jobs by nature, MSC sub-channels because k_chan_deinterleave (chan.cuh) has undone the time
// interleaver for the whole CIF.
//
// One warp per group, one LANE per trellis -- the same mapping as the decoder, so a warp stores whole 128-byte rows of the
// matrix.  Every lane keeps a window of 256 consecutive punctured symbols of its trellis in shared memory ([word][lane], bank =
// lane) as a ring of two units of 128 bytes (128-byte aligned in memory); the first words of the ring are repeated behind its
// end, so that a code period (at most 32 symbols) is read without wrapping.  8.5 KB per warp: 24 warps per SM.
//   A. refill, when the next VP_TILE steps may reach beyond the window: 8 independent LDG.128 per lane; each byte is read once.
//   B. de-puncture, one code period of 8 steps at a time, branch-free: the kept-symbol counts and offsets of the period come
//      from the code's two shift registers, a step's word is two LDS.32 + a funnel shift + a mask, and goes straight to the matrix.
// The first version fetched single bytes through the 16-way strided gather of a natural-order frame ring (one dependent L2
// round trip per step, 63 instructions per step and lane, every DRAM sector read six times).
// ---------------------------------------------------------------------------------------------------------------------------
#ifndef VP_UNIT
#define VP_UNIT 64u            // bytes per refill (64: 5.2 KB of window per warp, ten CTAs per SM; 128 was 0.03 ms slower per 1024-stream step)
#endif
#define VP_TILE (VP_UNIT / 4u) // steps per tile: a tile reads at most one unit's worth of symbols from the position of its first step
#define VP_WARPS 4u
#define VP_WORDS (VP_UNIT / 2u) // window of 2 units per lane, in words
#define VP_MIRROR 9u
#define VP_SMEM_BYTES (size_t(VP_WARPS) * (VP_WORDS + VP_MIRROR) * 32u * sizeof(uint32_t))

// puncturing state of one trellis: the segment that contains the current step
struct PrepSeg { uint32_t seg, seg_end, start, inb, cntw, K, pref_lo, pref_hi; };
__device__ __forceinline__ void prep_seg_load(PrepSeg& S, const VitJobDev* __restrict__ J, const uint32_t seg) {
    const uint32_t pi = J->seg_pi[seg];
    S.seg = seg;
    S.seg_end = J->seg_step_end[seg];
    S.start = seg ? J->seg_step_end[seg - 1u] : 0u;
    S.inb = J->seg_in_base[seg];
    S.cntw = c_pi_cnt[pi];
    S.K = c_pi_K[pi];
    S.pref_lo = uint32_t(c_pi_pref[pi]);
    S.pref_hi = uint32_t(c_pi_pref[pi] >> 32);
}
// index of the first punctured symbol of step t (t inside the segment)
__device__ __forceinline__ uint32_t prep_in_index(const PrepSeg& S, const uint32_t t) {
    const uint32_t u = t - S.start, g8 = u & 7u;
    const uint32_t pre = (((g8 & 4u) ? S.pref_hi : S.pref_lo) >> (8u * (g8 & 3u))) & 0xFFu;
    return S.inb + (u >> 3) * S.K + pre;
}

// mask of cnt bytes, cnt = 0..4: the clamped funnel shift gives 0 for a shift of 32
__device__ __forceinline__ uint32_t prep_mask(const uint32_t cnt) { return __funnelshift_rc(0xFFFFFFFFu, 0u, 32u - 8u * cnt); }
// the word of one step: `cnt` symbols from window position lam (counted in the lane's coordinate), the rest zero (punctured)
__device__ __forceinline__ uint32_t prep_word(const uint32_t (*L)[32], const uint32_t lane, const uint32_t lam, const uint32_t cnt) {
    const uint32_t wi = (lam >> 2) & (VP_WORDS - 1u);
    return __funnelshift_r(L[wi][lane], L[wi + 1u][lane], lam * 8u) & prep_mask(cnt);   // the shift uses the low 5 bits: (lam & 3) * 8
}

// `split` warps share a group: each takes a contiguous range of tiles.  A group alone keeps one warp busy for 1542 steps of a
// dependent refill -> de-puncture -> store chain, and a call has only as many groups as a quarter of the warps the GPU can
// hold; the step -> input index map (prep_in_index) is random access, so the walk can start anywhere.
__global__ void __launch_bounds__(VP_WARPS * 32)
k_vit_prep(const VitJobDev* __restrict__ jobs, VlPlan* __restrict__ plan, const uint32_t* __restrict__ list, uint32_t* __restrict__ sym, const GatherGeom G,
           const uint32_t split) {
    extern __shared__ __align__(16) uint32_t s_log_raw[];     // [VP_WARPS][VP_WORDS + VP_MIRROR][32]
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * VP_WARPS + w;
    const uint32_t g = gw / split, part = gw - g * split;
    if (g >= plan->n_groups) return;
    uint32_t list0, n_in, row0;
    vl_locate(plan, g, list0, n_in, row0);
    const bool have = lane < n_in;
    const VitJobDev* __restrict__ J = jobs + (have ? list[list0 + lane] : 0u);
    const uint32_t N = have ? J->total_steps : 0u;
    const uint32_t steps_g = __reduce_max_sync(FULL_MASK, N);
    const uint32_t padded = ((steps_g + VL_UNROLL - 1u) / VL_UNROLL) * VL_UNROLL + VL_UNROLL;
    uint32_t (*L)[32] = reinterpret_cast<uint32_t (*)[32]>(s_log_raw + size_t(w) * (VP_WORDS + VP_MIRROR) * 32u);
    // unit u of the window = the VP_UNIT bytes at base + VP_UNIT * u; positions are counted from `base` (lane coordinate)
    const uint8_t* base = nullptr;
    uint32_t origin = 0;                     // lane coordinate of the punctured symbol 0
    if (have) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(J->src);
        base = reinterpret_cast<const uint8_t*>(a & ~uintptr_t(VP_UNIT - 1u));
        origin = uint32_t(a & (VP_UNIT - 1u));
    }
    PrepSeg S;
    S.seg_end = 0u; S.start = 0u; S.inb = 0u; S.cntw = 0u; S.K = 0u; S.pref_lo = 0u; S.pref_hi = 0u; S.seg = 0u;
    uint32_t mask8[8];                       // byte masks of the 8 steps of the current segment's code period
#pragma unroll
    for (uint32_t g8 = 0; g8 < 8u; g8++) mask8[g8] = 0u;
    uint32_t mask_seg = 0xFFFFFFFFu;         // segment mask8 was built for
    bool odd = false;   // a segment that does not start on a code period (only dabgpu_viterbi_decode can build one: 128-bit blocks otherwise)
    if (have) {
        uint32_t seg = 0;
        while (seg < DABGPU_MAX_SEGMENTS - 1u && J->seg_step_end[seg] == 0u) seg++;   // skip empty leading segments
        prep_seg_load(S, J, seg);
        for (uint32_t k = 0; k + 1u < DABGPU_MAX_SEGMENTS; k++) odd = odd || ((J->seg_step_end[k] & 7u) != 0u && J->seg_step_end[k] < N);
    }
    const bool any_odd = __any_sync(FULL_MASK, odd);
    // this warp's tiles
    const uint32_t n_tiles = (padded + VP_TILE - 1u) / VP_TILE, per_part = (n_tiles + split - 1u) / split;
    const uint32_t t_begin = part * per_part * VP_TILE, t_end = min(padded, (part + 1u) * per_part * VP_TILE);
    uint32_t next_unit = 0u;                 // units below this one are in the window (the last two of them)
    if (have) {
        if (t_begin < N) {
            while (t_begin >= S.seg_end && S.seg < DABGPU_MAX_SEGMENTS - 1u) prep_seg_load(S, J, S.seg + 1u);
            next_unit = (origin + prep_in_index(S, t_begin)) / VP_UNIT;
        } else {
            next_unit = 0xFFFFFFFFu;         // beyond the end of this trellis: zeros only, nothing to fetch
        }
    }
    uint32_t* __restrict__ dst = sym + size_t(row0) * 32u + lane;
    bool m128 = false;                       // a -128 symbol went by (vl_branch: the decoder's short form needs |symbol| <= 127)

#pragma unroll 1
    for (uint32_t t0 = t_begin; t0 < t_end; t0 += VP_TILE) {
        // ---- A. refill: the tile reads at most 4 * VP_TILE = 128 symbols from the position of its first step ----
        uint32_t need_unit = 0;
        if (t0 < N) {
            while (t0 >= S.seg_end && S.seg < DABGPU_MAX_SEGMENTS - 1u) prep_seg_load(S, J, S.seg + 1u);
            need_unit = (origin + prep_in_index(S, t0) + 4u * VP_TILE - 1u) / VP_UNIT + 1u;
        }
        while (__any_sync(FULL_MASK, next_unit < need_unit)) {
            if (next_unit < need_unit) {
                const uint32_t u = next_unit++, h = (u & 1u) * (VP_UNIT / 4u);
                uint4 P[VP_UNIT / 16u];
#pragma unroll
                for (uint32_t j = 0; j < VP_UNIT / 16u; j++) P[j] = __ldg(reinterpret_cast<const uint4*>(base + (size_t(u) * VP_UNIT + 16u * j)));
#pragma unroll
                for (uint32_t j = 0; j < VP_UNIT / 16u; j++) {
                    const uint32_t wbase = h + 4u * j;
                    L[wbase + 0u][lane] = P[j].x; L[wbase + 1u][lane] = P[j].y; L[wbase + 2u][lane] = P[j].z; L[wbase + 3u][lane] = P[j].w;
                    // the -128 test runs on the punctured input (at most as many words as steps, half as many at rate 1/2); the
                    // unit may hold a few bytes of a neighbouring trellis: a false alarm only selects the general decoder form
                    m128 = m128 || vl_has_m128(P[j].x) || vl_has_m128(P[j].y) || vl_has_m128(P[j].z) || vl_has_m128(P[j].w);
                }
                if (h == 0u) {
#pragma unroll
                    for (uint32_t k = 0; k < VP_MIRROR; k++) L[VP_WORDS + k][lane] = L[k][lane];
                }
            }
        }
        __syncwarp();
        // ---- B. de-puncture VP_TILE steps ----
        if (any_odd) {   // general form, step by step
#pragma unroll 1
            for (uint32_t i = 0; i < VP_TILE; i++) {
                const uint32_t t = t0 + i;
                if (t >= padded) break;
                uint32_t word = 0;
                if (t < N) {
                    while (t >= S.seg_end && S.seg < DABGPU_MAX_SEGMENTS - 1u) prep_seg_load(S, J, S.seg + 1u);
                    const uint32_t g8 = (t - S.start) & 7u;
                    word = prep_word(L, lane, origin + prep_in_index(S, t), (S.cntw >> (4u * g8)) & 0xFu);
                }
                dst[size_t(t) * 32u] = word;
            }
            __syncwarp();
            continue;
        }
        // one code period (8 steps) at a time: every segment starts on a period (128-bit blocks = 32 steps)
#pragma unroll 1
        for (uint32_t p8 = 0; p8 < VP_TILE; p8 += 8u) {
            const uint32_t tp = t0 + p8;
            if (tp >= padded) break;                       // warp-uniform
            uint32_t words[8];
            if (tp + 8u <= N) {
                while (tp >= S.seg_end && S.seg < DABGPU_MAX_SEGMENTS - 1u) prep_seg_load(S, J, S.seg + 1u);
                if (mask_seg != S.seg) {                   // new code: the byte masks of its period (kept-symbol counts per step)
                    mask_seg = S.seg;
#pragma unroll
                    for (uint32_t g8 = 0; g8 < 8u; g8++) mask8[g8] = prep_mask((S.cntw >> (4u * g8)) & 0xFu);
                }
                const uint32_t lam = origin + S.inb + ((tp - S.start) >> 3) * S.K;
                const uint32_t* __restrict__ row = &L[(lam >> 2) & (VP_WORDS - 1u)][lane];
                const uint32_t b0 = lam & 3u;
#pragma unroll
                for (uint32_t g8 = 0; g8 < 8u; g8++) {
                    const uint32_t pre = (g8 & 4u) ? __byte_perm(S.pref_hi, 0u, 0x4440u + (g8 & 3u)) : __byte_perm(S.pref_lo, 0u, 0x4440u + g8);
                    const uint32_t b = b0 + pre;           // byte offset from `row`, < 36: inside the ring or its mirror
                    const uint32_t* __restrict__ q = row + (b >> 2) * 32u;
                    words[g8] = __funnelshift_r(q[0], q[32], b * 8u) & mask8[g8];
                }
            } else {
                // the last period of the trellis (tail bits) or beyond it: step by step, zero after the end
#pragma unroll
                for (uint32_t g8 = 0; g8 < 8u; g8++) {
                    const uint32_t t = tp + g8;
                    uint32_t word = 0;
                    if (t < N) {
                        while (t >= S.seg_end && S.seg < DABGPU_MAX_SEGMENTS - 1u) prep_seg_load(S, J, S.seg + 1u);
                        word = prep_word(L, lane, origin + prep_in_index(S, t), (S.cntw >> (4u * ((t - S.start) & 7u))) & 0xFu);
                    }
                    words[g8] = word;
                }
            }
#pragma unroll
            for (uint32_t g8 = 0; g8 < 8u; g8++)
                if (tp + g8 < padded) dst[size_t(tp + g8) * 32u] = words[g8];
        }
        __syncwarp();
    }
    if (__any_sync(FULL_MASK, m128) && lane == 0u) atomicOr(&plan->has_m128, 1u);
}

__device__ __forceinline__ void vl_emit_word(uint8_t* __restrict__ out, const uint32_t n_out_bytes, const uint32_t widx, const uint32_t v) {
    const uint32_t b0 = widx * 4u;
    if (b0 + 4u <= n_out_bytes && ((reinterpret_cast<uintptr_t>(out) & 3u) == 0)) {
        *reinterpret_cast<uint32_t*>(out + b0) = __byte_perm(v, 0u, 0x0123);
    } else {
#pragma unroll
        for (uint32_t q = 0; q < 4; q++)
            if (b0 + q < n_out_bytes) out[b0 + q] = uint8_t(v >> (24u - 8u * q));
    }
}

// Whole-block traceback from state 0 (viterbi_decoder_core.h:214-236): decoded bit b comes from the decision word of
// step b + 6.  All lanes walk the same rows (coalesced 256 B loads, two batches of rows in flight); a lane joins when the
// walk reaches the last decision word it needs (its history word stays 0 = state 0 until then).  The survivor state is the
// top of the history word h, which is at the same time the word of decoded bits being assembled (vl_traceback_step): 13
// instructions per step where the first version, with separate state and accumulator and a generic bit position, took 45.
// Rows 5..0 are walked as well (their bits are never emitted): no lower bound to test.
// Then energy dispersal (additive_scrambler.h:10-36) and the FIB CRCs.
// ALL: every lane that holds a trellis joins at the first row (the usual group: equal lengths), so the step needs no test
template <uint32_t K, bool ALL>
__device__ __forceinline__ void vl_tb_row(const uint2 d, const uint32_t t, const uint32_t join_t, uint32_t& h, const uint32_t n_out_bytes, const uint32_t flags,
                                          uint8_t* __restrict__ const out, const uint32_t* __restrict__ prbs_words) {
    if (ALL || t < join_t) h = vl_traceback_step<K>(d.x, d.y, h);
    if (((t - 6u) & 31u) == 0u && t >= 6u) {      // warp-uniform: 32 bits complete, bit b = t - 6 on top
        const uint32_t widx = (t - 6u) >> 5;
        uint32_t v = h;
        if (flags & VJ_DESCRAMBLE) v ^= prbs_words[widx];
        if (t < join_t) vl_emit_word(out, n_out_bytes, widx, v);
    }
}

template <uint32_t VL_TB_BLOCK, bool ALL>
__device__ __forceinline__ void vl_tb_block(const uint2 (&cur)[VL_TB_BLOCK], const uint32_t t0, const uint32_t join_t, uint32_t& h, const uint32_t n_out_bytes,
                                            const uint32_t flags, uint8_t* __restrict__ const out, const uint32_t* __restrict__ prbs_words) {
#pragma unroll
    for (int r = int(VL_TB_BLOCK) - 1; r >= 0; --r) {
        switch (r % VL_UNROLL) {
        case 0: vl_tb_row<0, ALL>(cur[r], t0 + uint32_t(r), join_t, h, n_out_bytes, flags, out, prbs_words); break;
        case 1: vl_tb_row<1, ALL>(cur[r], t0 + uint32_t(r), join_t, h, n_out_bytes, flags, out, prbs_words); break;
        case 2: vl_tb_row<2, ALL>(cur[r], t0 + uint32_t(r), join_t, h, n_out_bytes, flags, out, prbs_words); break;
        case 3: vl_tb_row<3, ALL>(cur[r], t0 + uint32_t(r), join_t, h, n_out_bytes, flags, out, prbs_words); break;
        default: vl_tb_row<4, ALL>(cur[r], t0 + uint32_t(r), join_t, h, n_out_bytes, flags, out, prbs_words); break;
        }
    }
}

template <uint32_t VL_TB_BLOCK>   // traceback rows fetched per batch (multiple of VL_UNROLL), two batches in flight
__device__ __forceinline__ void vl_traceback(const uint2* __restrict__ dec, const VitJobDev* __restrict__ J, const bool have, const uint32_t n_out_bytes,
                                             const uint32_t flags, uint8_t* __restrict__ const out, const uint32_t* __restrict__ prbs_words) {
    static_assert(VL_TB_BLOCK % VL_UNROLL == 0, "the layout index of a row must be a compile-time constant");
    const uint32_t nbits = n_out_bytes * 8u;
    const uint32_t top = __reduce_max_sync(FULL_MASK, nbits);      // rows top+5 .. 0 are walked
    if (top != 0u) {
        const uint32_t join_t = nbits != 0u ? nbits + 6u : 0u;
        // the usual group: every trellis has the same length, all lanes join at row top + 5.  Only the first block, which also
        // holds rows above it (padded steps of the forward pass), then needs the per-lane test.  Lanes without a trellis walk
        // along without storing anything.
        const bool same = __all_sync(FULL_MASK, join_t == 0u || join_t == top + 6u);
        uint32_t h = 0;
        const int first_block = int((top + 5u) / VL_TB_BLOCK);
        uint2 cur[VL_TB_BLOCK];
#pragma unroll
        for (int r = 0; r < int(VL_TB_BLOCK); r++) cur[r] = dec[size_t(uint32_t(first_block) * VL_TB_BLOCK + r) * 32u];
#pragma unroll 1
        for (int blk = first_block; blk >= 0; --blk) {
            uint2 nx[VL_TB_BLOCK];
            if (blk > 0) {
#pragma unroll
                for (int r = 0; r < int(VL_TB_BLOCK); r++) nx[r] = dec[size_t(uint32_t(blk - 1) * VL_TB_BLOCK + r) * 32u];
            }
            const uint32_t t0 = uint32_t(blk) * VL_TB_BLOCK;
            if (same && blk != first_block) vl_tb_block<VL_TB_BLOCK, true>(cur, t0, join_t, h, n_out_bytes, flags, out, prbs_words);
            else vl_tb_block<VL_TB_BLOCK, false>(cur, t0, join_t, h, n_out_bytes, flags, out, prbs_words);
            if (blk > 0) {
#pragma unroll
                for (int r = 0; r < int(VL_TB_BLOCK); r++) cur[r] = nx[r];
            }
        }
    }
    if (have && (flags & VJ_FIB_CRC)) {
        // FIB = 30 data bytes + CRC16, fic_decoder.cpp:98-116
        for (uint32_t f = 0; f < J->n_fibs; f++) {
            const uint8_t* fib = out + 32u * f;
            const uint16_t rx = uint16_t((uint16_t(fib[30]) << 8) | fib[31]);
            J->crc_ok[f] = (crc16_ccitt_dev(fib, 30) == rx) ? 1 : 0;
        }
    }
}

#define VL_WARPS_PER_BLOCK 4

// Forward pass of one group.  M128: the general branch-error form (a -128 symbol somewhere in the call); otherwise the short one.
template <bool M128>
__device__ __forceinline__ void vl_forward(const uint32_t* __restrict__ srow, uint2* __restrict__ dec, const uint32_t padded, const uint32_t N,
                                           unsigned long long* __restrict__ path_error, const VlConst kc) {
    VlState S;
    vl_reset(S);
    uint32_t final_rel = 0;
    uint32_t w[VL_UNROLL];
#pragma unroll
    for (int k = 0; k < VL_UNROLL; k++) w[k] = __ldg(srow + size_t(k) * 32u);
#pragma unroll 1
    for (uint32_t t0 = 0; t0 < padded; t0 += VL_UNROLL) {
        uint32_t wn[VL_UNROLL];
#pragma unroll
        for (int k = 0; k < VL_UNROLL; k++) wn[k] = __ldg(srow + size_t(t0 + VL_UNROLL + k) * 32u);
        uint2* __restrict__ drow = dec + size_t(t0) * 32u;
        vl_step5_emit<M128>(S, w, t0, N, [&](const int k, const uint32_t d0, const uint32_t d1) { drow[k * 32] = make_uint2(d0, d1); }, final_rel, kc);
#pragma unroll
        for (int k = 0; k < VL_UNROLL; k++) w[k] = wn[k];
    }
    if (path_error != nullptr) *path_error = vl_final_error(S, final_rel);
}

// Two builds of the same kernel.  <40, 1>: one CTA per SM = one decoder warp per SM sub-partition, for calls with up to one
// group per warp (256 streams x 76 trellises = 608 groups on 592 warps); the lone warp hides the latency of its traceback
// loads with 2 x 40 rows in flight (240 registers).  <10, 4>: 128 registers, four CTAs per SM, for larger calls: a single
// warp issues at most every other cycle, four warps per sub-partition bring the ACS loop from 540 to 300 cycles per step
// (scripts/micro/lane_fwd.cu) and hide the traceback latency by themselves.
template <uint32_t TB, int CTAS>
__global__ void __launch_bounds__(VL_WARPS_PER_BLOCK * 32, CTAS)
k_viterbi_lanes(const VitJobDev* __restrict__ jobs, VlPlan* __restrict__ plan, const uint32_t* __restrict__ list, const uint32_t* __restrict__ sym,
                uint2* __restrict__ scratch, const uint32_t scratch_rows, const uint32_t* __restrict__ prbs_words, const VlConst kc) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t slot = blockIdx.x * VL_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    uint2* __restrict__ dec = scratch + size_t(slot) * scratch_rows * 32u + lane;
    const bool m128 = plan->has_m128 != 0u;      // set by k_vit_prep, which has finished
    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(&plan->next_group, 1u);
        g = __shfl_sync(FULL_MASK, g, 0);
        if (g >= plan->n_groups) break;
        uint32_t list0, n_in, row0;
        vl_locate(plan, g, list0, n_in, row0);
        const bool have = lane < n_in;
        const VitJobDev* J = jobs + (have ? list[list0 + lane] : 0u);
        const uint32_t N = have ? J->total_steps : 0u;
        const uint32_t n_out_bytes = have ? J->n_out_bytes : 0u;
        const uint32_t flags = have ? J->flags : 0u;
        uint8_t* const out = have ? J->out : nullptr;
        const uint32_t steps_g = __reduce_max_sync(FULL_MASK, N);
        const uint32_t padded = ((steps_g + VL_UNROLL - 1u) / VL_UNROLL) * VL_UNROLL;
        const uint32_t* __restrict__ srow = sym + size_t(row0) * 32u + lane;

        // ---- forward pass: five trellis steps per iteration, symbols prefetched one iteration ahead ----
        if (m128) vl_forward<true>(srow, dec, padded, N, have ? J->path_error : nullptr, kc);
        else vl_forward<false>(srow, dec, padded, N, have ? J->path_error : nullptr, kc);

        vl_traceback<TB>(dec, J, have, n_out_bytes, flags, out, prbs_words);
        __syncwarp();
    }
}
