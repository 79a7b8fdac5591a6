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
//   k_vit_prep     time de-interleave + de-puncture (the loader of viterbi.cuh with the puncturing segment cached per trellis)
//                  into a [step][lane] word matrix per group, so that the decoder reads one coalesced 128-byte row per step
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
    uint32_t use_lanes, n_groups, n_active, oversize;
    uint32_t next_group;              // work counter of k_viterbi_lanes
    uint32_t pad_[3];
    uint32_t count[VL_BUCKETS];       // trellises per class
    uint32_t cursor[VL_BUCKETS];      // scatter cursors
    uint32_t list_base[VL_BUCKETS];   // first entry of the class in the job list
    uint32_t group_base[VL_BUCKETS];  // first group of the class (classes ordered longest first)
    uint32_t row_base[VL_BUCKETS];    // first symbol row of the class
};

__global__ void k_vit_count(const VitJobDev* __restrict__ jobs, const int n_jobs, VlPlan* __restrict__ plan) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_jobs) return;
    const uint32_t steps = jobs[gid].total_steps;
    if (steps == 0u) return;
    if (steps >= VL_MAX_STEPS) { plan->oversize = 1u; return; }
    atomicAdd(&plan->count[vl_bucket(steps)], 1u);
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
    if (gid >= n_jobs || !plan->use_lanes) return;
    const uint32_t steps = jobs[gid].total_steps;
    if (steps == 0u) return;
    const uint32_t b = vl_bucket(steps);
    list[plan->list_base[b] + atomicAdd(&plan->cursor[b], 1u)] = uint32_t(gid);
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

// Per-trellis puncturing state of k_vit_prep: the segment that contains the first step of the current 32-step chunk, with its
// tables, so that the common case (the whole chunk inside one segment) costs a dozen instructions per step instead of the
// segment search and table lookups of vit_load_step.
struct PrepSeg { uint32_t seg, seg_end, start, inb, cntw, K, pref_lo, pref_hi; };
__device__ __forceinline__ void prep_seg_load(PrepSeg& S, const VitJobDev& J, const uint32_t seg) {
    const uint32_t pi = J.seg_pi[seg];
    S.seg = seg;
    S.seg_end = J.seg_step_end[seg];
    S.start = seg ? J.seg_step_end[seg - 1u] : 0u;
    S.inb = J.seg_in_base[seg];
    S.cntw = c_pi_cnt[pi];
    S.K = c_pi_K[pi];
    S.pref_lo = uint32_t(c_pi_pref[pi]);
    S.pref_hi = uint32_t(c_pi_pref[pi] >> 32);
}

// One CTA per group; warp w fills the columns VP_JPW*w .. VP_JPW*w + VP_JPW-1 of the group's symbol matrix, lane = trellis
// step within a chunk of 32.  The pass is a chain of dependent L2 round trips per warp (one per trellis and chunk: the byte
// loads of a step are predicated on its puncturing count), so fewer trellises per warp shorten it, while more trellises
// per warp give full-sector stores.  Measured per 256-stream step (planning kernels included): 8 per warp 0.294 ms,
// 4 per warp 0.218 ms, 2 per warp 0.252 ms, 1 per warp 0.303 ms.
#define VP_JPW 4u
#define VP_WARPS (32u / VP_JPW)
__global__ void __launch_bounds__(VP_WARPS * 32)
k_vit_prep(const VitJobDev* __restrict__ jobs, const VlPlan* __restrict__ plan, const uint32_t* __restrict__ list, uint32_t* __restrict__ sym, const GatherGeom G) {
    __shared__ VitJobDev sJ[32];
    __shared__ uint32_t s_rowoff[32][16];
    __shared__ PrepSeg sSeg[32];
    const uint32_t g = blockIdx.x;
    if (g >= plan->n_groups) return;
    uint32_t list0, n_in, row0;
    vl_locate(plan, g, list0, n_in, row0);
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    if (threadIdx.x < 32u) {
        if (threadIdx.x < n_in) {
            sJ[threadIdx.x] = jobs[list[list0 + threadIdx.x]];
            uint32_t seg = 0;
            while (seg < DABGPU_MAX_SEGMENTS - 1u && sJ[threadIdx.x].seg_step_end[seg] == 0u) seg++;   // skip empty leading segments
            prep_seg_load(sSeg[threadIdx.x], sJ[threadIdx.x], seg);
        } else {
            sJ[threadIdx.x].total_steps = 0u;
            sSeg[threadIdx.x].seg_end = 0u;   // never on the fast path
        }
    }
    __syncthreads();
    uint32_t steps_g = sJ[lane].total_steps;
    steps_g = __reduce_max_sync(FULL_MASK, steps_g);
    const uint32_t padded = ((steps_g + VL_UNROLL - 1u) / VL_UNROLL) * VL_UNROLL + VL_UNROLL;
#pragma unroll 1
    for (uint32_t jj = 0; jj < VP_JPW; jj++) {
        const uint32_t q = VP_JPW * w + jj;
        if (sJ[q].total_steps != 0u) vit_fill_rowoff(sJ[q], G, s_rowoff[q], lane);
    }
    __syncwarp();
    uint32_t* dst = sym + (size_t(row0) * 32u + VP_JPW * w);
    // Measured alternatives: a planar soft-bit layout (positions r mod 16 contiguous per CIF) removed
    // the DRAM over-fetch of this pass (274 -> 70 MB) but not its time, and cost k_ofdm_demod 2 percent; issuing the byte loads
    // of the 8 trellises round by round (8 in flight per lane) did not help either: the pass is bound by its instruction count
    // (about 100 per trellis step and lane with vit_load_step), hence the cached segment state below.
#pragma unroll 1
    for (uint32_t t0 = 0; t0 < padded; t0 += 32u) {
        const uint32_t t = t0 + lane;
        uint32_t v[VP_JPW];
#pragma unroll
        for (uint32_t jj = 0; jj < VP_JPW; jj++) {
            const uint32_t q = VP_JPW * w + jj;
            const PrepSeg& S = sSeg[q];
            if (t0 + 32u <= S.seg_end) {
                // the whole chunk lies inside the cached segment (warp-uniform test): DAB_Viterbi_Decoder::depuncture_symbols
                // (dab_viterbi_decoder.cpp:131-181) with segment, code and tables read once per warp
                const uint32_t u = t - S.start, g8 = u & 7u;
                const uint32_t cnt = (S.cntw >> (4u * g8)) & 0xFu;
                const uint32_t pre = (((g8 & 4u) ? S.pref_hi : S.pref_lo) >> (8u * (g8 & 3u))) & 0xFFu;
                const uint32_t base = S.inb + (u >> 3) * S.K + pre;
                const int8_t* __restrict__ src = sJ[q].src;
                uint32_t word = 0;
#pragma unroll
                for (uint32_t r = 0; r < 4; r++) {
                    if (r < cnt) {
                        const uint32_t idx = base + r;
                        word |= uint32_t(uint8_t(__ldg(src + (size_t(s_rowoff[q][idx & 15u]) + idx)))) << (8u * r);
                    }
                }
                v[jj] = word;
            } else {
                v[jj] = (t < sJ[q].total_steps) ? vit_load_step(sJ[q], s_rowoff[q], t) : 0u;
            }
        }
        if (t < padded) *reinterpret_cast<uint4*>(dst + size_t(t) * 32u) = make_uint4(v[0], v[1], v[2], v[3]);
        // move the cached segments to the one that contains the first step of the next chunk
        __syncwarp();
        if (lane < VP_JPW) {
            const uint32_t q = VP_JPW * w + lane;
            PrepSeg& S = sSeg[q];
            if (S.seg_end != 0u && t0 + 32u >= S.seg_end && S.seg < DABGPU_MAX_SEGMENTS - 1u) {
                uint32_t seg = S.seg;
                while (seg < DABGPU_MAX_SEGMENTS - 1u && t0 + 32u >= sJ[q].seg_step_end[seg]) seg++;
                prep_seg_load(S, sJ[q], seg);
            }
        }
        __syncwarp();
    }
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
// walk reaches the last decision word it needs.  Then energy dispersal (additive_scrambler.h:10-36) and the FIB CRCs.
template <uint32_t VL_TB_BLOCK>   // traceback rows fetched per batch (multiple of VL_UNROLL), two batches in flight
__device__ __forceinline__ void vl_traceback(const uint2* __restrict__ dec, const VitJobDev* __restrict__ J, const bool have, const uint32_t n_out_bytes,
                                             const uint32_t flags, uint8_t* __restrict__ const out, const uint32_t* __restrict__ prbs_words) {
    const uint32_t nbits = n_out_bytes * 8u;
    const uint32_t top = __reduce_max_sync(FULL_MASK, nbits);      // rows top+5 .. 6 are walked
    if (top != 0u) {
        uint32_t state = 0, acc = 0;
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
#pragma unroll
            for (int r = int(VL_TB_BLOCK) - 1; r >= 0; --r) {
                const uint32_t t = uint32_t(blk) * VL_TB_BLOCK + uint32_t(r);
                if (t >= 6u && t < nbits + 6u) {
                    const uint32_t b = t - 6u;
                    const uint32_t bit = vl_decision(cur[r].x, cur[r].y, state, uint32_t(r) % VL_UNROLL);
                    state = (state >> 1) | (bit << 5);
                    acc = (acc >> 1) | (bit << 31);
                    if ((b & 31u) == 0u) {
                        const uint32_t widx = b >> 5;
                        uint32_t v = acc;
                        if (flags & VJ_DESCRAMBLE) v ^= prbs_words[widx];
                        vl_emit_word(out, n_out_bytes, widx, v);
                        acc = 0;
                    }
                }
            }
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
        {
            VlState S;
            vl_reset(S);
            uint64_t final_err = 0;
            uint32_t w[VL_UNROLL];
#pragma unroll
            for (int k = 0; k < VL_UNROLL; k++) w[k] = __ldg(srow + size_t(k) * 32u);
#pragma unroll 1
            for (uint32_t t0 = 0; t0 < padded; t0 += VL_UNROLL) {
                uint32_t wn[VL_UNROLL];
#pragma unroll
                for (int k = 0; k < VL_UNROLL; k++) wn[k] = __ldg(srow + size_t(t0 + VL_UNROLL + k) * 32u);
                uint32_t d[2 * VL_UNROLL];
                vl_step5(S, w, t0, N, d, final_err, kc);
#pragma unroll
                for (int k = 0; k < VL_UNROLL; k++) {
                    dec[size_t(t0 + k) * 32u] = make_uint2(d[2 * k], d[2 * k + 1]);
                    w[k] = wn[k];
                }
            }
            if (have && J->path_error != nullptr) *J->path_error = final_err;
        }

        vl_traceback<TB>(dec, J, have, n_out_bytes, flags, out, prbs_words);
        __syncwarp();
    }
}
