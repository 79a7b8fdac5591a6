// tests/host/lane_core_check.cpp -- TEST DRIVER (CPU): the lane-per-trellis Viterbi core (csrc/viterbi_lane_core.h, compiled
// here with the host emulation of the DPX instructions) against the oracle's restatement of the reference decoder
// (oracle/dab_oracle.c: dabo_vit_decode, pinned to VIT/x86/viterbi_decoder_avx_u16.h by tests/test_oracle_cpu.py).
// Usage: lane_core_check [trials]    exit code 0 = every trial bit-exact (decoded bytes and accumulated path error).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../sdrplusplus-dab-radio-plugin_b200/csrc/viterbi_lane_core.h"
#include "../../oracle/dab_oracle.h"

static const VlConst KC = {0xFFFFFFFFu, 2u, 4u, 16u, 256u, 0x10000u};

// returns the number of own-clamp activations seen (to prove that the saturating cases are exercised)
// general = the branch-error form that tolerates -128 symbols (vl_branch<K, true>); the short form needs |symbol| <= 127
static void lane_decode(const int8_t* soft4, uint32_t n_steps, uint8_t* out, uint32_t n_out_bytes, uint64_t* err, bool no_clamp, bool general = true) {
    VlState S;
    vl_reset(S);
    const uint32_t padded = ((n_steps + VL_UNROLL - 1) / VL_UNROLL) * VL_UNROLL;
    std::vector<uint32_t> d0(padded), d1(padded);
    uint32_t final_rel = 0;
    for (uint32_t t0 = 0; t0 < padded; t0 += VL_UNROLL) {
        uint32_t w[VL_UNROLL], dec[2 * VL_UNROLL];
        for (uint32_t k = 0; k < VL_UNROLL; k++) {
            w[k] = 0;
            if (t0 + k < n_steps) memcpy(&w[k], soft4 + 4 * size_t(t0 + k), 4);
        }
        if (no_clamp) S.CL = 0x7FFF7FFFu;
        if (general) vl_step5<true>(S, w, t0, n_steps, dec, final_rel, KC);
        else vl_step5<false>(S, w, t0, n_steps, dec, final_rel, KC);
        for (uint32_t k = 0; k < VL_UNROLL; k++) { d0[t0 + k] = dec[2 * k]; d1[t0 + k] = dec[2 * k + 1]; }
    }
    *err = vl_final_error(S, final_rel);
    // ViterbiDecoder_Core::chainback from state 0: decoded bit b comes from the decision word of step b + 6
    // twice: with the generic bit position (vl_decision) and with the history-word step the kernel uses (vl_traceback_step)
    memset(out, 0, n_out_bytes);
    uint32_t state = 0, h = 0;
    for (int b = int(n_out_bytes) * 8 - 1; b >= 0; --b) {
        const uint32_t t = uint32_t(b) + 6u;
        const uint32_t bit = vl_decision(d0[t], d1[t], state, t % VL_UNROLL);
        state = (state >> 1) | (bit << 5);
        switch (t % VL_UNROLL) {
        case 0: h = vl_traceback_step<0>(d0[t], d1[t], h); break;
        case 1: h = vl_traceback_step<1>(d0[t], d1[t], h); break;
        case 2: h = vl_traceback_step<2>(d0[t], d1[t], h); break;
        case 3: h = vl_traceback_step<3>(d0[t], d1[t], h); break;
        default: h = vl_traceback_step<4>(d0[t], d1[t], h); break;
        }
        if ((h >> 26) != state) { *err = ~0ull; return; }    // the two walks disagree: reported as a path-error mismatch
        out[b >> 3] |= uint8_t((h >> 31) << (7 - (b & 7)));
    }
}

static void encode(const std::vector<uint8_t>& bits, std::vector<int8_t>& sym) {
    static const unsigned G[4] = {109, 79, 83, 109};
    unsigned state = 0;
    sym.resize(bits.size() * 4);
    for (size_t i = 0; i < bits.size(); i++) {
        const unsigned reg = (state << 1) | bits[i];
        for (int r = 0; r < 4; r++) sym[4 * i + r] = __builtin_parity(reg & G[r]) ? 127 : -127;
        state = reg & 63u;
    }
}

int main(int argc, char** argv) {
    const int trials = argc > 1 ? atoi(argv[1]) : 400;
    std::mt19937 rng(12345);
    int bad = 0, clamp_mattered = 0, short_form = 0;
    static const uint32_t lengths[] = {774, 1542, 6, 7, 14, 46, 262, 3078, 6150, 102};
    for (int trial = 0; trial < trials; trial++) {
        const uint32_t n_steps = lengths[trial % 10];
        const uint32_t n_out = (n_steps - 6) / 8;
        const int kind = (trial / 10) % 8;
        std::vector<uint8_t> bits(n_steps, 0);
        for (uint32_t i = 0; i + 6 < n_steps; i++) bits[i] = rng() & 1u;
        std::vector<int8_t> sym;
        encode(bits, sym);
        std::vector<int8_t> soft(sym.size());
        std::normal_distribution<float> gauss(0.f, 1.f);
        for (size_t i = 0; i < sym.size(); i++) {
            int v;
            switch (kind) {
            case 0: v = int(sym[i] * 0.5f + 40.f * gauss(rng)); break;                 // moderate noise
            case 1: v = int(rng() % 256) - 128; break;                                 // garbage incl. -128
            case 2: v = (int(rng() % 3) - 1) * 127; break;                             // tie heavy
            case 3: v = (rng() & 1u) ? -128 : 127; break;                              // extreme values
            case 4: v = ((i / 4) % 97 < 60) ? sym[i] : -sym[i]; break;                 // clean with inverted bursts: large spread + fast growth
            case 5: v = int(sym[i] * 0.9f + 90.f * gauss(rng)); break;                 // heavy noise, clipped
            case 6: v = ((i / 4) % 41 < 33) ? sym[i] : int(rng() % 256) - 128; break;  // clean with garbage bursts
            default: v = 0; break;                                                      // all punctured
            }
            if (kind == 7 && (rng() % 5u) == 0) v = sym[i];
            soft[i] = int8_t(v < -128 ? -128 : (v > 127 ? 127 : v));
        }
        std::vector<uint8_t> exp(n_out + 1), got(n_out + 1), got_nc(n_out + 1);
        uint64_t exp_err = 0, got_err = 0, nc_err = 0;
        const int seg_pi[1] = {24};
        const int seg_bits[1] = {int(n_steps) * 4};
        const int rc = dabo_vit_decode(soft.data(), int(soft.size()), seg_pi, seg_bits, 1, exp.data(), int(n_out), &exp_err);
        if (rc < 0) { printf("oracle refused trial %d (rc %d)\n", trial, rc); return 2; }
        lane_decode(soft.data(), n_steps, got.data(), n_out, &got_err, false);
        lane_decode(soft.data(), n_steps, got_nc.data(), n_out, &nc_err, true);
        if (memcmp(got_nc.data(), got.data(), n_out) != 0 || nc_err != got_err) clamp_mattered++;
        // the short branch-error form, where it applies (no -128 anywhere; the kernel decides per call with vl_has_m128)
        bool m128 = false;
        for (size_t i = 0; i + 3 < soft.size(); i += 4) { uint32_t wv; memcpy(&wv, &soft[i], 4); m128 = m128 || vl_has_m128(wv); }
        bool truth = false;
        for (size_t i = 0; i < soft.size(); i++) truth = truth || soft[i] == -128;
        if (truth != m128) { printf("vl_has_m128 wrong in trial %d\n", trial); bad++; }
        if (!m128) {
            short_form++;
            std::vector<uint8_t> got_s(n_out + 1);
            uint64_t s_err = 0;
            lane_decode(soft.data(), n_steps, got_s.data(), n_out, &s_err, false, false);
            if (memcmp(exp.data(), got_s.data(), n_out) != 0 || exp_err != s_err) {
                bad++;
                if (bad <= 10) printf("MISMATCH (short form) trial %d kind %d steps %u\n", trial, kind, n_steps);
            }
        }
        if (memcmp(exp.data(), got.data(), n_out) != 0 || exp_err != got_err) {
            bad++;
            if (bad <= 10) printf("MISMATCH trial %d kind %d steps %u: err oracle %llu lane %llu\n", trial, kind, n_steps,
                                  (unsigned long long)exp_err, (unsigned long long)got_err);
        }
    }
    printf("lane_core_check: %d trials, %d mismatches, saturation clamp changed the result in %d trials, short branch-error form checked in %d\n", trials, bad,
           clamp_mattered, short_form);
    return bad ? 1 : 0;
}
