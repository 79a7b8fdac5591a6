// common.cuh -- shared declarations of libdabgpu (unity build: included once from dabgpu.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include <mutex>
#include "../../include/dabgpu.h"

#define FULL_MASK 0xffffffffu

static thread_local char g_last_error[512] = "";

static int set_error(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
#include <stdarg.h>
static int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                               \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return set_error(DABGPU_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                                    \
    } while (0)

// Frame geometry for the four transmission modes.
// Reference: ofdm/dab_ofdm_params_ref.cpp:10-58, dab/constants/dab_parameters.h:26-90
static int fill_params(int mode, dabgpu_params* p) {
    static const int T[4][9] = {
        // L, Tsym, Tnull, N, K, fic_syms, cifs, fibs_per_cif
        {76, 2552, 2656, 2048, 1536, 3, 4, 3, 0},
        {76, 638, 664, 512, 384, 3, 1, 3, 0},
        {153, 319, 345, 256, 192, 8, 1, 4, 0},
        {76, 1276, 1328, 1024, 768, 3, 2, 3, 0},
    };
    if (mode < 1 || mode > 4) return set_error(DABGPU_ERR_INVALID, "Invalid transmission mode %d", mode);
    const int* t = T[mode - 1];
    p->nb_frame_symbols = t[0];
    p->nb_symbol_period = t[1];
    p->nb_null_period = t[2];
    p->nb_fft = t[3];
    p->nb_cyclic_prefix = t[1] - t[3];
    p->nb_data_carriers = t[4];
    p->nb_frame_bits = (t[0] - 1) * 2 * t[4];
    p->nb_fic_bits = t[5] * 2 * t[4];
    p->nb_msc_bits = p->nb_frame_bits - p->nb_fic_bits;
    p->nb_cifs = t[6];
    p->nb_fibs_per_cif = t[7];
    p->nb_fib_group_bits = p->nb_fic_bits / t[6];
    p->nb_cif_bits = p->nb_msc_bits / t[6];
    p->nb_frame_samples = t[2] + t[0] * t[1];
    return DABGPU_OK;
}

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int alloc(size_t n) {
        if (n <= bytes && p) return DABGPU_OK;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        if (n == 0) return DABGPU_OK;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) return set_error(DABGPU_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e));
        bytes = n;
        return DABGPU_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int alloc(size_t n) {
        if (n <= bytes && p) return DABGPU_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
        if (n == 0) return DABGPU_OK;
        cudaError_t e = cudaMallocHost(&p, n);
        if (e != cudaSuccess) return set_error(DABGPU_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", n, cudaGetErrorString(e));
        bytes = n;
        return DABGPU_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Optional per-kernel-class device timing (CUDA events on the launching stream), used by bench.py for the
// roofline of the dominant kernel.  Off by default: no events are recorded on the product path.
enum { PROF_OFDM_CTL = 0, PROF_OFDM_DEMOD = 1, PROF_VITERBI = 2, PROF_DABPLUS = 3, PROF_CHAN_MISC = 4, PROF_CLASSES = 5 };
struct Profiler {
    bool on = false;
    struct Rec { int cls; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    double ms[PROF_CLASSES] = {0, 0, 0, 0, 0};
    unsigned long long n[PROF_CLASSES] = {0, 0, 0, 0, 0};
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void begin(int cls, cudaStream_t s) {
        if (!on) return;
        if (recs.size() >= 8192) collect();
        Rec r; r.cls = cls; r.a = get(); r.b = get();
        cudaEventRecord(r.a, s);
        recs.push_back(r);
    }
    void end(cudaStream_t s) {
        if (!on) return;
        cudaEventRecord(recs.back().b, s);
    }
    void collect() {
        for (auto& r : recs) {
            cudaEventSynchronize(r.b);
            float t = 0.0f;
            if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.cls] += t; n[r.cls]++; }
            pool.push_back(r.a); pool.push_back(r.b);
        }
        recs.clear();
    }
    void reset() { collect(); for (int i = 0; i < PROF_CLASSES; i++) { ms[i] = 0; n[i] = 0; } }
    void destroy() { collect(); for (auto e : pool) cudaEventDestroy(e); pool.clear(); }
};
