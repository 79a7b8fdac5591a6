// ofdm.cuh -- placeholder while the OFDM kernels are being brought up
#pragma once
#include "common.cuh"
struct OfdmState { uint64_t launches = 0; };
static int ofdm_init(OfdmState&, const dabgpu_config&, const dabgpu_params&, int, int8_t*, uint32_t*, dabgpu_frame_info*, unsigned long long*) { return DABGPU_OK; }
static void ofdm_destroy(OfdmState&) {}
static int ofdm_reset(OfdmState&, int, cudaStream_t) { return set_error(DABGPU_ERR_STATE, "OFDM stage not built"); }
static int ofdm_process(OfdmState&, const void*, size_t, int, int, int, int, cudaStream_t) { return set_error(DABGPU_ERR_STATE, "OFDM stage not built"); }
static int ofdm_attach(OfdmState&, const void*, size_t, size_t, cudaStream_t) { return set_error(DABGPU_ERR_STATE, "OFDM stage not built"); }
static int ofdm_advance(OfdmState&, int, int, int, int, cudaStream_t) { return set_error(DABGPU_ERR_STATE, "OFDM stage not built"); }
static int ofdm_get_status(OfdmState&, int, dabgpu_ofdm_status*, cudaStream_t) { return set_error(DABGPU_ERR_STATE, "OFDM stage not built"); }
static int ofdm_fetch_latest(OfdmState&, int, int, int8_t*, uint8_t*, cudaStream_t) { return set_error(DABGPU_ERR_STATE, "OFDM stage not built"); }
