/* oracle/fftw3.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Declaration shim for the three FFTW3f entry points the reference OFDM code
 * calls (vendor/DAB-Radio/src/ofdm/ofdm_demodulator.cpp:113-114, 227-228, 893,
 * 898; ofdm_modulator.cpp).  FFTW3 (libfftw3-dev, unpinned on Linux; vcpkg
 * fftw3 >= 3.3.10#3 on Windows) is a system dependency that is NOT vendored in
 * /root/reference and is not installed in this image, so the oracle build
 * links the reference sources against oracle/fft_shim.cpp instead.
 * Semantics relied upon by the reference: unnormalised transforms, plans made
 * on null pointers and re-used on arbitrary arrays, in-place allowed.
 */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef float fftwf_complex[2];
typedef struct fftwf_plan_s* fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex* in, fftwf_complex* out, int sign, unsigned flags);
void fftwf_execute_dft(const fftwf_plan p, fftwf_complex* in, fftwf_complex* out);
void fftwf_destroy_plan(fftwf_plan p);
#ifdef __cplusplus
}
#endif
