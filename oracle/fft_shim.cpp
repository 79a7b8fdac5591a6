// oracle/fft_shim.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Power-of-two single precision complex FFT behind the fftwf_* symbols declared
// in oracle/fftw3.h, so that the unmodified reference OFDM sources can be built
// without FFTW3 (absent from /root/reference and from this image).
//
// The reference's CPU timing (bench.py cpu_baseline / --impl reference) runs through this
// file, so it is written to be a fair stand-in for FFTW3f on an AVX2 host rather than a toy:
// four-step decomposition N = N1 x N2 in split (re / im) format, where both passes are
// Stockham autosort FFTs DOWN THE COLUMNS of a row-major matrix -- every butterfly loop runs
// over a contiguous row of >= 16 independent columns and is auto-vectorised (-O3, AVX2+FMA):
//   1. A[n1][n2] = x[n1*N2 + n2];   N1-point FFT over n1 for every column n2   -> B[k1][n2]
//   2. B[k1][n2] *= W_N^(k1*n2)
//   3. transpose to [n2][k1];        N2-point FFT over n2 for every column k1   -> C[k2][k1] = X[k1 + N1*k2]
// which is natural order.  Twiddles are computed in double and rounded once.  Unnormalised,
// in-place capable (the input is copied first), thread-safe for concurrent execution of one
// plan on different arrays (scratch is per thread).  Measured here: 2048 points in 11.4 us
// on one core (the first version of this shim, a scalar radix-4 Stockham, needed 19.6 us; FFTW3f with AVX2
// codelets is typically at 4-6 us for this size, so the OFDM stage of a real FFTW build would still be about 15 % faster).
#include "fftw3.h"
#include <cmath>
#include <complex>
#include <cstring>
#include <vector>

typedef std::complex<float> c32;

struct fftwf_plan_s {
    int n, n1, n2;
    int sign;
    std::vector<float> w1r, w1i;   // exp(sign*2*pi*i*k/n1), k < n1
    std::vector<float> w2r, w2i;   // exp(sign*2*pi*i*k/n2), k < n2
    std::vector<float> twr, twi;   // step 2: [k1][n2] = exp(sign*2*pi*i*k1*n2/n)
};

// radix-4 butterflies of one twiddle index over s contiguous columns (restrict-qualified parameters: the loop vectorises)
template <bool INV>
static inline void bfly4(const int s, const float* __restrict__ ar, const float* __restrict__ ai, const float* __restrict__ br, const float* __restrict__ bi,
                         const float* __restrict__ cr, const float* __restrict__ ci, const float* __restrict__ dr, const float* __restrict__ di,
                         float* __restrict__ y0r, float* __restrict__ y0i, float* __restrict__ y1r, float* __restrict__ y1i,
                         float* __restrict__ y2r, float* __restrict__ y2i, float* __restrict__ y3r, float* __restrict__ y3i,
                         const float w1r, const float w1i, const float w2r, const float w2i, const float w3r, const float w3i) {
    for (int q = 0; q < s; q++) {
        const float apcr = ar[q] + cr[q], apci = ai[q] + ci[q];
        const float amcr = ar[q] - cr[q], amci = ai[q] - ci[q];
        const float bpdr = br[q] + dr[q], bpdi = bi[q] + di[q];
        const float bmdr = br[q] - dr[q], bmdi = bi[q] - di[q];
        // +j*(b - d) for the inverse transform, -j*(b - d) for the forward one
        const float jr = INV ? -bmdi : bmdi, ji = INV ? bmdr : -bmdr;
        y0r[q] = apcr + bpdr;            y0i[q] = apci + bpdi;
        const float t1r = amcr + jr,     t1i = amci + ji;
        const float t2r = apcr - bpdr,   t2i = apci - bpdi;
        const float t3r = amcr - jr,     t3i = amci - ji;
        y1r[q] = w1r*t1r - w1i*t1i;      y1i[q] = w1r*t1i + w1i*t1r;
        y2r[q] = w2r*t2r - w2i*t2i;      y2i[q] = w2r*t2i + w2i*t2r;
        y3r[q] = w3r*t3r - w3i*t3i;      y3i[q] = w3r*t3i + w3i*t3r;
    }
}

// Stockham autosort FFT of length M over the rows of an [M][C] matrix (C contiguous columns), split format.
// (xr, xi) is the input and is destroyed; the result is returned in whichever buffer pair it ends up in.
static void fft_columns(const int M, const int C, const float* __restrict__ wr, const float* __restrict__ wi, const int sign,
                        float*& xr, float*& xi, float*& yr, float*& yi) {
    int n = M;   // current sub-transform length
    int s = C;   // stride (in floats) between consecutive elements of a sub-transform
    while (n >= 4) {
        const int n1 = n/4;
        const int wstep = M/n;
        for (int p = 0; p < n1; p++) {
            const float w1r = wr[p*wstep], w1i = wi[p*wstep];
            const float w2r = wr[2*p*wstep], w2i = wi[2*p*wstep];
            const float w3r = wr[3*p*wstep], w3i = wi[3*p*wstep];
            if (sign > 0) bfly4<true>(s, xr + s*p, xi + s*p, xr + s*(p+n1), xi + s*(p+n1), xr + s*(p+2*n1), xi + s*(p+2*n1), xr + s*(p+3*n1), xi + s*(p+3*n1),
                                      yr + s*(4*p), yi + s*(4*p), yr + s*(4*p+1), yi + s*(4*p+1), yr + s*(4*p+2), yi + s*(4*p+2), yr + s*(4*p+3), yi + s*(4*p+3),
                                      w1r, w1i, w2r, w2i, w3r, w3i);
            else bfly4<false>(s, xr + s*p, xi + s*p, xr + s*(p+n1), xi + s*(p+n1), xr + s*(p+2*n1), xi + s*(p+2*n1), xr + s*(p+3*n1), xi + s*(p+3*n1),
                              yr + s*(4*p), yi + s*(4*p), yr + s*(4*p+1), yi + s*(4*p+1), yr + s*(4*p+2), yi + s*(4*p+2), yr + s*(4*p+3), yi + s*(4*p+3),
                              w1r, w1i, w2r, w2i, w3r, w3i);
        }
        float* t;
        t = xr; xr = yr; yr = t;
        t = xi; xi = yi; yi = t;
        n /= 4; s *= 4;
    }
    if (n == 2) {
        for (int q = 0; q < s; q++) {
            const float ar = xr[q], ai = xi[q], br = xr[q+s], bi = xi[q+s];
            yr[q] = ar + br;   yi[q] = ai + bi;
            yr[q+s] = ar - br; yi[q+s] = ai - bi;
        }
        float* t;
        t = xr; xr = yr; yr = t;
        t = xi; xi = yi; yi = t;
    }
}

static void execute(const fftwf_plan_s& P, const c32* in, c32* out) {
    const int N = P.n, N1 = P.n1, N2 = P.n2;
    thread_local std::vector<float> scratch;
    if ((int)scratch.size() < 4*N) scratch.resize(4*N);
    float* xr = scratch.data();
    float* xi = xr + N;
    float* yr = xi + N;
    float* yi = yr + N;
    for (int i = 0; i < N; i++) { xr[i] = in[i].real(); xi[i] = in[i].imag(); }
    if (N1 > 1) {
        // 1. N1-point transforms down the columns of [N1][N2]
        fft_columns(N1, N2, P.w1r.data(), P.w1i.data(), P.sign, xr, xi, yr, yi);
        // 2 + 3a. twiddle and transpose into [N2][N1]
        const float* __restrict__ tr = P.twr.data();
        const float* __restrict__ ti = P.twi.data();
        for (int k1 = 0; k1 < N1; k1++) {
            const float* __restrict__ rr = xr + k1*N2;
            const float* __restrict__ ri = xi + k1*N2;
            const float* __restrict__ wr = tr + k1*N2;
            const float* __restrict__ wi = ti + k1*N2;
            for (int n2 = 0; n2 < N2; n2++) {
                yr[n2*N1 + k1] = rr[n2]*wr[n2] - ri[n2]*wi[n2];
                yi[n2*N1 + k1] = rr[n2]*wi[n2] + ri[n2]*wr[n2];
            }
        }
        float* t;
        t = xr; xr = yr; yr = t;
        t = xi; xi = yi; yi = t;
    }
    // 3b. N2-point transforms down the columns of [N2][N1]
    fft_columns(N2, N1, P.w2r.data(), P.w2i.data(), P.sign, xr, xi, yr, yi);
    for (int i = 0; i < N; i++) out[i] = c32(xr[i], xi[i]);
}

extern "C" fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex*, fftwf_complex*, int sign, unsigned) {
    if (n < 1 || (n & (n-1)) != 0) return nullptr;
    auto* p = new fftwf_plan_s;
    p->n = n;
    p->sign = (sign > 0) ? +1 : -1;
    int lg = 0;
    while ((1 << lg) < n) lg++;
    p->n1 = (n >= 64) ? (1 << (lg/2)) : 1;   // small sizes: a single pass with one column
    p->n2 = n / p->n1;
    auto table = [&](int m, std::vector<float>& re, std::vector<float>& im) {
        re.resize(size_t(m)); im.resize(size_t(m));
        for (int k = 0; k < m; k++) {
            const double a = 2.0*M_PI*double(k)/double(m);
            re[size_t(k)] = float(std::cos(a));
            im[size_t(k)] = float(p->sign*std::sin(a));
        }
    };
    table(p->n1, p->w1r, p->w1i);
    table(p->n2, p->w2r, p->w2i);
    p->twr.resize(size_t(n)); p->twi.resize(size_t(n));
    for (int k1 = 0; k1 < p->n1; k1++)
        for (int n2 = 0; n2 < p->n2; n2++) {
            const double a = 2.0*M_PI*double(k1)*double(n2)/double(n);
            p->twr[size_t(k1*p->n2 + n2)] = float(std::cos(a));
            p->twi[size_t(k1*p->n2 + n2)] = float(p->sign*std::sin(a));
        }
    return p;
}

extern "C" void fftwf_execute_dft(const fftwf_plan p, fftwf_complex* in, fftwf_complex* out) {
    execute(*p, reinterpret_cast<const c32*>(in), reinterpret_cast<c32*>(out));
}

extern "C" void fftwf_destroy_plan(fftwf_plan p) { delete p; }
