// oracle/fft_shim.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Power-of-two single precision complex FFT behind the fftwf_* symbols declared
// in oracle/fftw3.h, so that the unmodified reference OFDM sources can be built
// without FFTW3 (absent from /root/reference and from this image).
// Algorithm: Stockham autosort, radix-4 passes with one trailing radix-2 pass,
// twiddles computed in double and rounded once.  Unnormalised, thread-safe for
// concurrent execution of one plan on different arrays (scratch is per call).
#include "fftw3.h"
#include <cmath>
#include <complex>
#include <cstring>
#include <vector>

typedef std::complex<float> c32;

struct fftwf_plan_s {
    int n;
    int sign;
    std::vector<c32> tw;   // tw[k] = exp(sign*2*pi*i*k/n), k < n
};

static inline c32 mulj(const c32 v, const int sign) {
    // multiply by sign*j
    return (sign > 0) ? c32(-v.imag(), v.real()) : c32(v.imag(), -v.real());
}

static void stockham(const fftwf_plan_s& P, const c32* in, c32* out) {
    const int N = P.n;
    thread_local std::vector<c32> scratch;
    if ((int)scratch.size() < 2*N) scratch.resize(2*N);
    c32* x = scratch.data();
    c32* y = scratch.data() + N;
    std::memcpy(x, in, sizeof(c32)*N);
    const c32* W = P.tw.data();
    int n = N;   // current sub-transform length
    int s = 1;   // stride
    while (n >= 4) {
        const int n1 = n/4, n2 = n/2, n3 = n1+n2;
        const int wstep = N/n;
        for (int p = 0; p < n1; p++) {
            const c32 w1 = W[p*wstep];
            const c32 w2 = W[2*p*wstep];
            const c32 w3 = W[3*p*wstep];
            const c32* xa = x + s*p;
            const c32* xb = x + s*(p+n1);
            const c32* xc = x + s*(p+n2);
            const c32* xd = x + s*(p+n3);
            c32* y0 = y + s*(4*p+0);
            c32* y1 = y + s*(4*p+1);
            c32* y2 = y + s*(4*p+2);
            c32* y3 = y + s*(4*p+3);
            for (int q = 0; q < s; q++) {
                const c32 a = xa[q], b = xb[q], c = xc[q], d = xd[q];
                const c32 apc = a+c, amc = a-c, bpd = b+d;
                const c32 jbmd = mulj(b-d, P.sign);
                y0[q] = apc + bpd;
                y1[q] = w1*(amc + jbmd);
                y2[q] = w2*(apc - bpd);
                y3[q] = w3*(amc - jbmd);
            }
        }
        c32* t = x; x = y; y = t;
        n /= 4; s *= 4;
    }
    if (n == 2) {
        for (int q = 0; q < s; q++) {
            const c32 a = x[q], b = x[q+s];
            y[q] = a+b;
            y[q+s] = a-b;
        }
        c32* t = x; x = y; y = t;
    }
    std::memcpy(out, x, sizeof(c32)*N);
}

extern "C" fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex*, fftwf_complex*, int sign, unsigned) {
    if (n < 1 || (n & (n-1)) != 0) return nullptr;
    auto* p = new fftwf_plan_s;
    p->n = n;
    p->sign = (sign > 0) ? +1 : -1;
    p->tw.resize(n);
    for (int k = 0; k < n; k++) {
        const double a = 2.0*M_PI*double(k)/double(n);
        p->tw[k] = c32(float(std::cos(a)), float(p->sign*std::sin(a)));
    }
    return p;
}

extern "C" void fftwf_execute_dft(const fftwf_plan p, fftwf_complex* in, fftwf_complex* out) {
    stockham(*p, reinterpret_cast<const c32*>(in), reinterpret_cast<c32*>(out));
}

extern "C" void fftwf_destroy_plan(fftwf_plan p) { delete p; }
