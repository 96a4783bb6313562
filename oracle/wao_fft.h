// ORACLE — TEST INFRASTRUCTURE ONLY (see wao_core.h).
// Plain f32 FFTs for the oracle.  The reference delegates to the `realfft`/`rustfft` crates
// (Cargo.toml:45, not vendored); any correct unnormalised FFT reproduces them up to f32 rounding,
// which is far below the 1e-5 parity tolerance.  Twiddles are computed in f64 and rounded once.
#pragma once
#include <cmath>
#include <complex>
#include <vector>

namespace wao {

typedef std::complex<float> cf32;

// In-place iterative radix-2 complex FFT, size n = 2^k.  sign = -1 forward, +1 inverse (unnormalised).
class ComplexFFT {
  public:
    int n = 0;
    std::vector<cf32> tw;  // tw[k] = exp(-2*pi*i*k/n), k < n/2
    std::vector<int> rev;
    void init(int n_) {
        n = n_;
        tw.resize(n / 2);
        for (int k = 0; k < n / 2; k++) {
            double a = -2.0 * M_PI * (double)k / (double)n;
            tw[k] = cf32((float)std::cos(a), (float)std::sin(a));
        }
        rev.resize(n);
        int bits = 0;
        while ((1 << bits) < n) bits++;
        for (int i = 0; i < n; i++) {
            int r = 0;
            for (int b = 0; b < bits; b++)
                if (i & (1 << b)) r |= 1 << (bits - 1 - b);
            rev[i] = r;
        }
    }
    void run(cf32* x, int sign) const {
        for (int i = 0; i < n; i++)
            if (i < rev[i]) std::swap(x[i], x[rev[i]]);
        for (int len = 2; len <= n; len <<= 1) {
            int half = len >> 1, step = n / len;
            for (int i = 0; i < n; i += len) {
                for (int j = 0; j < half; j++) {
                    cf32 w = tw[j * step];
                    float wr = w.real(), wi = sign < 0 ? w.imag() : -w.imag();
                    cf32 b = x[i + j + half];
                    float br = b.real() * wr - b.imag() * wi;
                    float bi = b.real() * wi + b.imag() * wr;
                    cf32 a = x[i + j];
                    x[i + j] = cf32(a.real() + br, a.imag() + bi);
                    x[i + j + half] = cf32(a.real() - br, a.imag() - bi);
                }
            }
        }
    }
};

// Real FFT of even size n via one complex FFT of size n/2 (standard packing).
// forward: n reals -> n/2+1 complex bins (unnormalised).  inverse: n/2+1 bins -> n reals, unnormalised
// (caller divides by n), imaginary parts of bin 0 and bin n/2 are ignored like realfft does.
class RealFFT {
  public:
    int n = 0;
    ComplexFFT c;
    std::vector<cf32> w;  // exp(-2*pi*i*k/n), k <= n/4.. use n/2 entries
    mutable std::vector<cf32> tmp;
    void init(int n_) {
        n = n_;
        c.init(n / 2);
        w.resize(n / 2 + 1);
        for (int k = 0; k <= n / 2; k++) {
            double a = -2.0 * M_PI * (double)k / (double)n;
            w[k] = cf32((float)std::cos(a), (float)std::sin(a));
        }
        tmp.resize(n / 2);
    }
    void forward(const float* in, cf32* out) const {
        int h = n / 2;
        for (int i = 0; i < h; i++) tmp[i] = cf32(in[2 * i], in[2 * i + 1]);
        c.run(tmp.data(), -1);
        // X[k] = E[k] + w^k O[k];  E = (Z[k] + conj(Z[h-k]))/2, O = (Z[k] - conj(Z[h-k]))/(2i)
        out[0] = cf32(tmp[0].real() + tmp[0].imag(), 0.f);
        out[h] = cf32(tmp[0].real() - tmp[0].imag(), 0.f);
        for (int k = 1; k < h; k++) {
            cf32 zk = tmp[k], zc = std::conj(tmp[h - k]);
            cf32 e = 0.5f * (zk + zc);
            cf32 d = zk - zc;
            cf32 o = cf32(0.5f * d.imag(), -0.5f * d.real());  // d / (2i)
            cf32 t = cf32(w[k].real() * o.real() - w[k].imag() * o.imag(), w[k].real() * o.imag() + w[k].imag() * o.real());
            out[k] = e + t;
        }
    }
    void inverse(const cf32* in, float* out) const {
        int h = n / 2;
        // Z[k] = E[k] + i O[k], E[k] = (X[k] + conj(X[h-k]))/2, O[k] = conj(w^k) (X[k] - conj(X[h-k]))/2
        // unnormalised inverse of size n == 2 * (inverse of size h of Z) -> scale by 2
        for (int k = 0; k < h; k++) {
            cf32 xk = k == 0 ? cf32(in[0].real(), 0.f) : in[k];
            cf32 xc = (h - k) == h ? cf32(in[h].real(), 0.f) : std::conj(in[h - k]);
            if (k == 0) xc = cf32(in[h].real(), 0.f);
            cf32 e = xk + xc;
            cf32 d = xk - xc;
            cf32 wc = std::conj(w[k]);
            cf32 o = cf32(wc.real() * d.real() - wc.imag() * d.imag(), wc.real() * d.imag() + wc.imag() * d.real());
            tmp[k] = cf32(e.real() - o.imag(), e.imag() + o.real());  // e + i*o
        }
        c.run(tmp.data(), +1);
        for (int i = 0; i < h; i++) {
            out[2 * i] = tmp[i].real();
            out[2 * i + 1] = tmp[i].imag();
        }
    }
};

}  // namespace wao
