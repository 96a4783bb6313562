// Host-side (control-plane) math of the engine: per-node constants that the reference computes once per
// quantum from k-rate values and that the GPU kernels take as inputs.  f64/f32 glibc math in the same
// expressions as the reference so the constants are bit-identical to the CPU renderer's.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "wae_spatial.h"

namespace wae {
namespace hostmath {

static const double PI64 = 3.14159265358979323846;
static const float PI32 = 3.14159265358979323846f;
static const double F64_MAX = 1.7976931348623157e308;

struct BiquadCoefs {
    double b0, b1, b2, a1, a2;
};

inline BiquadCoefs norm(double b0, double b1, double b2, double a0, double a1, double a2) {
    double s = 1. / a0;  // normalize_coefs, src/node/biquad_filter.rs:28-40
    return BiquadCoefs{b0 * s, b1 * s, b2 * s, a1 * s, a2 * s};
}

// calculate_coefs, src/node/biquad_filter.rs:42-390 (WPT biquad-filters.js formulas)
inline BiquadCoefs biquad_coefs(int type, double sample_rate, double f0, double gain, double q) {
    const BiquadCoefs wire{1., 0., 0., 0., 0.}, zero{0., 0., 0., 0., 0.};
    double nyquist = sample_rate / 2.;
    double f = f0 / nyquist;
    f = f < 0. ? 0. : (f > 1. ? 1. : f);
    double w0 = PI64 * f, s = std::sin(w0), c = std::cos(w0);
    double A = std::pow(10., gain / 40.);
    switch (type) {
        case 0: {
            if (f == 1.) return wire;
            double a = s / (2. * std::pow(10., q / 20.)), beta = (1. - c) / 2.;
            return norm(beta, 2. * beta, beta, 1. + a, -2. * c, 1. - a);
        }
        case 1: {
            if (f == 1.) return zero;
            if (f == 0.) return wire;
            double a = s / (2. * std::pow(10., q / 20.)), beta = (1. + c) / 2.;
            return norm(beta, -2. * beta, beta, 1. + a, -2. * c, 1. - a);
        }
        case 2: {
            if (!(f > 0. && f < 1.)) return zero;
            if (!(q > 0.)) return wire;
            double a = s / (2. * q);
            return norm(a, 0., -a, 1. + a, -2. * c, 1. - a);
        }
        case 3: {
            if (!(f > 0. && f < 1.)) return wire;
            if (!(q > 0.)) return zero;
            double a = s / (2. * q);
            return norm(1., -2. * c, 1., 1. + a, -2. * c, 1. - a);
        }
        case 4: {
            if (!(f > 0. && f < 1.)) return wire;
            if (!(q > 0.)) return BiquadCoefs{-1., 0., 0., 0., 0.};
            double a = s / (2. * q);
            return norm(1. - a, -2. * c, 1. + a, 1. + a, -2. * c, 1. - a);
        }
        case 5: {
            if (!(f > 0. && f < 1.)) return wire;
            if (!(q > 0.)) return BiquadCoefs{A * A, 0., 0., 0., 0.};
            double a = s / (2. * q);
            return norm(1. + a * A, -2. * c, 1. - a * A, 1. + a / A, -2. * c, 1. - a / A);
        }
        case 6: {
            if (f == 1.) return BiquadCoefs{A * A, 0., 0., 0., 0.};
            if (f == 0.) return wire;
            double as = s / 2. * 1.41421356237309504880168872420969808;
            double t = 2. * as * std::sqrt(A), ap = A + 1., am = A - 1.;
            return norm(A * (ap - am * c + t), 2. * A * (am - ap * c), A * (ap - am * c - t), ap + am * c + t, -2. * (am + ap * c),
                        ap + am * c - t);
        }
        default: {
            if (f == 1.) return wire;
            if (!(f > 0.)) return BiquadCoefs{A * A, 0., 0., 0., 0.};
            double as = s / 2. * 1.41421356237309504880168872420969808;
            double t = 2. * as * std::sqrt(A), ap = A + 1., am = A - 1.;
            return norm(A * (ap + am * c + t), -2. * A * (am + ap * c), A * (ap + am * c - t), ap - am * c + t, 2. * (am - ap * c),
                        ap - am * c - t);
        }
    }
}

// get_computed_freq, src/node/biquad_filter.rs:393-399 (f32)
inline float biquad_computed_freq(float freq, float detune) { return detune != 0.f ? freq * exp2f(detune / 1200.f) : freq; }

// sine table, src/node/oscillator.rs:16-28
inline std::vector<float> sine_table() {
    std::vector<float> t(2048);
    for (int x = 0; x < 2048; x++) t[x] = sinf((float)x * 2.0f * PI32 * (1.f / 2048.f));
    return t;
}

inline double unroll_phase(double p) { return p >= 1. ? p - 1. : (p < 0. ? p + 1. : p); }

// Sample-accurate scheduling of AudioScheduledSourceNodes.  The reference walks quanta
// (current_time = frame / sample_rate, src/render/thread.rs:357-360) and, inside the quantum that contains a
// start/stop time, accumulates `current_time += dt` per frame (oscillator.rs:511-557, constant_source.rs:231-246).
// first_frame_at_or_after(T) returns the first frame whose accumulated time is >= T, reproducing that walk.
struct SchedClock {
    double sample_rate, dt;
    explicit SchedClock(float sr) : sample_rate((double)sr), dt(1. / (double)sr) {}
    double block_time(int64_t q) const { return (double)(q * 128) / sample_rate; }
    double next_block_time(int64_t q) const { return block_time(q) + dt * 128.; }
    // first quantum whose next_block_time is > T (i.e. the node is not skipped by `T >= next_block_time`)
    int64_t quantum_containing(double T) const {
        if (!(T < 1e15)) return std::numeric_limits<int64_t>::max() / 256;
        int64_t q = (int64_t)std::floor(T * sample_rate / 128.) - 2;
        if (q < 0) q = 0;
        while (!(T < next_block_time(q))) q++;
        return q;
    }
    // returns frame index; *time_out = accumulated time of that frame
    int64_t first_frame_at_or_after(double T, double* time_out = nullptr) const {
        int64_t q = quantum_containing(T);
        if (q >= std::numeric_limits<int64_t>::max() / 512) return std::numeric_limits<int64_t>::max();
        double t = block_time(q);
        for (int i = 0; i < 128; i++) {
            if (t >= T) {
                if (time_out) *time_out = t;
                return q * 128 + i;
            }
            t += dt;
        }
        if (time_out) *time_out = block_time(q + 1);
        return (q + 1) * 128;
    }
};

// get_stereo_gains, src/node/stereo_panner.rs:74-79
inline void stereo_gains(float x, float& gl, float& gr) {
    gl = sinf((1.f - x) * PI32 / 2.f);
    gr = sinf(x * PI32 / 2.f);
}

// spatial helpers (src/spatial.rs:205-299) live in wae_spatial.h, shared with the device code
using namespace ::wae::spatial;

}  // namespace hostmath
}  // namespace wae
