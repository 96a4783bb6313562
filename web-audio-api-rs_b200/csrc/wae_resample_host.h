// Host half of the over-sampled WaveShaper (src/node/waveshaper.rs:236-348,409-480): the anti-aliasing filter spectra of
// the synchronous FFT resamplers the reference builds with rubato::FftFixedInOut (128 -> 128 * factor frames up, back
// down).  Filter = Blackman-Harris^2 windowed sinc of fft_size_in taps, unit sum, divided by 2 * fft_size_in, transformed
// with a 2 * fft_size_in point real FFT; only the first min(fft_size_in, fft_size_out) = 128 bins are ever used.
// (rubato 0.16 is not vendored with the reference: published algorithm, parity unpinned — see DESIGN.md §6.)
#pragma once
#include <cmath>
#include <vector>
#include <vector_types.h>

namespace wae {

inline std::vector<float2> resampler_filter_bins(int fft_size_in, int fft_size_out, int bins) {
    const float PI32 = 3.14159265358979323846f;
    float cutoff = std::pow(0.4f, 16.0f / (float)fft_size_in);
    if (fft_size_in > fft_size_out) cutoff = cutoff * (float)fft_size_out / (float)fft_size_in;
    const int np = fft_size_in;
    std::vector<float> y(np);
    float sum = 0.f;
    for (int x = 0; x < np; x++) {
        const float xf = (float)x, nf = (float)np;
        const float w = 0.35875f - 0.48829f * std::cos(2.f * PI32 * xf / nf) + 0.14128f * std::cos(4.f * PI32 * xf / nf) -
                        0.01168f * std::cos(6.f * PI32 * xf / nf);
        const float t = (xf - (float)(np / 2)) * cutoff;
        const float sinc = t == 0.f ? 1.f : std::sin(t * PI32) / (t * PI32);
        const float val = (w * w) * sinc;
        sum += val;
        y[x] = val;
    }
    std::vector<float2> out(bins);
    const int N = 2 * fft_size_in;
    for (int k = 0; k < bins; k++) {
        double re = 0., im = 0.;
        for (int n = 0; n < np; n++) {
            const double v = (double)((y[n] / sum) / (float)N);
            const double a = -2.0 * 3.14159265358979323846 * (double)k * (double)n / (double)N;
            re += v * std::cos(a);
            im += v * std::sin(a);
        }
        out[k] = make_float2((float)re, (float)im);
    }
    return out;
}

}  // namespace wae
