// ORACLE — TEST INFRASTRUCTURE ONLY.  See wao_resampler.h (rubato 0.16 FftFixedInOut restated, parity unpinned).
#include "wao_resampler.h"

#include <cmath>
#include <numeric>

namespace wao {

static const float PI32 = 3.14159265358979323846f;

// rubato windows.rs: blackman_harris (periodic form, f32) squared
static std::vector<float> blackman_harris2(size_t npoints) {
    std::vector<float> w(npoints);
    const float pi2 = 2.f * PI32, pi4 = 4.f * PI32, pi6 = 6.f * PI32;
    const float np_f = (float)npoints;
    const float a = 0.35875f, b = 0.48829f, c = 0.14128f, d = 0.01168f;
    for (size_t x = 0; x < npoints; x++) {
        const float xf = (float)x;
        const float v = a - b * std::cos(pi2 * xf / np_f) + c * std::cos(pi4 * xf / np_f) - d * std::cos(pi6 * xf / np_f);
        w[x] = v * v;
    }
    return w;
}
static float sincf(float v) {
    if (v == 0.f) return 1.f;
    return std::sin(v * PI32) / (v * PI32);
}
// rubato sinc.rs: make_sincs(npoints, factor = 1, f_cutoff, BlackmanHarris2)[0]
std::vector<float> rubato_sinc_filter(size_t npoints, float f_cutoff) {
    std::vector<float> window = blackman_harris2(npoints);
    std::vector<float> y(npoints);
    float sum = 0.f;
    for (size_t x = 0; x < npoints; x++) {
        const float val = window[x] * sincf(((float)x - (float)(npoints / 2)) * f_cutoff);
        sum += val;
        y[x] = val;
    }
    for (size_t x = 0; x < npoints; x++) y[x] = y[x] / sum;
    return y;
}

FftFixedInOut::FftFixedInOut(size_t sample_rate_in, size_t sample_rate_out, size_t chunk_size_in, size_t channels) {
    const size_t gcd = std::gcd(sample_rate_in, sample_rate_out);
    const size_t fft_chunks = (size_t)std::ceil((float)chunk_size_in / (float)(sample_rate_in / gcd));
    fft_size_out = fft_chunks * sample_rate_out / gcd;
    fft_size_in = fft_chunks * sample_rate_in / gcd;
    // FftResampler::new: anti-aliasing cutoff and filter spectrum
    float cutoff = std::pow(0.4f, 16.0f / (float)fft_size_in);
    if (fft_size_in > fft_size_out) cutoff = cutoff * (float)fft_size_out / (float)fft_size_in;
    std::vector<float> sinc = rubato_sinc_filter(fft_size_in, cutoff);
    std::vector<float> filter_t(2 * fft_size_in, 0.f);
    for (size_t n = 0; n < fft_size_in; n++) filter_t[n] = sinc[n] / (float)(2 * fft_size_in);
    fft.init((int)(2 * fft_size_in));
    ifft.init((int)(2 * fft_size_out));
    filter_f.resize(fft_size_in + 1);
    fft.forward(filter_t.data(), filter_f.data());
    overlaps.assign(channels, std::vector<float>(fft_size_out, 0.f));
}

void FftFixedInOut::process(const std::vector<std::vector<float>>& in, std::vector<std::vector<float>>& out) {
    out.resize(in.size());
    std::vector<float> input_buf(2 * fft_size_in), output_buf(2 * fft_size_out);
    std::vector<cf32> input_f(fft_size_in + 1), output_f(fft_size_out + 1);
    const size_t new_len = fft_size_in < fft_size_out ? fft_size_in : fft_size_out;
    for (size_t c = 0; c < in.size(); c++) {
        for (size_t n = 0; n < fft_size_in; n++) input_buf[n] = in[c][n];
        for (size_t n = fft_size_in; n < 2 * fft_size_in; n++) input_buf[n] = 0.f;
        fft.forward(input_buf.data(), input_f.data());
        for (size_t k = 0; k < new_len; k++) output_f[k] = input_f[k] * filter_f[k];
        for (size_t k = new_len; k < output_f.size(); k++) output_f[k] = cf32(0.f, 0.f);
        ifft.inverse(output_f.data(), output_buf.data());
        out[c].resize(fft_size_out);
        for (size_t n = 0; n < fft_size_out; n++) out[c][n] = output_buf[n] + overlaps[c][n];
        for (size_t n = 0; n < fft_size_out; n++) overlaps[c][n] = output_buf[fft_size_out + n];
    }
}

}  // namespace wao
