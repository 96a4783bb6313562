// ORACLE — TEST INFRASTRUCTURE ONLY (see wao_core.h).
// Oscillator, BiquadFilter, IIRFilter, Gain, ConstantSource, AudioBufferSource renderers.
// Every function cites the reference lines it restates.
#include "wao_nodes.h"
#include <array>
#include <complex>

namespace wao {

static const double F64_MAX = 1.7976931348623157e308;

// ================================================================================================
// Oscillator — src/node/oscillator.rs
// ================================================================================================

// oscillator.rs:16-28: sine table entry = ((x as f32) * 2.0 * PI * (1. / 2048.)).sin() in f32
const float* precomputed_sine_table() {
    static std::vector<float> table;
    if (table.empty()) {
        table.resize(2048);
        const float pi = 3.14159265358979323846f;  // std::f32::consts::PI
        for (int x = 0; x < 2048; x++) table[x] = sinf((float)x * 2.0f * pi * (1.f / 2048.f));
    }
    return table.data();
}

// oscillator.rs:30-32
static inline double osc_computed_freq(float freq, float detune) { return (double)freq * std::exp2((double)detune / 1200.); }

// oscillator.rs:661-670
static inline double unroll_phase(double phase) {
    if (phase >= 1.) return phase - 1.;
    if (phase < 0.) return phase + 1.;
    return phase;
}
// oscillator.rs:673-675 (f64::rem_euclid(1.))
static inline double unroll_phase_unbounded(double phase) {
    double r = std::fmod(phase, 1.);
    return r < 0. ? r + 1. : r;
}
// oscillator.rs:645-659 (release semantics: polyBLEP enabled; the reference disables it only under cfg!(test))
static inline double poly_blep(double t, double dt) {
    if (t < dt) {
        t /= dt;
        return t + t - t * t - 1.0;
    } else if (t > 1.0 - dt) {
        t = (t - 1.0) / dt;
        return std::fma(t, t, t) + t + 1.0;
    }
    return 0.0;
}

// oscillator.rs:559-568 + generators :571-642
float OscillatorRenderer::generate_waveform_sample(double phase_incr) {
    switch (type) {
        case 0: {  // sine, :571-585
            double position = phase * 2048.;
            double floored = std::floor(position);
            size_t prev_index = (size_t)floored;
            size_t next_index = prev_index + 1;
            if (next_index == 2048) next_index = 0;
            float k = (float)(position - floored);
            return std::fma(sine_table[prev_index], 1.f - k, sine_table[next_index] * k);
        }
        case 2: {  // sawtooth, :588-595
            double ph = unroll_phase(phase + 0.5);
            double sample = 2.0 * ph - 1.0;
            sample -= poly_blep(ph, phase_incr);
            return (float)sample;
        }
        case 1: {  // square, :598-606
            double sample = phase < 0.5 ? 1.0 : -1.0;
            sample += poly_blep(phase, phase_incr);
            double shift_phase = unroll_phase(phase + 0.5);
            sample -= poly_blep(shift_phase, phase_incr);
            return (float)sample;
        }
        case 3: {  // triangle, :609-619
            double sample = -4. * phase + 2.;
            if (sample > 1.)
                sample = 2. - sample;
            else if (sample < -1.)
                sample = -2. - sample;
            return (float)sample;
        }
        default: {  // custom, :622-637
            size_t table_length = periodic_wave.size();
            double position = phase * (double)table_length;
            double floored = std::floor(position);
            size_t prev_index = (size_t)floored;
            size_t next_index = prev_index + 1;
            if (next_index == table_length) next_index = 0;
            float k = (float)(position - floored);
            return std::fma(periodic_wave[prev_index], 1.f - k, periodic_wave[next_index] * k);
        }
    }
}

// oscillator.rs:511-557
double OscillatorRenderer::generate_sample(float* output, bool outside_nyquist, double phase_incr, double current_time, double dt) {
    if (current_time < start_time || current_time >= stop_time) {
        *output = 0.f;
        return current_time + dt;
    }
    if (!started) {
        if (current_time > start_time) {
            double ratio = (current_time - start_time) / dt;
            phase = outside_nyquist ? unroll_phase_unbounded(phase_incr * ratio) : unroll_phase(phase_incr * ratio);
        }
        started = true;
    }
    *output = outside_nyquist ? 0.f : generate_waveform_sample(phase_incr);
    phase = outside_nyquist ? unroll_phase_unbounded(phase + phase_incr) : unroll_phase(phase + phase_incr);
    return current_time + dt;
}

// oscillator.rs:364-471
bool OscillatorRenderer::process(std::vector<Quantum>&, std::vector<Quantum>& outputs, const ParamValues& params, const Scope& scope) {
    Quantum& output = outputs[0];
    output.set_number_of_channels(1);
    double sample_rate = (double)scope.sample_rate;
    double dt = 1. / sample_rate;
    double next_block_time = scope.current_time + dt * (double)RQ;

    if (stop_time <= scope.current_time) {
        output.make_silent();
        return false;
    } else if (start_time >= next_block_time) {
        output.make_silent();
        if (stop_time <= next_block_time) return false;
        return start_time != F64_MAX;
    }
    float* channel_data = output.channel_mut(0).make_mut();
    ParamSlice frequency_values = params.get(frequency);
    ParamSlice detune_values = params.get(detune);
    double current_time = scope.current_time;
    if (!started && start_time < current_time) start_time = current_time;
    double nyquist = sample_rate / 2.;

    if (frequency_values.len == 1 && detune_values.len == 1) {
        double computed_freq = osc_computed_freq(frequency_values[0], detune_values[0]);
        double phase_incr = computed_freq / sample_rate;
        bool outside_nyquist = std::fabs(computed_freq) >= nyquist;
        bool fully_active = started && start_time <= scope.current_time && stop_time >= next_block_time;
        if (fully_active && !outside_nyquist) {
            for (int i = 0; i < RQ; i++) {
                channel_data[i] = generate_waveform_sample(phase_incr);
                phase = unroll_phase(phase + phase_incr);
            }
        } else {
            for (int i = 0; i < RQ; i++)
                current_time = generate_sample(&channel_data[i], outside_nyquist, phase_incr, current_time, dt);
        }
    } else {
        for (int i = 0; i < RQ; i++) {
            float freq = frequency_values[i % frequency_values.len];
            float det = detune_values[i % detune_values.len];
            double computed_freq = osc_computed_freq(freq, det);
            double phase_incr = computed_freq / sample_rate;
            bool outside_nyquist = std::fabs(computed_freq) >= nyquist;
            current_time = generate_sample(&channel_data[i], outside_nyquist, phase_incr, current_time, dt);
        }
    }
    if (stop_time <= next_block_time) return false;
    return true;
}

// ================================================================================================
// BiquadFilter — src/node/biquad_filter.rs
// ================================================================================================

// biquad_filter.rs:28-40
static BiquadCoefs normalize_coefs(double b0, double b1, double b2, double a0, double a1, double a2) {
    double scale = 1. / a0;
    return BiquadCoefs{b0 * scale, b1 * scale, b2 * scale, a1 * scale, a2 * scale};
}
static const BiquadCoefs WIRE{1., 0., 0., 0., 0.};
static const BiquadCoefs ZERO{0., 0., 0., 0., 0.};
static const double PI64 = 3.14159265358979323846;
static const double SQRT_2 = 1.41421356237309504880168872420969808;

// biquad_filter.rs:42-69
static BiquadCoefs lowpass_coefs(double freq, double q) {
    if (freq == 1.) return WIRE;
    double w0 = PI64 * freq;
    double alpha_q_db = std::sin(w0) / (2. * std::pow(10., q / 20.));
    double cos_w0 = std::cos(w0);
    double beta = (1. - cos_w0) / 2.;
    return normalize_coefs(beta, 2. * beta, beta, 1. + alpha_q_db, -2. * cos_w0, 1. - alpha_q_db);
}
// biquad_filter.rs:71-110
static BiquadCoefs highpass_coefs(double freq, double q) {
    if (freq == 1.) return ZERO;
    if (freq == 0.) return WIRE;
    double w0 = PI64 * freq;
    double alpha_q_db = std::sin(w0) / (2. * std::pow(10., q / 20.));
    double cos_w0 = std::cos(w0);
    double beta = (1. + cos_w0) / 2.;
    return normalize_coefs(beta, -2. * beta, beta, 1. + alpha_q_db, -2. * cos_w0, 1. - alpha_q_db);
}
// biquad_filter.rs:112-152
static BiquadCoefs bandpass_coefs(double freq, double q) {
    if (freq > 0. && freq < 1.) {
        if (q > 0.) {
            double w0 = PI64 * freq;
            double alpha_q = std::sin(w0) / (2. * q);
            double cos_w0 = std::cos(w0);
            return normalize_coefs(alpha_q, 0., -alpha_q, 1. + alpha_q, -2. * cos_w0, 1. - alpha_q);
        }
        return WIRE;
    }
    return ZERO;
}
// biquad_filter.rs:154-193
static BiquadCoefs notch_coefs(double freq, double q) {
    if (freq > 0. && freq < 1.) {
        if (q > 0.) {
            double w0 = PI64 * freq;
            double alpha_q = std::sin(w0) / (2. * q);
            double cos_w0 = std::cos(w0);
            return normalize_coefs(1., -2. * cos_w0, 1., 1. + alpha_q, -2. * cos_w0, 1. - alpha_q);
        }
        return ZERO;
    }
    return WIRE;
}
// biquad_filter.rs:195-231
static BiquadCoefs allpass_coefs(double freq, double q) {
    if (freq > 0. && freq < 1.) {
        if (q > 0.) {
            double w0 = PI64 * freq;
            double alpha_q = std::sin(w0) / (2. * q);
            double cos_w0 = std::cos(w0);
            return normalize_coefs(1. - alpha_q, -2. * cos_w0, 1. + alpha_q, 1. + alpha_q, -2. * cos_w0, 1. - alpha_q);
        }
        return BiquadCoefs{-1., 0., 0., 0., 0.};
    }
    return WIRE;
}
// biquad_filter.rs:233-276
static BiquadCoefs peaking_coefs(double freq, double q, double gain) {
    double A = std::pow(10., gain / 40.);
    if (freq > 0. && freq < 1.) {
        if (q > 0.) {
            double w0 = PI64 * freq;
            double alpha_q = std::sin(w0) / (2. * q);
            double cos_w0 = std::cos(w0);
            return normalize_coefs(1. + alpha_q * A, -2. * cos_w0, 1. - alpha_q * A, 1. + alpha_q / A, -2. * cos_w0, 1. - alpha_q / A);
        }
        return BiquadCoefs{A * A, 0., 0., 0., 0.};
    }
    return WIRE;
}
// biquad_filter.rs:278-320
static BiquadCoefs lowshelf_coefs(double freq, double gain) {
    double A = std::pow(10., gain / 40.);
    if (freq == 1.) return BiquadCoefs{A * A, 0., 0., 0., 0.};
    if (freq == 0.) return WIRE;
    double w0 = PI64 * freq;
    double cos_w0 = std::cos(w0);
    double alpha_s = std::sin(w0) / 2. * SQRT_2;
    double two_alpha_s_a_squared = 2. * alpha_s * std::sqrt(A);
    double a_plus_one = A + 1.;
    double a_minus_one = A - 1.;
    double b0 = A * (a_plus_one - a_minus_one * cos_w0 + two_alpha_s_a_squared);
    double b1 = 2. * A * (a_minus_one - a_plus_one * cos_w0);
    double b2 = A * (a_plus_one - a_minus_one * cos_w0 - two_alpha_s_a_squared);
    double a0 = a_plus_one + a_minus_one * cos_w0 + two_alpha_s_a_squared;
    double a1 = -2. * (a_minus_one + a_plus_one * cos_w0);
    double a2 = a_plus_one + a_minus_one * cos_w0 - two_alpha_s_a_squared;
    return normalize_coefs(b0, b1, b2, a0, a1, a2);
}
// biquad_filter.rs:322-364
static BiquadCoefs highshelf_coefs(double freq, double gain) {
    double A = std::pow(10., gain / 40.);
    if (freq == 1.) return WIRE;
    if (freq > 0.) {
        double w0 = PI64 * freq;
        double cos_w0 = std::cos(w0);
        double alpha_s = std::sin(w0) / 2. * SQRT_2;
        double two_alpha_s_a_squared = 2. * alpha_s * std::sqrt(A);
        double a_plus_one = A + 1.;
        double a_minus_one = A - 1.;
        double b0 = A * (a_plus_one + a_minus_one * cos_w0 + two_alpha_s_a_squared);
        double b1 = -2. * A * (a_minus_one + a_plus_one * cos_w0);
        double b2 = A * (a_plus_one + a_minus_one * cos_w0 - two_alpha_s_a_squared);
        double a0 = a_plus_one - a_minus_one * cos_w0 + two_alpha_s_a_squared;
        double a1 = 2. * (a_minus_one - a_plus_one * cos_w0);
        double a2 = a_plus_one - a_minus_one * cos_w0 - two_alpha_s_a_squared;
        return normalize_coefs(b0, b1, b2, a0, a1, a2);
    }
    return BiquadCoefs{A * A, 0., 0., 0., 0.};
}

// biquad_filter.rs:367-390
BiquadCoefs biquad_calculate_coefs(int type, double sample_rate, double f0, double gain, double q) {
    double nyquist = sample_rate / 2.;
    double norm_freq = f0 / nyquist;
    norm_freq = norm_freq < 0. ? 0. : (norm_freq > 1. ? 1. : norm_freq);  // clamp(0., 1.)
    switch (type) {
        case 0: return lowpass_coefs(norm_freq, q);
        case 1: return highpass_coefs(norm_freq, q);
        case 2: return bandpass_coefs(norm_freq, q);
        case 3: return notch_coefs(norm_freq, q);
        case 4: return allpass_coefs(norm_freq, q);
        case 5: return peaking_coefs(norm_freq, q, gain);
        case 6: return lowshelf_coefs(norm_freq, gain);
        default: return highshelf_coefs(norm_freq, gain);
    }
}

// biquad_filter.rs:393-399
float biquad_computed_freq(float freq, float detune) {
    if (detune != 0.f) return freq * exp2f(detune / 1200.f);
    return freq;
}

// biquad_filter.rs:670-737
void biquad_frequency_response(int type, float sample_rate, float frequency, float detune, float q, float gain,
                               const float* freq_hz, float* mag, float* phase, int n) {
    float n_quist = sample_rate / 2.f;
    float computed_freq = biquad_computed_freq(frequency, detune);
    BiquadCoefs c = biquad_calculate_coefs(type, (double)sample_rate, (double)computed_freq, (double)gain, (double)q);
    for (int i = 0; i < n; i++) {
        float freq = freq_hz[i];
        if (freq < 0.f || freq > n_quist) {
            mag[i] = NAN;
            phase[i] = NAN;
        } else {
            float f = freq / n_quist;
            double omega = -PI64 * (double)f;
            std::complex<double> z(std::cos(omega), std::sin(omega));
            std::complex<double> numerator = c.b0 + (c.b1 + c.b2 * z) * z;
            std::complex<double> denominator = std::complex<double>(1., 0.) + (c.a1 + c.a2 * z) * z;
            std::complex<double> response = numerator / denominator;
            mag[i] = (float)std::abs(response);
            phase[i] = (float)std::arg(response);
        }
    }
}

// biquad_filter.rs:764-899
bool BiquadFilterRenderer::process(std::vector<Quantum>& inputs, std::vector<Quantum>& outputs, const ParamValues& params, const Scope& scope) {
    const Quantum& input = inputs[0];
    Quantum& output = outputs[0];
    float sample_rate = scope.sample_rate;

    if (input.is_silent()) {
        bool ended = true;
        for (auto& v : xy)
            for (double s : v)
                if (is_normal(s)) ended = false;
        if (ended) {
            output.make_silent();
            return false;
        }
    }
    if (!input.is_silent()) {
        size_t num_channels = (size_t)input.number_of_channels();
        if (num_channels != xy.size()) {
            if (xy.size() > num_channels) xy.resize(num_channels);
            while (xy.size() < num_channels) xy.push_back({0., 0., 0., 0.});
        }
        output.set_number_of_channels((int)num_channels);
    } else {
        output.set_number_of_channels((int)xy.size());
    }

    ParamSlice frequency_v = params.get(frequency);
    ParamSlice detune_v = params.get(detune);
    ParamSlice q_v = params.get(q);
    ParamSlice gain_v = params.get(gain);
    double sample_rate_f64 = (double)sample_rate;
    float computed_freq = biquad_computed_freq(frequency_v[0], detune_v[0]);
    BiquadCoefs coef = biquad_calculate_coefs(type, sample_rate_f64, (double)computed_freq, (double)gain_v[0], (double)q_v[0]);
    BiquadCoefs coefs_list[RQ];
    for (int i = 0; i < RQ; i++) coefs_list[i] = coef;
    if (frequency_v.len != 1 || detune_v.len != 1 || q_v.len != 1 || gain_v.len != 1) {
        for (int i = 1; i < RQ; i++) {
            float f = frequency_v[i % frequency_v.len], d = detune_v[i % detune_v.len];
            float qq = q_v[i % q_v.len], g = gain_v[i % gain_v.len];
            float cf = biquad_computed_freq(f, d);
            coefs_list[i] = biquad_calculate_coefs(type, sample_rate_f64, (double)cf, (double)g, (double)qq);
        }
    }
    for (int ch = 0; ch < output.number_of_channels(); ch++) {
        // hold a clone of the input channel: output may alias input buffers otherwise
        Channel input_channel = input.is_silent() ? input.channel(0) : input.channel(ch);
        const float* in = input_channel.data();
        float* out = output.channel_mut(ch).make_mut();
        double x1 = xy[ch][0], x2 = xy[ch][1], y1 = xy[ch][2], y2 = xy[ch][3];
        for (int i = 0; i < RQ; i++) {
            const BiquadCoefs& c = coefs_list[i];
            double x = (double)in[i];
            double y = c.b0 * x + c.b1 * x1 + c.b2 * x2 - c.a1 * y1 - c.a2 * y2;
            if (!is_normal(y)) y = 0.;
            x2 = x1;
            x1 = x;
            y2 = y1;
            y1 = y;
            out[i] = (float)y;
        }
        xy[ch] = {x1, x2, y1, y2};
    }
    return true;
}

// ================================================================================================
// IIRFilter — src/node/iir_filter.rs
// ================================================================================================

// iir_filter.rs:282-320
IirFilterRenderer::IirFilterRenderer(std::vector<double> feedforward, std::vector<double> feedback) {
    if (feedforward.size() < feedback.size()) feedforward.resize(feedback.size(), 0.);
    if (feedforward.size() > feedback.size()) feedback.resize(feedforward.size(), 0.);
    double a0 = feedback[0];
    for (size_t i = 0; i < feedforward.size(); i++) norm_coeffs.emplace_back(feedforward[i] / a0, feedback[i] / a0);
    std::array<double, 20> z{};
    states.push_back(z);
    states.push_back(z);
}

// iir_filter.rs:323-414
bool IirFilterRenderer::process(std::vector<Quantum>& inputs, std::vector<Quantum>& outputs, const ParamValues&, const Scope&) {
    const Quantum& input = inputs[0];
    Quantum& output = outputs[0];
    size_t nc = norm_coeffs.size();
    if (input.is_silent()) {
        bool ended = true;
        for (auto& st : states) {
            for (size_t i = 0; i < nc; i++)
                if (is_normal(st[i])) ended = false;
            if (!ended) break;
        }
        if (ended) {
            output.make_silent();
            return false;
        }
    }
    if (!input.is_silent()) {
        size_t num_channels = (size_t)input.number_of_channels();
        if (num_channels != states.size()) {
            if (states.size() > num_channels) states.resize(num_channels);
            std::array<double, 20> z{};
            while (states.size() < num_channels) states.push_back(z);
        }
        output.set_number_of_channels((int)num_channels);
    } else {
        output.set_number_of_channels((int)states.size());
    }
    for (int ch = 0; ch < output.number_of_channels(); ch++) {
        Channel input_channel = input.is_silent() ? input.channel(0) : input.channel(ch);
        const float* in = input_channel.data();
        float* out = output.channel_mut(ch).make_mut();
        std::array<double, 20>& st = states[ch];
        for (int n = 0; n < RQ; n++) {
            double x = (double)in[n];
            double b0 = norm_coeffs[0].first;
            double last_state = st[0];
            double y = std::fma(b0, x, last_state);
            if (!is_normal(y)) y = 0.;
            for (size_t i = 0; i + 1 < nc; i++) {
                double b = norm_coeffs[i + 1].first, a = norm_coeffs[i + 1].second;
                double state = st[i + 1];
                st[i] = b * x - a * y + state;
            }
            out[n] = (float)y;
        }
    }
    return true;
}

// iir_filter.rs:221-265
void iir_frequency_response(const std::vector<double>& ff, const std::vector<double>& fb, float sample_rate_f32,
                            const float* freq_hz, float* mag, float* phase, int n) {
    double sample_rate = (double)sample_rate_f32;
    double nquist = sample_rate / 2.;
    for (int i = 0; i < n; i++) {
        double freq = (double)freq_hz[i];
        if (freq < 0. || freq > nquist) {
            mag[i] = NAN;
            phase[i] = NAN;
        } else {
            double z = -2.0 * PI64 * freq / sample_rate;
            std::complex<double> num(0., 0.), denom(0., 0.);
            for (size_t idx = 0; idx < ff.size(); idx++) num += std::polar(1.0, (double)idx * z) * ff[idx];
            for (size_t idx = 0; idx < fb.size(); idx++) denom += std::polar(1.0, (double)idx * z) * fb[idx];
            std::complex<double> response = num / denom;
            mag[i] = (float)std::abs(response);
            phase[i] = (float)std::arg(response);
        }
    }
}

// ================================================================================================
// Gain — src/node/gain.rs:130-199
// ================================================================================================
bool GainRenderer::process(std::vector<Quantum>& inputs, std::vector<Quantum>& outputs, const ParamValues& params, const Scope&) {
    const Quantum& input = inputs[0];
    Quantum& output = outputs[0];
    if (input.is_silent()) {
        output.make_silent();
        return false;
    }
    ParamSlice g = params.get(gain);
    if (g.len == 1) {
        float threshold = 1e-6f;
        float diff_to_zero = std::fabs(g[0]);
        if (diff_to_zero <= threshold) {
            output.make_silent();
            return false;
        }
        float diff_to_one = std::fabs(1.f - g[0]);
        if (diff_to_one <= threshold) {
            output = input;
            return false;
        }
    }
    output = input;
    if (g.len == 1) {
        float gv = g[0];
        for (int c = 0; c < output.number_of_channels(); c++) {
            float* o = output.channel_mut(c).make_mut();
            for (int i = 0; i < RQ; i++) o[i] *= gv;
        }
    } else {
        for (int c = 0; c < output.number_of_channels(); c++) {
            float* o = output.channel_mut(c).make_mut();
            for (int i = 0; i < RQ; i++) o[i] *= g[i % g.len];
        }
    }
    return false;
}

// ================================================================================================
// ConstantSource — src/node/constant_source.rs:190-262
// ================================================================================================
bool ConstantSourceRenderer::process(std::vector<Quantum>&, std::vector<Quantum>& outputs, const ParamValues& params, const Scope& scope) {
    Quantum& output = outputs[0];
    double dt = 1. / (double)scope.sample_rate;
    double next_block_time = scope.current_time + dt * (double)RQ;
    if (start_time >= next_block_time) {
        output.make_silent();
        if (stop_time <= next_block_time) return false;
        return start_time != F64_MAX;
    }
    output.force_mono();
    ParamSlice off = params.get(offset);
    float* o = output.channel_mut(0).make_mut();
    if (off.len == 1 && start_time <= scope.current_time && stop_time >= next_block_time) {
        for (int i = 0; i < RQ; i++) o[i] = off[0];
    } else {
        double current_time = scope.current_time;
        for (int i = 0; i < RQ; i++) {
            if (current_time < start_time || current_time >= stop_time)
                o[i] = 0.f;
            else
                o[i] = off[i % off.len];
            current_time += dt;
        }
    }
    return stop_time > next_block_time;
}

// ================================================================================================
// AudioBufferSource — src/node/audio_buffer_source.rs
// `almost` crate 0.2 (Cargo.toml:21, not vendored): equal(a,b) := a == b || |a-b| <= tol ||
// |a-b| <= max(|a|,|b|) * tol; zero(a) := |a| < tol; tol = sqrt(f64::EPSILON) = 1.4901161193847656e-8.
// ================================================================================================
static const double ALMOST_TOL = 1.4901161193847656e-8;
static inline bool almost_equal(double a, double b) {
    if (a == b) return true;
    if (std::isinf(a) || std::isinf(b) || std::isnan(a) || std::isnan(b)) return false;
    double abs_diff = std::fabs(b - a);
    if (abs_diff <= ALMOST_TOL) return true;
    double largest = std::max(std::fabs(a), std::fabs(b));
    return abs_diff <= largest * ALMOST_TOL;
}
static inline bool almost_zero(double a) { return std::fabs(a) < ALMOST_TOL; }

// audio_buffer_source.rs:400-417
void AudioBufferSourceRenderer::clamp_loop_boundaries() {
    if (buffer) {
        double duration = buffer->duration();
        if (loop_start < 0.)
            loop_start = 0.;
        else if (loop_start > duration)
            loop_start = duration;
        if (loop_end <= 0. || loop_end > duration) loop_end = duration;
    }
}

// audio_buffer_source.rs:421-845
bool AudioBufferSourceRenderer::process(std::vector<Quantum>&, std::vector<Quantum>& outputs, const ParamValues& params, const Scope& scope) {
    Quantum& output = outputs[0];
    if (ended) {
        output.make_silent();
        return false;
    }
    double sample_rate = (double)scope.sample_rate;
    double dt = 1. / sample_rate;
    double block_duration = dt * (double)RQ;
    double next_block_time = scope.current_time + block_duration;

    if (!buffer && start_time != F64_MAX) {
        output.make_silent();
        ended = true;
        return false;
    }
    if (start_time >= next_block_time) {
        output.make_silent();
        if (stop_time <= next_block_time) {
            ended = true;
            return false;
        }
        return start_time != F64_MAX;
    }
    if (!buffer) {
        output.make_silent();
        return false;
    }
    const AudioBuffer& buf = *buffer;
    double actual_loop_start = 0., actual_loop_end = 0.;
    double detune_v = (double)params.get(detune)[0];
    double playback_rate_v = (double)params.get(playback_rate)[0];
    double computed_playback_rate = playback_rate_v * std::exp2(detune_v / 1200.);
    size_t buffer_length = buf.length();
    double buffer_duration = buf.duration();
    double sampling_ratio = (double)buf.sample_rate / sample_rate;
    double buffer_time = this->buffer_time;

    output.set_number_of_channels(buf.number_of_channels());
    double block_time = scope.current_time;
    if (!started && start_time < block_time) start_time = block_time;
    if (start_time == block_time && offset == 0.) is_aligned = true;
    if (sampling_ratio != 1. || computed_playback_rate != 1.) is_aligned = false;
    if (loop_start != 0. || loop_end != buffer_duration) is_aligned = false;
    if (buffer_time + block_duration > duration || block_time + block_duration > stop_time) is_aligned = false;

    if (is_aligned) {
        // ---- fast track (:554-624)
        if (start_time == block_time) started = true;
        if (buffer_time + block_duration > buffer_duration) {
            size_t end_index = buf.length();
            bool has_loop_point = false;
            size_t loop_point_index = 0;
            for (int c = 0; c < buf.number_of_channels(); c++) {
                const float* bc = buf.channels[c].data();
                float* oc = output.channel_mut(c).make_mut();
                size_t start_index = (size_t)std::llround(buffer_time * sample_rate);  // f64::round: half away from zero
                size_t off = 0;
                for (size_t index = 0; index < (size_t)RQ; index++) {
                    size_t buffer_index = start_index + index - off;
                    if (buffer_index < end_index) {
                        oc[index] = bc[buffer_index];
                    } else {
                        if (is_looping && buffer_index >= end_index) {
                            has_loop_point = true;
                            loop_point_index = index;
                            start_index = 0;
                            off = index;
                            buffer_index = 0;
                        }
                        oc[index] = is_looping ? bc[buffer_index] : 0.f;
                    }
                }
            }
            if (has_loop_point)
                buffer_time = std::fmod((double)(RQ - loop_point_index) / sample_rate, buffer_duration);
            else
                buffer_time += block_duration;
        } else {
            size_t start_index = (size_t)std::llround(buffer_time * sample_rate);
            for (int c = 0; c < buf.number_of_channels(); c++) {
                float* oc = output.channel_mut(c).make_mut();
                std::memcpy(oc, buf.channels[c].data() + start_index, sizeof(float) * RQ);
            }
            buffer_time += block_duration;
        }
        buffer_time_elapsed += block_duration;
    } else {
        // ---- slow track (:625-823)
        if (is_looping) {
            if (loop_start >= 0. && loop_end > 0. && loop_start < loop_end) {
                actual_loop_start = loop_start;
                actual_loop_end = loop_end;
            } else {
                actual_loop_start = 0.;
                actual_loop_end = buffer_duration;
            }
        } else {
            entered_loop = false;
        }
        struct Info {
            bool some;
            size_t prev_frame_index;
            double k;
        };
        Info infos[RQ];
        for (int i = 0; i < RQ; i++) {
            infos[i].some = false;
            double current_time = block_time + (double)i * dt;
            if (!started && almost_equal(current_time, start_time)) start_time = current_time;
            if (almost_equal(buffer_time_elapsed, duration)) buffer_time_elapsed = duration;
            if (current_time < start_time || current_time >= stop_time || buffer_time_elapsed >= duration) continue;
            if (!started) {
                double delta = current_time - start_time;
                offset += delta * computed_playback_rate;
                offset = std::min(std::max(offset, 0.), buffer_duration);
                if (is_looping && computed_playback_rate >= 0. && offset > actual_loop_end) offset = actual_loop_end;
                if (is_looping && computed_playback_rate < 0. && offset < actual_loop_start) offset = actual_loop_start;
                buffer_time = offset;
                buffer_time_elapsed = std::fabs(delta * computed_playback_rate);
                started = true;
            }
            if (is_looping) {
                if (almost_equal(buffer_time, actual_loop_end)) buffer_time = actual_loop_end;
                if (almost_equal(buffer_time, actual_loop_start)) buffer_time = actual_loop_start;
                if (!entered_loop) {
                    if (offset < actual_loop_end && buffer_time >= actual_loop_start) entered_loop = true;
                    if (offset >= actual_loop_end && buffer_time < actual_loop_end) entered_loop = true;
                }
                if (entered_loop) {
                    while (buffer_time >= actual_loop_end) buffer_time -= actual_loop_end - actual_loop_start;
                    while (buffer_time < actual_loop_start) buffer_time += actual_loop_end - actual_loop_start;
                }
            }
            if (almost_zero(buffer_time)) buffer_time = 0.;
            if (buffer_time >= 0. && buffer_time < buffer_duration) {
                double position = buffer_time * sampling_ratio;
                double playhead = position * sample_rate;
                double playhead_floored = std::floor(playhead);
                size_t prev_frame_index = (size_t)playhead_floored;
                double k = playhead - playhead_floored;
                if (prev_frame_index < buffer_length) {
                    infos[i].some = true;
                    infos[i].prev_frame_index = prev_frame_index;
                    infos[i].k = k;
                }
            }
            double time_incr = dt * computed_playback_rate;
            buffer_time += time_incr;
            buffer_time_elapsed += std::fabs(time_incr);
        }
        for (int c = 0; c < buf.number_of_channels(); c++) {
            const std::vector<float>& bc = buf.channels[c];
            float* oc = output.channel_mut(c).make_mut();
            for (int i = 0; i < RQ; i++) {
                if (!infos[i].some) {
                    oc[i] = 0.f;
                    continue;
                }
                size_t pfi = infos[i].prev_frame_index;
                double k = infos[i].k;
                double prev_sample = (double)bc[pfi];
                double next_sample;
                if (pfi + 1 < bc.size()) {
                    next_sample = (double)bc[pfi + 1];
                } else if (is_looping) {
                    if (playback_rate_v >= 0.) {
                        double start_playhead = actual_loop_start * sample_rate;
                        size_t start_index = (std::floor(start_playhead) == start_playhead) ? (size_t)start_playhead : (size_t)start_playhead + 1;
                        next_sample = (double)bc[start_index];
                    } else {
                        double end_playhead = actual_loop_end * sample_rate;
                        size_t end_index = (size_t)end_playhead;
                        next_sample = (double)bc[end_index];
                    }
                } else {
                    if (almost_equal(k, 1.) || pfi == 0) {
                        next_sample = 0.;
                    } else {
                        float prev_prev_sample = bc[pfi - 1];
                        next_sample = 2. * prev_sample - (double)prev_prev_sample;
                    }
                }
                oc[i] = (float)std::fma(1. - k, prev_sample, k * next_sample);
            }
        }
    }
    this->buffer_time = buffer_time;
    if (next_block_time >= stop_time || buffer_time_elapsed >= duration ||
        (!is_looping && ((computed_playback_rate > 0. && buffer_time >= buffer_duration) || (computed_playback_rate < 0. && buffer_time < 0.)))) {
        ended = true;
    }
    return true;
}

}  // namespace wao
