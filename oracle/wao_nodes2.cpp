// ORACLE — TEST INFRASTRUCTURE ONLY (see wao_core.h).
#include "wao_nodes2.h"

namespace wao {

// ================================================================================================
// WaveShaper — src/node/waveshaper.rs
// ================================================================================================

// waveshaper.rs:555-572
float waveshaper_apply_curve(const std::vector<float>& curve, float input) {
    if (curve.empty()) return 0.f;
    float n = (float)curve.size();
    float v = (n - 1.f) / 2.0f * (input + 1.f);
    if (v <= 0.f) {
        return curve[0];
    } else if (v >= n - 1.f) {
        return curve[(size_t)(n - 1.f)];
    } else {
        float k = std::floor(v);
        float f = v - k;
        return (1.f - f) * curve[(size_t)k] + f * curve[(size_t)(k + 1.f)];
    }
}

// waveshaper.rs:480-503
void WaveShaperRenderer::set_curve(const float* c, size_t n) {
    if (!c) {
        has_curve = false;
        curve.clear();
        can_propagate_silence = true;
        return;
    }
    has_curve = true;
    curve.assign(c, c + n);
    if (n % 2 == 1) {
        can_propagate_silence = std::fabs(curve[n / 2]) < 1e-9f;
    } else {
        float a = curve[n / 2 - 1], b = curve[n / 2];
        can_propagate_silence = std::fabs((a + b) / 2.f) < 1e-9f;
    }
}

// waveshaper.rs:370-477 (OverSampleType::None)
bool WaveShaperRenderer::process(std::vector<Quantum>& inputs, std::vector<Quantum>& outputs, const ParamValues&, const Scope&) {
    const Quantum& input = inputs[0];
    Quantum& output = outputs[0];
    if (input.is_silent() && can_propagate_silence) {
        output.make_silent();
        return false;
    }
    output = input;
    if (has_curve && oversample == 0) {
        for (int c = 0; c < output.number_of_channels(); c++) {
            float* o = output.channel_mut(c).make_mut();
            for (int i = 0; i < RQ; i++) o[i] = waveshaper_apply_curve(curve, o[i]);
        }
    } else if (has_curve) {  // X2 / X4 (:409-480): up-sample, shape, down-sample; samplers rebuilt when the channel count changes
        const size_t factor = oversample == 1 ? 2 : 4;
        const size_t channels = (size_t)output.number_of_channels();
        if (channels != os_channels) {
            os_channels = channels;
            upsampler = FftFixedInOut(sample_rate, sample_rate * factor, RQ, channels);
            downsampler = FftFixedInOut(sample_rate * factor, sample_rate, RQ * factor, channels);
        }
        std::vector<std::vector<float>> in(channels), up, down;
        for (size_t c = 0; c < channels; c++) in[c].assign(output.channel((int)c).data(), output.channel((int)c).data() + RQ);
        upsampler.process(in, up);
        for (auto& ch : up)
            for (float& s : ch) s = waveshaper_apply_curve(curve, s);
        downsampler.process(up, down);
        for (size_t c = 0; c < channels; c++) {
            float* o = output.channel_mut((int)c).make_mut();
            for (int i = 0; i < RQ; i++) o[i] = down[c][i];
        }
    }
    return false;
}

// ================================================================================================
// Delay — src/node/delay.rs
// ================================================================================================

static void delay_check_ring(DelayShared& sh, const Quantum& q) {
    // delay.rs:395-407: fill the ring with silence up to its capacity
    if (sh.ring.size() < sh.capacity) {
        Quantum silence = q;
        silence.make_silent();
        sh.ring.resize(sh.capacity, silence);
    }
}

// delay.rs:428-461
bool DelayWriter::process(std::vector<Quantum>& inputs, std::vector<Quantum>& outputs, const ParamValues&, const Scope& scope) {
    Quantum input = inputs[0];
    Quantum& output = outputs[0];
    delay_check_ring(*sh, input);
    // delay.rs:470-488 check_ring_buffer_up_down_mix
    int buffer_ch = sh->ring[0].number_of_channels();
    int input_ch = input.number_of_channels();
    if (buffer_ch != input_ch)
        for (auto& q : sh->ring) q.mix(input_ch, SPEAKERS);
    sh->ring[index] = input;
    index = (index + 1) % sh->capacity;
    sh->latest_frame_written = scope.current_frame;
    output.make_silent();
    return false;
}

struct PlaybackInfo {
    size_t prev_block_index = 0, prev_frame_index = 0;
    float k = 0.f;
};

// delay.rs:688-743
static PlaybackInfo get_playback_infos(double delay, bool in_cycle, double sample_index, double quantum_duration,
                                       double sample_rate, int ring_size, int ring_index) {
    double clamped_delay = in_cycle ? std::max(delay, quantum_duration) : delay;
    double num_samples = clamped_delay * sample_rate;
    double position = sample_index - num_samples;
    double position_floored = std::floor(position);
    int num_frames = RQ;
    double block_offset = std::floor(position_floored / (double)num_frames);
    int prev_block_index = ring_index + (int)block_offset;
    if (prev_block_index < 0) prev_block_index += ring_size;
    int frame_offset = (int)position_floored % num_frames;
    if (frame_offset == 0) frame_offset = -num_frames;
    int prev_frame_index = frame_offset <= 0 ? num_frames + frame_offset : frame_offset;
    float k = (float)(position - position_floored);
    PlaybackInfo p;
    p.prev_block_index = (size_t)prev_block_index;
    p.prev_frame_index = (size_t)prev_frame_index;
    p.k = k;
    return p;
}

// delay.rs:515-684
bool DelayReader::process(std::vector<Quantum>&, std::vector<Quantum>& outputs, const ParamValues& params, const Scope& scope) {
    Quantum& output = outputs[0];
    delay_check_ring(*sh, output);
    std::vector<Quantum>& ring = sh->ring;
    int number_of_channels = ring[0].number_of_channels();
    output.set_number_of_channels(number_of_channels);
    if (!in_cycle) in_cycle = sh->latest_frame_written != scope.current_frame;

    ParamSlice delay = params.get(delay_time);
    double sample_rate = (double)scope.sample_rate;
    double dt = 1. / sample_rate;
    double quantum_duration = (double)RQ * dt;
    int ring_size = (int)ring.size();
    int ring_index = (int)index;
    PlaybackInfo infos[RQ];
    if (delay.len == 1) {
        infos[0] = get_playback_infos((double)delay[0], in_cycle, 0., quantum_duration, sample_rate, ring_size, ring_index);
        for (int i = 1; i < RQ; i++) {
            PlaybackInfo p = infos[i - 1];
            size_t prev_block_index = p.prev_block_index;
            size_t prev_frame_index = p.prev_frame_index + 1;
            if (prev_frame_index >= (size_t)RQ) {
                prev_block_index = (prev_block_index + 1) % ring.size();
                prev_frame_index = 0;
            }
            infos[i].prev_block_index = prev_block_index;
            infos[i].prev_frame_index = prev_frame_index;
            infos[i].k = p.k;
        }
    } else {
        for (int i = 0; i < RQ; i++)
            infos[i] = get_playback_infos((double)delay[i], in_cycle, (double)i, quantum_duration, sample_rate, ring_size, ring_index);
    }
    bool is_actively_processing = false;
    for (int c = 0; c < number_of_channels; c++) {
        float* o = output.channel_mut(c).make_mut();
        for (int i = 0; i < RQ; i++) {
            const PlaybackInfo& p = infos[i];
            size_t next_block_index = p.prev_block_index;
            size_t next_frame_index = p.prev_frame_index + 1;
            if (next_frame_index >= (size_t)RQ) {
                next_block_index = (next_block_index + 1) % ring.size();
                next_frame_index = 0;
            }
            float prev_sample = ring[p.prev_block_index].channel(c).data()[p.prev_frame_index];
            float next_sample = ring[next_block_index].channel(c).data()[next_frame_index];
            float value = std::fma(1.f - p.k, prev_sample, p.k * next_sample);
            if (is_normal(value)) is_actively_processing = true;
            o[i] = value;
        }
    }
    if (!is_actively_processing) output.make_silent();
    index = (index + 1) % sh->capacity;
    return true;
}

// ================================================================================================
// StereoPanner — src/node/stereo_panner.rs
// ================================================================================================

static const float PI32 = 3.14159265358979323846f;
// stereo_panner.rs:74-79
static inline void get_stereo_gains(float x, float& gl, float& gr) {
    gl = sinf((1.f - x) * PI32 / 2.f);
    gr = sinf(x * PI32 / 2.f);
}

// stereo_panner.rs:218-318
bool StereoPannerRenderer::process(std::vector<Quantum>& inputs, std::vector<Quantum>& outputs, const ParamValues& params, const Scope&) {
    const Quantum& input = inputs[0];
    Quantum& output = outputs[0];
    if (input.is_silent()) {
        output.make_silent();
        return false;
    }
    output.set_number_of_channels(2);
    ParamSlice pan_values = params.get(pan);
    Channel in0 = input.channel(0);
    Channel in1 = input.number_of_channels() > 1 ? input.channel(1) : input.channel(0);
    float* left = output.channel_mut(0).make_mut();
    float* right = output.channel_mut(1).make_mut();
    if (input.number_of_channels() == 1) {
        if (pan_values.len == 1) {
            float p = pan_values[0];
            float x = (p + 1.f) * 0.5f;
            float gl, gr;
            get_stereo_gains(x, gl, gr);
            for (int i = 0; i < RQ; i++) {
                left[i] = in0.data()[i] * gl;
                right[i] = in0.data()[i] * gr;
            }
        } else {
            for (int i = 0; i < RQ; i++) {
                float x = (pan_values[i] + 1.f) * 0.5f;
                float gl, gr;
                get_stereo_gains(x, gl, gr);
                left[i] = in0.data()[i] * gl;
                right[i] = in0.data()[i] * gr;
            }
        }
    } else if (input.number_of_channels() == 2) {
        if (pan_values.len == 1) {
            float p = pan_values[0];
            float x = p <= 0.f ? p + 1.f : p;
            float gl, gr;
            get_stereo_gains(x, gl, gr);
            for (int i = 0; i < RQ; i++) {
                float il = in0.data()[i], ir = in1.data()[i];
                if (p <= 0.f) {
                    left[i] = std::fma(ir, gl, il);
                    right[i] = ir * gr;
                } else {
                    left[i] = il * gl;
                    right[i] = std::fma(il, gr, ir);
                }
            }
        } else {
            for (int i = 0; i < RQ; i++) {
                float p = pan_values[i];
                float il = in0.data()[i], ir = in1.data()[i];
                float gl, gr;
                if (p <= 0.f) {
                    get_stereo_gains(p + 1.f, gl, gr);
                    left[i] = std::fma(ir, gl, il);
                    right[i] = ir * gr;
                } else {
                    get_stereo_gains(p, gl, gr);
                    left[i] = il * gl;
                    right[i] = std::fma(il, gr, ir);
                }
            }
        }
    }
    return false;
}

// ================================================================================================
// DynamicsCompressor — src/node/dynamics_compressor.rs
// ================================================================================================

// dynamics_compressor.rs:13-27
static inline float db_to_lin(float v) { return powf(10.0f, v / 20.f); }
static inline float lin_to_db(float v) { return v == 0.f ? -1000.f : 20.f * log10f(v); }
// test hooks for dynamics_compressor.rs:565-581 (test_db_to_lin / test_lin_to_db)
float compressor_db_to_lin(float v) { return db_to_lin(v); }
float compressor_lin_to_db(float v) { return lin_to_db(v); }

// dynamics_compressor.rs:330-478
bool DynamicsCompressorRenderer::process(std::vector<Quantum>& inputs, std::vector<Quantum>& outputs, const ParamValues& params, const Scope& scope) {
    Quantum input = inputs[0];
    Quantum& output = outputs[0];
    float sample_rate = scope.sample_rate;
    size_t ring_size = ring_capacity;
    if (ring.size() < ring_size) {
        Quantum silence = input;
        silence.make_silent();
        ring.resize(ring_size, silence);
    }
    float threshold_v = params.get(threshold)[0];
    float knee_v = params.get(knee)[0];
    float ratio_v = params.get(ratio)[0];
    float thr = knee_v > 0.f ? threshold_v + knee_v / 2.f : threshold_v;
    float half_knee = knee_v / 2.f;
    float knee_partial = (1.f / ratio_v - 1.f) / (2.f * knee_v);
    float attack_v = params.get(attack)[0];
    float release_v = params.get(release)[0];
    float attack_tau = expf(-1.f / (attack_v * sample_rate));
    float release_tau = expf(-1.f / (release_v * sample_rate));
    float full_range_gain = thr + (-thr / ratio_v);
    float full_range_makeup = 1.f / db_to_lin(full_range_gain);
    float makeup_gain = lin_to_db(powf(full_range_makeup, 0.6f));

    float prev = prev_detector_value;
    float reduction_gain = 0.f;
    float reduction_gains[RQ];
    for (int i = 0; i < RQ; i++) {
        float mx = -3.40282347e+38f;  // f32::MIN
        for (int c = 0; c < input.number_of_channels(); c++) {
            float s = std::fabs(input.channel(c).data()[i]);
            if (s > mx) mx = s;
        }
        float sample_db = lin_to_db(mx);
        float sample_attenuated;
        if (sample_db <= thr - half_knee) {
            sample_attenuated = sample_db;
        } else if (sample_db <= thr + half_knee) {
            float t = sample_db - thr + half_knee;
            sample_attenuated = sample_db + (t * t) * knee_partial;
        } else {
            sample_attenuated = thr + (sample_db - thr) / ratio_v;
        }
        float sample_attenuation = sample_db - sample_attenuated;
        float detector_value;
        if (sample_attenuation > prev)
            detector_value = attack_tau * prev + (1.f - attack_tau) * sample_attenuation;
        else
            detector_value = release_tau * prev + (1.f - release_tau) * sample_attenuation;
        reduction_gain = -detector_value + makeup_gain;
        reduction_gains[i] = db_to_lin(reduction_gain);
        prev = detector_value;
    }
    prev_detector_value = prev;
    reduction = reduction_gain;
    ring[ring_index] = input;
    size_t read_index = (ring_index + 1) % ring_size;
    ring_index = read_index;
    output = ring[read_index];
    if (output.is_silent()) {
        output.make_silent();
        return false;
    }
    for (int c = 0; c < output.number_of_channels(); c++) {
        float* o = output.channel_mut(c).make_mut();
        for (int i = 0; i < RQ; i++) o[i] *= reduction_gains[i];
    }
    return true;
}

// ================================================================================================
// Analyser — src/analysis.rs, src/node/analyser.rs
// ================================================================================================

// analysis.rs:13-24
std::vector<float> generate_blackman(size_t size) {
    float alpha = 0.16f;
    float a0 = (1.f - alpha) / 2.f;
    float a1 = 1.f / 2.f;
    float a2 = alpha / 2.f;
    std::vector<float> w(size);
    for (size_t i = 0; i < size; i++)
        w[i] = a0 - a1 * cosf(2.f * PI32 * (float)i / (float)size) + a2 * cosf(4.f * PI32 * (float)i / (float)size);
    return w;
}

void Analyser::write(const float* src, size_t len) {
    for (size_t i = 0; i < len; i++) ring[(write_index + i) % ANALYSER_RING] = src[i];
    write_index += len;
    if (write_index >= ANALYSER_RING) write_index -= ANALYSER_RING;
}

void Analyser::read(float* dst, size_t dst_len, size_t max_len) const {
    size_t len = std::min(dst_len, max_len);
    for (size_t i = 0; i < len; i++) dst[i] = ring[(ANALYSER_RING + write_index - len + i) % ANALYSER_RING];
}

void Analyser::set_fft_size(size_t n) {
    if (n != fft_size) {
        std::fill(last_fft_output.begin(), last_fft_output.end(), 0.f);
        blackman = generate_blackman(n);
        fft_size = n;
    }
}

void Analyser::compute_fft() {
    if (blackman.size() != fft_size) blackman = generate_blackman(fft_size);
    float smoothing = (float)smoothing_time_constant;
    std::vector<float> input(fft_size, 0.f);
    read(input.data(), fft_size, fft_size);
    for (size_t i = 0; i < fft_size; i++) input[i] *= blackman[i];
    RealFFT fft;
    fft.init((int)fft_size);
    std::vector<cf32> out(fft_size / 2 + 1);
    fft.forward(input.data(), out.data());
    float normalize_factor = 1.f / (float)fft_size;
    for (size_t k = 0; k < fft_size / 2; k++) {
        float norm = std::hypot(out[k].real(), out[k].imag()) * normalize_factor;
        float value = smoothing * last_fft_output[k] + (1.f - smoothing) * norm;
        last_fft_output[k] = std::isfinite(value) ? value : 0.f;
    }
}

void Analyser::get_float_time_domain_data(float* dst, size_t n) const { read(dst, n, fft_size); }

void Analyser::get_byte_time_domain_data(uint8_t* dst, size_t n) const {
    std::vector<float> tmp(n, 0.f);
    read(tmp.data(), n, fft_size);
    for (size_t i = 0; i < n; i++) {
        float scaled = 128.f * (1.f + tmp[i]);
        float clamped = scaled < 0.f ? 0.f : (scaled > 255.f ? 255.f : scaled);
        dst[i] = (uint8_t)clamped;
    }
}

void Analyser::get_float_frequency_data(float* dst, size_t n, double current_time) {
    if (current_time != last_fft_time) {
        compute_fft();
        last_fft_time = current_time;
    }
    size_t len = std::min(n, fft_size / 2);
    for (size_t i = 0; i < len; i++) dst[i] = 20.f * log10f(last_fft_output[i]);
}

void Analyser::get_byte_frequency_data(uint8_t* dst, size_t n, double current_time) {
    if (current_time != last_fft_time) {
        compute_fft();
        last_fft_time = current_time;
    }
    float mn = (float)min_decibels, mx = (float)max_decibels;
    size_t len = std::min(n, fft_size / 2);
    for (size_t i = 0; i < len; i++) {
        float db = 20.f * log10f(last_fft_output[i]);
        float scaled = 255.f / (mx - mn) * (db - mn);
        float clamped = scaled < 0.f ? 0.f : (scaled > 255.f ? 255.f : scaled);
        if (std::isnan(scaled)) clamped = 0.f;  // Rust: NaN as u8 == 0
        dst[i] = (uint8_t)clamped;
    }
}

// analyser.rs:267-294
bool AnalyserRenderer::process(std::vector<Quantum>& inputs, std::vector<Quantum>& outputs, const ParamValues&, const Scope&) {
    const Quantum& input = inputs[0];
    outputs[0] = input;
    Quantum mono = input;
    mono.mix(1, SPEAKERS);
    analyser->write(mono.channel(0).data(), RQ);
    return false;
}

// ================================================================================================
// ChannelMerger / ChannelSplitter
// ================================================================================================

// channel_merger.rs:146-171
bool ChannelMergerRenderer::process(std::vector<Quantum>& inputs, std::vector<Quantum>& outputs, const ParamValues&, const Scope&) {
    Quantum& output = outputs[0];
    bool any = false;
    for (auto& in : inputs)
        if (!in.is_silent()) any = true;
    if (any) {
        output.set_number_of_channels((int)inputs.size());
        for (size_t i = 0; i < inputs.size(); i++) output.channel_mut((int)i) = inputs[i].channel(0);
    } else {
        output.make_silent();
    }
    return false;
}

// channel_splitter.rs:183-208
bool ChannelSplitterRenderer::process(std::vector<Quantum>& inputs, std::vector<Quantum>& outputs, const ParamValues&, const Scope&) {
    const Quantum& input = inputs[0];
    for (size_t i = 0; i < outputs.size(); i++) {
        outputs[i].set_number_of_channels(1);
        if ((int)i < input.number_of_channels())
            outputs[i].channel_mut(0) = input.channel((int)i);
        else
            outputs[i].make_silent();
    }
    return false;
}

}  // namespace wao
