// ORACLE — TEST INFRASTRUCTURE ONLY (see wao_core.h).
// Renderer (processor) halves of the nodes on the hot path, SURVEY §8(a) rows a6..a18.
#pragma once
#include "wao_core.h"
#include "wao_fft.h"

namespace wao {

struct AudioBuffer {
    std::vector<std::vector<float>> channels;
    float sample_rate = 0.f;
    int number_of_channels() const { return (int)channels.size(); }
    size_t length() const { return channels.empty() ? 0 : channels[0].size(); }
    double duration() const { return (double)length() / (double)sample_rate; }  // src/buffer.rs:138-140
};

// ---- OscillatorRenderer, src/node/oscillator.rs:340-676 ---------------------------------------------
struct OscillatorRenderer : Processor {
    int type = 0;
    uint32_t frequency = 0, detune = 0;
    double phase = 0.;
    double start_time = 1.7976931348623157e308, stop_time = 1.7976931348623157e308;
    bool started = false;
    std::vector<float> periodic_wave;
    const float* sine_table = nullptr;
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override;
    const char* name() const override { return "OscillatorRenderer"; }

  private:
    double generate_sample(float* out, bool outside_nyquist, double phase_incr, double current_time, double dt);
    float generate_waveform_sample(double phase_incr);
};
const float* precomputed_sine_table();  // oscillator.rs:19-28

// ---- BiquadFilterRenderer, src/node/biquad_filter.rs:740-911 -----------------------------------------
struct BiquadCoefs {
    double b0, b1, b2, a1, a2;
};
BiquadCoefs biquad_calculate_coefs(int type, double sample_rate, double f0, double gain, double q);  // :367-390
float biquad_computed_freq(float freq, float detune);                                                // :393-399
// control-side helper used by the known-answer tests: BiquadFilterNode::get_frequency_response (:663-737)
void biquad_frequency_response(int type, float sample_rate, float frequency, float detune, float q, float gain,
                               const float* freq_hz, float* mag, float* phase, int n);
struct BiquadFilterRenderer : Processor {
    uint32_t q = 0, detune = 0, frequency = 0, gain = 0;
    int type = 0;
    std::vector<std::array<double, 4>> xy;
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override;
    const char* name() const override { return "BiquadFilterRenderer"; }
};

// ---- IirFilterRenderer, src/node/iir_filter.rs:269-414 -------------------------------------------------
struct IirFilterRenderer : Processor {
    std::vector<std::pair<double, double>> norm_coeffs;  // (b[n], a[n]) / a0
    std::vector<std::array<double, 20>> states;
    IirFilterRenderer(std::vector<double> feedforward, std::vector<double> feedback);
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override;
    const char* name() const override { return "IirFilterRenderer"; }
};
void iir_frequency_response(const std::vector<double>& ff, const std::vector<double>& fb, float sample_rate,
                            const float* freq_hz, float* mag, float* phase, int n);

// ---- GainRenderer, src/node/gain.rs:126-199 --------------------------------------------------------------
struct GainRenderer : Processor {
    uint32_t gain = 0;
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override;
    const char* name() const override { return "GainRenderer"; }
};

// ---- DestinationRenderer, src/node/destination.rs:143-163 --------------------------------------------------
struct DestinationRenderer : Processor {
    bool process(std::vector<Quantum>& in, std::vector<Quantum>& out, const ParamValues&, const Scope&) override {
        out[0] = in[0];
        return true;
    }
    bool has_side_effects() const override { return true; }
    const char* name() const override { return "DestinationRenderer"; }
};

// ---- ConstantSourceRenderer, src/node/constant_source.rs:176-262 ---------------------------------------------
struct ConstantSourceRenderer : Processor {
    uint32_t offset = 0;
    double start_time = 1.7976931348623157e308, stop_time = 1.7976931348623157e308;
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override;
    const char* name() const override { return "ConstantSourceRenderer"; }
};

// ---- AudioBufferSourceRenderer, src/node/audio_buffer_source.rs:351-866 -----------------------------------------
struct AudioBufferSourceRenderer : Processor {
    double start_time = 1.7976931348623157e308, stop_time = 1.7976931348623157e308;
    double offset = 0., duration = 1.7976931348623157e308;
    std::shared_ptr<AudioBuffer> buffer;
    uint32_t detune = 0, playback_rate = 0;
    bool is_looping = false;
    double loop_start = 0., loop_end = 0.;
    // AudioBufferRendererState (:351-372)
    double buffer_time = 0.;
    bool started = false, entered_loop = false;
    double buffer_time_elapsed = 0.;
    bool is_aligned = false, ended = false;
    void clamp_loop_boundaries();
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override;
    const char* name() const override { return "AudioBufferSourceRenderer"; }
};

// ---- fft-convolver 0.3 (external crate, restated) + ConvolverRenderer, src/node/convolver.rs ----------------------
struct FFTConvolver {
    size_t ir_len = 0, block_size = 0, seg_size = 0, seg_count = 0, fft_complex_size = 0;
    std::vector<std::vector<cf32>> segments, segments_ir;
    std::vector<float> fft_buffer, overlap, input_buffer;
    std::vector<cf32> pre_multiplied, conv;
    size_t current = 0, input_buffer_fill = 0;
    RealFFT fft;
    void init(size_t block_size, const float* ir, size_t ir_len);
    void process(const float* input, float* output, size_t len);
};
float convolver_normalize_buffer(const AudioBuffer& buffer);  // convolver.rs:16-53
struct ConvolverRenderer : Processor {
    bool has_convolvers = false;
    std::vector<FFTConvolver> convolvers;
    size_t impulse_length = 0;
    int impulse_number_of_channels = 0;
    size_t tail_count = 0;
    void set_buffer(const AudioBuffer& buffer, bool normalize);  // ConvolverNode::set_buffer, :259-317
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override;
    const char* name() const override { return "ConvolverRenderer"; }
};

}  // namespace wao
