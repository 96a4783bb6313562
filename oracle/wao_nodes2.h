// ORACLE — TEST INFRASTRUCTURE ONLY (see wao_core.h).
// WaveShaper, Delay (writer/reader), StereoPanner, DynamicsCompressor, Analyser, ChannelMerger/Splitter.
#pragma once
#include "wao_core.h"
#include "wao_resampler.h"
#include "wao_fft.h"

namespace wao {

// ---- WaveShaperRenderer, src/node/waveshaper.rs:358-572 (OverSampleType::None only; X2/X4 go through
// the un-vendored `rubato` crate: parity unpinned, not restated) ------------------------------------------
struct WaveShaperRenderer : Processor {
    bool has_curve = false;
    std::vector<float> curve;
    bool can_propagate_silence = true;
    int oversample = 0;      // 0 none, 1 X2, 2 X4 (waveshaper.rs:18-36)
    size_t sample_rate = 0;
    size_t os_channels = 0;  // channels the up / down samplers were built for (:411-424)
    FftFixedInOut upsampler, downsampler;
    void set_curve(const float* c, size_t n);  // onmessage, :480-503
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override;
    const char* name() const override { return "WaveShaperRenderer"; }
};
float compressor_db_to_lin(float v);  // dynamics_compressor.rs:13-19
float compressor_lin_to_db(float v);  // :21-27
float waveshaper_apply_curve(const std::vector<float>& curve, float input);  // :555-572

// ---- DelayWriter / DelayReader, src/node/delay.rs:376-743 --------------------------------------------------
struct DelayShared {
    std::vector<Quantum> ring;  // len == capacity once initialised
    size_t capacity = 0;
    uint64_t latest_frame_written = UINT64_MAX;
};
struct DelayWriter : Processor {
    std::shared_ptr<DelayShared> sh;
    size_t index = 0;
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override;
    bool has_side_effects() const override { return true; }
    const char* name() const override { return "DelayWriter"; }
};
struct DelayReader : Processor {
    std::shared_ptr<DelayShared> sh;
    uint32_t delay_time = 0;
    size_t index = 0;
    bool in_cycle = false;
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override;
    const char* name() const override { return "DelayReader"; }
};

// ---- StereoPannerRenderer, src/node/stereo_panner.rs:74-318 ------------------------------------------------
struct StereoPannerRenderer : Processor {
    uint32_t pan = 0;
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override;
    const char* name() const override { return "StereoPannerRenderer"; }
};

// ---- DynamicsCompressorRenderer, src/node/dynamics_compressor.rs:13-27,306-478 --------------------------------
struct DynamicsCompressorRenderer : Processor {
    uint32_t attack = 0, knee = 0, ratio = 0, release = 0, threshold = 0;
    float reduction = 0.f;
    std::vector<Quantum> ring;
    size_t ring_capacity = 0, ring_index = 0;
    float prev_detector_value = 0.f;
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override;
    const char* name() const override { return "DynamicsCompressorRenderer"; }
};

// ---- Analyser, src/node/analyser.rs:262-294 + src/analysis.rs ------------------------------------------------
constexpr size_t ANALYSER_RING = 32768 + 128;  // analysis.rs:74
struct Analyser {
    std::vector<float> ring = std::vector<float>(ANALYSER_RING, 0.f);
    size_t write_index = 0;
    size_t fft_size = 2048;
    double smoothing_time_constant = 0.8, min_decibels = -100., max_decibels = -30.;
    std::vector<float> last_fft_output = std::vector<float>(32768 / 2 + 1, 0.f);
    std::vector<float> blackman;
    double last_fft_time = -INFINITY;
    void write(const float* src, size_t len);                                   // analysis.rs:96-112
    void read(float* dst, size_t dst_len, size_t max_len) const;                // analysis.rs:114-127
    void set_fft_size(size_t n);                                                // analysis.rs:216-231
    void compute_fft();                                                         // analysis.rs:278-345
    void get_float_time_domain_data(float* dst, size_t n) const;                // analysis.rs:261-264
    void get_byte_time_domain_data(uint8_t* dst, size_t n) const;               // analysis.rs:266-276
    void get_float_frequency_data(float* dst, size_t n, double current_time);   // analysis.rs:347-369
    void get_byte_frequency_data(uint8_t* dst, size_t n, double current_time);  // analysis.rs:371-401
};
std::vector<float> generate_blackman(size_t size);  // analysis.rs:13-24
struct AnalyserRenderer : Processor {
    std::shared_ptr<Analyser> analyser;
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override;
    const char* name() const override { return "AnalyserRenderer"; }
};

// ---- ChannelMerger / ChannelSplitter, src/node/channel_merger.rs:146-171, channel_splitter.rs:183-208 ----------
struct ChannelMergerRenderer : Processor {
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override;
    const char* name() const override { return "ChannelMergerRenderer"; }
};
struct ChannelSplitterRenderer : Processor {
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override;
    const char* name() const override { return "ChannelSplitterRenderer"; }
};

}  // namespace wao
