// ORACLE — TEST INFRASTRUCTURE ONLY (see wao_core.h).
//
// C entry points of the oracle: the same graph-building surface as include/wae.h with the prefix wao_
// (control half of the reference: src/context/{base,concrete_base,offline}.rs + XxxNode::new of src/node/*.rs),
// plus wao_render (OfflineAudioContext::start_rendering_sync, src/context/offline.rs:157-185 ->
// src/render/thread.rs:260-302,355-396).
#include "../include/wae.h"
#include "wao_core.h"
#include "wao_param.h"
#include "wao_nodes.h"
#include "wao_nodes2.h"
#include "wao_panner.h"

#include <xmmintrin.h>
#include <pmmintrin.h>
#include <thread>
#include <atomic>
#include <chrono>

using namespace wao;

namespace {

thread_local std::string g_err;
int32_t fail(int32_t code, const std::string& msg) {
    g_err = msg;
    return code;
}

enum NodeKind {
    K_DEST, K_PARAM, K_OSC, K_BIQUAD, K_IIR, K_GAIN, K_ABSN, K_CONST, K_CONV, K_SHAPER, K_DELAY, K_SPANNER,
    K_PANNER, K_ANALYSER, K_COMP, K_MERGER, K_SPLITTER, K_LISTENER
};

struct NodeInfo {
    NodeKind kind;
    uint32_t out_id;  // id that carries the outputs (DelayNode: reader id)
    int n_inputs, n_outputs;
    std::vector<uint32_t> params;
    bool has_start = false;
    Processor* proc = nullptr;
    bool normalize = true;   // ConvolverNode::normalize (control side; used by the next set_buffer)
    bool has_buffer = false;  // AudioBufferSourceNode: a buffer was given
};

}  // namespace

struct wae_graph {  // the oracle's OfflineAudioContext
    Graph graph;
    uint32_t next_id = 11;  // ids 0..=10 reserved, src/context/mod.rs:24-40
    uint32_t channels;
    uint64_t length;
    float sample_rate;
    std::map<uint32_t, NodeInfo> info;
    bool listener_present = false;
    uint32_t listener_params[9] = {2, 3, 4, 5, 6, 7, 8, 9, 10};
    std::vector<std::pair<uint32_t, std::shared_ptr<Analyser>>> analysers;
    bool rendered = false;
    uint64_t frames_played = 0;
    std::vector<float> partial;           // [channels][length] PCM rendered before a suspend point
    std::vector<uint64_t> suspend_quanta;  // OfflineAudioContext::suspend_sync points already taken

    // BaseAudioContext::create_audio_param, src/context/base.rs:320-337: the param is its own graph node
    uint32_t create_param(uint32_t owner, const ParamDescriptor& d, float initial, bool fixed_id = false, uint32_t id = 0, bool send_set_value = true) {
        uint32_t pid = fixed_id ? id : next_id++;
        auto p = std::make_unique<ParamProcessor>(d);
        ParamProcessor* raw = p.get();
        ChannelConfig cfg{1, MODE_EXPLICIT, DISCRETE};  // src/param.rs:296-310
        graph.add_node(pid, std::move(p), 1, 1, cfg);
        NodeInfo ni{K_PARAM, pid, 1, 1, {}, false, raw};
        info[pid] = ni;
        // param.set_value(v) -> SetValue event (src/param.rs:403-426)
        if (send_set_value) {
            ParamEvent ev;
            ev.type = EV_SET_VALUE;
            ev.value = initial;
            ev.time = 0.;
            raw->handle_incoming_event(ev);
        }
        pending_param_edges.push_back({pid, owner});
        return pid;
    }
    std::vector<std::pair<uint32_t, uint32_t>> pending_param_edges;
    // ConcreteBaseAudioContext::register tail: RegisterNode, then the queued param->node ConnectNode messages
    void finish_register(uint32_t id, std::unique_ptr<Processor> proc, NodeKind kind, int n_in, int n_out, ChannelConfig cfg,
                         std::vector<uint32_t> params, uint32_t out_id) {
        Processor* raw = proc.get();
        graph.add_node(id, std::move(proc), n_in, n_out, cfg);
        NodeInfo ni{kind, out_id, n_in, n_out, params, false, raw};
        info[id] = ni;
        for (auto& e : pending_param_edges)
            if (e.second == id) graph.add_edge(e.first, 0, e.second, -1);
        pending_param_edges.erase(std::remove_if(pending_param_edges.begin(), pending_param_edges.end(),
                                                 [&](auto& e) { return e.second == id; }),
                                  pending_param_edges.end());
    }
    void ensure_listener();
};

static ChannelConfig resolve_cfg(const wae_channel_config& c, ChannelConfig def) {
    if (c.count == 0) return def;
    ChannelConfig r;
    r.count = (int)c.count;
    r.mode = (int)c.count_mode;
    r.interp = (int)c.interpretation;
    return r;
}

static const float F32_MAX = 3.40282347e+38f;

// AudioListener, src/spatial.rs:60-200 + ConcreteBaseAudioContext::ensure_audio_listener_present (:516-534)
void wae_graph::ensure_listener() {
    if (listener_present) return;
    listener_present = true;
    static const float defaults[9] = {0.f, 0.f, 0.f, 0.f, 0.f, -1.f, 0.f, 1.f, 0.f};
    for (int i = 0; i < 9; i++) {
        ParamDescriptor d{defaults[i], -F32_MAX, F32_MAX, true};
        create_param(1, d, defaults[i], true, 2 + (uint32_t)i, false);  // no set_value: spatial.rs:135-143
    }
    auto l = std::make_unique<ListenerRenderer>();
    for (int i = 0; i < 9; i++) l->params[i] = 2 + (uint32_t)i;
    ChannelConfig cfg;  // default
    std::vector<uint32_t> ps(listener_params, listener_params + 9);
    finish_register(1, std::move(l), K_LISTENER, 0, 9, ChannelConfig{1, MODE_EXPLICIT, DISCRETE}, ps, 1);  // spatial.rs:86-106
    graph.add_edge(1, 0, 0, -1);
}

extern "C" {

#define WAO_API __attribute__((visibility("default")))

WAO_API const char* wao_last_error(void) { return g_err.c_str(); }

WAO_API wae_status wao_graph_create(uint32_t number_of_channels, uint64_t length, float sample_rate, wae_graph** out) {
    // OfflineAudioContext::new asserts, src/context/offline.rs:78-84 + src/lib.rs asserts
    if (number_of_channels < 1 || number_of_channels > 32)
        return fail(WAE_NOT_SUPPORTED, "NotSupportedError - Invalid number of channels");
    if (length == 0) return fail(WAE_NOT_SUPPORTED, "NotSupportedError - Invalid length: 0");
    if (!(sample_rate >= 3000.f && sample_rate <= 768000.f))  // assert_valid_sample_rate, src/lib.rs
        return fail(WAE_NOT_SUPPORTED, "NotSupportedError - Invalid sample rate");
    auto* g = new wae_graph;
    g->channels = number_of_channels;
    g->length = length;
    g->sample_rate = sample_rate;
    // AudioDestinationNode::new, src/node/destination.rs:100-117
    ChannelConfig cfg{(int)number_of_channels, MODE_EXPLICIT, SPEAKERS};
    g->finish_register(0, std::make_unique<DestinationRenderer>(), K_DEST, 1, 1, cfg, {}, 0);
    *out = g;
    return WAE_OK;
}

WAO_API wae_status wao_graph_destroy(wae_graph* g) {
    delete g;
    return WAE_OK;
}

// OscillatorNode::new, src/node/oscillator.rs:211-275
WAO_API wae_status wao_create_oscillator(wae_graph* g, const wae_oscillator_options* o, wae_node_id* out) {
    uint32_t id = g->next_id++;
    float nyquist = g->sample_rate / 2.f;
    uint32_t f = g->create_param(id, ParamDescriptor{440.f, -nyquist, nyquist, true}, o->frequency);
    uint32_t d = g->create_param(id, ParamDescriptor{0.f, -153600.f, 153600.f, true}, o->detune);
    auto r = std::make_unique<OscillatorRenderer>();
    r->type = (int)o->type;
    r->frequency = f;
    r->detune = d;
    r->sine_table = precomputed_sine_table();
    if (o->type == WAE_OSC_CUSTOM) {
        if (!o->periodic_wave || o->periodic_wave_len == 0) return fail(WAE_INVALID_ARGUMENT, "custom oscillator needs a periodic wave table");
        r->periodic_wave.assign(o->periodic_wave, o->periodic_wave + o->periodic_wave_len);
    }
    ChannelConfig cfg;
    g->finish_register(id, std::move(r), K_OSC, 0, 1, cfg, {f, d}, id);
    *out = id;
    return WAE_OK;
}

// BiquadFilterNode::new, src/node/biquad_filter.rs:542-608
WAO_API wae_status wao_create_biquad_filter(wae_graph* g, const wae_biquad_options* o, wae_node_id* out) {
    if (o->type > 7) return fail(WAE_INVALID_ARGUMENT, "invalid biquad type");
    uint32_t id = g->next_id++;
    uint32_t q = g->create_param(id, ParamDescriptor{1.f, -F32_MAX, F32_MAX, true}, o->q);
    uint32_t d = g->create_param(id, ParamDescriptor{0.f, -153600.f, 153600.f, true}, o->detune);
    uint32_t f = g->create_param(id, ParamDescriptor{350.f, 0.f, g->sample_rate / 2.f, true}, o->frequency);
    uint32_t ga = g->create_param(id, ParamDescriptor{0.f, -F32_MAX, 40.f * log10f(F32_MAX), true}, o->gain);
    auto r = std::make_unique<BiquadFilterRenderer>();
    r->q = q;
    r->detune = d;
    r->frequency = f;
    r->gain = ga;
    r->type = (int)o->type;
    g->finish_register(id, std::move(r), K_BIQUAD, 1, 1, resolve_cfg(o->channel_config, ChannelConfig()), {q, d, f, ga}, id);
    *out = id;
    return WAE_OK;
}

// IIRFilterNode::new, src/node/iir_filter.rs:146-205 (asserts)
WAO_API wae_status wao_create_iir_filter(wae_graph* g, const wae_iir_options* o, wae_node_id* out) {
    if (o->feedforward_len == 0 || o->feedforward_len > 20) return fail(WAE_NOT_SUPPORTED, "NotSupportedError - invalid feedforward length");
    if (o->feedback_len == 0 || o->feedback_len > 20) return fail(WAE_NOT_SUPPORTED, "NotSupportedError - invalid feedback length");
    bool all_zero = true;
    for (uint32_t i = 0; i < o->feedforward_len; i++)
        if (o->feedforward[i] != 0.) all_zero = false;
    if (all_zero) return fail(WAE_INVALID_STATE, "InvalidStateError - all feedforward coefficients are zero");
    if (o->feedback[0] == 0.) return fail(WAE_INVALID_STATE, "InvalidStateError - first feedback coefficient is zero");
    uint32_t id = g->next_id++;
    auto r = std::make_unique<IirFilterRenderer>(std::vector<double>(o->feedforward, o->feedforward + o->feedforward_len),
                                                 std::vector<double>(o->feedback, o->feedback + o->feedback_len));
    g->finish_register(id, std::move(r), K_IIR, 1, 1, resolve_cfg(o->channel_config, ChannelConfig()), {}, id);
    *out = id;
    return WAE_OK;
}

// GainNode::new, src/node/gain.rs:86-117
WAO_API wae_status wao_create_gain(wae_graph* g, const wae_gain_options* o, wae_node_id* out) {
    uint32_t id = g->next_id++;
    uint32_t p = g->create_param(id, ParamDescriptor{1.f, -F32_MAX, F32_MAX, true}, o->gain);
    auto r = std::make_unique<GainRenderer>();
    r->gain = p;
    g->finish_register(id, std::move(r), K_GAIN, 1, 1, resolve_cfg(o->channel_config, ChannelConfig()), {p}, id);
    *out = id;
    return WAE_OK;
}

// AudioBuffer::new / ::from (src/buffer.rs:96-131): assert_valid_number_of_channels, assert_valid_buffer_length
static bool valid_buffer(const wae_audio_buffer* b) {
    if (b->number_of_channels < 1 || b->number_of_channels > 32 || !b->channels) {
        fail(WAE_NOT_SUPPORTED, "NotSupportedError - Invalid number of channels: " + std::to_string(b->number_of_channels) + " is outside range [1, 32]");
        return false;
    }
    if (b->length == 0) {
        fail(WAE_NOT_SUPPORTED, "NotSupportedError - Invalid length: 0 is less than or equal to minimum bound (0)");
        return false;
    }
    return true;
}
static std::shared_ptr<AudioBuffer> copy_buffer(const wae_audio_buffer* b) {
    auto ab = std::make_shared<AudioBuffer>();
    ab->sample_rate = b->sample_rate;
    for (uint32_t c = 0; c < b->number_of_channels; c++) ab->channels.emplace_back(b->channels[c], b->channels[c] + b->length);
    return ab;
}

// AudioBufferSourceNode::new, src/node/audio_buffer_source.rs:160-235
WAO_API wae_status wao_create_buffer_source(wae_graph* g, const wae_buffer_source_options* o, wae_node_id* out) {
    if (o->buffer && !valid_buffer(o->buffer)) return WAE_NOT_SUPPORTED;
    uint32_t id = g->next_id++;
    uint32_t d = g->create_param(id, ParamDescriptor{0.f, -F32_MAX, F32_MAX, false}, o->detune);
    uint32_t pr = g->create_param(id, ParamDescriptor{1.f, -F32_MAX, F32_MAX, false}, o->playback_rate);
    auto r = std::make_unique<AudioBufferSourceRenderer>();
    r->detune = d;
    r->playback_rate = pr;
    r->is_looping = o->loop != 0;
    r->loop_start = o->loop_start;
    r->loop_end = o->loop_end;
    if (o->buffer) {
        r->buffer = copy_buffer(o->buffer);
        r->clamp_loop_boundaries();
    }
    ChannelConfig cfg;
    g->finish_register(id, std::move(r), K_ABSN, 0, 1, cfg, {d, pr}, id);
    *out = id;
    return WAE_OK;
}

// ConstantSourceNode::new, src/node/constant_source.rs:138-170
WAO_API wae_status wao_create_constant_source(wae_graph* g, const wae_constant_source_options* o, wae_node_id* out) {
    uint32_t id = g->next_id++;
    uint32_t p = g->create_param(id, ParamDescriptor{1.f, -F32_MAX, F32_MAX, true}, o->offset);
    auto r = std::make_unique<ConstantSourceRenderer>();
    r->offset = p;
    ChannelConfig cfg;
    g->finish_register(id, std::move(r), K_CONST, 0, 1, cfg, {p}, id);
    *out = id;
    return WAE_OK;
}

// ConvolverNode::new + set_buffer, src/node/convolver.rs:199-317
WAO_API wae_status wao_create_convolver(wae_graph* g, const wae_convolver_options* o, wae_node_id* out) {
    ChannelConfig def{2, MODE_CLAMPED_MAX, SPEAKERS};
    ChannelConfig cfg = resolve_cfg(o->channel_config, def);
    if (cfg.count > 2) return fail(WAE_NOT_SUPPORTED, "NotSupportedError - ConvolverNode channel count cannot be greater than two");
    if (cfg.mode == MODE_MAX) return fail(WAE_NOT_SUPPORTED, "NotSupportedError - ConvolverNode channel count mode cannot be set to max");
    if (o->buffer) {
        if (o->buffer->sample_rate != g->sample_rate)
            return fail(WAE_NOT_SUPPORTED, "NotSupportedError - sample rate of the convolution buffer must match the audio context");
        uint32_t n = o->buffer->number_of_channels;
        if (!(n == 1 || n == 2 || n == 4))
            return fail(WAE_NOT_SUPPORTED, "NotSupportedError - the convolution buffer must consist of 1, 2 or 4 channels");
    }
    uint32_t id = g->next_id++;
    auto r = std::make_unique<ConvolverRenderer>();
    if (o->buffer) {
        auto ab = copy_buffer(o->buffer);
        r->set_buffer(*ab, !o->disable_normalization);
    }
    g->finish_register(id, std::move(r), K_CONV, 1, 1, cfg, {}, id);
    g->info[id].normalize = !o->disable_normalization;
    *out = id;
    return WAE_OK;
}

// WaveShaperNode::new, src/node/waveshaper.rs:190-260
WAO_API wae_status wao_create_wave_shaper(wae_graph* g, const wae_wave_shaper_options* o, wae_node_id* out) {
    if (o->oversample > WAE_OVERSAMPLE_X4) return fail(WAE_INVALID_ARGUMENT, "unknown oversample type");
    uint32_t id = g->next_id++;
    auto r = std::make_unique<WaveShaperRenderer>();
    if (o->curve) r->set_curve(o->curve, o->curve_len);
    r->oversample = (int)o->oversample;
    r->sample_rate = (size_t)g->sample_rate;  // `sample_rate as usize`, waveshaper.rs:148-150
    g->finish_register(id, std::move(r), K_SHAPER, 1, 1, resolve_cfg(o->channel_config, ChannelConfig()), {}, id);
    *out = id;
    return WAE_OK;
}

// DelayNode::new, src/node/delay.rs:283-368: writer = N, reader = N+1, delayTime = N+2
WAO_API wae_status wao_create_delay(wae_graph* g, const wae_delay_options* o, wae_node_id* out) {
    double max_delay_time = o->max_delay_time;
    if (!(max_delay_time > 0. && max_delay_time < 180.))
        return fail(WAE_NOT_SUPPORTED, "NotSupportedError - maxDelayTime MUST be greater than zero and less than three minutes");
    double sample_rate = (double)g->sample_rate;
    size_t num_quanta = (size_t)std::ceil(max_delay_time * sample_rate / (double)RQ);
    auto sh = std::make_shared<DelayShared>();
    sh->capacity = num_quanta + 1;
    uint32_t writer_id = g->next_id++;
    uint32_t reader_id = g->next_id++;
    uint32_t p = g->create_param(reader_id, ParamDescriptor{0.f, 0.f, (float)max_delay_time, true}, (float)o->delay_time);
    ChannelConfig cfg = resolve_cfg(o->channel_config, ChannelConfig());
    auto reader = std::make_unique<DelayReader>();
    reader->sh = sh;
    reader->delay_time = p;
    g->finish_register(reader_id, std::move(reader), K_DELAY, 1, 1, cfg, {p}, reader_id);
    auto writer = std::make_unique<DelayWriter>();
    writer->sh = sh;
    g->finish_register(writer_id, std::move(writer), K_DELAY, 1, 1, cfg, {p}, reader_id);
    g->graph.mark_cycle_breaker(writer_id);
    g->graph.add_edge(writer_id, 0, reader_id, 0);
    *out = writer_id;
    return WAE_OK;
}

// StereoPannerNode::new, src/node/stereo_panner.rs:163-200
WAO_API wae_status wao_create_stereo_panner(wae_graph* g, const wae_stereo_panner_options* o, wae_node_id* out) {
    ChannelConfig def{2, MODE_CLAMPED_MAX, SPEAKERS};
    ChannelConfig cfg = resolve_cfg(o->channel_config, def);
    if (cfg.mode == MODE_MAX) return fail(WAE_NOT_SUPPORTED, "NotSupportedError - StereoPannerNode channel count mode cannot be set to max");
    if (cfg.count > 2) return fail(WAE_NOT_SUPPORTED, "NotSupportedError - StereoPannerNode channel count cannot be greater than two");
    uint32_t id = g->next_id++;
    uint32_t p = g->create_param(id, ParamDescriptor{0.f, -1.f, 1.f, true}, o->pan);
    auto r = std::make_unique<StereoPannerRenderer>();
    r->pan = p;
    g->finish_register(id, std::move(r), K_SPANNER, 1, 1, cfg, {p}, id);
    *out = id;
    return WAE_OK;
}

// PannerNode::new, src/node/panner.rs:392-520
WAO_API wae_status wao_create_panner(wae_graph* g, const wae_panner_options* o, wae_node_id* out) {
    ChannelConfig def{2, MODE_CLAMPED_MAX, SPEAKERS};
    ChannelConfig cfg = resolve_cfg(o->channel_config, def);
    if (cfg.mode == MODE_MAX) return fail(WAE_NOT_SUPPORTED, "NotSupportedError - PannerNode channel count mode cannot be set to max");
    if (cfg.count > 2) return fail(WAE_NOT_SUPPORTED, "NotSupportedError - PannerNode channel count cannot be greater than two");
    if (o->ref_distance < 0.) return fail(WAE_INVALID_ARGUMENT, "RangeError - refDistance cannot be negative");
    if (o->max_distance <= 0.) return fail(WAE_INVALID_ARGUMENT, "RangeError - maxDistance must be strictly positive");
    if (o->rolloff_factor < 0.) return fail(WAE_INVALID_ARGUMENT, "RangeError - rolloffFactor cannot be negative");
    if (o->cone_outer_gain < 0. || o->cone_outer_gain > 1.) return fail(WAE_INVALID_STATE, "InvalidStateError - coneOuterGain must be in the range [0, 1]");
    if (o->panning_model == WAE_PANNING_HRTF) {
        std::string err;
        if (!hrtf_sphere_available(err)) return fail(WAE_UNSUPPORTED, err);
        if (!hrtf_state_new(g->sample_rate))
            return fail(WAE_UNSUPPORTED, "HRTF panning: the context sample rate differs from the HRIR sphere's (resampling the sphere is not restated)");
    }
    // the node id is taken first, then ensure_audio_listener_present (panner.rs:432), then the params
    uint32_t id = g->next_id++;
    g->ensure_listener();
    ParamDescriptor pd{0.f, -F32_MAX, F32_MAX, true};
    uint32_t px = g->create_param(id, pd, o->position_x);
    uint32_t py = g->create_param(id, pd, o->position_y);
    uint32_t pz = g->create_param(id, pd, o->position_z);
    ParamDescriptor ox{1.f, -F32_MAX, F32_MAX, true};
    uint32_t oxp = g->create_param(id, ox, o->orientation_x);
    uint32_t oyp = g->create_param(id, pd, o->orientation_y);
    uint32_t ozp = g->create_param(id, pd, o->orientation_z);
    auto r = std::make_unique<PannerRenderer>();
    r->position_x = px; r->position_y = py; r->position_z = pz;
    r->orientation_x = oxp; r->orientation_y = oyp; r->orientation_z = ozp;
    r->distance_model = (int)o->distance_model;
    r->ref_distance = o->ref_distance;
    r->max_distance = o->max_distance;
    r->rolloff_factor = o->rolloff_factor;
    r->cone_inner_angle = o->cone_inner_angle;
    r->cone_outer_angle = o->cone_outer_angle;
    r->cone_outer_gain = o->cone_outer_gain;
    if (o->panning_model == WAE_PANNING_HRTF) r->set_hrtf(g->sample_rate);
    g->finish_register(id, std::move(r), K_PANNER, 1, 1, cfg, {px, py, pz, oxp, oyp, ozp}, id);
    // context.base().connect_listener_to_panner(node.registration().id()): listener -> panner, hidden port
    g->graph.add_edge(1, 0, id, -1);
    *out = id;
    return WAE_OK;
}

// AnalyserNode::new, src/node/analyser.rs:130-175
WAO_API wae_status wao_create_analyser(wae_graph* g, const wae_analyser_options* o, wae_node_id* out) {
    uint32_t fft = o->fft_size ? o->fft_size : 2048;
    if ((fft & (fft - 1)) != 0) return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - Invalid fft size: not a power of two");
    if (fft < 32 || fft > 32768) return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - Invalid fft size: outside range [32, 32768]");
    double stc = o->fft_size ? o->smoothing_time_constant : 0.8;
    if (!(stc >= 0. && stc <= 1.)) return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - Invalid smoothing time constant");
    double mn = o->fft_size ? o->min_decibels : -100., mx = o->fft_size ? o->max_decibels : -30.;
    if (!(mn < mx)) return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - Invalid min decibels");
    uint32_t id = g->next_id++;
    auto a = std::make_shared<Analyser>();
    a->set_fft_size(fft);
    a->smoothing_time_constant = stc;
    a->min_decibels = mn;
    a->max_decibels = mx;
    auto r = std::make_unique<AnalyserRenderer>();
    r->analyser = a;
    g->analysers.emplace_back(id, a);
    g->finish_register(id, std::move(r), K_ANALYSER, 1, 1, resolve_cfg(o->channel_config, ChannelConfig()), {}, id);
    *out = id;
    return WAE_OK;
}

// DynamicsCompressorNode::new, src/node/dynamics_compressor.rs:130-260
WAO_API wae_status wao_create_dynamics_compressor(wae_graph* g, const wae_dynamics_compressor_options* o, wae_node_id* out) {
    ChannelConfig def{2, MODE_CLAMPED_MAX, SPEAKERS};
    ChannelConfig cfg = resolve_cfg(o->channel_config, def);
    if (cfg.count > 2) return fail(WAE_NOT_SUPPORTED, "NotSupportedError - DynamicsCompressorNode channel count cannot be greater than two");
    if (cfg.mode == MODE_MAX) return fail(WAE_NOT_SUPPORTED, "NotSupportedError - DynamicsCompressorNode channel count mode cannot be set to max");
    uint32_t id = g->next_id++;
    uint32_t at = g->create_param(id, ParamDescriptor{0.003f, 0.f, 1.f, false}, o->attack);
    uint32_t kn = g->create_param(id, ParamDescriptor{30.f, 0.f, 40.f, false}, o->knee);
    uint32_t ra = g->create_param(id, ParamDescriptor{12.f, 1.f, 20.f, false}, o->ratio);
    uint32_t re = g->create_param(id, ParamDescriptor{0.25f, 0.f, 1.f, false}, o->release);
    uint32_t th = g->create_param(id, ParamDescriptor{-24.f, -100.f, 0.f, false}, o->threshold);
    auto r = std::make_unique<DynamicsCompressorRenderer>();
    r->attack = at; r->knee = kn; r->ratio = ra; r->release = re; r->threshold = th;
    r->ring_capacity = (size_t)std::ceil(g->sample_rate * 0.006f / (float)RQ) + 1;
    g->finish_register(id, std::move(r), K_COMP, 1, 1, cfg, {at, kn, ra, re, th}, id);
    *out = id;
    return WAE_OK;
}

// ChannelMergerNode::new, src/node/channel_merger.rs:120-140
WAO_API wae_status wao_create_channel_merger(wae_graph* g, const wae_channel_merger_options* o, wae_node_id* out) {
    uint32_t n = o->number_of_inputs ? o->number_of_inputs : 6;
    if (n < 1 || n > 32) return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - Invalid number of inputs");
    uint32_t id = g->next_id++;
    ChannelConfig cfg{1, MODE_EXPLICIT, SPEAKERS};
    g->finish_register(id, std::make_unique<ChannelMergerRenderer>(), K_MERGER, (int)n, 1, cfg, {}, id);
    *out = id;
    return WAE_OK;
}

// ChannelSplitterNode::new, src/node/channel_splitter.rs:140-180
WAO_API wae_status wao_create_channel_splitter(wae_graph* g, const wae_channel_splitter_options* o, wae_node_id* out) {
    uint32_t n = o->number_of_outputs ? o->number_of_outputs : 6;
    if (n < 1 || n > 32) return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - Invalid number of outputs");
    uint32_t id = g->next_id++;
    ChannelConfig cfg{(int)n, MODE_EXPLICIT, DISCRETE};
    g->finish_register(id, std::make_unique<ChannelSplitterRenderer>(), K_SPLITTER, 1, (int)n, cfg, {}, id);
    *out = id;
    return WAE_OK;
}

// AudioNode::connect_from_output_to_input, src/node/audio_node.rs:259-289
WAO_API wae_status wao_connect(wae_graph* g, wae_node_id from, uint32_t output, wae_node_id to, uint32_t input) {
    auto fi = g->info.find(from), ti = g->info.find(to);
    if (fi == g->info.end() || ti == g->info.end() || fi->second.kind == K_PARAM || ti->second.kind == K_PARAM)
        return fail(WAE_INVALID_ARGUMENT, "InvalidAccessError - unknown node");
    if ((int)output >= fi->second.n_outputs) return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - output port " + std::to_string(output) + " is out of bounds");
    if ((int)input >= ti->second.n_inputs) return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - input port " + std::to_string(input) + " is out of bounds");
    g->graph.add_edge(fi->second.out_id, (int)output, to, (int)input);
    return WAE_OK;
}

WAO_API wae_status wao_connect_param(wae_graph* g, wae_node_id from, uint32_t output, wae_node_id to, uint32_t param_index) {
    if (to == 1) g->ensure_listener();  // BaseAudioContext::listener() creates it on first access (context/mod.rs)
    auto fi = g->info.find(from), ti = g->info.find(to);
    if (fi == g->info.end() || ti == g->info.end()) return fail(WAE_INVALID_ARGUMENT, "InvalidAccessError - unknown node");
    if ((int)output >= fi->second.n_outputs) return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - output port out of bounds");
    if (param_index >= ti->second.params.size()) return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - param index out of bounds");
    g->graph.add_edge(fi->second.out_id, (int)output, ti->second.params[param_index], 0);
    return WAE_OK;
}

WAO_API wae_status wao_disconnect(wae_graph* g, wae_node_id from) {
    auto fi = g->info.find(from);
    if (fi == g->info.end()) return fail(WAE_INVALID_ARGUMENT, "InvalidAccessError - unknown node");
    // keep the hidden writer->reader edge of a DelayNode: disconnect acts on the reader (delay.rs:160-163)
    g->graph.remove_edges_from(fi->second.out_id);
    return WAE_OK;
}

// ConcreteBaseAudioContext::disconnect(from, Option<output>, Option<to>, Option<input>) (src/context/concrete_base.rs:474-507) and the
// render side Graph::remove_edge (src/render/graph.rs:266-283)
static wae_status disconnect_matching(wae_graph* g, wae_node_id from, int32_t output, bool has_to, uint32_t to_id, int32_t input) {
    auto fi = g->info.find(from);
    if (fi == g->info.end() || fi->second.kind == K_PARAM) return fail(WAE_INVALID_ARGUMENT, "InvalidAccessError - unknown node");
    if (output >= fi->second.n_outputs)
        return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - output port " + std::to_string(output) + " is out of bounds");
    auto& out = g->graph.get(fi->second.out_id)->outgoing;
    const size_t before = out.size();
    out.erase(std::remove_if(out.begin(), out.end(),
                             [&](const Edge& e) {
                                 if (e.other_index < 0) return false;
                                 return (output < 0 || e.self_index == output) && (!has_to || e.other_id == to_id) && (input < 0 || e.other_index == input);
                             }),
              out.end());
    g->graph.ordered.clear();
    if (has_to && out.size() == before) return fail(WAE_INVALID_ARGUMENT, "InvalidAccessError - attempting to disconnect unconnected nodes");
    return WAE_OK;
}
WAO_API wae_status wao_disconnect_from(wae_graph* g, wae_node_id from, int32_t output, wae_node_id to, int32_t input) {
    if (to == WAE_NODE_NONE) return disconnect_matching(g, from, output, false, 0, input);
    auto ti = g->info.find(to);
    if (ti == g->info.end() || ti->second.kind == K_PARAM) return fail(WAE_INVALID_ARGUMENT, "InvalidAccessError - unknown node");
    if (input >= ti->second.n_inputs)
        return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - input port " + std::to_string(input) + " is out of bounds");
    return disconnect_matching(g, from, output, true, to, input);
}
WAO_API wae_status wao_disconnect_param(wae_graph* g, wae_node_id from, int32_t output, wae_node_id to, uint32_t param_index) {
    auto ti = g->info.find(to);
    if (ti == g->info.end()) return fail(WAE_INVALID_ARGUMENT, "InvalidAccessError - unknown node");
    if (param_index >= ti->second.params.size()) return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - param index out of bounds");
    return disconnect_matching(g, from, output, true, ti->second.params[param_index], -1);
}

static wae_status push_event(ParamProcessor* p, const wae_param_event* e) {
    ParamEvent ev;
    ev.type = (int)e->type;
    ev.value = e->value;
    ev.time = e->time;
    auto finite = [](float v) { return std::isfinite(v); };
    auto valid_time = [](double t) { return std::isfinite(t) && t >= 0.; };
    switch (e->type) {
        case WAE_EVENT_SET_VALUE:
            if (!finite(e->value)) return fail(WAE_INVALID_ARGUMENT, "TypeError - The provided value is non-finite.");
            ev.time = 0.;
            break;
        case WAE_EVENT_SET_VALUE_AT_TIME:
        case WAE_EVENT_LINEAR_RAMP_TO_VALUE_AT_TIME:
            if (!finite(e->value)) return fail(WAE_INVALID_ARGUMENT, "TypeError - The provided value is non-finite.");
            if (!valid_time(e->time)) return fail(WAE_INVALID_ARGUMENT, "RangeError - time should be positive");
            break;
        case WAE_EVENT_EXPONENTIAL_RAMP_TO_VALUE_AT_TIME:
            if (!finite(e->value)) return fail(WAE_INVALID_ARGUMENT, "TypeError - The provided value is non-finite.");
            if (e->value == 0.f) return fail(WAE_INVALID_ARGUMENT, "RangeError - value (0.0) should not be equal to zero");
            if (!valid_time(e->time)) return fail(WAE_INVALID_ARGUMENT, "RangeError - time should be positive");
            break;
        case WAE_EVENT_SET_TARGET_AT_TIME:
            if (!finite(e->value)) return fail(WAE_INVALID_ARGUMENT, "TypeError - The provided value is non-finite.");
            if (!valid_time(e->time) || !valid_time(e->aux)) return fail(WAE_INVALID_ARGUMENT, "RangeError - time should be positive");
            if (e->aux == 0.) {
                ev.type = EV_SET_VALUE_AT_TIME;  // param.rs:529-538
            } else {
                ev.has_time_constant = true;
                ev.time_constant = e->aux;
            }
            break;
        case WAE_EVENT_CANCEL_SCHEDULED_VALUES:
        case WAE_EVENT_CANCEL_AND_HOLD_AT_TIME:
            if (!valid_time(e->time)) return fail(WAE_INVALID_ARGUMENT, "RangeError - time should be positive");
            ev.value = 0.f;
            break;
        case WAE_EVENT_SET_VALUE_CURVE_AT_TIME:
            if (e->values_len < 2) return fail(WAE_INVALID_STATE, "InvalidStateError - sequence length should not be less than 2");
            if (!valid_time(e->time)) return fail(WAE_INVALID_ARGUMENT, "RangeError - time should be positive");
            if (!(std::isfinite(e->aux) && e->aux > 0.)) return fail(WAE_INVALID_ARGUMENT, "RangeError - duration should be strictly positive");
            ev.value = 0.f;
            ev.has_duration = true;
            ev.duration = e->aux;
            ev.values.assign(e->values, e->values + e->values_len);
            break;
        default: return fail(WAE_INVALID_ARGUMENT, "unknown event type");
    }
    std::string err = p->handle_incoming_event(std::move(ev));
    if (!err.empty()) return fail(WAE_NOT_SUPPORTED, err);
    return WAE_OK;
}

WAO_API wae_status wao_param_event_push(wae_graph* g, wae_node_id node, uint32_t param_index, const wae_param_event* e) {
    auto ni = g->info.find(node);
    if (ni == g->info.end() || param_index >= ni->second.params.size()) return fail(WAE_INVALID_ARGUMENT, "unknown param");
    auto* p = static_cast<ParamProcessor*>(g->info[ni->second.params[param_index]].proc);
    return push_event(p, e);
}

// test hooks: one AudioParamProcessor driven like the reference's unit tests (param.rs:1766-3545)
WAO_API wae_status wao_param_sim_create(uint32_t a_rate, float default_value, float min_value, float max_value, void** out) {
    *out = new ParamProcessor(ParamDescriptor{default_value, min_value, max_value, a_rate != 0});
    return WAE_OK;
}
WAO_API wae_status wao_param_sim_destroy(void* s) {
    delete static_cast<ParamProcessor*>(s);
    return WAE_OK;
}
WAO_API wae_status wao_param_sim_push(void* s, const wae_param_event* e) { return push_event(static_cast<ParamProcessor*>(s), e); }
WAO_API wae_status wao_param_sim_set_automation_rate(void* s, uint32_t a_rate) {
    static_cast<ParamProcessor*>(s)->a_rate = a_rate != 0;
    return WAE_OK;
}
WAO_API wae_status wao_param_sim_compute(void* s, double block_time, double dt, uint32_t count, float* out, uint32_t* len) {
    auto* p = static_cast<ParamProcessor*>(s);
    p->compute_buffer(block_time, dt, (int)count);
    for (int i = 0; i < p->buffer_len; i++) out[i] = p->buffer[i];
    *len = (uint32_t)p->buffer_len;
    return WAE_OK;
}

WAO_API wae_status wao_listener_param_event_push(wae_graph* g, uint32_t param_index, const wae_param_event* e) {
    if (param_index >= 9) return fail(WAE_INVALID_ARGUMENT, "unknown listener param");
    g->ensure_listener();
    auto* p = static_cast<ParamProcessor*>(g->info[2 + param_index].proc);
    return push_event(p, e);
}

WAO_API wae_status wao_param_set_automation_rate(wae_graph* g, wae_node_id node, uint32_t param_index, uint32_t rate) {
    auto ni = g->info.find(node);
    if (ni == g->info.end() || param_index >= ni->second.params.size()) return fail(WAE_INVALID_ARGUMENT, "unknown param");
    NodeKind k = ni->second.kind;
    auto* p = static_cast<ParamProcessor*>(g->info[ni->second.params[param_index]].proc);
    bool want_a = rate == WAE_AUTOMATION_RATE_A;
    // automation_rate_constrained params (ABSN, compressor): param.rs:349-353
    if ((k == K_ABSN || k == K_COMP) && want_a != p->a_rate)
        return fail(WAE_INVALID_STATE, "InvalidStateError - automation rate cannot be changed for this param");
    p->a_rate = want_a;
    return WAE_OK;
}

// AudioScheduledSourceNode::start_at / stop_at
WAO_API wae_status wao_source_start(wae_graph* g, wae_node_id node, double when, double offset, double duration) {
    auto ni = g->info.find(node);
    if (ni == g->info.end()) return fail(WAE_INVALID_ARGUMENT, "unknown node");
    NodeInfo& n = ni->second;
    if (!(n.kind == K_OSC || n.kind == K_ABSN || n.kind == K_CONST)) return fail(WAE_INVALID_ARGUMENT, "not a scheduled source node");
    if (!(std::isfinite(when) && when >= 0.)) return fail(WAE_INVALID_ARGUMENT, "RangeError - when should be positive");
    if (n.has_start) return fail(WAE_INVALID_STATE, "InvalidStateError - Cannot call `start` twice");
    n.has_start = true;
    if (n.kind == K_OSC) static_cast<OscillatorRenderer*>(n.proc)->start_time = when;
    if (n.kind == K_CONST) static_cast<ConstantSourceRenderer*>(n.proc)->start_time = when;
    if (n.kind == K_ABSN) {
        if (!(offset >= 0.) || !(duration >= 0.)) return fail(WAE_INVALID_ARGUMENT, "RangeError - offset/duration should be positive");
        auto* r = static_cast<AudioBufferSourceRenderer*>(n.proc);
        r->start_time = when;
        r->offset = offset;
        r->duration = duration >= 1e300 ? 1.7976931348623157e308 : duration;
        r->clamp_loop_boundaries();
    }
    return WAE_OK;
}

WAO_API wae_status wao_source_stop(wae_graph* g, wae_node_id node, double when) {
    auto ni = g->info.find(node);
    if (ni == g->info.end()) return fail(WAE_INVALID_ARGUMENT, "unknown node");
    NodeInfo& n = ni->second;
    if (!(n.kind == K_OSC || n.kind == K_ABSN || n.kind == K_CONST)) return fail(WAE_INVALID_ARGUMENT, "not a scheduled source node");
    if (!(std::isfinite(when) && when >= 0.)) return fail(WAE_INVALID_ARGUMENT, "RangeError - when should be positive");
    if (!n.has_start) return fail(WAE_INVALID_STATE, "InvalidStateError cannot stop before start");
    if (n.kind == K_OSC) static_cast<OscillatorRenderer*>(n.proc)->stop_time = when;
    if (n.kind == K_CONST) static_cast<ConstantSourceRenderer*>(n.proc)->stop_time = when;
    if (n.kind == K_ABSN) {
        auto* r = static_cast<AudioBufferSourceRenderer*>(n.proc);
        r->stop_time = when;
        r->clamp_loop_boundaries();
    }
    return WAE_OK;
}

WAO_API wae_status wao_oscillator_set_type(wae_graph* g, wae_node_id node, uint32_t type) {
    auto ni = g->info.find(node);
    if (ni == g->info.end() || ni->second.kind != K_OSC) return fail(WAE_INVALID_ARGUMENT, "not an oscillator");
    if (type == WAE_OSC_CUSTOM) return fail(WAE_INVALID_STATE, "InvalidStateError: Custom type cannot be set manually");
    auto* r = static_cast<OscillatorRenderer*>(ni->second.proc);
    if (r->type == WAE_OSC_CUSTOM) return WAE_OK;
    r->type = (int)type;
    return WAE_OK;
}

WAO_API wae_status wao_biquad_set_type(wae_graph* g, wae_node_id node, uint32_t type) {
    auto ni = g->info.find(node);
    if (ni == g->info.end() || ni->second.kind != K_BIQUAD || type > 7) return fail(WAE_INVALID_ARGUMENT, "not a biquad / bad type");
    static_cast<BiquadFilterRenderer*>(ni->second.proc)->type = (int)type;
    return WAE_OK;
}

// AudioNode::set_channel_count / _mode / _interpretation (src/node/audio_node.rs:417-441) with the per-node overrides:
// param.rs:325-333, spatial.rs:113-121, destination.rs:55-96, channel_merger.rs:39-110, channel_splitter.rs:36-134,
// convolver.rs:187-197, dynamics_compressor.rs:168-178, stereo_panner.rs:143-152, panner.rs:363-372.
// The render side applies them to the node's ChannelConfig (src/render/graph.rs:306-320).
static wae_status set_cfg_field(wae_graph* g, wae_node_id node, int field, uint32_t v) {
    auto ni = g->info.find(node);
    if (ni == g->info.end()) return fail(WAE_INVALID_ARGUMENT, "InvalidAccessError - unknown node");
    const NodeInfo& n = ni->second;
    static const char* what[3] = {"channel count", "channel count mode", "channel interpretation"};
    if (field == 1 && v > 2) return fail(WAE_INVALID_ARGUMENT, "unknown channel count mode");
    if (field == 2 && v > 1) return fail(WAE_INVALID_ARGUMENT, "unknown channel interpretation");
    bool store = true;
    switch (n.kind) {
        case K_PARAM: return fail(WAE_NOT_SUPPORTED, std::string("NotSupportedError - AudioParam has ") + what[field] + " constraints");
        case K_LISTENER: return fail(WAE_NOT_SUPPORTED, std::string("NotSupportedError - AudioListenerNode has ") + what[field] + " constraints");
        case K_DEST:
            if (field == 0 && v != g->channels)
                return fail(WAE_NOT_SUPPORTED, "NotSupportedError - not allowed to change OfflineAudioContext destination channel count");
            if (field == 1 && v != MODE_EXPLICIT) return fail(WAE_NOT_SUPPORTED, "InvalidStateError - AudioDestinationNode has channel count mode constraints");
            break;
        case K_MERGER:
            if (field == 0 && v != 1) return fail(WAE_NOT_SUPPORTED, "InvalidStateError - channel count of ChannelMergerNode must be equal to 1");
            if (field == 1 && v != MODE_EXPLICIT) return fail(WAE_NOT_SUPPORTED, "InvalidStateError - channel count of ChannelMergerNode must be set to Explicit");
            store = field != 0;
            break;
        case K_SPLITTER:
            if (field == 0 && v != (uint32_t)n.n_outputs)
                return fail(WAE_NOT_SUPPORTED, "InvalidStateError - channel count of ChannelSplitterNode must be equal to number of outputs");
            if (field == 1 && v != MODE_EXPLICIT) return fail(WAE_NOT_SUPPORTED, "InvalidStateError - channel count mode of ChannelSplitterNode must be set to Explicit");
            if (field == 2 && v != DISCRETE) return fail(WAE_NOT_SUPPORTED, "InvalidStateError - channel interpretation of ChannelSplitterNode must be set to Discrete");
            store = field != 0;
            break;
        case K_CONV: case K_COMP: case K_SPANNER: case K_PANNER: {
            const char* name = n.kind == K_CONV ? "ConvolverNode" : n.kind == K_COMP ? "DynamicsCompressorNode" : n.kind == K_SPANNER ? "StereoPannerNode" : "PannerNode";
            if (field == 0 && v > 2) return fail(WAE_NOT_SUPPORTED, std::string("NotSupportedError - ") + name + " channel count cannot be greater than two");
            if (field == 1 && v == MODE_MAX) return fail(WAE_NOT_SUPPORTED, std::string("NotSupportedError - ") + name + " channel count mode cannot be set to max");
            break;
        }
        default: break;
    }
    if (field == 0 && (v < 1 || v > 32))
        return fail(WAE_NOT_SUPPORTED, "NotSupportedError - Invalid number of channels: " + std::to_string(v) + " is outside range [1, 32]");
    if (!store) return WAE_OK;
    ChannelConfig& cfg = g->graph.get(node)->cfg;
    if (field == 0) cfg.count = (int)v;
    else if (field == 1) cfg.mode = (int)v;
    else cfg.interp = (int)v;
    return WAE_OK;
}
WAO_API wae_status wao_node_set_channel_count(wae_graph* g, wae_node_id node, uint32_t count) { return set_cfg_field(g, node, 0, count); }
WAO_API wae_status wao_node_set_channel_count_mode(wae_graph* g, wae_node_id node, uint32_t mode) { return set_cfg_field(g, node, 1, mode); }
WAO_API wae_status wao_node_set_channel_interpretation(wae_graph* g, wae_node_id node, uint32_t v) { return set_cfg_field(g, node, 2, v); }

// PeriodicWave::new (src/periodic_wave.rs:104-209), test hook for the wavetable tests (:278-345)
WAO_API wae_status wao_periodic_wave_table(const float* real, const float* imag, uint32_t len, uint32_t disable_normalization, float* table,
                                           uint32_t table_len) {
    const bool has_r = real != nullptr, has_i = imag != nullptr;
    if ((has_r || has_i) && len < 2) return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - `real` and `imag` length should at least 2");
    if (!table || table_len == 0) return fail(WAE_INVALID_ARGUMENT, "null table");
    static const float sine_r[2] = {0.f, 0.f}, sine_i[2] = {0.f, 1.f};  // no coefficients: the built-in sine (periodic_wave.rs:143-146)
    const uint32_t n = (has_r || has_i) ? len : 2;
    const float pi_2 = 2.f * 3.14159265358979323846f;
    for (uint32_t i = 0; i < table_len; i++) {
        float sample = 0.f;
        const float phase = pi_2 * (float)i / (float)table_len;
        for (uint32_t j = 1; j < n; j++) {
            const float re = has_r ? real[j] : ((has_r || has_i) ? 0.f : sine_r[j]);
            const float im = has_i ? imag[j] : ((has_r || has_i) ? 0.f : sine_i[j]);
            const float rad = phase * (float)j;
            const float contrib = re * std::cos(rad) + im * std::sin(rad);
            sample += contrib;
        }
        table[i] = sample;
    }
    if (!disable_normalization) {
        float mx = 0.f;
        for (uint32_t i = 0; i < table_len; i++) mx = std::fabs(table[i]) > mx ? std::fabs(table[i]) : mx;
        if (mx > 0.f) {
            const float norm = 1.f / mx;
            for (uint32_t i = 0; i < table_len; i++) table[i] *= norm;
        }
    }
    return WAE_OK;
}

// test hook: PannerRenderer's gains and spatial::azimuth_and_elevation for one source / listener configuration (panner.rs:927-986, spatial.rs:205-299)
WAO_API wae_status wao_spatial_params(uint32_t distance_model, const double* model6, const float* v15, float* out4) {
    PannerRenderer r;
    r.distance_model = (int)distance_model;
    r.ref_distance = model6[0]; r.max_distance = model6[1]; r.rolloff_factor = model6[2];
    r.cone_inner_angle = model6[3]; r.cone_outer_angle = model6[4]; r.cone_outer_gain = model6[5];
    out4[0] = r.dist_gain(v15, v15 + 6);
    out4[1] = r.cone_gain(v15, v15 + 3, v15 + 6);
    azimuth_and_elevation(v15, v15 + 6, v15 + 9, v15 + 12, out4[2], out4[3]);
    return WAE_OK;
}

static Analyser* find_analyser(wae_graph* g, wae_node_id node);
// ---- node attributes set after construction: the onmessage handlers of the renderers ----------------------------------------
// AudioBufferSourceNode::set_buffer (audio_buffer_source.rs:278-288, onmessage :856-872)
WAO_API wae_status wao_buffer_source_set_buffer(wae_graph* g, wae_node_id node, const wae_audio_buffer* buffer) {
    auto ni = g->info.find(node);
    if (ni == g->info.end() || ni->second.kind != K_ABSN || !buffer) return fail(WAE_INVALID_ARGUMENT, "not an AudioBufferSourceNode / null buffer");
    auto* r = static_cast<AudioBufferSourceRenderer*>(ni->second.proc);
    if (r->buffer) return fail(WAE_INVALID_STATE, "InvalidStateError - cannot assign buffer twice");
    if (!valid_buffer(buffer)) return WAE_NOT_SUPPORTED;
    r->buffer = copy_buffer(buffer);
    r->clamp_loop_boundaries();
    return WAE_OK;
}
// ConvolverNode::set_buffer (convolver.rs:259-317)
WAO_API wae_status wao_convolver_set_buffer(wae_graph* g, wae_node_id node, const wae_audio_buffer* buffer) {
    auto ni = g->info.find(node);
    if (ni == g->info.end() || ni->second.kind != K_CONV || !buffer) return fail(WAE_INVALID_ARGUMENT, "not a ConvolverNode / null buffer");
    if (buffer->sample_rate != g->sample_rate)
        return fail(WAE_NOT_SUPPORTED, "NotSupportedError - sample rate of the convolution buffer must match the audio context");
    const uint32_t c = buffer->number_of_channels;
    if (!(c == 1 || c == 2 || c == 4)) return fail(WAE_NOT_SUPPORTED, "NotSupportedError - the convolution buffer must consist of 1, 2 or 4 channels");
    auto ab = copy_buffer(buffer);
    static_cast<ConvolverRenderer*>(ni->second.proc)->set_buffer(*ab, ni->second.normalize);
    return WAE_OK;
}
// WaveShaperNode::set_curve (waveshaper.rs:203-213)
WAO_API wae_status wao_wave_shaper_set_curve(wae_graph* g, wae_node_id node, const float* curve, uint32_t len) {
    auto ni = g->info.find(node);
    if (ni == g->info.end() || ni->second.kind != K_SHAPER || (!curve && len)) return fail(WAE_INVALID_ARGUMENT, "not a WaveShaperNode / null curve");
    auto* r = static_cast<WaveShaperRenderer*>(ni->second.proc);
    if (r->has_curve) return fail(WAE_INVALID_STATE, "InvalidStateError - cannot assign curve twice");
    r->set_curve(curve, len);
    return WAE_OK;
}
// OscillatorNode::set_periodic_wave (oscillator.rs:334-337, onmessage :350-361)
WAO_API wae_status wao_oscillator_set_periodic_wave(wae_graph* g, wae_node_id node, const float* table, uint32_t len) {
    auto ni = g->info.find(node);
    if (ni == g->info.end() || ni->second.kind != K_OSC || !table || len == 0) return fail(WAE_INVALID_ARGUMENT, "not an OscillatorNode / empty wavetable");
    auto* r = static_cast<OscillatorRenderer*>(ni->second.proc);
    r->type = WAE_OSC_CUSTOM;
    r->periodic_wave.assign(table, table + len);
    return WAE_OK;
}
WAO_API wae_status wao_node_set_attribute(wae_graph* g, wae_node_id node, uint32_t attribute, double value) {
    auto ni = g->info.find(node);
    if (ni == g->info.end()) return fail(WAE_INVALID_ARGUMENT, "InvalidAccessError - unknown node");
    NodeInfo& n = ni->second;
    auto wrong = [&]() { return fail(WAE_INVALID_ARGUMENT, "this node has no such attribute"); };
    switch (attribute) {
        case WAE_ATTR_LOOP: case WAE_ATTR_LOOP_START: case WAE_ATTR_LOOP_END: {  // ControlMessage::Loop / LoopStart / LoopEnd, :873-893
            if (n.kind != K_ABSN) return wrong();
            auto* r = static_cast<AudioBufferSourceRenderer*>(n.proc);
            if (attribute == WAE_ATTR_LOOP) r->is_looping = value != 0.;
            else if (attribute == WAE_ATTR_LOOP_START) r->loop_start = value;
            else r->loop_end = value;
            if (attribute != WAE_ATTR_LOOP && r->buffer) r->clamp_loop_boundaries();
            return WAE_OK;
        }
        case WAE_ATTR_NORMALIZE:
            if (n.kind != K_CONV) return wrong();
            n.normalize = value != 0.;
            return WAE_OK;
        case WAE_ATTR_OVERSAMPLE:
            if (n.kind != K_SHAPER) return wrong();
            if (!(value == 0. || value == 1. || value == 2.)) return fail(WAE_INVALID_ARGUMENT, "unknown oversample type");
            static_cast<WaveShaperRenderer*>(n.proc)->oversample = (int)value;
            return WAE_OK;
        case WAE_ATTR_PANNING_MODEL: case WAE_ATTR_DISTANCE_MODEL: case WAE_ATTR_REF_DISTANCE: case WAE_ATTR_MAX_DISTANCE:
        case WAE_ATTR_ROLLOFF_FACTOR: case WAE_ATTR_CONE_INNER_ANGLE: case WAE_ATTR_CONE_OUTER_ANGLE: case WAE_ATTR_CONE_OUTER_GAIN: {
            if (n.kind != K_PANNER) return wrong();
            auto* r = static_cast<PannerRenderer*>(n.proc);
            switch (attribute) {
                case WAE_ATTR_PANNING_MODEL:
                    if (value == 1.) {
                        std::string err;
                        if (!hrtf_sphere_available(err)) return fail(WAE_UNSUPPORTED, err);
                        r->set_hrtf(g->sample_rate);
                        if (!r->hrtf_state) return fail(WAE_UNSUPPORTED, "HRTF panning: the HRIR sphere cannot be used at this sample rate");
                    } else if (value == 0.) {
                        r->hrtf_state.reset();
                    } else {
                        return fail(WAE_INVALID_ARGUMENT, "unknown panning model");
                    }
                    break;
                case WAE_ATTR_DISTANCE_MODEL:
                    if (!(value == 0. || value == 1. || value == 2.)) return fail(WAE_INVALID_ARGUMENT, "unknown distance model");
                    r->distance_model = (int)value;
                    break;
                case WAE_ATTR_REF_DISTANCE:
                    if (!(value >= 0.)) return fail(WAE_INVALID_ARGUMENT, "RangeError - refDistance cannot be negative");
                    r->ref_distance = value;
                    break;
                case WAE_ATTR_MAX_DISTANCE:
                    if (!(value > 0.)) return fail(WAE_INVALID_ARGUMENT, "RangeError - maxDistance must be strictly positive");
                    r->max_distance = value;
                    break;
                case WAE_ATTR_ROLLOFF_FACTOR:
                    if (!(value >= 0.)) return fail(WAE_INVALID_ARGUMENT, "RangeError - rolloffFactor cannot be negative");
                    r->rolloff_factor = value;
                    break;
                case WAE_ATTR_CONE_INNER_ANGLE: r->cone_inner_angle = value; break;
                case WAE_ATTR_CONE_OUTER_ANGLE: r->cone_outer_angle = value; break;
                default:
                    if (!(value >= 0. && value <= 1.)) return fail(WAE_INVALID_STATE, "InvalidStateError - coneOuterGain must be in the range [0, 1]");
                    r->cone_outer_gain = value;
            }
            return WAE_OK;
        }
        case WAE_ATTR_FFT_SIZE: case WAE_ATTR_SMOOTHING_TIME_CONSTANT: case WAE_ATTR_MIN_DECIBELS: case WAE_ATTR_MAX_DECIBELS: {
            if (n.kind != K_ANALYSER) return wrong();
            Analyser* a = find_analyser(g, node);
            if (attribute == WAE_ATTR_FFT_SIZE) {
                const uint64_t f = (uint64_t)value;
                if (!((double)f == value && f >= 32 && f <= 32768 && (f & (f - 1)) == 0))
                    return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - Invalid fft size: must be a power of two in [32, 32768]");
                a->set_fft_size((size_t)f);
            } else if (attribute == WAE_ATTR_SMOOTHING_TIME_CONSTANT) {
                if (!(value >= 0. && value <= 1.)) return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - Invalid smoothing time constant: must be in [0, 1]");
                a->smoothing_time_constant = value;
            } else if (attribute == WAE_ATTR_MIN_DECIBELS) {
                if (!(value < a->max_decibels)) return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - Invalid min decibels: must be less than max decibels");
                a->min_decibels = value;
            } else {
                if (!(value > a->min_decibels)) return fail(WAE_INVALID_ARGUMENT, "IndexSizeError - Invalid max decibels: must be greater than min decibels");
                a->max_decibels = value;
            }
            return WAE_OK;
        }
        default: return fail(WAE_INVALID_ARGUMENT, "unknown attribute");
    }
}

// ---- rendering --------------------------------------------------------------------------------------

struct NoDenormals {  // crate no_denormals: FTZ + DAZ while rendering (src/render/thread.rs:373-380)
    unsigned int saved;
    NoDenormals() {
        saved = _mm_getcsr();
        _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
        _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
    }
    ~NoDenormals() { _mm_setcsr(saved); }
};

// render_audiobuffer_sync + render_offline_quantum, src/render/thread.rs:260-302,355-396.
// out: planar [channels][length]
// render quanta until `until_frame` into dst ([channels][length] planar) — render_audiobuffer_sync's loop body
// (src/render/thread.rs:260-302,355-396)
static void render_until(wae_graph* g, float* dst, uint64_t until_frame) {
    const uint64_t length = g->length;
    while (g->frames_played < until_frame) {
        uint64_t current_frame = g->frames_played;
        g->frames_played += RQ;
        Scope scope{current_frame, (double)current_frame / (double)g->sample_rate, g->sample_rate};
        const Quantum* rendered;
        {
            NoDenormals guard;
            rendered = &g->graph.render(scope);
        }
        uint64_t remaining = std::min<uint64_t>(length - current_frame, RQ);
        for (uint32_t c = 0; c < g->channels; c++) {
            float* d = dst + (size_t)c * length + current_frame;
            if ((int)c < rendered->number_of_channels())
                std::memcpy(d, rendered->channel((int)c).data(), remaining * sizeof(float));
            else
                std::memset(d, 0, remaining * sizeof(float));
        }
    }
}

// OfflineAudioContext::suspend_sync (src/context/offline.rs:330-387; render loop: src/render/thread.rs:271-290): rendering
// pauses at the quantised frame, the caller then mutates the graph (that is the callback) and rendering resumes.  The oracle
// is a live graph, so it simply renders up to the suspend point now.
WAO_API wae_status wao_graph_suspend(wae_graph* g, double suspend_time) {
    if (g->rendered) return fail(WAE_INVALID_STATE, "InvalidStateError - cannot suspend when rendering has already started");
    if (!(suspend_time >= 0.)) return fail(WAE_INVALID_STATE, "InvalidStateError - suspendTime cannot be negative");
    const uint64_t quantum = (uint64_t)std::ceil(suspend_time * (double)g->sample_rate / (double)RQ);  // offline.rs:248-251
    const uint64_t total = (g->length + RQ - 1) / RQ;
    if (quantum * RQ <= g->frames_played && !(quantum == 0 && g->frames_played == 0))
        return fail(WAE_INVALID_STATE, "InvalidStateError - cannot suspend at a time that is not after the current time");
    if (quantum >= total) return fail(WAE_INVALID_STATE, "InvalidStateError - cannot suspend after the end of the rendering");
    for (uint64_t q : g->suspend_quanta)
        if (q == quantum) return fail(WAE_INVALID_STATE, "InvalidStateError - cannot suspend multiple times at the same render quantum");
    g->suspend_quanta.push_back(quantum);
    if (g->partial.empty()) g->partial.assign((size_t)g->channels * g->length, 0.f);
    render_until(g, g->partial.data(), quantum * RQ);
    return WAE_OK;
}

WAO_API wae_status wao_render(wae_graph* g, float* out) {
    if (g->rendered) return fail(WAE_INVALID_STATE, "InvalidStateError - Cannot call `startRendering` twice");
    g->rendered = true;
    if (!g->partial.empty()) std::memcpy(out, g->partial.data(), g->partial.size() * sizeof(float));
    render_until(g, out, (g->length + RQ - 1) / RQ * RQ);
    return WAE_OK;
}

// The reference renders one context per thread; users parallelise across contexts
// (examples/decode_multithreaded.rs:16-23).  CPU-baseline helper: render n graphs on `threads` workers.
WAO_API wae_status wao_render_many(wae_graph* const* graphs, uint32_t n, float* out, uint32_t threads, double* seconds) {
    if (n == 0) return WAE_OK;
    size_t stride = (size_t)graphs[0]->channels * graphs[0]->length;
    std::atomic<uint32_t> next{0};
    std::atomic<int> bad{0};
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        for (;;) {
            uint32_t i = next.fetch_add(1);
            if (i >= n) break;
            if (wao_render(graphs[i], out + stride * i) != WAE_OK) bad = 1;
        }
    };
    if (threads <= 1) {
        work();
    } else {
        std::vector<std::thread> ts;
        for (uint32_t t = 0; t < threads; t++) ts.emplace_back(work);
        for (auto& t : ts) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    return bad ? fail(WAE_INVALID_STATE, "a graph failed to render") : WAE_OK;
}

// graph introspection for the order tests (src/render/graph/test.rs)
WAO_API uint32_t wao_render_order(wae_graph* g, uint32_t* ids, uint32_t cap) {
    const auto& o = g->graph.order();
    uint32_t n = (uint32_t)o.size();
    for (uint32_t i = 0; i < n && i < cap; i++) ids[i] = o[i];
    return n;
}

// analyser read-out (control thread API), src/node/analyser.rs:222-264
static Analyser* find_analyser(wae_graph* g, wae_node_id node) {
    for (auto& a : g->analysers)
        if (a.first == node) return a.second.get();
    return nullptr;
}
WAO_API wae_status wao_analyser_get_float_time_domain_data(wae_graph* g, wae_node_id node, float* out, uint32_t len) {
    Analyser* a = find_analyser(g, node);
    if (!a) return fail(WAE_INVALID_ARGUMENT, "not an analyser");
    a->get_float_time_domain_data(out, len);
    return WAE_OK;
}
WAO_API wae_status wao_analyser_get_float_frequency_data(wae_graph* g, wae_node_id node, float* out, uint32_t len) {
    Analyser* a = find_analyser(g, node);
    if (!a) return fail(WAE_INVALID_ARGUMENT, "not an analyser");
    a->get_float_frequency_data(out, len, (double)g->frames_played / (double)g->sample_rate);
    return WAE_OK;
}
WAO_API wae_status wao_analyser_get_byte_frequency_data(wae_graph* g, wae_node_id node, uint8_t* out, uint32_t len) {
    Analyser* a = find_analyser(g, node);
    if (!a) return fail(WAE_INVALID_ARGUMENT, "not an analyser");
    a->get_byte_frequency_data(out, len, (double)g->frames_played / (double)g->sample_rate);
    return WAE_OK;
}
WAO_API wae_status wao_analyser_get_byte_time_domain_data(wae_graph* g, wae_node_id node, uint8_t* out, uint32_t len) {
    Analyser* a = find_analyser(g, node);
    if (!a) return fail(WAE_INVALID_ARGUMENT, "not an analyser");
    a->get_byte_time_domain_data(out, len);
    return WAE_OK;
}

// load_hrtf_processor (src/node/panner.rs:39-68): the HRIR sphere bytes (resources/IRC_1003_C.bin in the reference)
WAO_API wae_status wao_set_hrir_sphere(const void* data, uint64_t len) {
    std::string err;
    if (!hrtf_set_sphere(data, len, err)) return fail(WAE_INVALID_ARGUMENT, err);
    return WAE_OK;
}
// test hook: triangle + barycentric weights for a direction in the sphere's own coordinates
WAO_API int32_t wao_hrtf_locate(const float* pos, const uint32_t* faces, uint32_t n_faces, const float* dir, uint32_t* idx, float* k) {
    return hrtf_locate(pos, faces, n_faces, dir, idx, k) ? 1 : 0;
}
// DynamicsCompressorNode::reduction
WAO_API wae_status wao_compressor_reduction(wae_graph* g, wae_node_id node, float* out) {
    auto ni = g->info.find(node);
    if (ni == g->info.end() || ni->second.kind != K_COMP) return fail(WAE_INVALID_ARGUMENT, "not a compressor");
    *out = static_cast<DynamicsCompressorRenderer*>(ni->second.proc)->reduction;
    return WAE_OK;
}

// ---- known-answer helpers (control-side functions the reference's tests pin) ----------------------------
WAO_API void wao_biquad_coefs(uint32_t type, double sample_rate, double f0, double gain, double q, double* out5) {
    BiquadCoefs c = biquad_calculate_coefs((int)type, sample_rate, f0, gain, q);
    out5[0] = c.b0; out5[1] = c.b1; out5[2] = c.b2; out5[3] = c.a1; out5[4] = c.a2;
}
WAO_API void wao_biquad_frequency_response(uint32_t type, float sample_rate, float frequency, float detune, float q, float gain,
                                           const float* freq_hz, float* mag, float* phase, uint32_t n) {
    biquad_frequency_response((int)type, sample_rate, frequency, detune, q, gain, freq_hz, mag, phase, (int)n);
}
WAO_API void wao_iir_frequency_response(const double* ff, uint32_t nff, const double* fb, uint32_t nfb, float sample_rate,
                                        const float* freq_hz, float* mag, float* phase, uint32_t n) {
    iir_frequency_response(std::vector<double>(ff, ff + nff), std::vector<double>(fb, fb + nfb), sample_rate, freq_hz, mag, phase, (int)n);
}
WAO_API float wao_db_to_lin(float v) { return compressor_db_to_lin(v); }
WAO_API float wao_lin_to_db(float v) { return compressor_lin_to_db(v); }
WAO_API void wao_blackman(uint32_t size, float* out) {
    auto w = generate_blackman(size);
    std::memcpy(out, w.data(), size * sizeof(float));
}
WAO_API float wao_convolver_normalize(const wae_audio_buffer* b) { return convolver_normalize_buffer(*copy_buffer(b)); }
// AudioBuffer::resample, src/buffer.rs:311-363 (linear interpolation; the input side of configs C4/C5)
WAO_API uint64_t wao_resample_linear(const float* in, uint64_t len, float from_rate, float to_rate, float* out, uint64_t out_cap);
// AudioRenderQuantum::mix test hook (src/render/quantum.rs:803-1440): in [from][128] -> out [to][128]
WAO_API void wao_mix(const float* in, uint32_t from, uint32_t to, uint32_t interpretation, float* out) {
    Alloc alloc;
    Channel silence(&alloc.zeroes, &alloc);
    Quantum q(silence);
    q.set_number_of_channels((int)from);
    for (uint32_t c = 0; c < from; c++) std::memcpy(q.channel_mut((int)c).make_mut(), in + 128 * c, 128 * sizeof(float));
    q.mix((int)to, (int)interpretation);
    for (uint32_t c = 0; c < to; c++) std::memcpy(out + 128 * c, q.channel((int)c).data(), 128 * sizeof(float));
}

// test hook: one impulse response through the restated sphere resampler (wao_hrtf.cpp)
WAO_API wae_status wao_hrir_resample(const float* hrir, uint32_t len, double ratio, float* out, uint32_t cap, uint32_t* n) {
    std::vector<float> r = wao::resample_hrir(std::vector<float>(hrir, hrir + len), ratio);
    *n = (uint32_t)r.size();
    for (uint32_t i = 0; i < *n && i < cap; i++) out[i] = r[i];
    return WAE_OK;
}

}  // extern "C"
