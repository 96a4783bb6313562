// ORACLE — TEST INFRASTRUCTURE ONLY (see wao_core.h).
// Restates src/param.rs: sample formulas :64-120, process/mix_to_output :685-797,
// handle_incoming_event :799-1036, automation computations :1038-1498, compute_buffer :1500-1600.
#include "wao_param.h"

namespace wao {

static const float SNAP_TO_TARGET = 1e-10f;  // param.rs:22

// param.rs:64-75
static inline float linear_ramp_sample(double start_time, double duration, float start_value, float diff, double time) {
    double phase = (time - start_time) / duration;
    return std::fma(diff, (float)phase, start_value);
}
// param.rs:78-88
static inline float exp_ramp_sample(double start_time, double duration, float start_value, float ratio, double time) {
    double phase = (time - start_time) / duration;
    return start_value * std::pow(ratio, (float)phase);
}
// param.rs:91-101
static inline float set_target_sample(double start_time, double time_constant, float end_value, float diff, double time) {
    double exponent = -((time - start_time) / time_constant);
    return std::fma(diff, (float)std::exp(exponent), end_value);
}
// param.rs:105-120
static inline float value_curve_sample(double start_time, double duration, const std::vector<float>& values, double time) {
    if (time - start_time >= duration) return values[values.size() - 1];
    double position = (double)(values.size() - 1) * (time - start_time) / duration;
    size_t k = position > 0. ? (size_t)position : 0;  // Rust's `as usize` saturates: a time before the curve's start (negative position) -> 0
    float phase = (float)(position - std::floor(position));
    return std::fma(values[k + 1] - values[k], phase, values[k]);
}

static inline float clamp_nan(float v, float def, float mn, float mx) {
    if (std::isnan(v)) return def;
    // f32::max / f32::min semantics of Rust on non-NaN inputs
    v = v > mn ? v : mn;
    v = v < mx ? v : mx;
    return v;
}

void ParamProcessor::sort_timeline() {
    std::stable_sort(timeline.begin(), timeline.end(), [](const ParamEvent& a, const ParamEvent& b) { return a.time < b.time; });
}

// param.rs:685-702
bool ParamProcessor::process(std::vector<Quantum>& inputs, std::vector<Quantum>& outputs, const ParamValues&, const Scope& scope) {
    double period = 1. / (double)scope.sample_rate;
    compute_buffer(scope.current_time, period, RQ);
    mix_to_output(inputs[0], outputs[0]);
    return true;
}

// param.rs:739-797
void ParamProcessor::mix_to_output(const Quantum& input, Quantum& output) {
    if (buffer_len == 1 || !a_rate) {
        float value = buffer[0];
        if (input.is_silent() || !a_rate) {
            output.single_valued = true;
            value += input.channel(0).data()[0];
            value = clamp_nan(value, default_value, min_value, max_value);
            output.channel_mut(0).make_mut()[0] = value;
        } else {
            output.single_valued = false;
            output = input;
            output.single_valued = false;
            float* o = output.channel_mut(0).make_mut();
            for (int i = 0; i < RQ; i++) {
                o[i] += value;
                o[i] = clamp_nan(o[i], default_value, min_value, max_value);
            }
        }
    } else {
        output = input;
        output.single_valued = false;
        float* o = output.channel_mut(0).make_mut();
        for (int i = 0; i < RQ; i++) {
            o[i] += buffer[i];
            o[i] = clamp_nan(o[i], default_value, min_value, max_value);
        }
    }
}

// param.rs:799-1036
std::string ParamProcessor::handle_incoming_event(ParamEvent event) {
    if (event.type == EV_CANCEL_SCHEDULED) {
        if (!timeline.empty()) {
            const ParamEvent& cur = timeline.front();
            if ((cur.type == EV_LINEAR_RAMP || cur.type == EV_EXP_RAMP) && cur.time >= event.time) {
                if (has_last_event) intrinsic_value = last_event.value;
            }
        }
        timeline.erase(std::remove_if(timeline.begin(), timeline.end(), [&](const ParamEvent& q) { return !(q.time < event.time); }),
                       timeline.end());
        return "";
    }
    if (event.type == EV_CANCEL_AND_HOLD) {
        ParamEvent* e1 = nullptr;
        ParamEvent* e2 = nullptr;
        double t1 = -1.7976931348623157e308, t2 = 1.7976931348623157e308;
        sort_timeline();
        for (auto& q : timeline) {
            if (q.time >= t1 && q.time <= event.time) {
                t1 = q.time;
                e1 = &q;
            } else if (q.time < t2 && q.time > event.time) {
                t2 = q.time;
                e2 = &q;
            }
        }
        if (e2) {
            if (e2->type == EV_LINEAR_RAMP || e2->type == EV_EXP_RAMP) {
                e2->has_cancel_time = true;
                e2->cancel_time = event.time;
            }
        } else if (e1) {
            if (e1->type == EV_SET_TARGET) {
                e1->has_cancel_time = true;
                e1->cancel_time = event.time;
            } else if (e1->type == EV_SET_VALUE_CURVE) {
                if (event.time <= e1->time + e1->duration) {
                    e1->has_cancel_time = true;
                    e1->cancel_time = event.time;
                }
            }
        }
        timeline.erase(std::remove_if(timeline.begin(), timeline.end(),
                                      [&](const ParamEvent& q) {
                                          double t = q.has_cancel_time ? q.cancel_time : q.time;
                                          return !(t <= event.time);
                                      }),
                       timeline.end());
        return "";
    }
    if (event.type == EV_SET_VALUE_CURVE) {
        double start_time = event.time, end_time = start_time + event.duration;
        for (auto& q : timeline)
            if (!(q.time <= start_time || q.time >= end_time))
                return "NotSupportedError - scheduling SetValueCurveAtTime at time of another automation event";
    }
    if (event.type == EV_SET_VALUE_AT_TIME || event.type == EV_SET_VALUE || event.type == EV_LINEAR_RAMP ||
        event.type == EV_EXP_RAMP || event.type == EV_SET_TARGET) {
        for (auto& q : timeline) {
            if (q.type == EV_SET_VALUE_CURVE) {
                double start_time = q.time, end_time = start_time + q.duration;
                if (!(event.time <= start_time || event.time >= end_time))
                    return "NotSupportedError - scheduling automation event during SetValueCurveAtTime";
            }
        }
    }
    if (event.type == EV_SET_VALUE) intrinsic_value = event.value;

    if (timeline.empty() && !has_last_event && (event.type == EV_LINEAR_RAMP || event.type == EV_EXP_RAMP)) {
        ParamEvent sv;
        sv.type = EV_SET_VALUE;
        sv.value = intrinsic_value;
        sv.time = 0.;
        timeline.push_back(sv);
    }
    if (timeline.empty() && event.type == EV_SET_TARGET) {
        ParamEvent sv;
        sv.type = EV_SET_VALUE;
        sv.value = intrinsic_value;
        sv.time = 0.;
        timeline.push_back(sv);
    }
    timeline.push_back(std::move(event));
    sort_timeline();
    return "";
}

static inline int end_index_of(double end_time, double block_time, double dt, int count) {
    double v = std::max(end_time - block_time, 0.) / dt;
    double r = std::round(v);
    // Rust `as usize` saturates
    if (!(r < 4.0e9)) return count;
    int idx = (int)r;
    return std::min(idx, count);
}

// param.rs:1038-1091
bool ParamProcessor::compute_set_value(const BlockInfos& infos) {
    ParamEvent& event = timeline.front();
    double time = event.time;
    if (time == 0.) time = infos.block_time;
    if (infos.is_a_rate) {
        int end_index_clipped = end_index_of(time, infos.block_time, infos.dt, infos.count);
        for (int i = buffer_len; i < end_index_clipped; i++) push(intrinsic_value);
    }
    if (time > infos.next_block_time) return true;
    intrinsic_value = event.value;
    ParamEvent ev = std::move(timeline.front());
    timeline.erase(timeline.begin());
    ev.time = time;
    last_event = std::move(ev);
    has_last_event = true;
    return false;
}

// param.rs:1093-1170
bool ParamProcessor::compute_linear_ramp(const BlockInfos& infos) {
    ParamEvent& event = timeline.front();
    double start_time = last_event.time;
    double end_time = event.time;
    double duration = end_time - start_time;
    if (event.has_cancel_time) end_time = event.cancel_time;
    float start_value = last_event.value;
    float end_value = event.value;
    float diff = end_value - start_value;
    if (infos.is_a_rate) {
        int start_index = buffer_len;
        int end_index_clipped = end_index_of(end_time, infos.block_time, infos.dt, infos.count);
        if (end_index_clipped > start_index) {
            double time = std::fma((double)start_index, infos.dt, infos.block_time);
            float value = 0.f;
            for (int i = start_index; i < end_index_clipped; i++) {
                value = linear_ramp_sample(start_time, duration, start_value, diff, time);
                push(value);
                time += infos.dt;
            }
            intrinsic_value = value;
        }
    }
    if (end_time >= infos.next_block_time) {
        intrinsic_value = linear_ramp_sample(start_time, duration, start_value, diff, infos.next_block_time);
        return true;
    }
    if (event.has_cancel_time) {
        float value = linear_ramp_sample(start_time, duration, start_value, diff, end_time);
        intrinsic_value = value;
        ParamEvent le = std::move(timeline.front());
        timeline.erase(timeline.begin());
        le.time = end_time;
        le.value = value;
        last_event = std::move(le);
    } else {
        intrinsic_value = end_value;
        last_event = std::move(timeline.front());
        timeline.erase(timeline.begin());
    }
    has_last_event = true;
    return false;
}

// param.rs:1172-1272
bool ParamProcessor::compute_exp_ramp(const BlockInfos& infos) {
    ParamEvent& event = timeline.front();
    double start_time = last_event.time;
    double end_time = event.time;
    double duration = end_time - start_time;
    if (event.has_cancel_time) end_time = event.cancel_time;
    float start_value = last_event.value;
    float end_value = event.value;
    float ratio = end_value / start_value;
    if (start_value == 0.f || start_value * end_value < 0.f) {
        ParamEvent e;
        e.type = EV_SET_VALUE_AT_TIME;
        e.time = end_time;
        e.value = end_value;
        timeline.front() = e;
        return false;
    }
    if (infos.is_a_rate) {
        int start_index = buffer_len;
        int end_index_clipped = end_index_of(end_time, infos.block_time, infos.dt, infos.count);
        if (end_index_clipped > start_index) {
            double time = std::fma((double)start_index, infos.dt, infos.block_time);
            float value = 0.f;
            for (int i = start_index; i < end_index_clipped; i++) {
                value = exp_ramp_sample(start_time, duration, start_value, ratio, time);
                push(value);
                time += infos.dt;
            }
            intrinsic_value = value;
        }
    }
    if (end_time >= infos.next_block_time) {
        intrinsic_value = exp_ramp_sample(start_time, duration, start_value, ratio, infos.next_block_time);
        return true;
    }
    if (event.has_cancel_time) {
        float value = exp_ramp_sample(start_time, duration, start_value, ratio, end_time);
        intrinsic_value = value;
        ParamEvent le = std::move(timeline.front());
        timeline.erase(timeline.begin());
        le.time = end_time;
        le.value = value;
        last_event = std::move(le);
    } else {
        intrinsic_value = end_value;
        last_event = std::move(timeline.front());
        timeline.erase(timeline.begin());
    }
    has_last_event = true;
    return false;
}

// param.rs:1274-1427
bool ParamProcessor::compute_set_target(const BlockInfos& infos) {
    ParamEvent& event = timeline.front();
    double end_time = infos.next_block_time;
    bool ended = false;
    if (timeline.size() > 1) {
        const ParamEvent& next = timeline[1];
        if (next.type == EV_LINEAR_RAMP || next.type == EV_EXP_RAMP) {
            end_time = infos.block_time;
            ended = true;
        } else if (next.time < infos.next_block_time) {
            end_time = next.time;
            ended = true;
        }
    }
    if (event.has_cancel_time && event.cancel_time < infos.next_block_time) {
        end_time = event.cancel_time;
        ended = true;
    }
    double start_time = event.time;
    float start_value = last_event.value;
    float end_value = event.value;
    float diff = start_value - end_value;
    double time_constant = event.time_constant;
    if (infos.is_a_rate) {
        int start_index = buffer_len;
        int end_index_clipped = end_index_of(end_time, infos.block_time, infos.dt, infos.count);
        if (end_index_clipped > start_index) {
            double time = std::fma((double)start_index, infos.dt, infos.block_time);
            float value = 0.f;
            for (int i = start_index; i < end_index_clipped; i++) {
                value = (time - start_time < 0.) ? intrinsic_value
                                                 : set_target_sample(start_time, time_constant, end_value, diff, time);
                push(value);
                time += infos.dt;
            }
            intrinsic_value = value;
        }
    }
    if (!ended) {
        float value = set_target_sample(start_time, time_constant, end_value, diff, infos.next_block_time);
        float d = std::fabs(end_value - value);
        if (d < SNAP_TO_TARGET) {
            intrinsic_value = end_value;
            if (end_value == 0.f) {
                for (int i = 0; i < buffer_len; i++)
                    if (std::fpclassify(buffer[i]) == FP_SUBNORMAL) buffer[i] = 0.f;
            }
            ParamEvent e;
            e.type = EV_SET_VALUE_AT_TIME;
            e.time = infos.next_block_time;
            e.value = end_value;
            timeline.front() = e;
        } else {
            intrinsic_value = value;
        }
        return true;
    }
    float value = set_target_sample(start_time, time_constant, end_value, diff, end_time);
    intrinsic_value = value;
    ParamEvent ev = std::move(timeline.front());
    timeline.erase(timeline.begin());
    ev.time = end_time;
    ev.value = value;
    last_event = std::move(ev);
    has_last_event = true;
    return false;
}

// param.rs:1429-1498
bool ParamProcessor::compute_set_value_curve(const BlockInfos& infos) {
    ParamEvent& event = timeline.front();
    double start_time = event.time;
    double duration = event.duration;
    const std::vector<float>& values = event.values;
    double end_time = start_time + duration;
    if (event.has_cancel_time) end_time = event.cancel_time;
    if (infos.is_a_rate) {
        int start_index = buffer_len;
        int end_index_clipped = end_index_of(end_time, infos.block_time, infos.dt, infos.count);
        if (end_index_clipped > start_index) {
            double time = std::fma((double)start_index, infos.dt, infos.block_time);
            float value = 0.f;
            for (int i = start_index; i < end_index_clipped; i++) {
                value = (time < start_time) ? intrinsic_value : value_curve_sample(start_time, duration, values, time);
                push(value);
                time += infos.dt;
            }
            intrinsic_value = value;
        }
    }
    if (end_time >= infos.next_block_time) {
        intrinsic_value = value_curve_sample(start_time, duration, values, infos.next_block_time);
        return true;
    }
    float value = event.has_cancel_time ? value_curve_sample(start_time, duration, values, end_time) : values[values.size() - 1];
    ParamEvent le = std::move(timeline.front());
    timeline.erase(timeline.begin());
    le.time = end_time;
    le.value = value;
    intrinsic_value = value;
    last_event = std::move(le);
    has_last_event = true;
    return false;
}

// param.rs:1500-1600
void ParamProcessor::compute_buffer(double block_time, double dt, int count) {
    float clamped = intrinsic_value;
    clamped = clamped < min_value ? min_value : clamped;
    clamped = clamped > max_value ? max_value : clamped;
    current_value = clamped;
    buffer_len = 0;
    bool is_a_rate = a_rate;
    double next_block_time = std::fma(dt, (double)count, block_time);
    bool is_constant_block;
    if (timeline.empty()) {
        is_constant_block = true;
    } else {
        const ParamEvent& e = timeline.front();
        if (e.type != EV_LINEAR_RAMP && e.type != EV_EXP_RAMP)
            is_constant_block = e.time >= next_block_time;
        else
            is_constant_block = false;
    }
    if (!is_a_rate || is_constant_block) {
        push(intrinsic_value);
        if (is_constant_block) return;
    }
    BlockInfos infos{block_time, dt, count, is_a_rate, next_block_time};
    for (;;) {
        bool exit_loop;
        if (timeline.empty()) {
            if (is_a_rate)
                for (int i = buffer_len; i < count; i++) push(intrinsic_value);
            exit_loop = true;
        } else {
            switch (timeline.front().type) {
                case EV_SET_VALUE:
                case EV_SET_VALUE_AT_TIME: exit_loop = compute_set_value(infos); break;
                case EV_LINEAR_RAMP: exit_loop = compute_linear_ramp(infos); break;
                case EV_EXP_RAMP: exit_loop = compute_exp_ramp(infos); break;
                case EV_SET_TARGET: exit_loop = compute_set_target(infos); break;
                case EV_SET_VALUE_CURVE: exit_loop = compute_set_value_curve(infos); break;
                default: exit_loop = true;
            }
        }
        if (exit_loop) break;
    }
}

}  // namespace wao
