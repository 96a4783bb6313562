// AudioParamProcessor::compute_buffer (src/param.rs:1038-1600) — the per-quantum event state machine of one AudioParam:
// set_value / linear & exponential ramps / setTarget with snap-to-target / value curves / cancel_and_hold over the
// host-prepared timeline.  Shared by the device (k_param: one thread per automated param) and the host (the planner replays
// it up to a suspend point to learn the render-side state new events are inserted against, wae_engine.cu).
#pragma once
#include "../../include/wae.h"
#include "wae_device.h"
#include "wae_spatial.h"  // WAE_HD

#include <cmath>

namespace wae {

struct ParamCursor {
    const ParamInst& p;
    ParamState& s;
    WAE_HD bool empty() const { return s.head >= p.n_events; }
    WAE_HD ParamEvDev peek() const { return s.override_valid ? s.override_ev : p.events[s.head]; }
    WAE_HD bool has_next() const { return s.head + 1 < p.n_events; }
    WAE_HD ParamEvDev next() const { return p.events[s.head + 1]; }
    WAE_HD ParamEvDev pop() {
        ParamEvDev e = peek();
        s.head++;
        s.override_valid = 0;
        return e;
    }
    WAE_HD void replace_peek(const ParamEvDev& e) {
        s.override_ev = e;
        s.override_valid = 1;
    }
};

WAE_HD float par_linear(double t0, double dur, float v0, float diff, double t) { return fmaf(diff, (float)((t - t0) / dur), v0); }
WAE_HD float par_exp(double t0, double dur, float v0, float ratio, double t) { return v0 * powf(ratio, (float)((t - t0) / dur)); }
WAE_HD float par_target(double t0, double tau, float v1, float diff, double t) { return fmaf(diff, (float)exp(-((t - t0) / tau)), v1); }
WAE_HD float par_curve(double t0, double dur, const float* values, int n, double t) {
    if (t - t0 >= dur) return values[n - 1];
    double position = (double)(n - 1) * (t - t0) / dur;
    int k = (int)position;
    float phase = (float)(position - floor(position));
    return fmaf(values[k + 1] - values[k], phase, values[k]);
}
WAE_HD int par_end_index(double end_time, double block_time, double dt, int count) {
    double r = round(fmax(end_time - block_time, 0.) / dt);
    if (!(r < 4.0e9)) return count;
    int idx = (int)r;
    return idx < count ? idx : count;
}


// Fills buf with the intrinsic values of the quantum starting at block_time; returns how many were written: 1 (constant
// / k-rate block) or `count` (128 in the render path; the reference's unit tests use shorter blocks).  Advances the state (events popped, last event, intrinsic value).
WAE_HD int param_compute_buffer(const ParamInst& p, ParamState& st, double block_time, float* buf, const int count = 128) {
    ParamCursor tl{p, st};
    const double dt = 1. / (double)p.sample_rate;
    const double next_block_time = fma(dt, (double)count, block_time);
    int len = 0;
        // ---- compute_buffer (param.rs:1500-1600)
    bool is_constant_block;
    if (tl.empty()) {
        is_constant_block = true;
    } else {
        ParamEvDev e = tl.peek();
        is_constant_block = (e.type != WAE_EVENT_LINEAR_RAMP_TO_VALUE_AT_TIME && e.type != WAE_EVENT_EXPONENTIAL_RAMP_TO_VALUE_AT_TIME)
                                ? e.time >= next_block_time
                                : false;
    }
    if (!p.a_rate || is_constant_block) buf[len++] = st.intrinsic;
    if (!is_constant_block) {
        for (;;) {
            bool exit_loop;
            if (tl.empty()) {
                if (p.a_rate)
                    while (len < count) buf[len++] = st.intrinsic;
                exit_loop = true;
            } else {
                ParamEvDev ev = tl.peek();
                switch (ev.type) {
                    case WAE_EVENT_SET_VALUE:
                    case WAE_EVENT_SET_VALUE_AT_TIME: {  // param.rs:1038-1091
                        double time = ev.time == 0. ? block_time : ev.time;
                        if (p.a_rate) {
                            int e = par_end_index(time, block_time, dt, count);
                            while (len < e) buf[len++] = st.intrinsic;
                        }
                        if (time > next_block_time) {
                            exit_loop = true;
                            break;
                        }
                        st.intrinsic = ev.value;
                        ParamEvDev l = tl.pop();
                        l.time = time;
                        st.last = l;
                        st.has_last = 1;
                        exit_loop = false;
                        break;
                    }
                    case WAE_EVENT_LINEAR_RAMP_TO_VALUE_AT_TIME:
                    case WAE_EVENT_EXPONENTIAL_RAMP_TO_VALUE_AT_TIME: {  // param.rs:1093-1272
                        const bool lin = ev.type == WAE_EVENT_LINEAR_RAMP_TO_VALUE_AT_TIME;
                        const double start_time = st.last.time;
                        double end_time = ev.time;
                        const double duration = end_time - start_time;
                        if (ev.has_cancel) end_time = ev.cancel_time;
                        const float v0 = st.last.value, v1 = ev.value;
                        const float k = lin ? v1 - v0 : v1 / v0;
                        if (!lin && (v0 == 0.f || v0 * v1 < 0.f)) {  // degenerate exponential ramp -> SetValueAtTime
                            ParamEvDev r{};
                            r.type = WAE_EVENT_SET_VALUE_AT_TIME;
                            r.time = end_time;
                            r.value = v1;
                            tl.replace_peek(r);
                            exit_loop = false;
                            break;
                        }
                        if (p.a_rate) {
                            int e = par_end_index(end_time, block_time, dt, count);
                            if (e > len) {
                                double time = fma((double)len, dt, block_time);
                                float value = 0.f;
                                while (len < e) {
                                    value = lin ? par_linear(start_time, duration, v0, k, time) : par_exp(start_time, duration, v0, k, time);
                                    buf[len++] = value;
                                    time += dt;
                                }
                                st.intrinsic = value;
                            }
                        }
                        if (end_time >= next_block_time) {
                            st.intrinsic = lin ? par_linear(start_time, duration, v0, k, next_block_time)
                                               : par_exp(start_time, duration, v0, k, next_block_time);
                            exit_loop = true;
                            break;
                        }
                        if (ev.has_cancel) {
                            float value = lin ? par_linear(start_time, duration, v0, k, end_time) : par_exp(start_time, duration, v0, k, end_time);
                            st.intrinsic = value;
                            ParamEvDev l = tl.pop();
                            l.time = end_time;
                            l.value = value;
                            st.last = l;
                        } else {
                            st.intrinsic = v1;
                            st.last = tl.pop();
                        }
                        st.has_last = 1;
                        exit_loop = false;
                        break;
                    }
                    case WAE_EVENT_SET_TARGET_AT_TIME: {  // param.rs:1274-1427
                        double end_time = next_block_time;
                        bool ended = false;
                        if (tl.has_next()) {
                            ParamEvDev nx = tl.next();
                            if (nx.type == WAE_EVENT_LINEAR_RAMP_TO_VALUE_AT_TIME || nx.type == WAE_EVENT_EXPONENTIAL_RAMP_TO_VALUE_AT_TIME) {
                                end_time = block_time;
                                ended = true;
                            } else if (nx.time < next_block_time) {
                                end_time = nx.time;
                                ended = true;
                            }
                        }
                        if (ev.has_cancel && ev.cancel_time < next_block_time) {
                            end_time = ev.cancel_time;
                            ended = true;
                        }
                        const double start_time = ev.time;
                        const float v0 = st.last.value, v1 = ev.value;
                        const float diff = v0 - v1;
                        const double tau = ev.aux;
                        if (p.a_rate) {
                            int e = par_end_index(end_time, block_time, dt, count);
                            if (e > len) {
                                double time = fma((double)len, dt, block_time);
                                float value = 0.f;
                                while (len < e) {
                                    value = (time - start_time < 0.) ? st.intrinsic : par_target(start_time, tau, v1, diff, time);
                                    buf[len++] = value;
                                    time += dt;
                                }
                                st.intrinsic = value;
                            }
                        }
                        if (!ended) {
                            float value = par_target(start_time, tau, v1, diff, next_block_time);
                            if (fabsf(v1 - value) < 1e-10f) {  // SNAP_TO_TARGET, param.rs:22
                                st.intrinsic = v1;
                                if (v1 == 0.f)
                                    for (int i = 0; i < len; i++)
                                        if (buf[i] != 0.f && fabsf(buf[i]) < 1.17549435e-38f) buf[i] = 0.f;
                                ParamEvDev r{};
                                r.type = WAE_EVENT_SET_VALUE_AT_TIME;
                                r.time = next_block_time;
                                r.value = v1;
                                tl.replace_peek(r);
                            } else {
                                st.intrinsic = value;
                            }
                            exit_loop = true;
                            break;
                        }
                        float value = par_target(start_time, tau, v1, diff, end_time);
                        st.intrinsic = value;
                        ParamEvDev l = tl.pop();
                        l.time = end_time;
                        l.value = value;
                        st.last = l;
                        st.has_last = 1;
                        exit_loop = false;
                        break;
                    }
                    case WAE_EVENT_SET_VALUE_CURVE_AT_TIME: {  // param.rs:1429-1498
                        const double start_time = ev.time, duration = ev.aux;
                        const float* values = p.curves + ev.values_off;
                        const int nv = ev.values_len;
                        double end_time = start_time + duration;
                        if (ev.has_cancel) end_time = ev.cancel_time;
                        if (p.a_rate) {
                            int e = par_end_index(end_time, block_time, dt, count);
                            if (e > len) {
                                double time = fma((double)len, dt, block_time);
                                float value = 0.f;
                                while (len < e) {
                                    value = time < start_time ? st.intrinsic : par_curve(start_time, duration, values, nv, time);
                                    buf[len++] = value;
                                    time += dt;
                                }
                                st.intrinsic = value;
                            }
                        }
                        if (end_time >= next_block_time) {
                            st.intrinsic = par_curve(start_time, duration, values, nv, next_block_time);
                            exit_loop = true;
                            break;
                        }
                        float value = ev.has_cancel ? par_curve(start_time, duration, values, nv, end_time) : values[nv - 1];
                        ParamEvDev l = tl.pop();
                        l.time = end_time;
                        l.value = value;
                        st.intrinsic = value;
                        st.last = l;
                        st.has_last = 1;
                        exit_loop = false;
                        break;
                    }
                    default: exit_loop = true;
                }
            }
            if (exit_loop) break;
        }
    }
    return len;
}

}  // namespace wae
