// The AudioParam state machine of wae_param_core.h with the per-frame fill loops turned into calls on a sink.
//
// Why: in every branch of AudioParamProcessor::compute_buffer (src/param.rs:1093-1498) the intrinsic value left behind is a closed form
// of the event (its value at next_block_time / end_time / the target) and never the last frame the fill loop produced, so the walk
// over the events is cheap and serial while the 128 frame values of a quantum are independent of each other.  A sink that records the
// fills lets a warp evaluate them in parallel (k_param_parallel, the default: WAE_OPT_PARAM_PARALLEL = 1); a sink that evaluates them on
// the spot reproduces param_compute_buffer.  Frame times are the reference's running sum (`time += dt` from the fill's first frame), so
// a lane that starts at frame i re-accumulates i - first additions and gets bit-identical times.
//
// STATUS: the walker and both sinks are exercised on the host by the reference's param.rs tests (tests/test_param_timeline.py, third
// implementation "engine-walk"); the kernel that uses the recording sink is the default and is compared bit for bit with k_param on
// hardware (tests/test_gpu_criterion_and_setters.py).  param_compute_buffer stays the code k_param (option = 0) and the suspend replay run.
#pragma once
#include "wae_param_core.h"

#include <cstring>

namespace wae {

enum { PF_CONST = 0, PF_LINEAR = 1, PF_EXP = 2, PF_TARGET = 3, PF_CURVE = 4 };
struct ParamFill {       // frames [first, last) of the quantum
    int first, last, kind, n;
    float a, b, pre;     // ramps: v0, k | target: v1, diff, value before the start time | curve: pre = value before the start time
    double t0, d;        // ramps: start time, duration | target: start time, time constant | curve: start time, duration
    double time0;        // time of frame `first`
    const float* values; // curve
};

WAE_HD float param_fill_at(const ParamFill& f, double time) {
    switch (f.kind) {
        case PF_LINEAR: return par_linear(f.t0, f.d, f.a, f.b, time);
        case PF_EXP: return par_exp(f.t0, f.d, f.a, f.b, time);
        case PF_TARGET: return (time - f.t0 < 0.) ? f.pre : par_target(f.t0, f.d, f.a, f.b, time);
        default: return time < f.t0 ? f.pre : par_curve(f.t0, f.d, f.values, f.n, time);
    }
}
// frames [from, to) of a fill, from >= f.first: the time of `from` is re-accumulated from the fill's first frame
WAE_HD void param_fill_range(const ParamFill& f, int from, int to, double dt, float* buf) {
    double time = f.time0;
    for (int i = f.first; i < from; i++) time += dt;
    for (int i = from; i < to; i++) {
        buf[i] = param_fill_at(f, time);
        time += dt;
    }
}

// evaluates every fill on the spot: param_walk<SerialSink> == param_compute_buffer
struct SerialSink {
    float* buf;
    double dt;
    WAE_HD void constant(int first, int last, float v) {
        for (int i = first; i < last; i++) buf[i] = v;
    }
    WAE_HD void emit(const ParamFill& f) { param_fill_range(f, f.first, f.last, dt, buf); }
    WAE_HD void ramp(int first, int last, bool lin, double t0, double dur, float v0, float k, double time0, double) {
        emit(ParamFill{first, last, lin ? PF_LINEAR : PF_EXP, 0, v0, k, 0.f, t0, dur, time0, nullptr});
    }
    WAE_HD void target(int first, int last, double t0, double tau, float v1, float diff, float pre, double time0, double) {
        emit(ParamFill{first, last, PF_TARGET, 0, v1, diff, pre, t0, tau, time0, nullptr});
    }
    WAE_HD void curve(int first, int last, double t0, double dur, const float* values, int n, float pre, double time0, double) {
        emit(ParamFill{first, last, PF_CURVE, n, 0.f, 0.f, pre, t0, dur, time0, values});
    }
    WAE_HD void flush_subnormals(int len) {
        for (int i = 0; i < len; i++)
            if (buf[i] != 0.f && fabsf(buf[i]) < 1.17549435e-38f) buf[i] = 0.f;
    }
};

// writes the constant fills, records the computed ones (a quantum rarely holds more than two); when the record is full the fill is
// evaluated on the spot like SerialSink does.  After the walk: evaluate fills[0..n) in parallel, then apply `flush` to buf[0..flush).
struct RecordSink {
    static constexpr int kMax = 4;
    float* buf;
    double dt;
    ParamFill fills[kMax];
    int n = 0;
    int flush = 0;
    WAE_HD void constant(int first, int last, float v) {
        for (int i = first; i < last; i++) buf[i] = v;
    }
    WAE_HD void emit(const ParamFill& f) {
        if (n < kMax) fills[n++] = f;
        else param_fill_range(f, f.first, f.last, dt, buf);
    }
    WAE_HD void ramp(int first, int last, bool lin, double t0, double dur, float v0, float k, double time0, double) {
        emit(ParamFill{first, last, lin ? PF_LINEAR : PF_EXP, 0, v0, k, 0.f, t0, dur, time0, nullptr});
    }
    WAE_HD void target(int first, int last, double t0, double tau, float v1, float diff, float pre, double time0, double) {
        emit(ParamFill{first, last, PF_TARGET, 0, v1, diff, pre, t0, tau, time0, nullptr});
    }
    WAE_HD void curve(int first, int last, double t0, double dur, const float* values, int n_values, float pre, double time0, double) {
        emit(ParamFill{first, last, PF_CURVE, n_values, 0.f, 0.f, pre, t0, dur, time0, values});
    }
    WAE_HD void flush_subnormals(int len) { flush = len > flush ? len : flush; }
    // host-side completion (the kernel spreads this over the lanes)
    WAE_HD void finish() {
        for (int k = 0; k < n; k++) param_fill_range(fills[k], fills[k].first, fills[k].last, dt, buf);
        for (int i = 0; i < flush; i++)
            if (buf[i] != 0.f && fabsf(buf[i]) < 1.17549435e-38f) buf[i] = 0.f;
    }
};

template <class Sink>
WAE_HD int param_walk(const ParamInst& p, ParamState& st, double block_time, Sink& sink, const int count = 128) {
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
    if (!p.a_rate || is_constant_block) {
        sink.constant(len, len + 1, st.intrinsic);
        len++;
    }
    if (!is_constant_block) {
        for (;;) {
            bool exit_loop;
            if (tl.empty()) {
                if (p.a_rate && len < count) {
                    sink.constant(len, count, st.intrinsic);
                    len = count;
                }
                exit_loop = true;
            } else {
                ParamEvDev ev = tl.peek();
                switch (ev.type) {
                    case WAE_EVENT_SET_VALUE:
                    case WAE_EVENT_SET_VALUE_AT_TIME: {  // param.rs:1038-1091
                        double time = ev.time == 0. ? block_time : ev.time;
                        if (p.a_rate) {
                            int e = par_end_index(time, block_time, dt, count);
                            if (e > len) {
                                sink.constant(len, e, st.intrinsic);
                                len = e;
                            }
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
                                // (the value the loop of compute_buffer leaves in `intrinsic_value` is overwritten below in every path)
                                sink.ramp(len, e, lin, start_time, duration, v0, k, fma((double)len, dt, block_time), dt);
                                len = e;
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
                                sink.target(len, e, start_time, tau, v1, diff, st.intrinsic, fma((double)len, dt, block_time), dt);
                                len = e;
                            }
                        }
                        if (!ended) {
                            float value = par_target(start_time, tau, v1, diff, next_block_time);
                            if (fabsf(v1 - value) < 1e-10f) {  // SNAP_TO_TARGET, param.rs:22
                                st.intrinsic = v1;
                                if (v1 == 0.f) sink.flush_subnormals(len);
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
                                sink.curve(len, e, start_time, duration, values, nv, st.intrinsic, fma((double)len, dt, block_time), dt);
                                len = e;
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


// ---- speculation (k_param_spec): the state a quantum is walked from, predicted without walking the quanta before it ------------------
WAE_HD unsigned par_bits(float v) {
#ifdef __CUDA_ARCH__
    return (unsigned)__float_as_int(v);
#else
    unsigned u;
    memcpy(&u, &v, sizeof u);
    return u;
#endif
}
// The intrinsic value the walk leaves behind after a quantum that ended at `nbt` while the event at the head of the queue stayed there:
// the closed forms param_walk assigns at next_block_time, or nothing in a constant block.  A wrong guess is harmless: the caller compares
// the predicted state with the one the walk of the previous quantum really left (param_state_equal) and drops what was built on it.
WAE_HD float param_predict_intrinsic(const ParamInst& p, ParamState& st, double nbt) {
    ParamCursor tl{p, st};
    if (tl.empty()) return st.intrinsic;
    const ParamEvDev ev = tl.peek();
    const bool lin = ev.type == WAE_EVENT_LINEAR_RAMP_TO_VALUE_AT_TIME;
    if (lin || ev.type == WAE_EVENT_EXPONENTIAL_RAMP_TO_VALUE_AT_TIME) {
        const double start = st.last.time, duration = ev.time - start;
        const float v0 = st.last.value, v1 = ev.value;
        if (!lin && (v0 == 0.f || v0 * v1 < 0.f)) return st.intrinsic;
        return lin ? par_linear(start, duration, v0, v1 - v0, nbt) : par_exp(start, duration, v0, v1 / v0, nbt);
    }
    if (ev.time >= nbt) return st.intrinsic;  // the quantum that ended at nbt was a constant block
    if (ev.type == WAE_EVENT_SET_TARGET_AT_TIME) return par_target(ev.time, ev.aux, ev.value, st.last.value - ev.value, nbt);
    if (ev.type == WAE_EVENT_SET_VALUE_CURVE_AT_TIME) return par_curve(ev.time, ev.aux, p.curves + ev.values_off, ev.values_len, nbt);
    return st.intrinsic;
}
WAE_HD bool param_ev_equal(const ParamEvDev& a, const ParamEvDev& b) {
    return a.type == b.type && par_bits(a.value) == par_bits(b.value) && a.time == b.time && a.aux == b.aux &&
           a.cancel_time == b.cancel_time && a.has_cancel == b.has_cancel && a.values_off == b.values_off && a.values_len == b.values_len;
}
// did the walk of the previous quantum leave exactly the state this quantum was walked from?
WAE_HD bool param_state_equal(const ParamState& a, const ParamState& b) {
    if (par_bits(a.intrinsic) != par_bits(b.intrinsic) || a.head != b.head || a.has_last != b.has_last || a.override_valid != b.override_valid)
        return false;
    if (a.has_last && !param_ev_equal(a.last, b.last)) return false;
    if (a.override_valid && !param_ev_equal(a.override_ev, b.override_ev)) return false;
    return true;
}

}  // namespace wae
