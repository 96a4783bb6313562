// Host half of AudioParam automation: folds the control-thread event stream of one AudioParam into the sorted event
// timeline the render side starts with — AudioParamProcessor::handle_incoming_event (src/param.rs:799-1036): cancel
// / cancel-and-hold editing, the implicit SetValue before a first ramp or setTarget, set_value updating the intrinsic
// value, stable sort by time.  The per-quantum evaluation of that timeline (compute_buffer, param.rs:1500-1600) runs
// on the GPU (k_param in wae_kernels.cu).
#pragma once
#include "wae_device.h"
#include "wae_graph.h"

#include <algorithm>
#include <string>
#include <vector>

namespace wae {

struct ParamTimeline {
    std::vector<ParamEvDev> events;  // the render side's event queue: sorted, values_off indexes `curves`
    std::vector<float> curves;
    float intrinsic = 0.f;
    bool has_last = false;  // AudioParamProcessor::last_event: the last event the render side popped (none before rendering
    ParamEvDev last{};      // started; after a suspend point: what the replay of the state machine says)
    std::string error;      // the reference's panic text when the event stream is invalid
};

// handle_incoming_event (param.rs:799-1036) for `n` events in arrival order, against the queue / last event / intrinsic
// value in `tl`
inline void fold_param_events(ParamTimeline& tl, const ParamEv* evs, size_t n) {
    std::vector<ParamEvDev>& q = tl.events;
    auto make = [&](const ParamEv& e) {
        ParamEvDev d{};
        d.type = e.type;
        d.value = e.value;
        d.time = e.time;
        d.aux = e.aux;
        d.has_cancel = 0;
        d.cancel_time = 0.;
        d.values_off = 0;
        d.values_len = 0;
        if (!e.values.empty()) {
            d.values_off = (int32_t)tl.curves.size();
            d.values_len = (int32_t)e.values.size();
            tl.curves.insert(tl.curves.end(), e.values.begin(), e.values.end());
        }
        return d;
    };
    auto sort_q = [&]() { std::stable_sort(q.begin(), q.end(), [](const ParamEvDev& a, const ParamEvDev& b) { return a.time < b.time; }); };
    for (size_t ei = 0; ei < n; ei++) {
        const ParamEv& in = evs[ei];
        ParamEvDev ev = make(in);
        if (ev.type == WAE_EVENT_CANCEL_SCHEDULED_VALUES) {  // param.rs:812-866
            // in the middle of a ramp (only possible once rendering has started): the value from before the ramp is restored
            if (!q.empty() && tl.has_last &&
                (q.front().type == WAE_EVENT_LINEAR_RAMP_TO_VALUE_AT_TIME || q.front().type == WAE_EVENT_EXPONENTIAL_RAMP_TO_VALUE_AT_TIME) &&
                q.front().time >= ev.time)
                tl.intrinsic = tl.last.value;
            q.erase(std::remove_if(q.begin(), q.end(), [&](const ParamEvDev& x) { return !(x.time < ev.time); }), q.end());
            continue;
        }
        if (ev.type == WAE_EVENT_CANCEL_AND_HOLD_AT_TIME) {  // param.rs:868-938
            ParamEvDev *e1 = nullptr, *e2 = nullptr;
            double t1 = -1.7976931348623157e308, t2 = 1.7976931348623157e308;
            sort_q();
            for (auto& x : q) {
                if (x.time >= t1 && x.time <= ev.time) {
                    t1 = x.time;
                    e1 = &x;
                } else if (x.time < t2 && x.time > ev.time) {
                    t2 = x.time;
                    e2 = &x;
                }
            }
            if (e2) {
                if (e2->type == WAE_EVENT_LINEAR_RAMP_TO_VALUE_AT_TIME || e2->type == WAE_EVENT_EXPONENTIAL_RAMP_TO_VALUE_AT_TIME) {
                    e2->has_cancel = 1;
                    e2->cancel_time = ev.time;
                }
            } else if (e1) {
                if (e1->type == WAE_EVENT_SET_TARGET_AT_TIME) {
                    e1->has_cancel = 1;
                    e1->cancel_time = ev.time;
                } else if (e1->type == WAE_EVENT_SET_VALUE_CURVE_AT_TIME && ev.time <= e1->time + e1->aux) {
                    e1->has_cancel = 1;
                    e1->cancel_time = ev.time;
                }
            }
            q.erase(std::remove_if(q.begin(), q.end(),
                                   [&](const ParamEvDev& x) {
                                       double t = x.has_cancel ? x.cancel_time : x.time;
                                       return !(t <= ev.time);
                                   }),
                    q.end());
            continue;
        }
        if (ev.type == WAE_EVENT_SET_VALUE_CURVE_AT_TIME) {  // param.rs:950-962
            double t0 = ev.time, t1 = t0 + ev.aux;
            for (auto& x : q)
                if (!(x.time <= t0 || x.time >= t1)) {
                    tl.error = "NotSupportedError - scheduling SetValueCurveAtTime at time of another automation event";
                    return;
                }
        } else {  // param.rs:964-986
            for (auto& x : q)
                if (x.type == WAE_EVENT_SET_VALUE_CURVE_AT_TIME) {
                    double t0 = x.time, t1 = t0 + x.aux;
                    if (!(ev.time <= t0 || ev.time >= t1)) {
                        tl.error = "NotSupportedError - scheduling automation event during SetValueCurveAtTime";
                        return;
                    }
                }
        }
        if (ev.type == WAE_EVENT_SET_VALUE) tl.intrinsic = ev.value;  // param.rs:988-990
        bool ramp = ev.type == WAE_EVENT_LINEAR_RAMP_TO_VALUE_AT_TIME || ev.type == WAE_EVENT_EXPONENTIAL_RAMP_TO_VALUE_AT_TIME;
        if (q.empty() && ((ramp && !tl.has_last) || ev.type == WAE_EVENT_SET_TARGET_AT_TIME)) {  // param.rs:992-1030: implicit SetValue
            ParamEvDev sv{};
            sv.type = WAE_EVENT_SET_VALUE;
            sv.value = tl.intrinsic;
            sv.time = 0.;
            q.push_back(sv);
        }
        // (the queue is sorted at this point: an event that arrives in time order — nearly all do — is in place where it was appended)
        const bool in_order = q.empty() || !(ev.time < q.back().time);
        q.push_back(ev);
        if (!in_order) sort_q();
    }
}

// the timeline a param starts the render with: every control message is drained before the first quantum
inline ParamTimeline build_param_timeline(const Param& p) {
    ParamTimeline tl;
    tl.intrinsic = p.default_value;
    fold_param_events(tl, p.events.data(), p.events.size());
    return tl;
}

}  // namespace wae
