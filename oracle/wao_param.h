// ORACLE — TEST INFRASTRUCTURE ONLY (see wao_core.h).
// AudioParamProcessor: restates src/param.rs:664-1600 (render half of AudioParam).
#pragma once
#include "wao_core.h"

namespace wao {

enum EventType {
    EV_SET_VALUE = 0,
    EV_SET_VALUE_AT_TIME = 1,
    EV_LINEAR_RAMP = 2,
    EV_EXP_RAMP = 3,
    EV_CANCEL_SCHEDULED = 4,
    EV_SET_TARGET = 5,
    EV_CANCEL_AND_HOLD = 6,
    EV_SET_VALUE_CURVE = 7
};

struct ParamEvent {
    int type = EV_SET_VALUE;
    float value = 0.f;
    double time = 0.;
    bool has_time_constant = false;
    double time_constant = 0.;
    bool has_cancel_time = false;
    double cancel_time = 0.;
    bool has_duration = false;
    double duration = 0.;
    std::vector<float> values;
};

struct ParamDescriptor {
    float default_value, min_value, max_value;
    bool a_rate;
};

class ParamProcessor : public Processor {
  public:
    float default_value, min_value, max_value;
    float intrinsic_value;
    bool a_rate;
    float current_value;
    std::vector<ParamEvent> timeline;  // AudioParamEventTimeline (sorted, stable)
    bool has_last_event = false;
    ParamEvent last_event;
    float buffer[RQ];
    int buffer_len = 0;

    explicit ParamProcessor(const ParamDescriptor& d)
        : default_value(d.default_value), min_value(d.min_value), max_value(d.max_value),
          intrinsic_value(d.default_value), a_rate(d.a_rate), current_value(d.default_value) {}

    bool process(std::vector<Quantum>& inputs, std::vector<Quantum>& outputs, const ParamValues&, const Scope& scope) override;
    const char* name() const override { return "AudioParamProcessor"; }

    // returns empty string on success, else the reference's panic message
    std::string handle_incoming_event(ParamEvent ev);
    void compute_buffer(double block_time, double dt, int count);

  private:
    struct BlockInfos {
        double block_time, dt;
        int count;
        bool is_a_rate;
        double next_block_time;
    };
    void push(float v) { buffer[buffer_len++] = v; }
    void sort_timeline();
    void mix_to_output(const Quantum& input, Quantum& output);
    bool compute_set_value(const BlockInfos&);
    bool compute_linear_ramp(const BlockInfos&);
    bool compute_exp_ramp(const BlockInfos&);
    bool compute_set_target(const BlockInfos&);
    bool compute_set_value_curve(const BlockInfos&);
};

}  // namespace wao
