// ORACLE — TEST INFRASTRUCTURE ONLY (see wao_core.h).
// AudioListener (src/spatial.rs), PannerRenderer (src/node/panner.rs:640-1057) and spatial helpers.
#pragma once
#include "wao_core.h"
#include "wao_fft.h"
#include <string>

namespace wao {

// spatial.rs:173-185: the listener renderer does nothing, it only orders its params
struct ListenerRenderer : Processor {
    uint32_t params[9];
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override { return true; }
    const char* name() const override { return "ListenerRenderer"; }
};

// spatial.rs:205-299
void azimuth_and_elevation(const float sp[3], const float lp[3], const float lf[3], const float lu[3], float& az, float& el);
float spatial_distance(const float sp[3], const float lp[3]);
float spatial_angle(const float sp[3], const float so[3], const float lp[3]);

bool hrtf_sphere_available(std::string& why);
bool hrtf_set_sphere(const void* data, uint64_t len, std::string& err);
bool hrtf_locate(const float* pos, const uint32_t* faces, size_t n_faces, const float dir[3], uint32_t idx[3], float k[3]);

struct HrtfState;  // defined in wao_hrtf.cpp

struct PannerRenderer : Processor {
    uint32_t position_x = 0, position_y = 0, position_z = 0, orientation_x = 0, orientation_y = 0, orientation_z = 0;
    int distance_model = 1;
    double ref_distance = 1., max_distance = 10000., rolloff_factor = 1.;
    double cone_inner_angle = 360., cone_outer_angle = 360., cone_outer_gain = 0.;
    std::shared_ptr<HrtfState> hrtf_state;
    size_t tail_time_counter = 0;
    void set_hrtf(float sample_rate);
    bool process(std::vector<Quantum>&, std::vector<Quantum>&, const ParamValues&, const Scope&) override;
    const char* name() const override { return "PannerRenderer"; }
    float cone_gain(const float sp[3], const float so[3], const float lp[3]) const;  // panner.rs:927-952
    float dist_gain(const float sp[3], const float lp[3]) const;                     // panner.rs:954-986
};

// HRTF processing hooks (wao_hrtf.cpp)
size_t hrtf_tail_time_samples(const HrtfState&);
// source: 128 mono samples; out: 128 interleaved (l, r) pairs
void hrtf_process(HrtfState&, const float* source, float new_distance_gain, const float projected_source[3], float* out_lr);
std::shared_ptr<HrtfState> hrtf_state_new(float sample_rate);
// one impulse response resampled as HrirSphere::new does for a context rate other than the data's (ratio = context / data)
std::vector<float> resample_hrir(const std::vector<float>& hrir, double ratio);

}  // namespace wao
