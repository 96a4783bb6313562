// ORACLE — TEST INFRASTRUCTURE ONLY (see wao_core.h).
// `vecmath` (Cargo.toml, not vendored) 3-vector helpers restated: square_len = x*x+y*y+z*z, len = sqrt,
// normalized = v * (1/len), dot, cross — trivial closed forms.
#include "wao_panner.h"

namespace wao {

static const float PI32 = 3.14159265358979323846f;
static const float F32_MIN_POSITIVE = 1.17549435e-38f;

static inline float sq_len(const float a[3]) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }
static inline float dot(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void sub(const float a[3], const float b[3], float o[3]) {
    o[0] = a[0] - b[0];
    o[1] = a[1] - b[1];
    o[2] = a[2] - b[2];
}
static inline void normalized(const float a[3], float o[3]) {
    float inv = 1.f / std::sqrt(sq_len(a));
    o[0] = a[0] * inv;
    o[1] = a[1] * inv;
    o[2] = a[2] * inv;
}
static inline void cross(const float a[3], const float b[3], float o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// spatial.rs:205-270
void azimuth_and_elevation(const float sp[3], const float lp[3], const float lf[3], const float lu[3], float& az, float& el) {
    float relative_pos[3];
    sub(sp, lp, relative_pos);
    if (sq_len(relative_pos) <= F32_MIN_POSITIVE) {
        az = 0.f;
        el = 0.f;
        return;
    }
    float source_listener[3];
    normalized(relative_pos, source_listener);
    float listener_right[3];
    cross(lf, lu, listener_right);
    if (sq_len(listener_right) == 0.f) {
        az = 0.f;
        el = 0.f;
        return;
    }
    float right_norm[3], forward_norm[3], up[3];
    normalized(listener_right, right_norm);
    normalized(lf, forward_norm);
    cross(right_norm, forward_norm, up);
    float elevation = 90.f - 180.f * acosf(dot(source_listener, up)) / PI32;
    if (elevation > 90.f)
        elevation = 180.f - elevation;
    else if (elevation < -90.f)
        elevation = -180.f - elevation;
    float up_projection = dot(source_listener, up);
    float projected[3] = {source_listener[0] - up[0] * up_projection, source_listener[1] - up[1] * up_projection,
                          source_listener[2] - up[2] * up_projection};
    if (sq_len(projected) == 0.f) {
        az = 0.f;
        el = elevation;
        return;
    }
    float pn[3];
    normalized(projected, pn);
    float azimuth = 180.f * acosf(dot(pn, right_norm)) / PI32;
    float front_back = dot(pn, forward_norm);
    if (front_back < 0.f) azimuth = 360.f - azimuth;
    if (azimuth >= 0.f && azimuth <= 270.f)
        azimuth = 90.f - azimuth;
    else
        azimuth = 450.f - azimuth;
    az = azimuth;
    el = elevation;
}

// spatial.rs:272-274
float spatial_distance(const float sp[3], const float lp[3]) {
    float d[3];
    sub(sp, lp, d);
    return std::sqrt(sq_len(d));
}

// spatial.rs:276-299
float spatial_angle(const float sp[3], const float so[3], const float lp[3]) {
    if (sq_len(so) == 0.f) return 0.f;
    float nso[3];
    normalized(so, nso);
    float rel[3];
    sub(sp, lp, rel);
    if (sq_len(rel) <= F32_MIN_POSITIVE) return 0.f;
    float sl[3];
    normalized(rel, sl);
    float angle = 180.f * acosf(dot(sl, nso)) / PI32;
    return std::fabs(angle);
}

// panner.rs:927-952
float PannerRenderer::cone_gain(const float sp[3], const float so[3], const float lp[3]) const {
    float abs_inner_angle = (float)std::fabs(cone_inner_angle) / 2.f;
    float abs_outer_angle = (float)std::fabs(cone_outer_angle) / 2.f;
    if (abs_inner_angle >= 180.f && abs_outer_angle >= 180.f) return 1.f;
    float outer_gain = (float)cone_outer_gain;
    float abs_angle = spatial_angle(sp, so, lp);
    if (abs_angle < abs_inner_angle) return 1.f;
    if (abs_angle >= abs_outer_angle) return outer_gain;
    float x = (abs_angle - abs_inner_angle) / (abs_outer_angle - abs_inner_angle);
    return (1.f - x) + outer_gain * x;
}

static inline double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// panner.rs:954-986
float PannerRenderer::dist_gain(const float sp[3], const float lp[3]) const {
    double distance = (double)spatial_distance(sp, lp);
    double g;
    switch (distance_model) {
        case 0: {
            double rolloff = clampd(rolloff_factor, 0., 1.);
            double d2ref = std::min(ref_distance, max_distance);
            double d2max = std::max(ref_distance, max_distance);
            double d_clamped = clampd(distance, d2ref, d2max);
            g = 1. - rolloff * (d_clamped - d2ref) / (d2max - d2ref);
            break;
        }
        case 1: {
            double rolloff = std::max(rolloff_factor, 0.);
            if (distance > 0.)
                g = ref_distance / (ref_distance + rolloff * (std::max(ref_distance, distance) - ref_distance));
            else
                g = 1.;
            break;
        }
        default: {
            double rolloff = std::max(rolloff_factor, 0.);
            g = std::pow(std::max(distance, ref_distance) / ref_distance, -rolloff);
        }
    }
    return (float)g;
}

struct SpatialParams {
    float dist_gain, cone_gain, azimuth, elevation;
};

static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// panner.rs:988-1017
static inline void apply_mono_to_stereo_gain(const SpatialParams& p, float& l, float& r) {
    float azimuth = clampf(p.azimuth, -180.f, 180.f);
    if (azimuth < -90.f)
        azimuth = -180.f - azimuth;
    else if (azimuth > 90.f)
        azimuth = 180.f - azimuth;
    float x = (azimuth + 90.f) / 180.f;
    float gain_l = cosf(x * PI32 / 2.f);
    float gain_r = sinf(x * PI32 / 2.f);
    l *= gain_l * p.dist_gain * p.cone_gain;
    r *= gain_r * p.dist_gain * p.cone_gain;
}

// panner.rs:1019-1057
static inline void apply_stereo_to_stereo_gain(const SpatialParams& p, float il, float ir, float& ol, float& orr) {
    float azimuth = clampf(p.azimuth, -180.f, 180.f);
    if (azimuth < -90.f)
        azimuth = -180.f - azimuth;
    else if (azimuth > 90.f)
        azimuth = 180.f - azimuth;
    float x = azimuth <= 0.f ? (azimuth + 90.f) / 90.f : azimuth / 90.f;
    float gain_l = cosf(x * PI32 / 2.f);
    float gain_r = sinf(x * PI32 / 2.f);
    if (azimuth <= 0.f) {
        ol = (il + ir * gain_l) * p.dist_gain * p.cone_gain;
        orr = ir * gain_r * p.dist_gain * p.cone_gain;
    } else {
        ol = il * gain_l * p.dist_gain * p.cone_gain;
        orr = (ir + il * gain_r) * p.dist_gain * p.cone_gain;
    }
}

void PannerRenderer::set_hrtf(float sample_rate) { hrtf_state = hrtf_state_new(sample_rate); }

// panner.rs:685-904
bool PannerRenderer::process(std::vector<Quantum>& inputs, std::vector<Quantum>& outputs, const ParamValues& params, const Scope&) {
    const Quantum& input = inputs[0];
    Quantum& output = outputs[0];
    if (input.is_silent()) {
        bool tail_time = hrtf_state ? hrtf_tail_time_samples(*hrtf_state) > tail_time_counter : false;
        if (!tail_time) {
            output.make_silent();
            return false;
        }
        tail_time_counter += RQ;
    }
    ParamSlice v[15] = {params.get(position_x), params.get(position_y), params.get(position_z),
                        params.get(orientation_x), params.get(orientation_y), params.get(orientation_z),
                        params.get(2), params.get(3), params.get(4), params.get(5), params.get(6), params.get(7),
                        params.get(8), params.get(9), params.get(10)};
    auto spatial_at = [&](int i) {
        float sp[3] = {v[0][i % v[0].len], v[1][i % v[1].len], v[2][i % v[2].len]};
        float so[3] = {v[3][i % v[3].len], v[4][i % v[4].len], v[5][i % v[5].len]};
        float lp[3] = {v[6][i % v[6].len], v[7][i % v[7].len], v[8][i % v[8].len]};
        float lf[3] = {v[9][i % v[9].len], v[10][i % v[10].len], v[11][i % v[11].len]};
        float lu[3] = {v[12][i % v[12].len], v[13][i % v[13].len], v[14][i % v[14].len]};
        SpatialParams p;
        p.dist_gain = dist_gain(sp, lp);
        p.cone_gain = cone_gain(sp, so, lp);
        azimuth_and_elevation(sp, lp, lf, lu, p.azimuth, p.elevation);
        return p;
    };
    if (hrtf_state) {
        SpatialParams p = spatial_at(0);
        float new_distance_gain = p.cone_gain * p.dist_gain;
        float az_rad = p.azimuth * PI32 / 180.f;
        float el_rad = p.elevation * PI32 / 180.f;
        float x = sinf(az_rad) * cosf(el_rad);
        float z = cosf(az_rad) * cosf(el_rad);
        float y = sinf(el_rad);
        float projected_source[3] = {x, y, z};
        if (std::fabs(x) <= 1e-6f && std::fabs(y) <= 1e-6f && std::fabs(z) <= 1e-6f) {
            projected_source[0] = 0.f;
            projected_source[1] = 0.f;
            projected_source[2] = 1.f;
        }
        output = input;
        float overall_gain_correction = 1.f;
        if (output.number_of_channels() == 2) {
            overall_gain_correction *= 2.f;
            output.mix(1, SPEAKERS);
        }
        float lr[2 * RQ];
        Channel src = output.channel(0);
        hrtf_process(*hrtf_state, src.data(), new_distance_gain, projected_source, lr);
        output.set_number_of_channels(2);
        float* left = output.channel_mut(0).make_mut();
        float* right = output.channel_mut(1).make_mut();
        for (int i = 0; i < RQ; i++) {
            left[i] = overall_gain_correction * lr[2 * i];
            right[i] = overall_gain_correction * lr[2 * i + 1];
        }
    } else {
        bool single_valued = true;
        for (int k = 6; k < 15; k++)
            if (v[k].len != 1) single_valued = false;
        if (input.number_of_channels() == 1) {
            output = input;
            output.mix(2, SPEAKERS);
            float* left = output.channel_mut(0).make_mut();
            float* right = output.channel_mut(1).make_mut();
            if (single_valued) {
                SpatialParams p = spatial_at(0);
                for (int i = 0; i < RQ; i++) apply_mono_to_stereo_gain(p, left[i], right[i]);
            } else {
                for (int i = 0; i < RQ; i++) {
                    SpatialParams p = spatial_at(i);
                    apply_mono_to_stereo_gain(p, left[i], right[i]);
                }
            }
        } else {
            Channel il = input.channel(0), ir = input.channel(1);
            output.set_number_of_channels(2);
            float* left = output.channel_mut(0).make_mut();
            float* right = output.channel_mut(1).make_mut();
            if (single_valued) {
                SpatialParams p = spatial_at(0);
                for (int i = 0; i < RQ; i++) apply_stereo_to_stereo_gain(p, il.data()[i], ir.data()[i], left[i], right[i]);
            } else {
                for (int i = 0; i < RQ; i++) {
                    SpatialParams p = spatial_at(i);
                    apply_stereo_to_stereo_gain(p, il.data()[i], ir.data()[i], left[i], right[i]);
                }
            }
        }
    }
    return hrtf_state != nullptr;
}

}  // namespace wao
