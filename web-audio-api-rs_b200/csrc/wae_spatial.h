// Spatialisation math shared by the planner (static panners, host) and the kernels (moving sources / listener, device):
// src/spatial.rs:205-299 (azimuth / elevation / cone angle; vecmath restated: normalised = v * (1 / len)) and
// PannerRenderer::dist_gain / cone_gain (src/node/panner.rs:927-986).
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define WAE_HD __host__ __device__ inline
#else
#define WAE_HD inline
#endif

namespace wae {
namespace spatial {

constexpr float kPi = 3.14159265358979323846f;

WAE_HD float sq_len(const float a[3]) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }
WAE_HD float dot3(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
WAE_HD void sub3(const float a[3], const float b[3], float o[3]) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
WAE_HD void norm3(const float a[3], float o[3]) {
    float inv = 1.f / sqrtf(sq_len(a));
    o[0] = a[0] * inv; o[1] = a[1] * inv; o[2] = a[2] * inv;
}
WAE_HD void cross3(const float a[3], const float b[3], float o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
// spatial.rs:205-268
WAE_HD void azimuth_elevation(const float sp[3], const float lp[3], const float lf[3], const float lu[3], float& az, float& el) {
    const float MINP = 1.17549435e-38f;
    az = 0.f;
    el = 0.f;
    float rel[3];
    sub3(sp, lp, rel);
    if (sq_len(rel) <= MINP) return;
    float sl[3], right[3];
    norm3(rel, sl);
    cross3(lf, lu, right);
    if (sq_len(right) == 0.f) return;
    float rn[3], fn[3], up[3];
    norm3(right, rn);
    norm3(lf, fn);
    cross3(rn, fn, up);
    float elevation = 90.f - 180.f * acosf(dot3(sl, up)) / kPi;
    if (elevation > 90.f) elevation = 180.f - elevation;
    else if (elevation < -90.f) elevation = -180.f - elevation;
    float upp = dot3(sl, up);
    float proj[3] = {sl[0] - up[0] * upp, sl[1] - up[1] * upp, sl[2] - up[2] * upp};
    if (sq_len(proj) == 0.f) {
        el = elevation;
        return;
    }
    float pn[3];
    norm3(proj, pn);
    float azimuth = 180.f * acosf(dot3(pn, rn)) / kPi;
    if (dot3(pn, fn) < 0.f) azimuth = 360.f - azimuth;
    azimuth = (azimuth >= 0.f && azimuth <= 270.f) ? 90.f - azimuth : 450.f - azimuth;
    az = azimuth;
    el = elevation;
}
// spatial.rs:283-299
WAE_HD float cone_angle(const float sp[3], const float so[3], const float lp[3]) {
    const float MINP = 1.17549435e-38f;
    if (sq_len(so) == 0.f) return 0.f;
    float nso[3], rel[3], sl[3];
    norm3(so, nso);
    sub3(sp, lp, rel);
    if (sq_len(rel) <= MINP) return 0.f;
    norm3(rel, sl);
    return fabsf(180.f * acosf(dot3(sl, nso)) / kPi);
}

// the PannerNode attributes the renderer reads (panner.rs:670-683)
struct PanModel {
    int32_t distance_model;  // 0 linear, 1 inverse, 2 exponential
    int32_t pad;
    double ref_distance, max_distance, rolloff_factor;
    double cone_inner_angle, cone_outer_angle, cone_outer_gain;
};
struct SpatialParams {
    float dist_gain, cone_gain, azimuth, elevation;
};

// panner.rs:954-986
WAE_HD float dist_gain(const PanModel& m, const float sp[3], const float lp[3]) {
    float rel[3];
    sub3(sp, lp, rel);
    const double distance = (double)sqrtf(sq_len(rel));
    double gd;
    if (m.distance_model == 0) {
        const double ro = fmin(fmax(m.rolloff_factor, 0.), 1.);
        const double d2ref = fmin(m.ref_distance, m.max_distance), d2max = fmax(m.ref_distance, m.max_distance);
        const double dc = distance < d2ref ? d2ref : (distance > d2max ? d2max : distance);
        gd = 1. - ro * (dc - d2ref) / (d2max - d2ref);
    } else if (m.distance_model == 1) {
        const double ro = fmax(m.rolloff_factor, 0.);
        gd = distance > 0. ? m.ref_distance / (m.ref_distance + ro * (fmax(m.ref_distance, distance) - m.ref_distance)) : 1.;
    } else {
        const double ro = fmax(m.rolloff_factor, 0.);
        gd = pow(fmax(distance, m.ref_distance) / m.ref_distance, -ro);
    }
    return (float)gd;
}
// panner.rs:927-952
WAE_HD float cone_gain(const PanModel& m, const float sp[3], const float so[3], const float lp[3]) {
    const float inner = (float)fabs(m.cone_inner_angle) / 2.f, outer = (float)fabs(m.cone_outer_angle) / 2.f;
    if (inner >= 180.f && outer >= 180.f) return 1.f;
    const float og = (float)m.cone_outer_gain;
    const float a = cone_angle(sp, so, lp);
    if (a < inner) return 1.f;
    if (a >= outer) return og;
    const float x = (a - inner) / (outer - inner);
    return (1.f - x) + og * x;
}
// v = source position, source orientation, listener position / forward / up (the 15 a-rate params of panner.rs:714-780)
WAE_HD SpatialParams spatial_params(const PanModel& m, const float v[15]) {
    SpatialParams p;
    p.dist_gain = dist_gain(m, v, v + 6);
    p.cone_gain = cone_gain(m, v, v + 3, v + 6);
    azimuth_elevation(v, v + 6, v + 9, v + 12, p.azimuth, p.elevation);
    return p;
}
// direction of the source in the listener frame (panner.rs:792-802)
WAE_HD void projected_source(const SpatialParams& p, float out[3]) {
    const float az = p.azimuth * kPi / 180.f, el = p.elevation * kPi / 180.f;
    float x = sinf(az) * cosf(el), z = cosf(az) * cosf(el), y = sinf(el);
    if (fabsf(x) <= 1e-6f && fabsf(y) <= 1e-6f && fabsf(z) <= 1e-6f) x = 0.f, y = 0.f, z = 1.f;
    out[0] = x, out[1] = y, out[2] = z;
}

// Sphere triangle crossed by the segment 0 -> 10 d and the barycentric weights of the hit point (the hrtf crate's
// ray / triangle query + get_barycentric_coords); hits on an edge or vertex resolve to the face where the hit is most
// interior.  pos: [vertex][3], tri: 3 indices per face.
WAE_HD bool hrir_locate(const float* pos, const uint32_t* tri, int n_faces, const float d[3], uint32_t v[3], float w[3]) {
    const float r[3] = {d[0] * 10.f, d[1] * 10.f, d[2] * 10.f};
    bool any = false;
    float best = -3.0e38f;
    for (int f = 0; f < n_faces; f++) {
        const float *A = pos + 3 * tri[3 * f], *B = pos + 3 * tri[3 * f + 1], *C = pos + 3 * tri[3 * f + 2];
        const float e0[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
        const float e1[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
        float n[3];
        cross3(e0, e1, n);
        const float plane_d = -dot3(A, n);
        const float rn = dot3(r, n);
        if (rn == 0.f) continue;
        const float t = -plane_d / rn;
        if (!(t >= 0.f && t <= 1.f)) continue;
        const float q[3] = {r[0] * t - A[0], r[1] * t - A[1], r[2] * t - A[2]};
        const float d00 = dot3(e0, e0), d01 = dot3(e0, e1), d11 = dot3(e1, e1), d20 = dot3(q, e0), d21 = dot3(q, e1);
        const float den = d00 * d11 - d01 * d01;
        if (den == 0.f) continue;
        const float wb = (d11 * d20 - d01 * d21) / den;
        const float wc = (d00 * d21 - d01 * d20) / den;
        const float wa = 1.f - wb - wc;
        const float inside = fminf(wa, fminf(wb, wc));
        if (inside > best) {
            best = inside;
            any = true;
            v[0] = tri[3 * f], v[1] = tri[3 * f + 1], v[2] = tri[3 * f + 2];
            w[0] = wa, w[1] = wb, w[2] = wc;
        }
    }
    return any;
}

}  // namespace spatial
}  // namespace wae
