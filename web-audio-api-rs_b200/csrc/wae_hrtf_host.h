// Host half of HRTF panning (PanningModelType::HRTF, src/node/panner.rs:39-68,215-271): the HRIR sphere container the
// reference embeds (resources/IRC_1003_C.bin, handed over by the caller through wae_engine_set_hrir_sphere), and the
// per-direction blend of the three vertex responses of the sphere triangle the source direction crosses — the job of the
// third-party `hrtf` 0.8.1 crate's HrirSphere / sample_bilinear.  The convolution itself runs on the GPU (k_hrtf_fir).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace wae {

struct HrirSphere {
    uint32_t sample_rate = 0;
    uint32_t taps = 0;  // samples per impulse response
    struct Vertex {
        float p[3];
        uint32_t left, right;  // offsets into `ir`
    };
    std::vector<Vertex> vertices;
    std::vector<uint32_t> tri;  // 3 vertex indices per face
    std::vector<float> ir;

    // layout: "HRIR" | u32 rate | u32 taps | u32 #vertices | u32 #indices | indices | {xyz, left[taps], right[taps]} per vertex
    bool parse(const uint8_t* bytes, uint64_t size, std::string& err) {
        struct Reader {
            const uint8_t* b;
            uint64_t n, at = 0;
            bool take(void* dst, uint64_t k) {
                if (at + k > n) return false;
                std::memcpy(dst, b + at, k);
                at += k;
                return true;
            }
        } rd{bytes, size};
        char magic[4];
        uint32_t nv = 0, ni = 0;
        if (!bytes || !rd.take(magic, 4) || std::memcmp(magic, "HRIR", 4) != 0) return err = "HRIR sphere: bad magic", false;
        if (!rd.take(&sample_rate, 4) || !rd.take(&taps, 4) || !rd.take(&nv, 4) || !rd.take(&ni, 4)) return err = "HRIR sphere: truncated header", false;
        if (taps == 0 || taps > 4096 || nv == 0 || ni == 0 || ni % 3) return err = "HRIR sphere: unsupported header values", false;
        tri.resize(ni);
        if (!rd.take(tri.data(), 4ull * ni)) return err = "HRIR sphere: truncated index table", false;
        for (uint32_t t : tri)
            if (t >= nv) return err = "HRIR sphere: face index out of range", false;
        vertices.resize(nv);
        ir.resize(2ull * nv * taps);
        for (uint32_t v = 0; v < nv; v++) {
            Vertex& vx = vertices[v];
            vx.left = 2u * v * taps;
            vx.right = vx.left + taps;
            if (!rd.take(vx.p, 12) || !rd.take(&ir[vx.left], 4ull * taps) || !rd.take(&ir[vx.right], 4ull * taps))
                return err = "HRIR sphere: truncated vertex data", false;
        }
        return true;
    }

    // Blend weights of direction d (sphere coordinates): the face whose plane the segment 0 -> 10 d crosses inside the
    // triangle; hits on an edge / vertex resolve to the face where the hit is most interior.
    bool locate(const float d[3], uint32_t v[3], float w[3]) const {
        const float r[3] = {d[0] * 10.f, d[1] * 10.f, d[2] * 10.f};
        auto dot = [](const float* x, const float* y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; };
        bool any = false;
        float best = -3.0e38f;
        for (size_t f = 0; f + 2 < tri.size(); f += 3) {
            const float *A = vertices[tri[f]].p, *B = vertices[tri[f + 1]].p, *C = vertices[tri[f + 2]].p;
            const float e0[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
            const float e1[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
            const float n[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
            const float plane_d = -dot(A, n);
            const float rn = dot(r, n);
            if (rn == 0.f) continue;
            const float t = -plane_d / rn;
            if (!(t >= 0.f && t <= 1.f)) continue;
            const float hit[3] = {r[0] * t, r[1] * t, r[2] * t};
            const float q[3] = {hit[0] - A[0], hit[1] - A[1], hit[2] - A[2]};
            const float d00 = dot(e0, e0), d01 = dot(e0, e1), d11 = dot(e1, e1), d20 = dot(q, e0), d21 = dot(q, e1);
            const float den = d00 * d11 - d01 * d01;
            if (den == 0.f) continue;
            const float wb = (d11 * d20 - d01 * d21) / den;
            const float wc = (d00 * d21 - d01 * d20) / den;
            const float wa = 1.f - wb - wc;
            const float inside = std::fmin(wa, std::fmin(wb, wc));
            if (inside > best) {
                best = inside;
                any = true;
                v[0] = tri[f], v[1] = tri[f + 1], v[2] = tri[f + 2];
                w[0] = wa, w[1] = wb, w[2] = wc;
            }
        }
        return any;
    }

    // blended (left | right) responses for direction d -> out[2 * taps]; false: no face found
    bool blend(const float d[3], float* out) const {
        uint32_t v[3];
        float w[3];
        if (!locate(d, v, w)) return false;
        for (int ear = 0; ear < 2; ear++) {
            const float* a = &ir[ear ? vertices[v[0]].right : vertices[v[0]].left];
            const float* b = &ir[ear ? vertices[v[1]].right : vertices[v[1]].left];
            const float* c = &ir[ear ? vertices[v[2]].right : vertices[v[2]].left];
            for (uint32_t i = 0; i < taps; i++) out[ear * taps + i] = a[i] * w[0] + b[i] * w[1] + c[i] * w[2];
        }
        return true;
    }
};

}  // namespace wae
