// Host half of HRTF panning (PanningModelType::HRTF, src/node/panner.rs:39-68,215-271): the HRIR sphere container the
// reference embeds (resources/IRC_1003_C.bin, handed over by the caller through wae_engine_set_hrir_sphere), and the
// per-direction blend of the three vertex responses of the sphere triangle the source direction crosses — the job of the
// third-party `hrtf` 0.8.1 crate's HrirSphere / sample_bilinear.  The convolution itself runs on the GPU (k_hrtf_fir).
#pragma once
#include "wae_spatial.h"

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
        finish();
        return true;
    }

    std::vector<float> pos;  // [vertex][3]
    void finish() {
        pos.resize(3 * vertices.size());
        for (size_t v = 0; v < vertices.size(); v++) std::memcpy(&pos[3 * v], vertices[v].p, 12);
    }
    // blend weights of direction d (sphere coordinates)
    bool locate(const float d[3], uint32_t v[3], float w[3]) const {
        return spatial::hrir_locate(pos.data(), tri.data(), (int)(tri.size() / 3), d, v, w);
    }
};

}  // namespace wae
