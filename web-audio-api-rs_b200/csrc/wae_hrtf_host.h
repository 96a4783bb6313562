// Host half of HRTF panning (PanningModelType::HRTF, src/node/panner.rs:39-68,215-271): the HRIR sphere container the
// reference embeds (resources/IRC_1003_C.bin, handed over by the caller through wae_engine_set_hrir_sphere), and the
// per-direction blend of the three vertex responses of the sphere triangle the source direction crosses — the job of the
// third-party `hrtf` 0.8.1 crate's HrirSphere / sample_bilinear.  The convolution itself runs on the GPU (k_hrtf_fir).
#pragma once
#include "wae_spatial.h"

#include <algorithm>
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
    // HrirSphere::new(reader, context_rate) of the crate: when the context runs at another rate than the data (the embedded sphere is
    // 44.1 kHz, contexts usually 48 kHz) every response is resampled once with rubato's asynchronous sinc resampler — see
    // hrir_resample() below.  Geometry is shared, `taps` becomes the resampled length.
    HrirSphere at_rate(uint32_t rate) const;

    // blend weights of direction d (sphere coordinates)
    bool locate(const float d[3], uint32_t v[3], float w[3]) const {
        return spatial::hrir_locate(pos.data(), tri.data(), (int)(tri.size() / 3), d, v, w);
    }
};

// One impulse response through rubato's SincFixedIn as hrtf 0.8.1 drives it (a single `process` call over the whole response;
// sinc_len 256, f_cutoff 0.95, 160 sub-sample phases, cubic interpolation between the 4 nearest phases, BlackmanHarris2 window).
// Restated from the published algorithm; the crates are not part of the reference checkout (parity unpinned).
struct SincBank {
    static constexpr int kLen = 256, kPhases = 160;
    std::vector<float> taps;  // [phase][kLen]
    explicit SincBank(float cutoff) : taps((size_t)kLen * kPhases) {
        const int total = kLen * kPhases;
        const float pi = 3.14159265358979323846f, n = (float)total;
        std::vector<float> proto((size_t)total);
        float norm = 0.f;
        for (int i = 0; i < total; i++) {
            const float t = (float)i;
            const float win = 0.35875f - 0.48829f * std::cos(2.f * pi * t / n) + 0.14128f * std::cos(4.f * pi * t / n) - 0.01168f * std::cos(6.f * pi * t / n);
            const float arg = (t - (float)(total / 2)) * cutoff / (float)kPhases;
            const float snc = arg == 0.f ? 1.f : std::sin(arg * pi) / (arg * pi);
            proto[(size_t)i] = win * win * snc;
            norm += proto[(size_t)i];
        }
        norm /= (float)kPhases;
        for (int tap = 0; tap < kLen; tap++)
            for (int ph = 0; ph < kPhases; ph++) taps[(size_t)(kPhases - ph - 1) * kLen + tap] = proto[(size_t)(kPhases * tap + ph)] / norm;
    }
    float dot(const float* x, int phase) const {
        const float* c = &taps[(size_t)phase * kLen];
        float acc = 0.f;
        for (int j = 0; j < kLen; j++) acc += x[j] * c[j];
        return acc;
    }
};

inline std::vector<float> hrir_resample(const SincBank& bank, const float* hrir, size_t len, double ratio) {
    const int L = SincBank::kLen, P = SincBank::kPhases;
    std::vector<float> padded(len + 2 * (size_t)L, 0.f);  // two filter lengths of silence before the response
    std::copy(hrir, hrir + len, padded.begin() + 2 * L);
    const double step = 1.0 / ratio, stop = (double)((int64_t)len - (L + 1));
    std::vector<float> out;
    if (!(-(double)(L / 2) < stop)) return out;  // shorter than half a filter: the resampler emits nothing
    out.reserve((size_t)((double)len * ratio) + 16);
    for (double pos = -(double)(L / 2) + step; ; pos += step) {  // the resampler advances first, then emits
        const double whole = std::floor(pos);
        const int64_t base = (int64_t)whole;
        const int64_t phase0 = (int64_t)std::floor((pos - whole) * (double)P) - 1;
        const double fine = pos * (double)P;
        const float x = (float)(fine - std::floor(fine));
        float y[4];
        for (int k = 0; k < 4; k++) {
            int64_t ph = phase0 + k, at = base;
            if (ph < 0) { ph += P; at -= 1; }
            else if (ph >= P) { ph -= P; at += 1; }
            y[k] = bank.dot(padded.data() + (at + 2 * L), (int)ph);
        }
        const float c1 = -(1.f / 3.f) * y[0] - 0.5f * y[1] + y[2] - (1.f / 6.f) * y[3];
        const float c2 = 0.5f * (y[0] + y[2]) - y[1];
        const float c3 = 0.5f * (y[1] - y[2]) + (1.f / 6.f) * (y[3] - y[0]);
        const float x2 = x * x;
        out.push_back(y[1] + c1 * x + c2 * x2 + c3 * x2 * x);
        if (!(pos < stop)) break;
    }
    return out;
}

inline HrirSphere HrirSphere::at_rate(uint32_t rate) const {
    HrirSphere r;
    r.sample_rate = rate;
    r.vertices = vertices;
    r.tri = tri;
    r.pos = pos;
    const double ratio = (double)rate / (double)sample_rate;
    const SincBank bank(ratio >= 1.0 ? 0.95f : 0.95f * (float)ratio);
    for (size_t v = 0; v < vertices.size(); v++) {
        std::vector<float> left = hrir_resample(bank, &ir[vertices[v].left], taps, ratio);
        std::vector<float> right = hrir_resample(bank, &ir[vertices[v].right], taps, ratio);
        if (v == 0) {
            r.taps = (uint32_t)left.size();
            r.ir.resize(2ull * vertices.size() * r.taps);
        }
        left.resize(r.taps, 0.f);
        right.resize(r.taps, 0.f);
        r.vertices[v].left = (uint32_t)(2ull * v * r.taps);
        r.vertices[v].right = r.vertices[v].left + r.taps;
        std::memcpy(&r.ir[r.vertices[v].left], left.data(), r.taps * sizeof(float));
        std::memcpy(&r.ir[r.vertices[v].right], right.data(), r.taps * sizeof(float));
    }
    return r;
}

}  // namespace wae
