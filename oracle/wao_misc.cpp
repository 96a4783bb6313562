// ORACLE — TEST INFRASTRUCTURE ONLY (see wao_core.h).
#include <cstdint>
#include <cmath>
#include <cstddef>

// AudioBuffer::resample, src/buffer.rs:311-363: linear interpolation keeping first and last sample,
// target_length = ceil(len * to/from).  Returns the target length (writes at most out_cap samples).
extern "C" __attribute__((visibility("default"))) uint64_t wao_resample_linear(const float* in, uint64_t len, float from_rate,
                                                                              float to_rate, float* out, uint64_t out_cap) {
    if (std::fabs(from_rate - to_rate) <= 0.1f || len == 0) {
        for (uint64_t i = 0; i < len && i < out_cap; i++) out[i] = in[i];
        return len;
    }
    double ratio = (double)to_rate / (double)from_rate;
    uint64_t target_length = (uint64_t)std::ceil((double)len * ratio);
    for (uint64_t i = 0; i < target_length && i < out_cap; i++) {
        double position = (double)i / (double)(target_length - 1);
        double playhead = position * (double)(len - 1);
        double playhead_floored = std::floor(playhead);
        uint64_t prev_index = (uint64_t)playhead_floored;
        uint64_t next_index = prev_index + 1 < len - 1 ? prev_index + 1 : len - 1;
        float k = (float)(playhead - playhead_floored);
        float k_inv = 1.f - k;
        out[i] = k_inv * in[prev_index] + k * in[next_index];
    }
    return target_length;
}
