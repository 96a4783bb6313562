// ORACLE — TEST INFRASTRUCTURE ONLY (see wao_core.h).
//
// HRTF panning: the reference delegates to the third-party crate `hrtf = "0.8.1"` (Cargo.toml:41; call
// sites src/node/panner.rs:39-68,239-271) whose source is NOT in /root/reference, and at 48 kHz that crate
// resamples the 44.1 kHz HRIR sphere with rubato's sinc resampler (also not vendored).
// PARITY UNPINNED: the only reference test (panner.rs:1225-1269) asserts "output != input" and "tail is
// non-zero".  Not restated in this round: hrtf_sphere_available() reports why, wao_create_panner returns
// WAE_UNSUPPORTED for panningModel = HRTF.
#include "wao_panner.h"

namespace wao {

struct HrtfState {
    size_t len = 0;
};

bool hrtf_sphere_available(std::string& why) {
    why = "HRTF panning depends on the un-vendored hrtf 0.8.1 crate (parity unpinned); not restated yet";
    return false;
}
size_t hrtf_tail_time_samples(const HrtfState& s) { return s.len; }
void hrtf_process(HrtfState&, const float*, float, const float*, float* out_lr) {
    for (int i = 0; i < 2 * RQ; i++) out_lr[i] = 0.f;
}
std::shared_ptr<HrtfState> hrtf_state_new(float) { return nullptr; }

}  // namespace wao
