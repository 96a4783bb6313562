"""HRTF panning at a context rate other than the HRIR sphere's (the embedded sphere is 44.1 kHz data, contexts usually run at 48 kHz):
the hrtf crate resamples every impulse response once at load time with rubato's asynchronous sinc resampler.  Neither crate is in the
reference checkout (parity unpinned); what is checked here is that the two independent statements of that algorithm —
the oracle's (oracle/wao_hrtf.cpp) and the product's host code (csrc/wae_hrtf_host.h) — agree sample for sample, and that the
result IS a band-limited resampling (a tone keeps its frequency and amplitude, with the resampler's half-filter-length delay)."""
import ctypes as C

import numpy as np
import pytest

fp = C.POINTER(C.c_float)


def resample(api, x, ratio):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(int(len(x) * ratio) + 64, np.float32)
    n = C.c_uint32(0)
    api.check(api.hrir_resample(x.ctypes.data_as(fp), len(x), float(ratio), out.ctypes.data_as(fp), len(out), C.byref(n)))
    return out[:n.value]


@pytest.mark.parametrize("ratio", [48000 / 44100, 96000 / 44100, 27000 / 44100, 32000 / 44100])
def test_the_two_statements_agree(pkg, oracle, ratio):
    import os
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "web-audio-api-rs_b200", "libwae_b200.so")
    if not os.path.exists(so):
        pytest.skip("libwae_b200.so is not built")
    rng = np.random.default_rng(7)
    hrir = (rng.standard_normal(512) * np.exp(-np.arange(512) / 60.0)).astype(np.float32)
    a, b = resample(oracle.api, hrir, ratio), resample(pkg.api(), hrir, ratio)
    assert len(a) == len(b) and len(a) > 100
    assert np.abs(a - b).max() <= 1e-6 * np.abs(a).max()


def test_length_and_delay_of_the_resampled_response(host_api):
    # 512 taps at 44.1 kHz -> 48 kHz: the resampler starts half a filter (128 input frames) early and stops 257 frames before the
    # end of the chunk, so (512 - 257 + 128) * 48000 / 44100 ~ 417 output frames come out and an impulse at input frame k lands at
    # output frame k * ratio (the 128-frame head start cancels the filter's 128-frame centre)
    ratio = 48000 / 44100
    x = np.zeros(512, np.float32)
    x[100] = 1.0
    y = resample(host_api, x, ratio)
    assert 415 <= len(y) <= 419
    assert abs(int(np.argmax(np.abs(y))) - 100 * ratio) <= 1.0
    assert abs(float(y.sum()) - ratio) <= 2e-3  # unit DC gain for signals: one input sample spreads over `ratio` output samples


def test_a_tone_keeps_its_frequency_and_amplitude(host_api):
    ratio = 48000 / 44100
    n = np.arange(512)
    x = np.sin(2 * np.pi * 3000.0 / 44100.0 * n).astype(np.float32)
    y = resample(host_api, x, ratio)
    mid = slice(140, len(y) - 10)  # the first ~128 * ratio frames ramp in through the zero history
    m = np.arange(len(y))[mid]
    basis = np.stack([np.sin(2 * np.pi * 3000.0 / 48000.0 * m), np.cos(2 * np.pi * 3000.0 / 48000.0 * m)], axis=1)
    coef, *_ = np.linalg.lstsq(basis, y[mid].astype(np.float64), rcond=None)
    assert abs(float(np.hypot(*coef)) - 1.0) <= 1e-2                    # same amplitude
    assert np.abs(basis @ coef - y[mid]).max() <= 5e-3                  # nothing but that 3 kHz tone
    assert abs(float(np.arctan2(coef[1], coef[0]))) <= 0.1              # sub-sample time alignment with the input
