"""AudioBuffer::resample — the input side of the path (decode_audio_data resamples every asset to the context rate) — pinned on the
reference's unit tests (src/buffer.rs:716-816).  `resample(x, from_rate, to_rate)` is the oracle's restatement here and the GPU
kernel behind wae_resample_linear in tests/test_gpu_reference_cases.py."""
import ctypes as C

import numpy as np
import pytest

fp = C.POINTER(C.c_float)


@pytest.fixture
def resample(oracle):
    def run(x, from_rate, to_rate):
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros(int(np.ceil(len(x) * (to_rate / from_rate))) + 8, np.float32)
        n = oracle.api.resample_linear(x.ctypes.data_as(fp), len(x), float(from_rate), float(to_rate), out.ctypes.data_as(fp), len(out))
        return out[:n]
    return run


def check_up_and_downsample(resample):  # buffer.rs:735-770 test_upsample, test_downsample
    up = resample([1.0, 2.0, 3.0, 4.0, 5.0], 48000.0, 96000.0)
    want = np.float32(1.0) + np.float32(4.0 / 9.0) * np.arange(10, dtype=np.float32)  # (5 - 1) / (10 - 1)
    assert len(up) == 10 and np.abs(up - want).max() <= 1e-6
    down = resample([1.0, 2.0, 3.0, 4.0, 5.0], 96000.0, 48000.0)
    assert np.array_equal(down, np.array([1.0, 3.0, 5.0], np.float32))


def check_resample_stereo(resample, source_sr):  # buffer.rs:772-816 test_resample_stereo: one period of sin / cos, abs <= 1e-3
    target_sr = 44100
    two_pi = np.float32(2.0) * np.float32(np.pi)
    phase = np.arange(source_sr, dtype=np.float32) / np.float32(source_sr) * two_pi
    want_phase = np.arange(target_sr, dtype=np.float32) / np.float32(target_sr) * two_pi
    for fn in (np.sin, np.cos):
        got = resample(fn(phase).astype(np.float32), float(source_sr), float(target_sr))
        assert len(got) == target_sr
        assert np.abs(got - fn(want_phase).astype(np.float32)).max() <= 1e-3


def check_resample_edge_cases(resample):  # buffer.rs:724-733 test_resample_from_empty, :313-318 "very similar" rates are left alone
    assert len(resample(np.zeros(0, np.float32), 48000.0, 44100.0)) == 0
    x = np.arange(7, dtype=np.float32)
    assert np.array_equal(resample(x, 48000.0, 48000.05), x)
    one = resample([3.0], 48000.0, 96000.0)  # a single frame: both output frames are that frame
    assert len(one) == 2 and not np.isnan(one).any()


def test_up_and_downsample(resample):
    check_up_and_downsample(resample)


@pytest.mark.parametrize("source_sr", [22500, 38000, 48000, 96000])
def test_resample_stereo(resample, source_sr):
    check_resample_stereo(resample, source_sr)


def test_resample_edge_cases(resample):
    check_resample_edge_cases(resample)
