"""The reference's Analyser unit tests (src/analysis.rs:414-868) restated at graph level: the ring buffer is fed by rendering an
AudioBufferSource through an AnalyserNode instead of calling AnalyserRingBuffer::write directly, the read-outs are the node's.
Every function names the `#[test]` it restates and takes any backend (tests/test_gpu_reference_cases.py reruns them on CUDA)."""
import numpy as np
import pytest

RQ = 128


def _analyse(pkg, be, signal, sr=44100.0, **opts):
    signal = np.asarray(signal, np.float32)
    n = -(-len(signal) // RQ) * RQ
    c = pkg.OfflineAudioContext(1, n, sr, be)
    src = c.create_buffer_source(pkg.AudioBuffer([signal], sr))
    a = c.create_analyser(**opts)
    src.connect(a)
    a.connect(c.destination())
    src.start()
    out = c.start_rendering_sync().get_channel_data(0)
    assert np.array_equal(out[:len(signal)], signal)  # the analyser is a pass-through
    return a, c


def test_time_domain_data_vs_fft_size(pkg, oracle):  # analysis.rs:655-691 test_get_float_time_domain_data_vs_fft_size
    a, _c = _analyse(pkg, oracle, np.ones(RQ), fft_size=32)
    dst = np.zeros(RQ, np.float32)  # dst is bigger than fft_size: only fft_size values are written
    a.get_float_time_domain_data(out=dst)
    assert np.array_equal(dst, np.concatenate([np.ones(32), np.zeros(96)]).astype(np.float32))
    a, _c = _analyse(pkg, oracle, np.ones(RQ), fft_size=128)
    assert np.array_equal(a.get_float_time_domain_data(16), np.ones(16, np.float32))  # dst is smaller than fft_size


def test_time_domain_data_is_the_most_recent_window(pkg, oracle):  # analysis.rs:439-590 ring buffer write / read (wrap included)
    # 300 quanta wrap the 32768-frame ring buffer (RING_BUFFER_SIZE = MAX_FFT_SIZE, analysis.rs:19-24) more than once
    sig = (np.arange(300 * RQ) % 1000).astype(np.float32) / np.float32(1000.0)
    a, _c = _analyse(pkg, oracle, sig, fft_size=2048)
    assert np.array_equal(a.get_float_time_domain_data(), sig[-2048:])
    assert np.array_equal(a.get_float_time_domain_data(100), sig[-100:])  # a shorter array gets the most recent len frames (:113-126)


def test_byte_time_domain_data(pkg, oracle):  # analysis.rs:693-718 get_byte_time_domain_data
    a, _c = _analyse(pkg, oracle, np.ones(RQ))
    assert np.array_equal(a.get_byte_time_domain_data(RQ)[-RQ:], np.full(RQ, 255, np.uint8))
    a, _c = _analyse(pkg, oracle, np.ones(2048), fft_size=128)
    assert np.array_equal(a.get_byte_time_domain_data(RQ), np.full(RQ, 255, np.uint8))
    a, _c = _analyse(pkg, oracle, -np.ones(2048), fft_size=128)
    assert np.array_equal(a.get_byte_time_domain_data(RQ), np.zeros(RQ, np.uint8))


@pytest.mark.parametrize("bins", [range(1, 32), range(32, 64), range(64, 96), range(96, 128)], ids=["1-31", "32-63", "64-95", "96-127"])
def test_float_frequency_data_peaks_at_the_sine_bin(pkg, oracle, bins):  # analysis.rs:720-769 test_get_float_frequency_data
    sr, fft_size = 44100.0, 1024
    res = np.float32(43.066)
    i = np.arange(fft_size, dtype=np.float32)
    for k in bins:
        freq = res * np.float32(k)
        phase = freq * i / np.float32(sr)
        sig = np.sin(phase * np.float32(2.0) * np.float32(np.pi), dtype=np.float32)
        a, _c = _analyse(pkg, oracle, sig, sr, fft_size=fft_size, smoothing_time_constant=0.8)
        db = a.get_float_frequency_data()
        assert len(db) == fft_size // 2 and int(np.argmax(db)) == k and np.sum(db == db[k]) == 1


def test_frequency_data_vs_frequency_bin_count(pkg, oracle):
    # analysis.rs:771-806: silence -> -inf dB / byte 0 in the fft_size / 2 bins, the rest of the caller's array is left alone
    a, _c = _analyse(pkg, oracle, np.zeros(RQ), fft_size=RQ)
    assert a.frequency_bin_count() == RQ // 2
    bins = np.full(RQ, -1.0, np.float32)
    a.get_float_frequency_data(out=bins)
    assert np.all(np.isneginf(bins[:RQ // 2])) and np.array_equal(bins[RQ // 2:], np.full(RQ // 2, -1.0, np.float32))
    a, _c = _analyse(pkg, oracle, np.zeros(RQ), fft_size=RQ)
    bins = np.full(RQ, 255, np.uint8)
    a.get_byte_frequency_data(out=bins)
    assert np.array_equal(bins[:RQ // 2], np.zeros(RQ // 2, np.uint8)) and np.array_equal(bins[RQ // 2:], np.full(RQ // 2, 255, np.uint8))


def test_option_constraints(pkg, oracle):  # analysis.rs:592-653 fft size / smoothing / decibel constraints
    c = pkg.OfflineAudioContext(1, RQ, 44100.0, oracle)
    c.create_analyser(fft_size=32)
    c.create_analyser(fft_size=32768)
    c.create_analyser(min_decibels=-20.0, max_decibels=10.0)
    for bad in [dict(fft_size=13), dict(fft_size=16), dict(fft_size=65536), dict(smoothing_time_constant=-1.0), dict(smoothing_time_constant=2.0),
                dict(min_decibels=-30.0, max_decibels=-30.0), dict(min_decibels=-100.0, max_decibels=-100.0), dict(min_decibels=0.0, max_decibels=-30.0)]:
        with pytest.raises(pkg.WaeError):
            c.create_analyser(**bad)
