"""Pins the oracle's AudioBufferSourceRenderer against the reference's own unit tests (src/node/audio_buffer_source.rs
:974-1890); every function names the `#[test]` it restates.  The functions take any backend: tests/test_gpu_reference_cases.py
runs the same cases through the CUDA engine.  (Tests about events / `onended` / set_buffer inside a suspend callback are
control-plane behaviour and are not restated.)"""
import numpy as np
import pytest

RQ = 128
PI = np.float32(np.pi)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def ctx(pkg, be, ch, length, sr):
    return pkg.OfflineAudioContext(ch, length, sr, be)


def buf(pkg, data, sr, length=None):
    chans = []
    for d in data:
        a = np.zeros(length if length is not None else len(d), np.float32)
        a[:len(d)] = d
        chans.append(a)
    return pkg.AudioBuffer(chans, sr)


def render1(pkg, be, length, sr, buffer, setup, ch=1, **opts):
    c = ctx(pkg, be, ch, length, sr)
    s = c.create_buffer_source(buffer, **opts)
    s.connect(c.destination())
    setup(s, c)
    return c.start_rendering_sync()


def test_sub_quantum_start_1(pkg, oracle):  # :974-994
    out = render1(pkg, oracle, RQ, 48000.0, buf(pkg, [[1.0]], 48000.0), lambda s, c: s.start_at(1.0 / 48000.0)).get_channel_data(0)
    want = np.zeros(RQ, np.float32)
    want[1] = 1.0
    assert np.array_equal(out, want)


def test_sub_quantum_start_2(pkg, oracle):  # :997-1033
    sr = 44100.0
    c = ctx(pkg, oracle, 2, int(4 * sr), sr)
    dirac = buf(pkg, [[1.0], [1.0]], sr, 512)
    offsets = [0, 3, 512, 517, 1000, 1005, 20000, 21234, 37590]
    for i in offsets:
        s = c.create_buffer_source(dirac)
        s.connect(c.destination())
        s.start_at(i / sr)
    res = c.start_rendering_sync()
    left, right = res.get_channel_data(0), res.get_channel_data(1)
    assert np.array_equal(left, right)
    assert all(left[i] != 0.0 for i in offsets)


def test_sub_sample_start(pkg, oracle):  # :1036-1056
    out = render1(pkg, oracle, RQ, 48000.0, buf(pkg, [[1.0]], 48000.0), lambda s, c: s.start_at(1.5 / 48000.0)).get_channel_data(0)
    want = np.zeros(RQ, np.float32)
    want[2] = 0.5
    assert np.array_equal(out, want)


def test_sub_quantum_stop(pkg, oracle):  # :1059-1100 (fast track, slow track)
    sr = 48000.0
    out = render1(pkg, oracle, RQ, sr, buf(pkg, [[0, 0, 0, 0, 1.0]], sr, RQ), lambda s, c: (s.start_at(0.0), s.stop_at(4.0 / sr))).get_channel_data(0)
    assert np.array_equal(out, np.zeros(RQ, np.float32))
    out = render1(pkg, oracle, RQ, sr, buf(pkg, [[0, 0, 0, 1.0]], sr, RQ), lambda s, c: (s.start_at(1.0 / sr), s.stop_at(4.0 / sr))).get_channel_data(0)
    assert np.array_equal(out, np.zeros(RQ, np.float32))


def test_sub_sample_stop(pkg, oracle):  # :1103-1148 (fast track, slow track)
    sr = 48000.0
    b = buf(pkg, [[0, 0, 0, 0, 1.0, 1.0]], sr, RQ)
    out = render1(pkg, oracle, RQ, sr, b, lambda s, c: (s.start_at(0.0), s.stop_at(4.5 / sr))).get_channel_data(0)
    want = np.zeros(RQ, np.float32)
    want[4] = 1.0
    assert np.array_equal(out, want)
    out = render1(pkg, oracle, RQ, sr, b, lambda s, c: (s.start_at(1.0 / sr), s.stop_at(5.5 / sr))).get_channel_data(0)
    want = np.zeros(RQ, np.float32)
    want[5] = 1.0
    assert np.array_equal(out, want)


def test_start_in_the_past(pkg, oracle):  # :1151-1172: start_at(0) issued from a suspend callback at frame 128
    sr = 48000.0
    c = ctx(pkg, oracle, 1, 2 * RQ, sr)
    dirac = buf(pkg, [[1.0]], sr)

    def cb(context):
        s = context.create_buffer_source(dirac)
        s.connect(context.destination())
        s.start_at(0.0)

    c.suspend_sync(128.0 / sr, cb)
    out = c.start_rendering_sync().get_channel_data(0)
    want = np.zeros(2 * RQ, np.float32)
    want[128] = 1.0
    assert np.array_equal(out, want)


def _sine(n, sr, freq=1.0, scale=2.0):
    i = np.arange(n, dtype=np.float32)
    return np.sin(np.float32(freq) * i / np.float32(sr) * np.float32(scale) * PI).astype(np.float32)


@pytest.mark.parametrize("buf_sr", [22500, 38000, 43800, 48000, 96000])
def test_audio_buffer_resampling(pkg, oracle, buf_sr):  # :1175-1217
    base = 44100
    out = render1(pkg, oracle, base, float(base), pkg.AudioBuffer([_sine(buf_sr, buf_sr)], float(buf_sr)), lambda s, c: s.start_at(0.0)).get_channel_data(0)
    assert np.abs(out - _sine(base, base)).max() <= 1e-6


def test_playback_rate_and_detune(pkg, oracle):  # :1220-1255 test_playback_rate, :1294-1329 test_detune
    sr = 44100
    sine = _sine(sr, sr)
    want = _sine(sr, sr, scale=1.0)
    out = render1(pkg, oracle, sr, float(sr), pkg.AudioBuffer([sine], float(sr)), lambda s, c: (s.playback_rate.set_value(0.5), s.start())).get_channel_data(0)
    assert np.abs(out - want).max() <= 1e-6
    out = render1(pkg, oracle, sr, float(sr), pkg.AudioBuffer([sine], float(sr)), lambda s, c: (s.detune.set_value(-1200.0), s.start())).get_channel_data(0)
    assert np.abs(out - want).max() <= 1e-6


def test_negative_playback_rate(pkg, oracle):  # :1258-1291
    sr = 44100
    sine = _sine(sr, sr)
    b = pkg.AudioBuffer([sine], float(sr))
    out = render1(pkg, oracle, sr, float(sr), b, lambda s, c: (s.playback_rate.set_value(-1.0), s.start_at_with_offset(0.0, sr / float(sr)))).get_channel_data(0)
    want = np.concatenate([[0.0], sine[::-1][:-1]]).astype(np.float32)
    assert np.abs(out - want).max() <= 1e-6


def test_end_of_file(pkg, oracle):  # :1332-1381 (fast track, slow track 1), :1837-1889 (fast track 2, slow track 2)
    sr = 48000.0
    data = np.zeros(129, np.float32)
    data[0] = data[128] = 1.0
    out = render1(pkg, oracle, 2 * RQ, sr, pkg.AudioBuffer([data], sr), lambda s, c: s.start_at(0.0)).get_channel_data(0)
    want = np.zeros(256, np.float32)
    want[0] = want[128] = 1.0
    assert np.array_equal(out, want)
    out = render1(pkg, oracle, 2 * RQ, sr, pkg.AudioBuffer([data], sr), lambda s, c: s.start_at(1.0 / sr)).get_channel_data(0)
    want = np.zeros(256, np.float32)
    want[1] = want[129] = 1.0
    assert np.abs(out - want).max() <= 1e-10
    b5 = buf(pkg, [[1.0]], sr, 5)
    out = render1(pkg, oracle, RQ, sr, b5, lambda s, c: (s.start_at(0.0), s.stop_at(125.0 / sr))).get_channel_data(0)
    want = np.zeros(RQ, np.float32)
    want[0] = 1.0
    assert np.array_equal(out, want)
    out = render1(pkg, oracle, RQ, sr, b5, lambda s, c: (s.start_at(1.0 / sr), s.stop_at(125.0 / sr))).get_channel_data(0)
    want = np.zeros(RQ, np.float32)
    want[1] = 1.0
    assert np.array_equal(out, want)


def test_with_duration_and_offset(pkg, oracle):  # :1384-1506 test_with_duration_0 / _1 / _2, test_with_offset
    sr = 48000.0
    b = buf(pkg, [[0, 0, 0, 0, 1.0, 1.0]], sr, RQ)
    for start, offset, duration, idx in [(0.0, 0.0, 4.5 / sr, 4), (1.0 / sr, 0.0, 4.5 / sr, 5), (0.0, 1.0 / sr, 3.5 / sr, 3)]:
        out = render1(pkg, oracle, RQ, sr, b, lambda s, c: s.start_at_with_offset_and_duration(start, offset, duration)).get_channel_data(0)
        want = np.zeros(RQ, np.float32)
        want[idx] = 1.0
        assert np.array_equal(out, want), (start, offset, duration)
    sr = 32768.0
    a, e = 3.1, 37.2
    out = render1(pkg, oracle, RQ, sr, pkg.AudioBuffer([np.ones(RQ, np.float32)], sr),
                  lambda s, c: s.start_at_with_offset_and_duration(a / sr, 0.0, (e - a) / sr)).get_channel_data(0)
    want = np.ones(RQ, np.float32)
    want[:int(np.floor(a)) + 1] = 0.0
    want[int(np.ceil(e)):] = 0.0
    assert np.array_equal(out, want)


def test_reverse_playback_with_duration(pkg, oracle):  # :1537-1555
    sr = 48000.0
    b = pkg.AudioBuffer([f32([1, 2, 3, 4, 5])], sr)
    out = render1(pkg, oracle, RQ, sr, b, lambda s, c: (s.playback_rate.set_value(-1.0), s.start_at_with_offset_and_duration(0.0, 5 / sr, 2.0 / sr))).get_channel_data(0)
    want = np.zeros(RQ, np.float32)
    want[1] = 5.0
    assert np.array_equal(out, want)


def test_offset_larger_than_buffer_duration(pkg, oracle):  # :1558-1573 (the source is not even connected)
    sr = 48000.0
    c = ctx(pkg, oracle, 1, RQ, sr)
    s = c.create_buffer_source(pkg.AudioBuffer([np.ones(13, np.float32)], sr))
    s.start_at_with_offset(0.0, 64.0 / sr)
    assert np.array_equal(c.start_rendering_sync().get_channel_data(0), np.zeros(RQ, np.float32))


LOOP_LENS = [RQ // 2 - 1, RQ // 2, RQ // 2 + 1, RQ - 1, RQ, RQ + 1, 2 * RQ - 1, 2 * RQ, 2 * RQ + 1]


@pytest.mark.parametrize("buffer_len", LOOP_LENS)
def test_track_loop_mono(pkg, oracle, buffer_len):  # :1576-1651 test_fast_track_loop_mono / test_slow_track_loop_mono
    sr, n = 48000.0, 4 * RQ
    out = render1(pkg, oracle, n, sr, buf(pkg, [[1.0]], sr, buffer_len), lambda s, c: s.start(), loop=True).get_channel_data(0)
    want = np.zeros(n, np.float32)
    want[0::buffer_len] = 1.0
    assert np.abs(out - want).max() <= 1e-10
    out = render1(pkg, oracle, n, sr, buf(pkg, [[1.0]], sr, buffer_len), lambda s, c: s.start_at(1.0 / sr), loop=True).get_channel_data(0)
    want = np.zeros(n, np.float32)
    want[1::buffer_len] = 1.0
    assert np.abs(out - want).max() <= 1e-9


@pytest.mark.parametrize("buffer_len", LOOP_LENS)
def test_track_loop_stereo(pkg, oracle, buffer_len):  # :1654-1755 test_fast_track_loop_stereo / test_slow_track_loop_stereo
    sr, n = 48000.0, 4 * RQ
    for first, tol, setup in [(0, 1e-10, lambda s, c: s.start()), (1, 1e-9, lambda s, c: s.start_at(1.0 / sr))]:
        res = render1(pkg, oracle, n, sr, buf(pkg, [[1.0], [0.0, 1.0]], sr, buffer_len), setup, ch=2, loop=True)
        left, right = np.zeros(n, np.float32), np.zeros(n, np.float32)
        for i in range(first, n, buffer_len):
            left[i] = 1.0
            if i < n - 1:
                right[i + 1] = 1.0
        assert np.abs(res.get_channel_data(0) - left).max() <= tol
        assert np.abs(res.get_channel_data(1) - right).max() <= tol


def test_reverse_loop_boundaries(pkg, oracle):  # :1758-1777
    sr = 48000.0
    b = pkg.AudioBuffer([f32([1, 2, 3, 4, 5])], sr)
    out = render1(pkg, oracle, RQ, sr, b, lambda s, c: (s.playback_rate.set_value(-1.0), s.start_at_with_offset(0.0, 3.0 / sr)),
                  loop=True, loop_start=1.0 / sr, loop_end=4.0 / sr).get_channel_data(0)
    assert np.array_equal(out[:8], f32([4, 3, 2, 4, 3, 2, 4, 3]))


@pytest.mark.parametrize("loop_start,loop_end,error", [(-2.0, -1.0, 0.0), (-1.0, -2.0, 0.0), (0.0, 0.0, 0.0), (-1.0, 2.0, 0.0), (2.0, -1.0, 1e-10),
                                                      (1.0, 1.0, 1e-10), (2.0, 3.0, 1e-10), (3.0, 2.0, 1e-10)])
def test_loop_out_of_bounds(pkg, oracle, loop_start, loop_end, error):  # :1780-1834
    sr = 48000.0
    n = 4800
    out = render1(pkg, oracle, n, sr, buf(pkg, [[1.0]], sr, 500), lambda s, c: s.start(), loop=True, loop_start=loop_start, loop_end=loop_end).get_channel_data(0)
    want = np.zeros(n, np.float32)
    want[0::500] = 1.0
    assert np.abs(out - want).max() <= error


def test_construct_with_options_and_run(pkg, oracle):  # :896-915
    sr = 44100.0
    c = pkg.OfflineAudioContext(1, RQ, sr, oracle)
    src = c.create_buffer_source(pkg.AudioBuffer([np.ones(RQ, np.float32)], sr))
    src.connect(c.destination())
    src.start()
    assert np.array_equal(c.start_rendering_sync().get_channel_data(0), np.ones(RQ, np.float32))


def test_null_buffer_start_ends_before_start_time(pkg, oracle):  # :1509-1534: a buffer given after the source has already "ended" stays unheard
    sr = 48000.0
    c = pkg.OfflineAudioContext(1, int(sr), sr, oracle)
    src = c.create_buffer_source()
    src.connect(c.destination())
    src.start_at(0.75)
    c.suspend_sync(0.5, lambda ctx: src.set_buffer(pkg.AudioBuffer([np.ones(1, np.float32)], sr)))
    out = c.start_rendering_sync().get_channel_data(0)
    assert not out.any()


def test_loop_no_restart_suspend(pkg, oracle):  # :1892-1917: set_loop(true) after the buffer has played out does not restart it
    sr = 48000.0
    c = pkg.OfflineAudioContext(1, RQ * 2, sr, oracle)
    src = c.create_buffer_source()
    src.connect(c.destination())
    src.set_buffer(pkg.AudioBuffer([np.ones(1, np.float32)], sr))
    src.start_at(0.0)
    c.suspend_sync(RQ / sr, lambda ctx: src.set_loop(True))
    out = c.start_rendering_sync().get_channel_data(0)
    want = np.zeros(RQ * 2, np.float32)
    want[0] = 1.0
    assert np.array_equal(out, want)


@pytest.mark.parametrize("sample_rate,buffer_rate,threshold", [(44100.0, 44100.0, 9.0957e-5), (44100.0, 43800.0, 3.8986e-3)])
def test_subsample_buffer_stitching(pkg, oracle, sample_rate, buffer_rate, threshold):
    # :1987-2043 (ported there from wpt sub-sample-buffer-stitching.html): a sine cut into 30-frame buffers, each started at its
    # sub-sample time, adds up to the sine again
    buffer_length, frequency = 30, 440.0
    length = buffer_length * 15
    c = pkg.OfflineAudioContext(2, length, sample_rate, oracle)
    i = np.arange(length, dtype=np.float32)
    omega = np.float32(2.0) * np.float32(np.pi) / np.float32(buffer_rate) * np.float32(frequency)
    wave = np.sin(omega * i, dtype=np.float32)
    for k in range(0, length, buffer_length):
        src = c.create_buffer_source(pkg.AudioBuffer([wave[k:k + buffer_length].copy()], buffer_rate))
        src.connect(c.destination())
        src.start_at(k / buffer_rate)
    omega = np.float32(2.0) * np.float32(np.pi) / np.float32(sample_rate) * np.float32(frequency)
    want = np.sin(omega * i, dtype=np.float32)
    got = c.start_rendering_sync().get_channel_data(0)
    assert np.abs(got - want).max() <= threshold
