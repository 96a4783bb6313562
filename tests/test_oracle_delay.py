"""Pins the oracle's DelayWriter / DelayReader against the reference's own unit tests (src/node/delay.rs:766-1200); every
function names the `#[test]` it restates and takes any backend (tests/test_gpu_reference_cases.py reruns them on the CUDA
engine)."""
import numpy as np

RQ = 128
SR = 48000.0


def _dirac_through_delay(pkg, be, length, max_delay, delay_time, ch=1, start=0.0, feedback_gain=None, data=None):
    c = pkg.OfflineAudioContext(ch, length, SR, be)
    d = c.create_delay(max_delay_time=max_delay)
    d.delay_time.set_value(delay_time)
    d.connect(c.destination())
    if feedback_gain is not None:
        g = c.create_gain()
        g.gain.set_value(feedback_gain)
        d.connect(g)
        g.connect(d)
    s = c.create_buffer_source(pkg.AudioBuffer(data if data is not None else [np.ones(1, np.float32)], SR))
    s.connect(d)
    s.start_at(start)
    return c.start_rendering_sync()


def _expect(length, **at):
    e = np.zeros(length, np.float32)
    for k, v in at.items():
        e[int(k[1:])] = v
    return e


def test_sample_accurate(pkg, oracle):  # :766-792
    for n in (128, 131, 197):
        out = _dirac_through_delay(pkg, oracle, 256, 2.0, n / SR).get_channel_data(0)
        assert np.abs(out - _expect(256, **{f"i{n}": 1.0})).max() <= 1e-5


def test_sub_sample_accurate(pkg, oracle):  # :794-848 test_sub_sample_accurate_1 / _2
    out = _dirac_through_delay(pkg, oracle, 256, 2.0, 128.5 / SR).get_channel_data(0)
    assert np.abs(out - _expect(256, i128=0.5, i129=0.5)).max() <= 1e-5
    out = _dirac_through_delay(pkg, oracle, 256, 2.0, 128.8 / SR).get_channel_data(0)
    assert np.abs(out - _expect(256, i128=0.2, i129=0.8)).max() <= 1e-5


def test_multichannel(pkg, oracle):  # :850-881
    two = [np.zeros(256, np.float32), np.zeros(256, np.float32)]
    two[0][0] = 1.0
    two[1][1] = 1.0
    res = _dirac_through_delay(pkg, oracle, 256, 2.0, 128 / SR, ch=2, data=two)
    assert np.abs(res.get_channel_data(0) - _expect(256, i128=1.0)).max() <= 1e-5
    assert np.abs(res.get_channel_data(1) - _expect(256, i129=1.0)).max() <= 1e-5


def test_input_number_of_channels_change(pkg, oracle):  # :883-924: a mono source, then a stereo one, into the same delay
    c = pkg.OfflineAudioContext(2, 3 * RQ, SR, oracle)
    d = c.create_delay(max_delay_time=2.0)
    d.delay_time.set_value(128 / SR)
    d.connect(c.destination())
    one = np.zeros(128, np.float32)
    one[0] = 1.0
    s1 = c.create_buffer_source(pkg.AudioBuffer([one], SR))
    s1.connect(d)
    s1.start_at(0.0)
    two = [np.zeros(256, np.float32), np.zeros(256, np.float32)]
    two[0][0] = 1.0
    two[1][1] = 1.0
    s2 = c.create_buffer_source(pkg.AudioBuffer(two, SR))
    s2.connect(d)
    s2.start_at(128 / SR)
    res = c.start_rendering_sync()
    assert np.abs(res.get_channel_data(0) - _expect(384, i128=1.0, i256=1.0)).max() <= 1e-5
    assert np.abs(res.get_channel_data(1) - _expect(384, i128=1.0, i257=1.0)).max() <= 1e-5


def test_node_stays_alive_long_enough(pkg, oracle):  # :926-960 (the control handles are dropped before rendering)
    out = _dirac_through_delay(pkg, oracle, 5 * RQ, 1.0, 128 / SR, start=128 * 3 / SR).get_channel_data(0)
    assert np.abs(out - _expect(5 * RQ, i512=1.0)).max() <= 1e-5


def test_subquantum_delay(pkg, oracle):  # :962-988: every delay below one render quantum
    for i in range(128):
        out = _dirac_through_delay(pkg, oracle, 128, 1.0, float(np.float32(i) / np.float32(SR))).get_channel_data(0)
        assert np.abs(out - _expect(128, **{f"i{i}": 1.0})).max() <= 1e-5, i


def test_min_delay_when_in_loop(pkg, oracle):  # :990-1023: inside a cycle the delay is at least one quantum
    out = _dirac_through_delay(pkg, oracle, 256, 1.0, 1 / SR, feedback_gain=0.0).get_channel_data(0)
    assert np.array_equal(out, _expect(256, i128=1.0))


def test_max_delay(pkg, oracle):  # :1025-1074: delayTime == maxDelayTime, 1 s and 1.5 s at 44.1 kHz, bit-exact copy
    sr = 44100.0
    tone = np.sin(np.float32(20.0) * np.float32(2.0) * np.float32(np.pi) * np.arange(2 * 44100, dtype=np.float32) / np.float32(sr)).astype(np.float32)
    for seconds in (1.0, 1.5):
        c = pkg.OfflineAudioContext(1, 4 * 44100, sr, oracle)
        s = c.create_buffer_source(pkg.AudioBuffer([tone], sr))
        d = c.create_delay(max_delay_time=seconds)
        d.delay_time.set_value(seconds)
        s.connect(d)
        d.connect(c.destination())
        s.start_at(0.0)
        out = c.start_rendering_sync().get_channel_data(0)
        n0 = int(seconds * sr)
        assert np.all(out[:n0] == 0.0)
        assert np.array_equal(out[n0:n0 + len(tone)], tone)
        assert np.all(out[n0 + len(tone):] == 0.0)


def test_max_delay_smaller_than_quantum_size(pkg, oracle):  # :1076-1119 (in a loop: clamped up to one quantum)
    out = _dirac_through_delay(pkg, oracle, 256, 64 / SR, 64 / SR, feedback_gain=0.0).get_channel_data(0)
    assert np.array_equal(out, _expect(256, i128=1.0))


def test_max_delay_multiple_of_quantum_size(pkg, oracle):  # :1121-1175 _1 / _2
    out = _dirac_through_delay(pkg, oracle, 256, 128 / SR, 128 / SR).get_channel_data(0)
    assert np.abs(out - _expect(256, i128=1.0)).max() <= 1e-5
    out = _dirac_through_delay(pkg, oracle, 384, 256 / SR, 256 / SR).get_channel_data(0)
    assert np.abs(out - _expect(384, i256=1.0)).max() <= 1e-5


def test_subquantum_delay_dynamic_lifetime(pkg, oracle):  # :1177-1200
    c = pkg.OfflineAudioContext(1, 3 * RQ, SR, oracle)
    d = c.create_delay(max_delay_time=1.0)
    d.delay_time.set_value(float(np.float32(64.0) / np.float32(SR)))
    d.connect(c.destination())
    s = c.create_constant_source()
    s.connect(d)
    s.start_at(0.0)
    s.stop_at(120.0 / SR)
    out = c.start_rendering_sync().get_channel_data(0)
    want = np.zeros(3 * RQ, np.float32)
    want[64:64 + 120] = 1.0
    assert np.abs(out - want).max() <= 1e-5
