"""Third statements.  The oracle and the library were written by the same hand from the same reading of the reference; wherever a published
formula or a stock numerical routine exists, the oracle is checked here against THAT instead: the W3C Web Audio specification's biquad
coefficient formulas run through scipy.signal.lfilter in f64, scipy's IIR recursion, numpy's FFT for the analyser's spectrum (Blackman window
alpha 0.16, |X| / N, smoothing, dB: src/analysis.rs:14-24,281-369), closed forms of the AudioParam curves.  None of this code was derived from
oracle/ or csrc/.  (The GPU suite compares the CUDA engine with the oracle on the same kinds of graphs: tests/test_gpu_parity.py.)"""
import numpy as np
import pytest

scipy_signal = pytest.importorskip("scipy.signal")

RQ = 128


def _render_through(pkg, be, sr, x, make_node):
    n = -(-len(x) // RQ) * RQ
    c = pkg.OfflineAudioContext(1, n, sr, be)
    src = c.create_buffer_source(pkg.AudioBuffer([np.asarray(x, np.float32)], sr))
    node = make_node(c)
    src.connect(node)
    node.connect(c.destination())
    src.start()
    return c.start_rendering_sync().get_channel_data(0)[:len(x)].astype(np.float64)


def w3c_biquad(kind, fs, f0, q, gain_db):
    """https://webaudio.github.io/web-audio-api/#filters-characteristics (b0, b1, b2), (a0, a1, a2)"""
    a = 10.0 ** (gain_db / 40.0)
    w0 = 2.0 * np.pi * f0 / fs
    cos, sin = np.cos(w0), np.sin(w0)
    aq = sin / (2.0 * q)
    aqdb = sin / (2.0 * 10.0 ** (q / 20.0))
    a_s = sin / 2.0 * np.sqrt((a + 1.0 / a) * (1.0 / 1.0 - 1.0) + 2.0)
    r = np.sqrt(a)
    if kind == "lowpass":
        return ((1 - cos) / 2, 1 - cos, (1 - cos) / 2), (1 + aqdb, -2 * cos, 1 - aqdb)
    if kind == "highpass":
        return ((1 + cos) / 2, -(1 + cos), (1 + cos) / 2), (1 + aqdb, -2 * cos, 1 - aqdb)
    if kind == "bandpass":
        return (aq, 0.0, -aq), (1 + aq, -2 * cos, 1 - aq)
    if kind == "notch":
        return (1.0, -2 * cos, 1.0), (1 + aq, -2 * cos, 1 - aq)
    if kind == "allpass":
        return (1 - aq, -2 * cos, 1 + aq), (1 + aq, -2 * cos, 1 - aq)
    if kind == "peaking":
        return (1 + aq * a, -2 * cos, 1 - aq * a), (1 + aq / a, -2 * cos, 1 - aq / a)
    if kind == "lowshelf":
        return ((a * ((a + 1) - (a - 1) * cos + 2 * a_s * r), 2 * a * ((a - 1) - (a + 1) * cos), a * ((a + 1) - (a - 1) * cos - 2 * a_s * r)),
                ((a + 1) + (a - 1) * cos + 2 * a_s * r, -2 * ((a - 1) + (a + 1) * cos), (a + 1) + (a - 1) * cos - 2 * a_s * r))
    if kind == "highshelf":
        return ((a * ((a + 1) + (a - 1) * cos + 2 * a_s * r), -2 * a * ((a - 1) + (a + 1) * cos), a * ((a + 1) + (a - 1) * cos - 2 * a_s * r)),
                ((a + 1) - (a - 1) * cos + 2 * a_s * r, 2 * ((a - 1) - (a + 1) * cos), (a + 1) - (a - 1) * cos - 2 * a_s * r))
    raise ValueError(kind)


KINDS = ["lowpass", "highpass", "bandpass", "notch", "allpass", "peaking", "lowshelf", "highshelf"]


@pytest.mark.parametrize("kind", KINDS)
def test_biquad_render_vs_the_specification_and_scipy(pkg, oracle, kind):
    rng = np.random.default_rng(KINDS.index(kind))
    sr = 48000.0
    x = rng.uniform(-1, 1, 4096).astype(np.float32)
    for f0, q, gain in [(350.0, 1.0, 0.0), (1234.5, 4.0, 6.0), (9000.0, 0.7, -9.0), (60.0, 12.0, 3.0)]:
        b, a = w3c_biquad(kind, sr, f0, q, gain)
        want = scipy_signal.lfilter(np.array(b) / a[0], np.array(a) / a[0], x.astype(np.float64))
        got = _render_through(pkg, oracle, sr, x, lambda c: c.create_biquad_filter(type_=KINDS.index(kind), frequency=f0, q=q, gain=gain))
        # f64 recursion on both sides, f32 samples out: a few ulp of the output's magnitude
        assert np.abs(got - want).max() <= 4e-7 * max(1.0, np.abs(want).max()), (kind, f0, q, gain)


def test_biquad_detune_is_a_frequency_ratio(pkg, oracle):
    # computedFrequency = frequency * 2^(detune / 1200) (spec; biquad_filter.rs:753-757)
    rng = np.random.default_rng(9)
    sr = 44100.0
    x = rng.uniform(-1, 1, 2048).astype(np.float32)
    b, a = w3c_biquad("lowpass", sr, 500.0 * 2.0 ** (700.0 / 1200.0), 2.0, 0.0)
    want = scipy_signal.lfilter(np.array(b) / a[0], np.array(a) / a[0], x.astype(np.float64))
    got = _render_through(pkg, oracle, sr, x, lambda c: c.create_biquad_filter(type_=0, frequency=500.0, q=2.0, detune=700.0))
    assert np.abs(got - want).max() <= 2e-6


@pytest.mark.parametrize("order", [1, 2, 3, 5, 8])
def test_iir_render_vs_scipy(pkg, oracle, order):
    rng = np.random.default_rng(100 + order)
    sr = 48000.0
    x = rng.uniform(-1, 1, 3000).astype(np.float32)
    b, a = scipy_signal.butter(order, 0.2 + 0.05 * order)          # a stable filter of that order
    want = scipy_signal.lfilter(b, a, x.astype(np.float64))
    got = _render_through(pkg, oracle, sr, x, lambda c: c.create_iir_filter(list(b), list(a)))
    assert np.abs(got - want).max() <= 1e-6
    # unnormalised feedback[0] (iir_filter.rs:282-309 divides everything by it)
    got = _render_through(pkg, oracle, sr, x, lambda c: c.create_iir_filter(list(b * 2.5), list(a * 2.5)))
    assert np.abs(got - want).max() <= 1e-6


@pytest.mark.parametrize("fft_size", [256, 2048])
def test_analyser_spectrum_vs_numpy(pkg, oracle, fft_size):
    # the most recent fft_size frames, Blackman window (alpha 0.16, a0 - a1 cos(2 pi i / N) + a2 cos(4 pi i / N): the periodic form),
    # |rfft| / N, smoothed from an all-zero history: tau * 0 + (1 - tau) * |X_k| / N, in dB
    import test_oracle_analyser as AN
    rng = np.random.default_rng(fft_size)
    sr = 44100.0
    sig = (0.5 * np.sin(2 * np.pi * 1000.0 * np.arange(fft_size + 384) / sr) + 0.1 * rng.uniform(-1, 1, fft_size + 384)).astype(np.float32)
    for tau in (0.0, 0.8):
        a, _c = AN._analyse(pkg, oracle, sig, sr, fft_size=fft_size, smoothing_time_constant=tau)
        got = a.get_float_frequency_data()
        window = sig[len(sig) - fft_size:].astype(np.float64)   # (len(sig) is a whole number of quanta)
        i = np.arange(fft_size)
        w = 0.42 - 0.5 * np.cos(2 * np.pi * i / fft_size) + 0.08 * np.cos(4 * np.pi * i / fft_size)
        mag = np.abs(np.fft.rfft(window * w))[:fft_size // 2] / fft_size
        want = 20 * np.log10((1 - tau) * mag)
        assert len(got) == fft_size // 2
        loud = want > -90.0
        assert loud.sum() > fft_size // 8
        assert np.abs(got[loud] - want[loud]).max() <= 2e-2, tau   # f32 transform vs f64: hundredths of a dB at -90 dB


def test_audioparam_curves_vs_closed_forms(pkg, oracle):
    # spec section 1.6 (and param.rs:64-140): v(t) of a linear ramp, an exponential ramp and setTarget, sampled per frame (a-rate gain)
    sr = 48000.0
    n = RQ * 40
    t = np.arange(n) / sr
    t0, t1 = 0.01, 0.09

    def run(schedule):
        c = pkg.OfflineAudioContext(1, n, sr, oracle)
        s = c.create_constant_source(1.0)
        g = c.create_gain(1.0)
        schedule(g.gain)
        s.connect(g)
        g.connect(c.destination())
        s.start()
        return c.start_rendering_sync().get_channel_data(0).astype(np.float64)
    got = run(lambda p: p.set_value_at_time(0.25, t0).linear_ramp_to_value_at_time(2.0, t1))
    want = np.where(t < t0, 1.0, np.where(t < t1, 0.25 + (2.0 - 0.25) * (t - t0) / (t1 - t0), 2.0))
    assert np.abs(got - want).max() <= 2e-6
    got = run(lambda p: p.set_value_at_time(0.25, t0).exponential_ramp_to_value_at_time(2.0, t1))
    want = np.where(t < t0, 1.0, np.where(t < t1, 0.25 * (2.0 / 0.25) ** ((t - t0) / (t1 - t0)), 2.0))
    assert np.abs(got - want).max() <= 2e-6
    tc = 0.02
    got = run(lambda p: p.set_value_at_time(0.25, 0.0).set_target_at_time(2.0, t0, tc))
    curve = 2.0 + (0.25 - 2.0) * np.exp(-(t - t0) / tc)
    want = np.where(t < t0, 0.25, curve)
    late = t > t0 + 0.5 / sr   # (the frame AT t0 belongs to whichever side the reference's accumulated clock puts it)
    assert np.abs(got - want)[late].max() <= 2e-6 and np.all(got[:RQ] == 0.25)
    # Between the first quantum and the start time the reference does NOT hold the previous value.  The quantum that consumes the event
    # before the setTarget (here the first one) goes on to the pending setTarget, fills its remaining frames with the held value and then
    # stores the curve evaluated at the NEXT block's time as the intrinsic value (param.rs:1353-1355, 1372-1405) — with t < t0 the exponent
    # is positive: 2 - 1.75 e^{+(t0 - 128 / sr) / tc} = -0.525.  The following quanta are "constant blocks" (event.time >= next_block_time,
    # param.rs:1530-1548) and output that intrinsic value, and so do the waiting frames of the quantum the curve starts in.  The oracle
    # (and the engine's param core, tests/test_param_timeline.py) follow the reference here, not the specification's "hold".
    waiting = (t < t0 - 0.5 / sr) & (np.arange(n) >= RQ)
    quirk = 2.0 + (0.25 - 2.0) * np.exp(-(RQ / sr - t0) / tc)
    assert waiting.sum() > 300 and quirk < -0.5 and np.abs(got - quirk)[waiting].max() <= 2e-6
