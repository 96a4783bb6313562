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


def _render_channels(pkg, be, sr, channels, make_node, out_channels=2):
    n = -(-len(channels[0]) // RQ) * RQ
    c = pkg.OfflineAudioContext(out_channels, n, sr, be)
    src = c.create_buffer_source(pkg.AudioBuffer([np.asarray(ch, np.float32) for ch in channels], sr))
    node = make_node(c)
    src.connect(node)
    node.connect(c.destination())
    src.start()
    a = c.start_rendering_sync()
    return np.array([a.get_channel_data(i)[:len(channels[0])] for i in range(out_channels)], np.float64)


@pytest.mark.parametrize("delay_frames", [0.0, 1.0, 40.25, 127.5, 128.0, 300.75, 1000.5])
def test_delay_vs_linear_interpolation(pkg, oracle, delay_frames):
    # spec: y(t) = x(t - delayTime); between samples the reference interpolates linearly (delay.rs:640-700)
    rng = np.random.default_rng(5)
    sr = 48000.0
    x = rng.uniform(-1, 1, 2048).astype(np.float32)
    got = _render_through(pkg, oracle, sr, x, lambda c: c.create_delay(1.0, delay_frames / sr))
    d = float(np.float32(delay_frames / sr)) * sr   # delayTime is an f32 AudioParam
    k = int(np.floor(d))
    f = d - k
    xp = np.concatenate([np.zeros(k + 2), x.astype(np.float64)])
    idx = np.arange(len(x)) + 2
    want = (1 - f) * xp[idx] + f * xp[idx - 1]
    assert np.abs(got - want).max() <= 2e-6, delay_frames


@pytest.mark.parametrize("pan", [-1.0, -0.3, 0.0, 0.45, 1.0])
def test_stereo_panner_vs_the_specification(pkg, oracle, pan):
    # https://webaudio.github.io/web-audio-api/#stereopanner-algorithm
    rng = np.random.default_rng(6)
    sr = 48000.0
    left, right = rng.uniform(-1, 1, (2, 512)).astype(np.float32)
    got = _render_channels(pkg, oracle, sr, [left], lambda c: c.create_stereo_panner(pan))
    x = (pan + 1) / 2
    assert np.abs(got[0] - left * np.cos(x * np.pi / 2)).max() <= 2e-7 and np.abs(got[1] - left * np.sin(x * np.pi / 2)).max() <= 2e-7
    got = _render_channels(pkg, oracle, sr, [left, right], lambda c: c.create_stereo_panner(pan))
    x = pan + 1 if pan <= 0 else pan
    gl, gr = np.cos(x * np.pi / 2), np.sin(x * np.pi / 2)
    if pan <= 0:
        want = [left + right * gl, right * gr]
    else:
        want = [left * gl, right + left * gr]
    assert np.abs(got[0] - want[0]).max() <= 4e-7 and np.abs(got[1] - want[1]).max() <= 4e-7


@pytest.mark.parametrize("points", [2, 3, 64, 1025])
def test_wave_shaper_vs_the_specification(pkg, oracle, points):
    # https://webaudio.github.io/web-audio-api/#WaveShaperNode-attributes: v = (N - 1) / 2 * (x + 1), k = floor(v), f = v - k,
    # y = (1 - f) curve[k] + f curve[k + 1]; x <= -1 -> curve[0], x >= 1 -> curve[N - 1]
    rng = np.random.default_rng(points)
    sr = 48000.0
    curve = rng.uniform(-1, 1, points).astype(np.float32)
    x = np.concatenate([rng.uniform(-1.3, 1.3, 1000), [-1.0, 1.0, 0.0, -2.0, 2.0]]).astype(np.float32)
    got = _render_through(pkg, oracle, sr, x, lambda c: c.create_wave_shaper(curve=curve))
    v = (points - 1) / 2 * (x.astype(np.float64) + 1)
    k = np.clip(np.floor(v).astype(int), 0, points - 2)
    f = v - k
    want = (1 - f) * curve[k] + f * curve[k + 1]
    want = np.where(x <= -1, curve[0], np.where(x >= 1, curve[-1], want))
    # the reference computes v in f32: half an ulp of v (up to N - 1) times the local slope of the curve (up to 2 per point)
    assert np.abs(got - want).max() <= 1e-6 + 2.5e-7 * points


@pytest.mark.parametrize("model", ["linear", "inverse", "exponential"])
def test_panner_distance_gain_vs_the_specification(pkg, oracle, model):
    # https://webaudio.github.io/web-audio-api/#enumdef-distancemodeltype; source straight ahead of the default listener (azimuth 0): the
    # equal-power law gives a mono input cos(pi / 4) on both sides
    sr = 48000.0
    x = np.full(256, 0.5, np.float32)
    ref, mx, roll = 2.0, 40.0, 0.8 if model == "linear" else 1.7
    for d in [0.5, 2.0, 7.0, 39.0, 90.0]:
        got = _render_channels(pkg, oracle, sr, [x], lambda c: c.create_panner(distance_model={"linear": pkg.LINEAR, "inverse": pkg.INVERSE, "exponential": pkg.EXPONENTIAL}[model],
                                                                           position=(0.0, 0.0, -d), ref_distance=ref, max_distance=mx, rolloff_factor=roll))
        if model == "linear":
            g = 1 - roll * (min(max(d, ref), mx) - ref) / (mx - ref)
        elif model == "inverse":
            g = ref / (ref + roll * (max(d, ref) - ref))
        else:
            g = (max(d, ref) / ref) ** (-roll)
        want = 0.5 * np.cos(np.pi / 4) * g
        assert np.abs(got - want).max() <= 1e-6, (model, d)


def test_convolver_normalisation_vs_the_specification(pkg, oracle):
    # https://webaudio.github.io/web-audio-api/#dom-convolvernode-normalize (calculateNormalizationScale)
    import ctypes as C
    import importlib
    B = importlib.import_module(pkg.__name__ + "._binding")
    rng = np.random.default_rng(12)
    for n_ch, length, sr, amp in [(1, 1000, 44100.0, 0.3), (2, 4096, 48000.0, 1.0), (4, 777, 96000.0, 0.01), (1, 64, 48000.0, 1e-7)]:
        data = (amp * rng.uniform(-1, 1, (n_ch, length))).astype(np.float32)
        rows = [np.ascontiguousarray(r) for r in data]
        ptrs = (C.POINTER(C.c_float) * n_ch)(*[B.fptr(r) for r in rows])
        got = oracle.api.convolver_normalize(C.byref(B.AudioBufferDesc(n_ch, length, sr, ptrs)))
        power = np.sqrt(np.sum(data.astype(np.float64) ** 2) / (n_ch * length))
        power = max(power, 0.000125)
        want = 1 / power * 0.00125 * (44100.0 / sr) * (0.5 if n_ch == 4 else 1.0)
        assert abs(got - want) <= 2e-6 * want, (n_ch, length)


def _spec_azimuth(src, listener=(0.0, 0.0, 0.0), forward=(0.0, 0.0, -1.0), up=(0.0, 1.0, 0.0)):
    """https://webaudio.github.io/web-audio-api/#azimuth-elevation"""
    s = np.array(src, float) - np.array(listener, float)
    if not s.any():
        return 0.0
    s /= np.linalg.norm(s)
    f, u = np.array(forward, float), np.array(up, float)
    right = np.cross(f, u)
    right /= np.linalg.norm(right)
    f /= np.linalg.norm(f)
    up2 = np.cross(right, f)
    proj = s - np.dot(s, up2) * up2
    if not proj.any():
        return 0.0
    proj /= np.linalg.norm(proj)
    az = np.degrees(np.arccos(np.clip(np.dot(proj, right), -1, 1)))
    if np.dot(proj, f) < 0:
        az = 360 - az
    return 90 - az if 0 <= az <= 270 else 450 - az


@pytest.mark.parametrize("stereo", [False, True])
def test_equal_power_panner_vs_the_specification(pkg, oracle, stereo):
    # https://webaudio.github.io/web-audio-api/#Spatialization-equal-power-panning; sources on the unit circle and off the horizontal plane
    # (distance 1 = refDistance: no distance attenuation, full cones)
    rng = np.random.default_rng(21)
    sr = 48000.0
    left, right = (0.5 * rng.uniform(-1, 1, (2, 256))).astype(np.float32)
    for pos in [(0.0, 0.0, -1.0), (1.0, 0.0, 0.0), (-1.0, 0.0, 0.0), (0.6, 0.0, -0.8), (-0.6, 0.0, 0.8), (0.0, 0.0, 1.0), (0.5, 0.7, -0.5), (-0.3, -0.4, 0.2)]:
        p = np.array(pos) / np.linalg.norm(pos)
        got = _render_channels(pkg, oracle, sr, [left, right] if stereo else [left], lambda c: c.create_panner(position=tuple(float(v) for v in p)))
        az = float(np.clip(_spec_azimuth(p), -180, 180))
        az = -180 - az if az < -90 else (180 - az if az > 90 else az)
        if not stereo:
            x = (az + 90) / 180
            want = [left * np.cos(x * np.pi / 2), left * np.sin(x * np.pi / 2)]
        else:
            x = (az + 90) / 90 if az <= 0 else az / 90
            gl, gr = np.cos(x * np.pi / 2), np.sin(x * np.pi / 2)
            want = [left + right * gl, right * gr] if az <= 0 else [left * gl, right + left * gr]
        assert np.abs(got[0] - want[0]).max() <= 2e-6 and np.abs(got[1] - want[1]).max() <= 2e-6, (pos, az)


def test_panner_cone_gain_vs_the_specification(pkg, oracle):
    # https://webaudio.github.io/web-audio-api/#Spatialization-sound-cones — the gain law as specified, the ANGLE as the reference measures it:
    # spatial.rs:277-299 takes the angle between the source's orientation and the vector from the LISTENER to the SOURCE (source_position -
    # listener_position), where the specification uses the vector from the source to the listener: a source facing away from the listener is
    # "on axis" for the reference.  Oracle and engine follow the reference (the first attempt at this test, written from the specification,
    # found the outer gain at angle 0).  Source at (0, 0, -1); `angle` degrees between its orientation and (0, 0, -1).
    sr = 48000.0
    x = np.full(256, 0.5, np.float32)
    inner, outer, outer_gain = 60.0, 200.0, 0.25
    for angle in [0.0, 20.0, 30.0, 65.0, 99.9, 100.0, 150.0, 180.0]:
        a = np.radians(angle)
        orientation = (float(np.sin(a)), 0.0, float(-np.cos(a)))
        got = _render_channels(pkg, oracle, sr, [x], lambda c: c.create_panner(position=(0.0, 0.0, -1.0), orientation=orientation, cone_inner_angle=inner,
                                                                           cone_outer_angle=outer, cone_outer_gain=outer_gain))
        if angle <= inner / 2:
            g = 1.0
        elif angle >= outer / 2:
            g = outer_gain
        else:
            t = (angle - inner / 2) / (outer / 2 - inner / 2)
            g = (1 - t) + outer_gain * t
        want = 0.5 * np.cos(np.pi / 4) * g
        assert np.abs(got - want).max() <= 3e-6, (angle, g)


def test_value_curve_vs_the_specification(pkg, oracle):
    # https://webaudio.github.io/web-audio-api/#dom-audioparam-setvaluecurveattime: k = floor((N - 1) / T_D (t - T_0)),
    # v(t) = V[k] + (V[k + 1] - V[k]) ((N - 1) / T_D (t - T_0) - k); after T_0 + T_D the last value holds
    sr = 48000.0
    n = RQ * 30
    t = np.arange(n) / sr
    values = np.array([0.0, 1.0, -0.5, 0.25, 0.75], np.float32)
    t0, dur = 0.001, 0.045   # (starts inside the first quantum: a curve that is still pending at the end of a quantum is sampled early by the
    #                           reference, like the pending setTarget above — tests/test_param_timeline.py::test_a_pending_value_curve_...)
    c = pkg.OfflineAudioContext(1, n, sr, oracle)
    s = c.create_constant_source(1.0)
    g = c.create_gain(0.5)
    g.gain.set_value_curve_at_time(values, t0, dur)
    s.connect(g)
    g.connect(c.destination())
    s.start()
    got = c.start_rendering_sync().get_channel_data(0).astype(np.float64)
    pos = (len(values) - 1) / dur * (t - t0)
    k = np.clip(np.floor(pos).astype(int), 0, len(values) - 2)
    curve = values[k] + (values[k + 1] - values[k]) * (pos - k)
    inside = (t > t0 + 1 / sr) & (t < t0 + dur - 1 / sr)
    assert np.all(got[t < t0 - 1 / sr] == 0.5)
    assert np.abs(got - curve)[inside].max() <= 1e-5
    assert np.all(got[t > t0 + dur + 1 / sr] == values[-1])


def test_sine_oscillator_and_detune_vs_numpy(pkg, oracle):
    # computedOscFrequency = frequency * 2^(detune / 1200); a sine of that frequency from phase 0 (the reference reads a 2048-point table with
    # linear interpolation: 2.4e-6 worst case for a sine + the f32 phase accumulation over the render)
    sr = 48000.0
    n = RQ * 32
    for f, det in [(440.0, 0.0), (1000.0, 700.0), (93.7, -1200.0)]:
        c = pkg.OfflineAudioContext(1, n, sr, oracle)
        o = c.create_oscillator(frequency=f, detune=det)
        o.connect(c.destination())
        o.start()
        got = c.start_rendering_sync().get_channel_data(0).astype(np.float64)
        want = np.sin(2 * np.pi * f * 2.0 ** (det / 1200.0) * np.arange(n) / sr)
        assert np.abs(got - want).max() <= 2e-4, (f, det)


def test_looping_buffer_source_vs_modular_indexing(pkg, oracle):
    # https://webaudio.github.io/web-audio-api/#playback-AudioBufferSourceNode at playbackRate 1 with sample-aligned loop points: the
    # playhead runs to loopEnd and wraps to loopStart
    rng = np.random.default_rng(31)
    sr = 48000.0
    buf = rng.uniform(-1, 1, 700).astype(np.float32)
    ls, le = 100, 420
    n = RQ * 12
    c = pkg.OfflineAudioContext(1, n, sr, oracle)
    s = c.create_buffer_source(pkg.AudioBuffer([buf], sr), loop=True, loop_start=ls / sr, loop_end=le / sr)
    s.connect(c.destination())
    s.start()
    got = c.start_rendering_sync().get_channel_data(0)
    idx = np.arange(n)
    idx = np.where(idx < le, idx, ls + (idx - le) % (le - ls))
    # (the frame AT every wrap belongs to whichever side the reference's accumulated f64 playhead puts it, interpolated: not asserted)
    at_wrap = (np.arange(n) >= le) & ((np.arange(n) - le) % (le - ls) == 0)
    assert at_wrap.sum() == 4 and np.array_equal(got[~at_wrap], buf[idx][~at_wrap])


def test_dynamics_compressor_vs_the_published_design(pkg, oracle):
    # The node's design is public: the specification's compression curve and makeup gain (full-range gain = the curve applied to 1.0, makeup =
    # its inverse to the power 0.6), the gain computer and the branching peak detector of Giannoulis, Massberg and Reiss, "Digital Dynamic Range
    # Compressor Design" (JAES 2012: eq. 4, 7, 16 — the paper dynamics_compressor.rs:355-360 cites), a 6 ms look-ahead in whole render quanta.
    # Restated in f64 from the paper and the spec (the reference works in f32: 1e-4 relative).
    rng = np.random.default_rng(77)
    sr = 48000.0
    n = RQ * 60
    t = np.arange(n) / sr
    x = (np.where((t > 0.02) & (t < 0.09), 0.9, 0.05) * np.sin(2 * np.pi * 330.0 * t) + 0.01 * rng.uniform(-1, 1, n)).astype(np.float32)
    for thr, knee, ratio, att, rel in [(-24.0, 30.0, 12.0, 0.003, 0.25), (-30.0, 0.0, 4.0, 0.01, 0.05), (-12.0, 6.0, 20.0, 0.001, 0.1)]:
        got = _render_through(pkg, oracle, sr, x, lambda c: c.create_dynamics_compressor(attack=att, knee=knee, ratio=ratio, release=rel, threshold=thr))
        T = thr + knee / 2 if knee > 0 else thr            # the knee is centred on the shifted threshold (paper's W around T)
        a_tau, r_tau = np.exp(-1 / (att * sr)), np.exp(-1 / (rel * sr))
        makeup_db = 20 * np.log10((1 / 10 ** ((T - T / ratio) / 20)) ** 0.6)
        xs = np.abs(x.astype(np.float64))
        xg = np.where(xs == 0, -1000.0, 20 * np.log10(np.maximum(xs, 1e-300)))
        yg = np.where(2 * (xg - T) < -knee, xg, np.where(2 * np.abs(xg - T) <= knee, xg + (1 / ratio - 1) * (xg - T + knee / 2) ** 2 / (2 * knee if knee > 0 else 1), T + (xg - T) / ratio))
        xl = xg - yg
        gain = np.empty(n)
        yl = 0.0
        for i in range(n):
            yl = a_tau * yl + (1 - a_tau) * xl[i] if xl[i] > yl else r_tau * yl + (1 - r_tau) * xl[i]
            gain[i] = 10 ** ((-yl + makeup_db) / 20)
        delay = (int(np.ceil(sr * 0.006 / RQ)) + 1 - 1) * RQ   # ring of ceil(6 ms / quantum) + 1 quanta, read one slot ahead of the write
        delayed = np.concatenate([np.zeros(delay), x.astype(np.float64)])[:n]
        want = delayed * gain
        assert np.abs(got - want).max() <= 2e-4 * max(1.0, np.abs(want).max()), (thr, knee, ratio)
        assert np.abs(want).max() > 0.3   # (the loud burst came through, attenuated)


def _spec_mix(x, m):
    """https://webaudio.github.io/web-audio-api/#channel-up-mixing-and-down-mixing, "speakers": x [k][n] -> [m][n]; layouts 1, 2, 4 (L R SL SR),
    6 (L R C LFE SL SR); other combinations are "discrete" (copy what fits, zero the rest)"""
    k = len(x)
    z = np.zeros_like(x[0])
    s = np.sqrt(0.5)
    if k == m:
        return list(x)
    table = {
        (1, 2): lambda: [x[0], x[0]], (1, 4): lambda: [x[0], x[0], z, z], (1, 6): lambda: [z, z, x[0], z, z, z],
        (2, 4): lambda: [x[0], x[1], z, z], (2, 6): lambda: [x[0], x[1], z, z, z, z], (4, 6): lambda: [x[0], x[1], z, z, x[2], x[3]],
        (2, 1): lambda: [0.5 * (x[0] + x[1])], (4, 1): lambda: [0.25 * (x[0] + x[1] + x[2] + x[3])],
        (6, 1): lambda: [s * (x[0] + x[1]) + x[2] + 0.5 * (x[4] + x[5])],
        (4, 2): lambda: [0.5 * (x[0] + x[2]), 0.5 * (x[1] + x[3])],
        (6, 2): lambda: [x[0] + s * (x[2] + x[4]), x[1] + s * (x[2] + x[5])],
        (6, 4): lambda: [x[0] + s * x[2], x[1] + s * x[2], x[4], x[5]],
    }
    if (k, m) in table:
        return table[(k, m)]()
    return [x[i] if i < k else z for i in range(m)]


@pytest.mark.parametrize("k", [1, 2, 3, 4, 6])
def test_speaker_mixing_vs_the_specification(pkg, oracle, k):
    rng = np.random.default_rng(40 + k)
    sr = 48000.0
    x = rng.uniform(-1, 1, (k, RQ * 2)).astype(np.float32)
    for m in [1, 2, 3, 4, 6]:
        c = pkg.OfflineAudioContext(m, RQ * 2, sr, oracle)
        src = c.create_buffer_source(pkg.AudioBuffer(list(x), sr))
        port = c.create_gain(1.0, cfg=pkg.channel_config(m, pkg.EXPLICIT, pkg.SPEAKERS))
        src.connect(port)
        port.connect(c.destination())
        src.start()
        a = c.start_rendering_sync()
        got = np.array([a.get_channel_data(i) for i in range(m)], np.float64)
        want = np.array(_spec_mix(x.astype(np.float64), m))
        assert np.abs(got - want).max() <= 3e-7, (k, m)


@pytest.mark.parametrize("kind", KINDS)
def test_biquad_frequency_response_of_both_libraries_vs_the_specification(pkg, host_api, kind):
    # BiquadFilterNode::get_frequency_response — the coefficient code of the oracle AND of the product (csrc/wae_hostmath.h: what the planner
    # feeds the kernels with) — against H(e^{jw}) of the specification's coefficients (scipy.signal.freqz)
    import ctypes as C
    sr = 44100.0
    f = np.geomspace(20.0, 20000.0, 64).astype(np.float32)
    fp = C.POINTER(C.c_float)
    for f0, q, gain, det in [(350.0, 1.0, 0.0, 0.0), (2000.0, 7.0, 9.0, 300.0), (80.0, 0.5, -12.0, -700.0)]:
        mag = np.zeros(len(f), np.float32)
        phase = np.zeros(len(f), np.float32)
        host_api.biquad_frequency_response(KINDS.index(kind), sr, f0, det, q, gain, f.ctypes.data_as(fp), mag.ctypes.data_as(fp), phase.ctypes.data_as(fp), len(f))
        b, a = w3c_biquad(kind, sr, f0 * 2.0 ** (det / 1200.0), q, gain)
        _w, h = scipy_signal.freqz(np.array(b) / a[0], np.array(a) / a[0], worN=2 * np.pi * f.astype(np.float64) / sr)
        assert np.abs(mag - np.abs(h)).max() <= 2e-5 * max(1.0, np.abs(h).max()), (kind, f0)
        dphi = np.angle(np.exp(1j * (phase.astype(np.float64) - np.angle(h))))
        assert np.abs(dphi[np.abs(h) > 1e-4]).max() <= 2e-4, (kind, f0)


def test_spatial_math_of_both_libraries_vs_the_specification(host_api):
    # wae_spatial_params / wao_spatial_params: what the planner and the moving-source kernels evaluate (csrc/wae_spatial.h) and what the oracle's
    # PannerNode evaluates — against the specification's azimuth / elevation algorithm and distance models, random sources and listeners
    import ctypes as C
    dp, fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
    rng = np.random.default_rng(404)
    worst_az = worst_el = 0.0
    for _ in range(400):
        src = rng.uniform(-5, 5, 3)
        lp = rng.uniform(-2, 2, 3)
        fwd = rng.uniform(-1, 1, 3)
        fwd /= np.linalg.norm(fwd)
        up = np.cross(np.cross(fwd, rng.uniform(-1, 1, 3)), fwd)   # perpendicular to forward
        if np.linalg.norm(up) < 0.2 or np.linalg.norm(src - lp) < 0.3:
            continue
        up /= np.linalg.norm(up)
        ref, mx, roll = 0.5 + rng.uniform(0, 2), 6.0 + rng.uniform(0, 4), rng.uniform(0.1, 1.0)
        model = int(rng.integers(0, 3))   # 0 linear, 1 inverse, 2 exponential (include/wae.h)
        m6 = np.array([ref, mx, roll, 360.0, 360.0, 0.0], np.float64)
        v15 = np.concatenate([src, [0.0, 0.0, 0.0], lp, fwd, up]).astype(np.float32)
        out = np.zeros(4, np.float32)
        host_api.check(host_api.spatial_params(model, m6.ctypes.data_as(dp), v15.ctypes.data_as(fp), out.ctypes.data_as(fp)))
        s64, l64, f64, u64 = (v15[0:3].astype(np.float64), v15[6:9].astype(np.float64), v15[9:12].astype(np.float64), v15[12:15].astype(np.float64))
        d = np.linalg.norm(s64 - l64)
        if model == 0:
            g = 1 - roll * (min(max(d, ref), mx) - ref) / (mx - ref)
        elif model == 1:
            g = ref / (ref + roll * (max(d, ref) - ref))
        else:
            g = (max(d, ref) / ref) ** (-roll)
        assert abs(out[0] - g) <= 2e-5, (model, d, out[0], g)
        az = _spec_azimuth(s64, l64, f64, u64)
        right = np.cross(f64, u64)
        up2 = np.cross(right / np.linalg.norm(right), f64 / np.linalg.norm(f64))
        sl = (s64 - l64) / d
        el = 90 - np.degrees(np.arccos(np.clip(np.dot(sl, up2), -1, 1)))   # spec: 90 - angle(sourceListener, up), folded into [-90, 90]
        el = 180 - el if el > 90 else (-180 - el if el < -90 else el)
        horizontal = np.linalg.norm(sl - np.dot(sl, up2) * up2)
        if horizontal > 0.2:   # (the azimuth is ill-conditioned right above / below the listener)
            worst_az = max(worst_az, abs(np.angle(np.exp(1j * np.radians(out[2] - az)))))
        worst_el = max(worst_el, abs(out[3] - el))
    assert worst_az <= np.radians(0.02) and worst_el <= 0.02, (worst_az, worst_el)


def test_periodic_wave_tables_of_both_libraries_vs_fourier_synthesis(host_api):
    # https://webaudio.github.io/web-audio-api/#waveform-generation: x(t) = sum_{k >= 1} real[k] cos(2 pi k t) + imag[k] sin(2 pi k t) (the
    # DC terms are ignored), normalised to a peak of 1 unless disabled — numpy's inverse FFT against both libraries' wavetables
    import test_node_setters as NS
    rng = np.random.default_rng(77)
    n = 2048
    for harmonics in (2, 5, 33, 200):
        real = rng.uniform(-1, 1, harmonics).astype(np.float32)
        imag = rng.uniform(-1, 1, harmonics).astype(np.float32)
        spec = np.zeros(n // 2 + 1, np.complex128)
        spec[1:harmonics] = (real[1:].astype(np.float64) - 1j * imag[1:].astype(np.float64)) * n / 2
        want = np.fft.irfft(spec, n)
        got = NS._wave(host_api, real, imag, disable_normalization=True, n=n) if "n" in NS._wave.__code__.co_varnames else NS._wave(host_api, real, imag, disable_normalization=True)
        # the reference accumulates in f32 with phases up to 2 pi k: an ulp of the phase (k * 1.2e-7 * 2 pi) per term, k terms
        tol = 2e-5 + 2e-7 * harmonics * harmonics
        assert np.abs(got - want).max() <= tol, harmonics
        got = NS._wave(host_api, real, imag)
        assert np.abs(got - want / np.abs(want).max()).max() <= tol, harmonics


@pytest.mark.parametrize("rate,buffer_sr", [(0.5, 48000.0), (1.37, 48000.0), (2.0, 48000.0), (1.0, 32000.0), (0.8, 44100.0)])
def test_buffer_source_resampling_vs_linear_interpolation(pkg, oracle, rate, buffer_sr):
    # the playhead advances by playbackRate * bufferRate / contextRate buffer frames per output frame; between frames the reference interpolates
    # linearly (audio_buffer_source.rs:727-800) — numpy.interp at the same positions
    rng = np.random.default_rng(int(rate * 100))
    sr = 48000.0
    buf = np.cumsum(rng.uniform(-0.05, 0.05, 6000)).astype(np.float32)   # smooth enough that a 1e-9 playhead error stays below 1e-6
    n = RQ * 16
    c = pkg.OfflineAudioContext(1, n, sr, oracle)
    s = c.create_buffer_source(pkg.AudioBuffer([buf], buffer_sr), playback_rate=rate)
    s.connect(c.destination())
    s.start()
    got = c.start_rendering_sync().get_channel_data(0).astype(np.float64)
    pos = np.arange(n) * rate * buffer_sr / sr
    inside = pos < len(buf) - 1
    want = np.interp(pos[inside], np.arange(len(buf)), buf.astype(np.float64))
    assert inside.sum() > 1500 and np.abs(got[inside] - want).max() <= 2e-6
