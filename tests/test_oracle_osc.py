"""Pins the oracle's OscillatorRenderer against the reference's own unit tests (src/node/oscillator.rs:806-1455); every
function names the `#[test]` it restates and takes any backend (tests/test_gpu_reference_cases.py reruns them on CUDA).

The reference's square_raw / sawtooth_raw expectations hold only under `cfg!(test)`, which switches polyBLEP off
(oscillator.rs:592,600,603); the oracle restates the release build (polyBLEP on), so those two compare away from the
band-limited steps only (see also test_oracle_kat.py::test_oscillator_polyblep_values)."""
import numpy as np

SR = 44100
TWO_PI = 2.0 * np.pi


def _render(pkg, be, length, sr, setup):
    c = pkg.OfflineAudioContext(1, length, float(sr), be)
    setup(c)
    return c.start_rendering_sync().get_channel_data(0)


def _accumulated_phase(n, incr, start=0.0, wrap=True):
    out = np.empty(n, np.float64)
    phase = start
    for i in range(n):
        out[i] = phase
        phase += incr
        if wrap and phase >= 1.0:
            phase -= 1.0
    return out


def _osc(c, freq, type_=None, start=0.0, **kw):
    o = c.create_oscillator(**kw)
    o.connect(c.destination())
    o.frequency.set_value(freq)
    if type_ is not None:
        o.set_type(type_)
    o.start_at(start)
    return o


def periodic_wave_table(real, imag, normalize, size=2048):
    """PeriodicWave::generate_wavetable + normalize (src/periodic_wave.rs:163-209), f32 arithmetic."""
    real, imag = np.asarray(real, np.float32), np.asarray(imag, np.float32)
    i = np.arange(size, dtype=np.float32)
    phase = np.float32(2.0) * np.float32(np.pi) * i / np.float32(size)
    table = np.zeros(size, np.float32)
    for j in range(1, len(real)):
        rad = phase * np.float32(j)
        table = table + (real[j] * np.cos(rad, dtype=np.float32) + imag[j] * np.sin(rad, dtype=np.float32))
    if normalize:
        m = np.abs(table).max()
        if m > 0:
            table = table * (np.float32(1.0) / m)
    return table.astype(np.float32)


def test_sine_raw(pkg, oracle):  # :806-840 sine_raw, :842-869 sine_raw_exact_phase
    for i in range(5):
        freq = float(np.float32(10.0) ** np.float32(i))
        out = _render(pkg, oracle, SR, SR, lambda c: _osc(c, freq))
        want = np.sin(_accumulated_phase(SR, freq / SR) * TWO_PI).astype(np.float32)
        assert np.abs(out - want).max() <= 1e-5
        exact = np.sin(freq * np.arange(SR, dtype=np.float64) / SR * TWO_PI).astype(np.float32)
        assert np.abs(out - exact).max() <= 1e-5


def test_square_and_sawtooth_raw_away_from_the_steps(pkg, oracle):  # :871-907 square_raw, :956-996 sawtooth_raw
    for i in range(5):
        freq = float(np.float32(10.0) ** np.float32(i))
        incr = freq / SR
        ph = _accumulated_phase(SR, incr)
        edge = np.minimum.reduce([ph, np.abs(ph - 0.5), 1.0 - ph]) < 1.01 * incr   # inside a polyBLEP window
        if edge.all():  # 10 kHz: every frame lies in a window
            continue
        out = _render(pkg, oracle, SR, SR, lambda c: _osc(c, freq, pkg.SQUARE))
        assert np.abs(out - np.where(ph < 0.5, 1.0, -1.0))[~edge].max() <= 1e-6
        out = _render(pkg, oracle, SR, SR, lambda c: _osc(c, freq, pkg.SAWTOOTH))
        off = np.where(ph + 0.5 >= 1.0, ph - 0.5, ph + 0.5)
        assert np.abs(out - (2.0 * off - 1.0))[~edge].max() <= 1e-6


def test_periodic_wave(pkg, oracle):  # :998-1045 periodic_wave_1f, :1047-1093 periodic_wave_2f
    for i in range(5):
        freq = float(np.float32(10.0) ** np.float32(i))
        ph = _accumulated_phase(SR, freq / SR)
        t1 = periodic_wave_table([0.0, 0.0], [0.0, 1.0], True)
        out = _render(pkg, oracle, SR, SR, lambda c: _osc(c, freq, periodic_wave=t1))
        assert np.abs(out - np.sin(ph * TWO_PI).astype(np.float32)).max() <= 1e-5
        t2 = periodic_wave_table([0.0, 0.0, 0.0], [0.0, 0.5, 0.5], False)
        out = _render(pkg, oracle, SR, SR, lambda c: _osc(c, freq, periodic_wave=t2))
        want = 0.5 * np.sin(ph * TWO_PI) + 0.5 * np.sin(2.0 * ph * TWO_PI)
        assert np.abs(out - want.astype(np.float32)).max() <= 1e-5


def test_sub_quantum_and_sub_sample_start(pkg, oracle):  # :1135-1165 osc_sub_quantum_start, :1167-1197 osc_sub_sample_start
    out = _render(pkg, oracle, SR, SR, lambda c: _osc(c, 1.25, start=2.0 / SR))
    want = np.concatenate([[0.0, 0.0], np.sin(_accumulated_phase(SR - 2, 1.25 / SR, wrap=False) * TWO_PI)]).astype(np.float32)
    assert np.abs(out - want).max() <= 1e-5
    sr = 96000
    incr = 1.0 / sr
    out = _render(pkg, oracle, sr, sr, lambda c: _osc(c, 1.0, start=1.3 / sr))
    want = np.concatenate([[0.0, 0.0], np.sin(_accumulated_phase(sr - 2, incr, start=0.7 * incr, wrap=False) * TWO_PI)]).astype(np.float32)
    assert np.abs(out - want).max() <= 1e-5


def test_sub_quantum_and_sub_sample_stop(pkg, oracle):  # :1199-1229 osc_sub_quantum_stop, :1278-1308 osc_sub_sample_stop
    for freq, stop, n_on in [(2345.6, 6.0, 6), (8910.1, 19.4, 20)]:
        freq = float(np.float32(freq))

        def setup(c):
            o = _osc(c, freq)
            o.stop_at(stop / SR)

        out = _render(pkg, oracle, SR, SR, setup)
        want = np.zeros(SR, np.float32)
        want[:n_on] = np.sin(_accumulated_phase(n_on, freq / SR, wrap=False) * TWO_PI)
        assert np.abs(out - want).max() <= 1e-5


def test_stop_disarms_future_start(pkg, oracle):  # :1231-1246
    def setup(c):
        o = c.create_oscillator()
        o.connect(c.destination())
        o.start_at(2.0 / SR)
        o.stop_at(0.0)

    assert np.array_equal(_render(pkg, oracle, 128, SR, setup), np.zeros(128, np.float32))


def test_start_in_the_past(pkg, oracle):  # :1310-1342: start_at(0) issued from a suspend callback at frame 128: phase 0 there
    freq = float(np.float32(8910.1))
    c = pkg.OfflineAudioContext(1, SR, float(SR), oracle)
    c.suspend_sync(128.0 / SR, lambda ctx: _osc(ctx, freq))
    out = c.start_rendering_sync().get_channel_data(0)
    want = np.zeros(SR, np.float32)
    want[128:] = np.sin(_accumulated_phase(SR - 128, freq / SR, wrap=False) * TWO_PI)
    assert np.abs(out - want).max() <= 1e-5


def test_computed_frequency_outside_nyquist_is_silent(pkg, oracle):  # :1344-1381 above / below nyquist, :1384-1406 re-entering
    for freq in (20000.0, -20000.0):
        def setup(c, freq=freq):
            o = _osc(c, freq)
            o.detune.set_value(1200.0)

        assert np.abs(_render(pkg, oracle, 128, SR, setup)).max() <= 1e-5

    def setup2(c):
        o = _osc(c, 20000.0)
        o.detune.set_value(2400.0)
        o.detune.set_value_at_time(0.0, 128.0 / SR)

    out = _render(pkg, oracle, 256, SR, setup2)
    assert np.abs(out[:128]).max() <= 1e-5 and np.isfinite(out[128:]).all() and np.any(out[128:] != 0.0)


def test_delayed_start_and_negative_frequency(pkg, oracle):  # :1409-1428 delayed start, :1430-1455 sine_negative_frequency
    out = _render(pkg, oracle, 256, SR, lambda c: _osc(c, 440.0, start=128.0 / SR))
    assert np.abs(out[:128]).max() <= 1e-5 and np.any(out[128:] != 0.0)
    out = _render(pkg, oracle, SR, SR, lambda c: _osc(c, -100.0))
    want = np.sin(-100.0 * np.arange(SR, dtype=np.float64) / SR * TWO_PI).astype(np.float32)
    assert np.abs(out - want).max() <= 1e-5
