"""Layouts that change during a render: the reference's per-quantum channel counts and its "silent" quanta.

In the reference an edge whose producer is silent — a source that has not started or has ended, a delay line or compressor ring that is
still empty, a filter whose tail has rung out — carries ONE silent channel (src/render/quantum.rs:109-111,512-517), so a stereo edge
changes its channel count over time and its consumers react: a biquad gains a channel that starts from zero state or drops one
(biquad_filter.rs:798-815), a delay line re-mixes its whole ring (delay.rs:470-488), panners switch between their mono and stereo laws,
a splitter output falls silent, the mixer folds the edges that exist (quantum.rs:532-569).  The engine follows this with per-quantum
layout tracks computed on the device (csrc/wae_device.h: BufRef::meta); these tests build the situations one by one and compare the CUDA
render with the oracle at the north_star tolerance."""
import numpy as np
import pytest

import graphs as G

pytestmark = pytest.mark.gpu
TOL = 1e-5
N = 128 * 60 + 17


def maxdiff(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())


def stereo_noise(seed, frames, amp=0.5):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-amp, amp, (2, frames))).astype(np.float32)


def check(pkg, engine, oracle, build, n=3, tol=TOL):
    gpu = G.render(pkg, [build(engine.backend, g) for g in range(n)])
    cpu = G.render(pkg, [build(oracle, g) for g in range(n)])
    assert np.isfinite(cpu).all()
    err = maxdiff(gpu, cpu)
    assert err <= tol, err
    return gpu, cpu


def finite_stereo(pkg, c, seed, frames, start=0.0):
    s = c.create_buffer_source(pkg.AudioBuffer(list(stereo_noise(seed, frames)), G.SR))
    s.start_at(start)
    return s


def mono_tone(pkg, c, f=330.0, start=0.0, stop=None):
    o = c.create_oscillator(frequency=f)
    o.start_at(start)
    if stop is not None:
        o.stop_at(stop)
    return o


def test_biquad_drops_a_channel_when_the_stereo_source_ends(pkg, engine, oracle):
    """stereo buffer (ends) + mono tone (goes on) -> biquad: at the end of the buffer the input turns mono, the filter truncates its
    state to one channel and the right output becomes the up-mix of the left (biquad_filter.rs:798-815)"""
    def build(be, g):
        c = pkg.OfflineAudioContext(2, N, G.SR, be)
        bq = c.create_biquad_filter(type_=pkg.LOWPASS, frequency=300.0 + 100 * g, q=6.0)
        finite_stereo(pkg, c, 10 + g, 128 * 20 + 31 * g).connect(bq)
        mono_tone(pkg, c, 220.0 + 10 * g).connect(bq)
        bq.connect(c.destination())
        return c
    gpu, _ = check(pkg, engine, oracle, build)
    assert np.array_equal(gpu[0, 0, 128 * 30:], gpu[0, 1, 128 * 30:])  # one channel left: L == R at the destination


def test_biquad_gains_a_channel_that_starts_from_zero_state(pkg, engine, oracle):
    def build(be, g):
        c = pkg.OfflineAudioContext(2, N, G.SR, be)
        bq = c.create_biquad_filter(type_=pkg.BANDPASS, frequency=500.0, q=12.0)
        mono_tone(pkg, c, 480.0).connect(bq)
        finite_stereo(pkg, c, 20 + g, 128 * 15, start=(128 * (10 + g) + 40) / G.SR).connect(bq)
        bq.connect(c.destination())
        return c
    check(pkg, engine, oracle, build)


def test_iir_and_automated_biquad_follow_the_channel_count(pkg, engine, oracle):
    def build(be, g):
        c = pkg.OfflineAudioContext(2, N, G.SR, be)
        iir = c.create_iir_filter([0.2, 0.3, 0.1], [1.0, -0.5, 0.3])
        bq = c.create_biquad_filter(type_=pkg.PEAKING, frequency=900.0, q=4.0, gain=6.0)
        bq.frequency.linear_ramp_to_value_at_time(2500.0, 0.1)
        src = finite_stereo(pkg, c, 30 + g, 128 * 18 + 5, start=(128 * 3) / G.SR)
        tone = mono_tone(pkg, c, 150.0)
        for f in (iir, bq):
            src.connect(f)
            tone.connect(f)
            f.connect(c.destination())
        return c
    check(pkg, engine, oracle, build)


def test_filter_tail_end_is_seen_by_the_next_filter(pkg, engine, oracle):
    """the moment a filter's tail has rung out (its f64 state holds no normal value) is data dependent; until then its output keeps two
    channels, afterwards it is one silent channel — which truncates the state of a slowly decaying filter behind it"""
    def build(be, g):
        c = pkg.OfflineAudioContext(2, 128 * 700, G.SR, be)
        fast = c.create_biquad_filter(type_=pkg.LOWPASS, frequency=9000.0, q=0.5)
        slow = c.create_biquad_filter(type_=pkg.BANDPASS, frequency=120.0, q=30.0)
        finite_stereo(pkg, c, 40 + g, 128 * 10 + 77).connect(fast)
        fast.connect(slow)
        mono_tone(pkg, c, 117.0).connect(slow)
        slow.connect(c.destination())
        return c
    check(pkg, engine, oracle, build, n=2)


def test_delay_line_collapses_to_mono_when_its_input_does(pkg, engine, oracle):
    def build(be, g):
        c = pkg.OfflineAudioContext(2, N, G.SR, be)
        d = c.create_delay(max_delay_time=0.2, delay_time=0.031 + 0.004 * g)
        finite_stereo(pkg, c, 50 + g, 128 * 22 + 9).connect(d)
        d.connect(c.destination())
        return c
    gpu, _ = check(pkg, engine, oracle, build)
    # what is still in the line when the source ends comes out as its mono down-mix
    assert np.array_equal(gpu[0, 0, 128 * 24:], gpu[0, 1, 128 * 24:])


def test_delay_line_with_a_late_stereo_source_and_a_tone(pkg, engine, oracle):
    def build(be, g):
        c = pkg.OfflineAudioContext(2, N, G.SR, be)
        d = c.create_delay(max_delay_time=0.2, delay_time=0.0123)
        finite_stereo(pkg, c, 60 + g, 128 * 12, start=(128 * 9 + 3) / G.SR).connect(d)
        mono_tone(pkg, c, 200.0, stop=(128 * 40) / G.SR).connect(d)
        bq = c.create_biquad_filter(type_=pkg.HIGHPASS, frequency=100.0, q=2.0)
        d.connect(bq)
        bq.connect(c.destination())
        return c
    check(pkg, engine, oracle, build)


def test_feedback_delay_with_a_finite_stereo_source(pkg, engine, oracle):
    def build(be, g):
        c = pkg.OfflineAudioContext(2, 128 * 80, G.SR, be)
        d = c.create_delay(max_delay_time=0.1, delay_time=0.013)
        fb = c.create_gain(0.5)
        finite_stereo(pkg, c, 70 + g, 128 * 16 + 21).connect(d)
        d.connect(fb)
        fb.connect(d)
        d.connect(c.destination())
        return c
    check(pkg, engine, oracle, build, n=2)


def test_panners_switch_between_their_mono_and_stereo_laws(pkg, engine, oracle):
    def build(be, g):
        c = pkg.OfflineAudioContext(2, N, G.SR, be)
        sp = c.create_stereo_panner(pan=-0.4 + 0.3 * g)
        pn = c.create_panner(position=(1.5, 0.2 * g, -1.0), distance_model=1)
        mv = c.create_panner(position=(-2.0, 0.0, 1.0))
        mv.position_x.linear_ramp_to_value_at_time(2.0, 0.12)
        src = finite_stereo(pkg, c, 80 + g, 128 * 20, start=(128 * 6) / G.SR)
        tone = mono_tone(pkg, c, 300.0, stop=(128 * 45) / G.SR)
        for p in (sp, pn, mv):
            src.connect(p)
            tone.connect(p)
            p.connect(c.destination())
        return c
    check(pkg, engine, oracle, build)


def test_compressor_analyser_and_gain_carry_the_layout(pkg, engine, oracle):
    def build(be, g):
        c = pkg.OfflineAudioContext(2, N, G.SR, be)
        comp = c.create_dynamics_compressor()
        an = c.create_analyser(fft_size=256)
        gn = c.create_gain(0.7)
        gn.gain.set_target_at_time(0.2, 0.05, 0.02)
        bq = c.create_biquad_filter(type_=pkg.LOWPASS, frequency=2000.0, q=3.0)
        finite_stereo(pkg, c, 90 + g, 128 * 25 + 3).connect(comp)
        comp.connect(an)
        an.connect(gn)
        gn.connect(bq)
        mono_tone(pkg, c, 700.0).connect(bq)
        bq.connect(c.destination())
        return c
    check(pkg, engine, oracle, build)


def test_zero_gain_silences_and_nonzero_curve_unsilences(pkg, engine, oracle):
    def build(be, g):
        c = pkg.OfflineAudioContext(2, N, G.SR, be)
        mute = c.create_gain(0.0)
        offset_curve = (np.linspace(-1, 1, 9) * 0.5 + 0.25).astype(np.float32)  # maps 0 to 0.25: a silent input still sounds (on ONE channel)
        sh = c.create_wave_shaper(curve=offset_curve)
        bq = c.create_biquad_filter(type_=pkg.LOWPASS, frequency=800.0, q=5.0)
        src = finite_stereo(pkg, c, 100 + g, 128 * 14 + 60)
        src.connect(mute)
        src.connect(sh)
        mute.connect(bq)
        sh.connect(bq)
        bq.connect(c.destination())
        return c
    check(pkg, engine, oracle, build)


@pytest.mark.parametrize("oversample", [1, 2])
def test_oversampled_shaper_freezes_while_its_input_is_silent(pkg, engine, oracle, oversample):
    """waveshaper.rs:395-398: silent input + curve through 0 -> early return, the resamplers are not fed; when the next source starts
    they continue from the state the first one left"""
    def build(be, g):
        c = pkg.OfflineAudioContext(2, N, G.SR, be)
        sh = c.create_wave_shaper(curve=np.tanh(np.linspace(-2, 2, 33)).astype(np.float32), oversample=oversample)
        rng = np.random.default_rng(110 + g)
        a = c.create_buffer_source(pkg.AudioBuffer([rng.uniform(-0.8, 0.8, 128 * 9 + 40).astype(np.float32)], G.SR))
        b2 = c.create_buffer_source(pkg.AudioBuffer([rng.uniform(-0.8, 0.8, 128 * 11).astype(np.float32)], G.SR))
        a.connect(sh)
        b2.connect(sh)
        a.start()
        b2.start_at((128 * 25 + 5) / G.SR)
        sh.connect(c.destination())
        return c
    check(pkg, engine, oracle, build)


def test_convolver_keeps_its_tail_then_falls_silent(pkg, engine, oracle):
    def build(be, g):
        c = pkg.OfflineAudioContext(2, 8192 * 2 + 128 * 5, G.SR, be)
        ir = G.synthetic_ir(1500, 2, decay=0.01, seed=5 + g)
        cv = c.create_convolver(pkg.AudioBuffer(ir, G.SR))
        bq = c.create_biquad_filter(type_=pkg.LOWPASS, frequency=3000.0, q=2.0)
        finite_stereo(pkg, c, 120 + g, 128 * 30 + 11).connect(cv)
        cv.connect(bq)
        mono_tone(pkg, c, 250.0).connect(bq)
        bq.connect(c.destination())
        return c
    check(pkg, engine, oracle, build, n=2)


def test_splitter_and_merger_with_sources_that_end(pkg, engine, oracle):
    def build(be, g):
        c = pkg.OfflineAudioContext(2, N, G.SR, be)
        sp = c.create_channel_splitter(2)
        mg = c.create_channel_merger(2)
        finite_stereo(pkg, c, 130 + g, 128 * 17 + 1).connect(sp)
        mono_tone(pkg, c, 410.0, start=(128 * 8) / G.SR, stop=(128 * 50) / G.SR).connect(sp)
        bq = c.create_biquad_filter(type_=pkg.LOWPASS, frequency=1200.0, q=4.0)
        sp.connect_from_output_to_input(bq, 1, 0)      # the right channel: silent whenever the splitter's input is mono
        bq.connect_from_output_to_input(mg, 0, 1)
        sp.connect_from_output_to_input(mg, 0, 0)
        mg.connect(c.destination())
        return c
    check(pkg, engine, oracle, build)


def test_explicit_and_clamped_ports_with_silent_inputs(pkg, engine, oracle):
    def build(be, g):
        c = pkg.OfflineAudioContext(2, N, G.SR, be)
        ex = c.create_biquad_filter(type_=pkg.LOWPASS, frequency=700.0, q=3.0, cfg=pkg.context.channel_config(2, pkg.EXPLICIT, pkg.SPEAKERS))
        cl = c.create_biquad_filter(type_=pkg.LOWPASS, frequency=900.0, q=3.0, cfg=pkg.context.channel_config(1, pkg.CLAMPED_MAX, pkg.SPEAKERS))
        dd = c.create_delay(max_delay_time=0.1, delay_time=0.01, cfg=pkg.context.channel_config(2, pkg.EXPLICIT, pkg.DISCRETE))
        src = finite_stereo(pkg, c, 140 + g, 128 * 13 + 77, start=(128 * 2) / G.SR)
        tone = mono_tone(pkg, c, 180.0, stop=(128 * 35) / G.SR)
        for node in (ex, cl, dd):
            src.connect(node)
            tone.connect(node)
            node.connect(c.destination())
        return c
    check(pkg, engine, oracle, build)


@pytest.mark.parametrize("order", [(1, 2, 6), (2, 1, 6), (6, 1, 2), (1, 2, 8)])
def test_mixing_three_layouts_folds_edge_by_edge(pkg, engine, oracle, order):
    """quantum.rs:532-569: every add first brings the running sum to max(sum, edge) channels — mono then stereo then 5.1 puts the mono
    into L / R (1 -> 2 -> 6), not into C (1 -> 6); more than 6 channels are discrete"""
    def build(be, g):
        top = max(order)
        c = pkg.OfflineAudioContext(top, 128 * 6, G.SR, be)
        mix = c.create_gain(1.0)  # Max / speakers
        rng = np.random.default_rng(150 + g)
        nodes = []
        for ch in order:  # created in this order; the mixer sums the LAST created first
            s = c.create_buffer_source(pkg.AudioBuffer(list(rng.uniform(-0.3, 0.3, (ch, 128 * 6)).astype(np.float32)), G.SR))
            s.start()
            nodes.append(s)
        for s in nodes:
            s.connect(mix)
        dest = c.destination()
        mix.connect(dest)
        return c
    n_ch = max(order)
    gpu = np.stack([np.stack(b.channels) for b in pkg.render_batch([build(engine.backend, 0)])])
    cpu = np.stack([np.stack(b.channels) for b in pkg.render_batch([build(oracle, 0)])])
    assert gpu.shape[1] == n_ch
    assert maxdiff(gpu, cpu) <= TOL
