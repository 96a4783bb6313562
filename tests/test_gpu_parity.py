"""Parity tests proper: the CUDA engine (through the C ABI) vs the CPU oracle on the same seeded graphs.
Tolerance = the north_star contract: 1e-5 absolute on f32 PCM (TOL below); several cases are bit-exact."""
import numpy as np
import pytest

import graphs as G

pytestmark = pytest.mark.gpu
TOL = 1e-5  # BASELINE.json north_star: "match the reference ... within 1e-5 f32"


def both(pkg, engine, oracle, build, n=1):
    gpu = G.render(pkg, [build(engine.backend, g) for g in range(n)])
    cpu = G.render(pkg, [build(oracle, g) for g in range(n)])
    return gpu, cpu


def maxdiff(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())


def test_c1_osc_biquad(pkg, engine, oracle):
    # BASELINE configs[0]: 1 s of osc -> biquad -> destination at 48 kHz stereo
    gpu, cpu = both(pkg, engine, oracle, lambda be, g: G.c1_osc_biquad(pkg, be, 48000))
    assert gpu.shape == (1, 2, 48000)
    assert maxdiff(gpu, cpu) <= TOL
    assert np.array_equal(gpu[0, 0], gpu[0, 1])  # mono fan-in is up-mixed by copy


def test_c1_partial_last_quantum(pkg, engine, oracle):
    gpu, cpu = both(pkg, engine, oracle, lambda be, g: G.c1_osc_biquad(pkg, be, 555, 44100.0))
    assert gpu.shape == (1, 2, 555)
    assert maxdiff(gpu, cpu) <= TOL


@pytest.mark.parametrize("serial", [0, 1])
def test_c2_buffer_biquad_gain(pkg, engine, oracle, serial):
    engine.set_option(pkg.OPT_SERIAL_FILTERS, serial)
    try:
        n, length = 24, 128 * 97 + 5
        gpu, cpu = both(pkg, engine, oracle, lambda be, g: G.c2_buffer_biquad_gain(pkg, be, g, length), n)
        assert maxdiff(gpu, cpu) <= TOL
        if serial:  # the serial kernel keeps the reference's f64 operation order
            assert maxdiff(gpu, cpu) <= 1e-7
    finally:
        engine.set_option(pkg.OPT_SERIAL_FILTERS, 0)


@pytest.mark.parametrize("chunk", [128, 1024, 2048, 4096 + 128, 0])
def test_chunk_invariance(pkg, engine, oracle, chunk):
    # size-independent property: the result must not depend on how the render is cut into time chunks
    engine.set_option(pkg.OPT_CHUNK_FRAMES, chunk)
    try:
        length = 128 * 131
        gpu, cpu = both(pkg, engine, oracle, lambda be, g: G.c2_buffer_biquad_gain(pkg, be, g, length), 5)
        assert maxdiff(gpu, cpu) <= TOL
    finally:
        engine.set_option(pkg.OPT_CHUNK_FRAMES, 0)


def test_one_shot_render_batch_abi(pkg, engine, oracle):
    """wae_render_batch(HOST): prepare + render + D2H in ONE call, the entry point the Rust binding uses."""
    import ctypes as C
    n_graphs, length = 5, 128 * 37 + 11
    ctxs = [G.c2_buffer_biquad_gain(pkg, engine.backend, g, length) for g in range(n_graphs)]
    out = np.zeros((n_graphs, 2, length), np.float32)
    arr = (C.c_void_p * n_graphs)(*[c._g for c in ctxs])
    api = engine.backend.api
    api.check(api.render_batch(engine.backend.engine, arr, n_graphs, out.ctypes.data_as(C.c_void_p), 0))
    cpu = G.render(pkg, [G.c2_buffer_biquad_gain(pkg, oracle, g, length) for g in range(n_graphs)])
    assert maxdiff(out, cpu) <= TOL
    # the pipelined path (pinned mirror built on first use) gives the same PCM
    b = pkg.Batch([G.c2_buffer_biquad_gain(pkg, engine.backend, g, length) for g in range(n_graphs)])
    out2 = np.zeros_like(out)
    b.run_pipelined(out2.ctypes.data_as(C.c_void_p))
    b.run_pipelined(out2.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, out2)


def test_c3_many_voices_summation_order(pkg, engine, oracle):
    # C3 scaled down: 300 voices summed at the destination in the reference's order (last-created first)
    gpu, cpu = both(pkg, engine, oracle, lambda be, g: G.c3_many_voices(pkg, be, 300, 128 * 40))
    scale = float(np.abs(cpu).max())
    assert scale > 1.0
    assert maxdiff(gpu, cpu) <= TOL * max(1.0, scale / 8)  # abs 1e-5 at amplitudes up to 8, relative above


@pytest.mark.parametrize("typ", ["sine", "square", "sawtooth", "triangle"])
def test_oscillator_types_and_schedule(pkg, engine, oracle, typ):
    t = {"sine": 0, "square": 1, "sawtooth": 2, "triangle": 3}[typ]

    def build(be, g):
        c = pkg.OfflineAudioContext(1, 128 * 20, G.SR, be)
        osc = c.create_oscillator(type_=t, frequency=443.7, detune=35.0)
        osc.connect(c.destination())
        osc.start_at(0.0113)   # sub-sample start inside quantum 4
        osc.stop_at(0.0402)
        return c

    gpu, cpu = both(pkg, engine, oracle, build)
    assert np.array_equal(gpu == 0, cpu == 0)  # same start/stop frames
    assert maxdiff(gpu, cpu) <= TOL


def test_oscillator_outside_nyquist_and_negative(pkg, engine, oracle):
    def build(be, g):
        c = pkg.OfflineAudioContext(1, 128 * 8, G.SR, be)
        for f in (23999.0, -440.0, 24000.0):
            o = c.create_oscillator(frequency=f)
            o.connect(c.destination())
            o.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build)
    assert maxdiff(gpu, cpu) <= TOL


@pytest.mark.parametrize("btype", range(8))
def test_biquad_all_types(pkg, engine, oracle, btype):
    def build(be, g):
        pcm = G.c2_source(g, 128 * 30)
        c = pkg.OfflineAudioContext(2, 128 * 30, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        b = c.create_biquad_filter(type_=btype, frequency=1200.0, q=2.5, gain=6.0, detune=120.0)
        s.connect(b)
        b.connect(c.destination())
        s.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 2)
    assert maxdiff(gpu, cpu) <= TOL


def test_low_frequency_high_q_biquad_long(pkg, engine, oracle):
    # poles close to the unit circle + many scan tiles: the stress case of the time-parallel recurrence
    def build(be, g):
        pcm = G.c2_source(g, 128 * 400)
        c = pkg.OfflineAudioContext(1, 128 * 400, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0]], G.SR))
        b = c.create_biquad_filter(type_=pkg.BANDPASS, frequency=25.0, q=30.0)
        s.connect(b)
        b.connect(c.destination())
        s.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build)
    assert maxdiff(gpu, cpu) <= TOL


def test_iir_filter(pkg, engine, oracle):
    def build(be, g):
        pcm = G.c2_source(g, 128 * 25)
        c = pkg.OfflineAudioContext(2, 128 * 25, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        f = c.create_iir_filter([0.0675, 0.1349, 0.0675, 0.01], [1.0, -1.1430, 0.4128])
        s.connect(f)
        f.connect(c.destination())
        s.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 3)
    assert maxdiff(gpu, cpu) <= 1e-7  # same serial f64 order


def test_second_order_iir_rides_the_biquad_scan(pkg, engine, oracle):
    """An IIRFilterNode of order <= 2 fed by a constant layout is lowered to the fused chain's time-parallel biquad (normalised by a0);
    the reference runs it through its own f64 IIR loop (src/node/iir_filter.rs:206-241), so the two agree to the scan's tolerance."""
    def build(be, g):
        pcm = G.c2_source(g, 128 * 40)
        c = pkg.OfflineAudioContext(2, 128 * 40, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        f = c.create_iir_filter([0.135, 0.2698, 0.135], [2.0, -2.2860, 0.8256]) if g % 2 == 0 else c.create_iir_filter([0.5, 0.5], [1.0, -0.2])
        s.connect(f)
        f.connect(c.destination())
        s.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 4)
    assert maxdiff(gpu, cpu) <= TOL


def test_second_order_iir_keeps_its_memory_across_a_suspend_point(pkg, engine, oracle):
    """A render cut by suspend_sync is planned segment by segment; a source that covers the first segment entirely but ends in the second
    gives the filter a constant layout in one plan and a changing one in the other.  The filter memory has to survive the cut, so such
    graphs keep the serial IIR kernel in every segment (regression: fuzz seeds 31 / 33 lost the state at the suspend frame)."""
    def build(be, g):
        n = 128 * 40
        pcm = G.c2_source(g, 128 * 30)                       # ends at quantum 30 of 40
        c = pkg.OfflineAudioContext(2, n, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        f = c.create_iir_filter([0.2, 0.3, 0.1], [1.0, -0.4, 0.2])
        s.connect(f)
        f.connect(c.destination())
        s.start()
        if g != 1:
            c.suspend_sync(128 * 18 / G.SR, lambda ctx: None)  # graph 1: no cut, same batch
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 3)
    assert maxdiff(gpu, cpu) <= TOL


def test_gain_shaper_panner_chain(pkg, engine, oracle):
    curve = np.tanh(np.linspace(-3, 3, 1024)).astype(np.float32)

    def build(be, g):
        c = pkg.OfflineAudioContext(2, 128 * 16, G.SR, be)
        o = c.create_oscillator(type_=pkg.SAWTOOTH, frequency=330.0 + 10 * g)
        ws = c.create_wave_shaper(curve)
        gn = c.create_gain(0.7)
        sp = c.create_stereo_panner(pan=-0.4 + 0.3 * g)
        o.connect(ws)
        ws.connect(gn)
        gn.connect(sp)
        sp.connect(c.destination())
        o.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 3)
    assert maxdiff(gpu, cpu) <= TOL


def test_stereo_panner_stereo_input(pkg, engine, oracle):
    def build(be, g):
        pcm = G.c2_source(g, 128 * 8)
        c = pkg.OfflineAudioContext(2, 128 * 8, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        sp = c.create_stereo_panner(pan=[-0.6, 0.0, 0.8][g])
        s.connect(sp)
        sp.connect(c.destination())
        s.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 3)
    assert maxdiff(gpu, cpu) <= TOL


def test_equal_power_panner(pkg, engine, oracle):
    def build(be, g):
        c = pkg.OfflineAudioContext(2, 128 * 8, G.SR, be)
        o = c.create_oscillator(frequency=500.0)
        p = c.create_panner(position=[(3.0, 1.0, -2.0), (-4.0, 0.0, 0.5), (0.0, 0.0, 0.0)][g], distance_model=g,
                            cone_inner_angle=60.0, cone_outer_angle=120.0, cone_outer_gain=0.3, max_distance=50.0)
        o.connect(p)
        p.connect(c.destination())
        o.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 3)
    assert maxdiff(gpu, cpu) <= TOL


@pytest.mark.parametrize("taps,chunk", [(256, 0), (512, 0), (250, 0), (256, 128), (256, 1024)])
def test_hrtf_panner(pkg, engine, oracle, taps, chunk):
    """PanningModelType::HRTF (panner.rs:215-271,781-830): static source / listener, mono and stereo inputs, all
    distance models, cone gain; synthetic HRIR sphere at the context rate (the real one cannot travel)."""
    data = G.synthetic_hrir_sphere(int(G.SR), taps)
    oracle.set_hrir_sphere(data)
    engine.backend.set_hrir_sphere(data)
    positions = [(3.0, 1.0, -2.0), (-4.0, 0.0, 0.5), (0.0, 0.0, 0.0), (0.2, -6.0, 0.1), (1.0, 0.0, 0.0), (0.0, 0.0, -1.0)]

    def build(be, g):
        n = 128 * 21 + 40
        pcm = G.c2_source(g, n) * np.float32(0.5)
        c = pkg.OfflineAudioContext(2, n, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]] if g % 2 else [pcm[0]], G.SR))
        p = c.create_panner(panning_model=pkg.context.HRTF, position=positions[g], distance_model=g % 3,
                            cone_inner_angle=60.0, cone_outer_angle=120.0, cone_outer_gain=0.3, max_distance=50.0,
                            orientation=(0.0, 1.0, 0.5))
        s.connect(p)
        p.connect(c.destination())
        s.start()
        return c

    engine.set_option(pkg.OPT_CHUNK_FRAMES, chunk)
    try:
        gpu, cpu = both(pkg, engine, oracle, build, len(positions))
    finally:
        engine.set_option(pkg.OPT_CHUNK_FRAMES, 0)
    assert float(np.abs(cpu).max()) > 0.05
    assert maxdiff(gpu, cpu) <= TOL


def test_static_hrtf_panner_is_lowered_to_the_convolver_kernels(pkg, engine, oracle):
    """A static source heard by a static listener through a constant-layout input is one fixed pair of impulse responses: the planner hands
    it to the time-batched convolver kernels (the hrtf crate itself convolves by FFT overlap-save); a moving source, or an input whose layout
    changes (a source that ends), keeps the per-quantum FIR kernel.  Same PCM as the oracle either way."""
    data = G.synthetic_hrir_sphere(int(G.SR), 256)
    oracle.set_hrir_sphere(data)
    engine.backend.set_hrir_sphere(data)
    n = 128 * 40

    def build(be, kind):
        c = pkg.OfflineAudioContext(2, n, G.SR, be)
        pcm = G.c2_source(kind, n) * np.float32(0.5)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        p = c.create_panner(panning_model=pkg.context.HRTF, position=(2.0, 0.5, -1.5))
        s.connect(p)
        p.connect(c.destination())
        s.start()
        if kind == 1:
            p.position_x.linear_ramp_to_value_at_time(-3.0, n / G.SR)  # moving source
        if kind == 2:
            s.stop_at(0.03)  # the input goes silent: per-quantum layout
        return c

    import os
    want = {0: os.environ.get("WAE_HRTF_FFT", "1") != "0", 1: False, 2: False}
    for kind in (0, 1, 2):
        batch = pkg.Batch([build(engine.backend, kind)])
        names = {name for name, _t, _k in batch.stage_times()}
        batch.destroy()
        assert ("k_conv_mac_ifft" in names) == want[kind], (kind, names)
        assert ("k_hrtf_fir" in names) == (not want[kind]), (kind, names)
        gpu = G.render(pkg, [build(engine.backend, kind)])
        cpu = G.render(pkg, [build(oracle, kind)])
        assert maxdiff(gpu, cpu) <= TOL, kind


@pytest.mark.parametrize("model", ["equalpower", "hrtf"])
@pytest.mark.parametrize("who", ["source", "listener", "both", "audio_rate"])
def test_panner_moving_source_and_listener(pkg, engine, oracle, model, who):
    """Automated panner / listener params (panner.rs:714-780): a-rate spatial params for equal-power panning — including
    the reference's rule that only the LISTENER params decide between the single-valued and the per-frame path
    (panner.rs:833-841) — and first-value-per-quantum (k-rate) for HRTF (panner.rs:781-788)."""
    if model == "hrtf":
        data = G.synthetic_hrir_sphere(int(G.SR), 128, subdivisions=1)
        oracle.set_hrir_sphere(data)
        engine.backend.set_hrir_sphere(data)

    def build(be, g):
        n = 128 * 24
        pcm = G.c2_source(g, n) * np.float32(0.5)
        c = pkg.OfflineAudioContext(2, n, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]] if g % 2 else [pcm[0]], G.SR))
        p = c.create_panner(panning_model=pkg.context.HRTF if model == "hrtf" else pkg.context.EQUALPOWER,
                            position=(-3.0, 0.5, 1.0), distance_model=g % 3, max_distance=40.0,
                            cone_inner_angle=50.0, cone_outer_angle=140.0, cone_outer_gain=0.2, orientation=(0.0, 0.2, -1.0))
        if who in ("source", "both"):
            p.position_x.linear_ramp_to_value_at_time(4.0, 0.05)
            p.position_z.set_target_at_time(-6.0, 0.01, 0.02)
            p.orientation_x.set_value_at_time(1.0, 0.03)
        if who in ("listener", "both"):
            l = c.listener()
            l.position_x.linear_ramp_to_value_at_time(2.0, 0.04)
            l.forward_x.set_value_at_time(0.6, 0.02)
            l.up_z.linear_ramp_to_value_at_time(0.3, 0.06)
        if who == "audio_rate":
            lfo = c.create_oscillator(frequency=13.0)
            amp = c.create_gain(gain=2.5)
            lfo.connect(amp)
            amp.connect(p.position_x)
            amp.connect(c.listener().position_y)
            lfo.start()
        s.connect(p)
        p.connect(c.destination())
        s.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 3)
    assert float(np.abs(cpu).max()) > 0.02
    assert maxdiff(gpu, cpu) <= (3e-5 if model == "equalpower" else TOL)  # acosf / atan chains: few-ulp libm differences on angles


def test_c5_full_chain(pkg, engine, oracle):
    """configs[4] scaled down: Oscillator -> WaveShaper -> Biquad -> Convolver -> Panner(HRTF) -> Analyser -> destination."""
    data = G.synthetic_hrir_sphere(int(G.SR), 512)
    oracle.set_hrir_sphere(data)
    engine.backend.set_hrir_sphere(data)
    ir = G.synthetic_ir(5000, 2, decay=0.05)
    n = 128 * 100 + 77
    cg = [G.c5_full_chain(pkg, engine.backend, g, n, ir) for g in range(4)]
    cc = [G.c5_full_chain(pkg, oracle, g, n, ir) for g in range(4)]
    gpu, cpu = G.render(pkg, cg), G.render(pkg, cc)
    assert float(np.abs(cpu).max()) > 1e-3
    assert maxdiff(gpu, cpu) <= TOL
    for a, b in zip(cg, cc):
        fa, fb = a._test_analyser.get_float_frequency_data(), b._test_analyser.get_float_frequency_data()
        assert np.abs(10.0 ** (fa / 20) - 10.0 ** (fb / 20)).max() <= 1e-6


@pytest.mark.parametrize("sphere_rate,ctx_rate", [(44100, 48000.0), (44100, 96000.0), (48000, 32000.0)])
def test_hrtf_panner_resamples_the_sphere_to_the_context_rate(pkg, engine, oracle, sphere_rate, ctx_rate):
    """HrirSphere::new(reader, context_rate): the reference's embedded sphere is 44.1 kHz data, its contexts usually 48 kHz — every
    response is resampled once (asynchronous sinc resampler; csrc/wae_hrtf_host.h vs oracle/wao_hrtf.cpp) before the panner runs."""
    data = G.synthetic_hrir_sphere(sphere_rate, 384)
    oracle.set_hrir_sphere(data)
    engine.backend.set_hrir_sphere(data)
    positions = [(3.0, 1.0, -2.0), (-4.0, 0.0, 0.5), (0.2, -6.0, 0.1)]

    def build(be, g):
        n = 128 * 15 + 7
        pcm = G.c2_source(g, n) * np.float32(0.5)
        c = pkg.OfflineAudioContext(2, n, ctx_rate, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]] if g % 2 else [pcm[0]], ctx_rate))
        p = c.create_panner(panning_model=pkg.context.HRTF, position=positions[g])
        s.connect(p)
        p.connect(c.destination())
        s.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, len(positions))
    assert float(np.abs(cpu).max()) > 0.01
    assert maxdiff(gpu, cpu) <= TOL


def test_hrtf_panner_rejects_a_malformed_sphere(pkg, engine):
    with pytest.raises(pkg.WaeError):
        engine.backend.set_hrir_sphere(b"HRIX" + bytes(64))


def test_delay_node(pkg, engine, oracle):
    def build(be, g):
        pcm = G.c2_source(g, 128 * 40)
        c = pkg.OfflineAudioContext(2, 128 * 40, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        d = c.create_delay(max_delay_time=0.5, delay_time=[0.0, 0.00101, 0.0213][g])
        s.connect(d)
        d.connect(c.destination())
        s.connect(c.destination())
        s.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 3)
    assert maxdiff(gpu, cpu) <= TOL


def test_dynamics_compressor(pkg, engine, oracle):
    def build(be, g):
        pcm = G.c2_source(g, 128 * 30) * np.float32(0.9)
        c = pkg.OfflineAudioContext(2, 128 * 30, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        d = c.create_dynamics_compressor()
        s.connect(d)
        d.connect(c.destination())
        s.start()
        c._test_comp = d
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 2)
    assert maxdiff(gpu, cpu) <= 5e-5  # f32 log10/pow chains differ by a few ulp between glibc and CUDA libm
    cg, cc = build(engine.backend, 1), build(oracle, 1)
    G.render(pkg, [cg]), G.render(pkg, [cc])
    assert abs(cg._test_comp.reduction() - cc._test_comp.reduction()) <= 1e-3  # dB


def test_dynamics_compressor_automated_params(pkg, engine, oracle):
    """k-rate automation of threshold / knee / ratio / attack / release (dynamics_compressor.rs:352-391)."""
    def build(be, g):
        n = 128 * 40
        pcm = G.c2_source(g, n) * np.float32(0.9)
        c = pkg.OfflineAudioContext(2, n, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        d = c.create_dynamics_compressor()
        d.threshold.linear_ramp_to_value_at_time(-50.0, 0.08)
        d.ratio.set_value_at_time(4.0, 0.02)
        d.knee.set_target_at_time(5.0, 0.01, 0.02)
        d.attack.set_value_at_time(0.05, 0.03)
        d.release.exponential_ramp_to_value_at_time(0.05, 0.1)
        s.connect(d)
        d.connect(c.destination())
        s.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 2)
    assert maxdiff(gpu, cpu) <= 5e-5


def test_analyser_passthrough_and_time_domain(pkg, engine, oracle):
    def build(be, g):
        c = pkg.OfflineAudioContext(1, 128 * 30 + 17, G.SR, be)
        o = c.create_oscillator(frequency=700.0)
        a = c.create_analyser(fft_size=1024)
        o.connect(a)
        a.connect(c.destination())
        o.start()
        c._test_analyser = a
        return c

    cg, cc = build(engine.backend, 0), build(oracle, 0)
    gpu = G.render(pkg, [cg])
    cpu = G.render(pkg, [cc])
    assert maxdiff(gpu, cpu) <= TOL
    tg = cg._test_analyser.get_float_time_domain_data()
    tc = cc._test_analyser.get_float_time_domain_data()
    assert maxdiff(tg, tc) <= TOL


@pytest.mark.parametrize("fft_size,smoothing", [(32, 0.8), (1024, 0.0), (2048, 0.8), (32768, 0.3)])
def test_analyser_frequency_data(pkg, engine, oracle, fft_size, smoothing):
    """Analyser::compute_fft + get_float_frequency_data (src/analysis.rs:278-369) after the render."""
    def build(be):
        c = pkg.OfflineAudioContext(1, 128 * 300 + 5, G.SR, be)
        o1 = c.create_oscillator(frequency=700.0)
        o2 = c.create_oscillator(type_=pkg.SAWTOOTH, frequency=3111.0)
        g = c.create_gain(gain=0.25)
        a = c.create_analyser(fft_size=fft_size, smoothing_time_constant=smoothing)
        o1.connect(a)
        o2.connect(g)
        g.connect(a)
        a.connect(c.destination())
        o1.start()
        o2.start()
        c._test_analyser = a
        return c

    cg, cc = build(engine.backend), build(oracle)
    assert maxdiff(G.render(pkg, [cg]), G.render(pkg, [cc])) <= TOL
    fg = cg._test_analyser.get_float_frequency_data()
    fc = cc._test_analyser.get_float_frequency_data()
    assert fg.shape == fc.shape == (fft_size // 2,)
    lin_g, lin_c = 10.0 ** (fg.astype(np.float64) / 20), 10.0 ** (fc.astype(np.float64) / 20)
    assert np.abs(lin_g - lin_c).max() <= 1e-6          # magnitudes (the quantity the FFT produces)
    loud = lin_c > 1e-4
    assert loud.any() and np.abs(fg[loud] - fc[loud]).max() <= 1e-2   # dB where the bin is above the noise floor
    # a second read at the same current_time returns the cached spectrum (analysis.rs:353-361)
    assert np.array_equal(cg._test_analyser.get_float_frequency_data(), fg)
    # byte read-outs (analysis.rs:266-276, 371-401): floor() of a scaled value, so allow one count where the float sits on a step
    bg, bc = cg._test_analyser.get_byte_frequency_data(), cc._test_analyser.get_byte_frequency_data()
    assert np.abs(bg.astype(int) - bc.astype(int)).max() <= 1 and int(bc.max()) > 0
    tg, tc = cg._test_analyser.get_byte_time_domain_data(), cc._test_analyser.get_byte_time_domain_data()
    assert np.abs(tg.astype(int) - tc.astype(int)).max() <= 1 and tg.std() > 1


@pytest.mark.parametrize("n,src,dst", [(1, 44100, 48000), (5, 48000, 44100), (1000, 44100, 48000), (48000, 96000, 48000),
                                       (12345, 8000, 48000), (777, 48000, 48000.05)])
def test_resample_linear(pkg, engine, oracle, n, src, dst):
    """AudioBuffer::resample (src/buffer.rs:311-363) on the GPU vs the oracle."""
    import ctypes as C
    x = np.random.default_rng(n).standard_normal(n).astype(np.float32)
    got = engine.resample(x, src, dst)
    want = np.zeros(len(got) + 8, np.float32)
    fp = C.POINTER(C.c_float)
    m = oracle.api.resample_linear(x.ctypes.data_as(fp), n, float(src), float(dst), want.ctypes.data_as(fp), len(want))
    assert m == len(got)
    assert maxdiff(got, want[:m]) <= 1e-6


def test_channel_splitter_merger(pkg, engine, oracle):
    def build(be, g):
        pcm = G.c2_source(g, 128 * 6)
        c = pkg.OfflineAudioContext(2, 128 * 6, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        sp = c.create_channel_splitter(2)
        mg = c.create_channel_merger(2)
        s.connect(sp)
        sp.connect_from_output_to_input(mg, 0, 1)  # swap left / right
        sp.connect_from_output_to_input(mg, 1, 0)
        mg.connect(c.destination())
        s.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 1)
    assert np.array_equal(gpu, cpu)


@pytest.mark.parametrize("in_ch,ir_ch", [(1, 1), (1, 2), (2, 1), (2, 2), (2, 4), (1, 4)])
def test_convolver_channel_routing(pkg, engine, oracle, in_ch, ir_ch):
    # the six routings of src/node/convolver.rs:378-487
    ir = G.synthetic_ir(3000, ir_ch, seed=5)

    def build(be, g):
        length = 128 * 60
        pcm = G.c2_source(g, length)
        c = pkg.OfflineAudioContext(2, length, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[i] for i in range(in_ch)], G.SR))
        cv = c.create_convolver(pkg.AudioBuffer(ir, G.SR))
        s.connect(cv)
        cv.connect(c.destination())
        s.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 2)
    assert maxdiff(gpu, cpu) <= TOL


@pytest.mark.parametrize("chunk", [1024, 4096, 0])
def test_c4_convolver_long_ir(pkg, engine, oracle, chunk):
    engine.set_option(pkg.OPT_CHUNK_FRAMES, chunk)
    try:
        ir = G.synthetic_ir(20000, 2)  # 20 partitions of 1024
        gpu, cpu = both(pkg, engine, oracle, lambda be, g: G.c4_convolver(pkg, be, g, 128 * 250 + 77, ir), 3)
        assert maxdiff(gpu, cpu) <= TOL
    finally:
        engine.set_option(pkg.OPT_CHUNK_FRAMES, 0)


@pytest.mark.parametrize("chunk", [8192, 0])
def test_convolver_reads_the_source_buffer_in_place_and_writes_the_destination(pkg, engine, oracle, chunk):
    """source -> Convolver -> destination with a buffer that covers the whole (quantum-padded) render: no copy into the arena on the way
    in, no mix on the way out — the inverse transforms stop at the render length (odd, so the last float4 is cut)."""
    engine.set_option(pkg.OPT_CHUNK_FRAMES, chunk)
    try:
        n = 8192 * 2 + 3001
        ir = G.synthetic_ir(9000, 2)

        def build(be, g):
            rng = np.random.default_rng(900 + g)
            pcm = rng.uniform(-0.3, 0.3, (2, n + 500)).astype(np.float32)   # longer than the padded render
            c = pkg.OfflineAudioContext(2, n, G.SR, be)
            src = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
            cv = c.create_convolver(pkg.AudioBuffer(ir if g % 2 == 0 else [ir[0]], G.SR))  # stereo / mono response (two paths each)
            src.connect(cv)
            cv.connect(c.destination())
            src.start()
            return c

        gpu, cpu = both(pkg, engine, oracle, build, 4)
        assert gpu.shape == (4, 2, n) and maxdiff(gpu, cpu) <= TOL
    finally:
        engine.set_option(pkg.OPT_CHUNK_FRAMES, 0)


def test_convolver_linearity(pkg, engine):
    # size-independent property at a larger size: conv(a + b) == conv(a) + conv(b) within f32 rounding
    ir = G.synthetic_ir(48000, 2)
    length = 128 * 600
    rng = np.random.default_rng(3)
    a = rng.uniform(-0.5, 0.5, (2, length)).astype(np.float32)
    bb = rng.uniform(-0.5, 0.5, (2, length)).astype(np.float32)

    def build(pcm):
        c = pkg.OfflineAudioContext(2, length, G.SR, engine.backend)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        cv = c.create_convolver(pkg.AudioBuffer(ir, G.SR))
        s.connect(cv)
        cv.connect(c.destination())
        s.start()
        return c

    out = G.render(pkg, [build(a), build(bb), build(a + bb)])
    assert maxdiff(out[0] + out[1], out[2]) <= TOL


def test_north_star_graph(pkg, engine, oracle):
    ir = G.synthetic_ir(9000, 2)
    gpu, cpu = both(pkg, engine, oracle, lambda be, g: G.north_star_voices_convolver(pkg, be, 40, 128 * 100, ir, seed=g), 2)
    assert maxdiff(gpu, cpu) <= TOL


def test_unsupported_is_reported_not_faked(pkg, engine):
    # a ConvolverNode inside a DelayNode feedback loop is not lowered: reported, never approximated
    c = pkg.OfflineAudioContext(2, 256, G.SR, engine.backend)
    src = c.create_constant_source()
    g = c.create_gain(0.5)
    d = c.create_delay(1.0, 0.01)
    conv = c.create_convolver(pkg.AudioBuffer([np.ones(8, np.float32)], G.SR))
    src.connect(g)
    g.connect(d)
    d.connect(conv)
    conv.connect(g)
    g.connect(c.destination())
    src.start()
    with pytest.raises(pkg.WaeError) as e:
        c.start_rendering_sync()
    assert e.value.status == 4  # WAE_UNSUPPORTED -> the caller falls back to the CPU renderer


@pytest.mark.parametrize("oversample", [1, 2])
@pytest.mark.parametrize("chunk", [0, 128, 1024])
def test_waveshaper_oversampled(pkg, engine, oracle, oversample, chunk):
    """OverSampleType::X2 / X4 (waveshaper.rs:409-480): FFT up-sampler -> curve -> FFT down-sampler, mono and stereo."""
    def build(be, g):
        n = 128 * 24
        pcm = G.c2_source(g, n) * np.float32(1.4)
        c = pkg.OfflineAudioContext(2, n, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]] if g else [pcm[0]], G.SR))
        x = np.linspace(-1.0, 1.0, 33)
        sh = c.create_wave_shaper(curve=np.tanh(3.0 * x).astype(np.float32), oversample=oversample)
        s.connect(sh)
        sh.connect(c.destination())
        s.start()
        return c

    engine.set_option(pkg.OPT_CHUNK_FRAMES, chunk)
    try:
        gpu, cpu = both(pkg, engine, oracle, build, 2)
    finally:
        engine.set_option(pkg.OPT_CHUNK_FRAMES, 0)
    assert float(np.abs(cpu).max()) > 0.3
    assert maxdiff(gpu, cpu) <= TOL


def test_offline_rs_cases_on_gpu(pkg, engine):
    # tests/offline.rs:10-46 and :48-81 on the CUDA path (exact)
    c = pkg.OfflineAudioContext(2, 555, 44100.0, engine.backend)
    c1 = c.create_constant_source()
    c1.offset.set_value(2.0)
    c1.connect(c.destination())
    c2 = c.create_constant_source()
    c2.offset.set_value(-4.0)
    c2.connect(c.destination())
    c1.start()
    c2.start()
    out = c.start_rendering_sync()
    assert np.array_equal(out.get_channel_data(0), np.full(555, -2.0, np.float32))
    assert np.array_equal(out.get_channel_data(1), np.full(555, -2.0, np.float32))
    sr = 48000.0
    c = pkg.OfflineAudioContext(1, 512, sr, engine.backend)
    osc = c.create_oscillator(type_=pkg.SQUARE, frequency=0.0)
    osc.connect(c.destination())
    osc.start_at(128.0 / sr)
    osc.stop_at(128.0 * 3.0 / sr)
    out = c.start_rendering_sync().get_channel_data(0)
    assert np.array_equal(out, np.concatenate([np.zeros(128), np.ones(256), np.zeros(128)]).astype(np.float32))


def test_cycle_is_muted_and_cycle_breaker_feeds_back(pkg, engine, oracle):
    # tests/offline.rs:170-203 test_cycle and :205-244 test_cycle_breaker on the CUDA path
    c = pkg.OfflineAudioContext(1, 128, 48000.0, engine.backend)
    cycle1 = c.create_gain()
    cycle1.connect(c.destination())
    cycle2 = c.create_gain()
    cycle2.connect(cycle1)
    cycle1.connect(cycle2)
    sc = c.create_constant_source()
    sc.offset.set_value(1.0)
    sc.connect(cycle1)
    other = c.create_constant_source()
    other.offset.set_value(2.0)
    other.connect(c.destination())
    sc.start()
    other.start()
    out = c.start_rendering_sync().get_channel_data(0)
    assert np.array_equal(out, np.full(128, 2.0, np.float32))

    sr = 48000.0
    c = pkg.OfflineAudioContext(1, 128 * 3, sr, engine.backend)
    delay = c.create_delay(1.0 / sr)
    delay.delay_time.set_value(1.0 / sr)
    delay.connect(c.destination())
    delay.connect(delay)
    source = c.create_constant_source()
    source.offset.set_value(1.0)
    source.connect(delay)
    source.connect(c.destination())
    source.start()
    out = c.start_rendering_sync().get_channel_data(0)
    assert np.array_equal(out[:128], np.full(128, 1.0, np.float32))
    assert np.array_equal(out[128:256], np.full(128, 2.0, np.float32))
    assert np.array_equal(out[256:], np.full(128, 3.0, np.float32))


def test_feedback_delay_echo(pkg, engine, oracle):
    # a classic feedback echo: source -> delay -> gain(0.6) -> back into the delay; stereo source, 25 quanta
    def build(be, g):
        pcm = G.c2_source(g, 128 * 25)
        pcm[:, 128 * 3:] = 0
        c = pkg.OfflineAudioContext(2, 128 * 25, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        d = c.create_delay(max_delay_time=0.05, delay_time=0.0071)
        fb = c.create_gain(0.6)
        s.connect(d)
        d.connect(fb)
        fb.connect(d)
        d.connect(c.destination())
        s.connect(c.destination())
        s.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 2)
    assert maxdiff(gpu, cpu) <= TOL


@pytest.mark.parametrize("chunk", [0, 1024])
def test_feedback_echo_into_reverb_and_mixed_batch(pkg, engine, oracle, chunk):
    """A DelayNode feedback loop followed by a ConvolverNode, a biquad and an HRTF panner, in one batch with feedback-free
    graphs: the cyclic levels are replayed quantum by quantum inside a chunk, everything downstream runs on whole chunks."""
    data = G.synthetic_hrir_sphere(int(G.SR), 128, subdivisions=1)
    oracle.set_hrir_sphere(data)
    engine.backend.set_hrir_sphere(data)
    ir = G.synthetic_ir(3000, 2, decay=0.03)
    n = 128 * 70 + 33

    def build(be, g):
        if g == 2:
            return G.c4_convolver(pkg, be, g, n, ir)
        if g == 3:
            return G.c2_buffer_biquad_gain(pkg, be, g, n)
        pcm = G.c2_source(g, n)
        pcm[:, 128 * 5:] = 0
        c = pkg.OfflineAudioContext(2, n, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        lp = c.create_biquad_filter(type_=pkg.LOWPASS, frequency=3000.0)
        d = c.create_delay(max_delay_time=0.05, delay_time=[0.0071, 0.0029][g])
        fb = c.create_gain(0.55)
        s.connect(lp)
        lp.connect(d)
        d.connect(fb)
        fb.connect(d)
        cv = c.create_convolver(pkg.AudioBuffer(ir, G.SR))
        d.connect(cv)
        hp = c.create_biquad_filter(type_=pkg.HIGHPASS, frequency=150.0)
        cv.connect(hp)
        pn = c.create_panner(panning_model=pkg.context.HRTF, position=(2.0, 0.5, -1.0))
        hp.connect(pn)
        pn.connect(c.destination())
        s.connect(c.destination())
        s.start()
        return c

    engine.set_option(pkg.OPT_CHUNK_FRAMES, chunk)
    try:
        gpu, cpu = both(pkg, engine, oracle, build, 4)
    finally:
        engine.set_option(pkg.OPT_CHUNK_FRAMES, 0)
    assert maxdiff(gpu, cpu) <= TOL


def test_convolver_around_feedback(pkg, engine, oracle):
    """A reverb FEEDING an echo loop runs time-batched before the per-quantum part; a ConvolverNode INSIDE the loop is reported."""
    ir = G.synthetic_ir(2000, 1, decay=0.02)
    n = 128 * 70

    def build(be, g):
        pcm = G.c2_source(g, n)
        pcm[:, 128 * 4:] = 0
        c = pkg.OfflineAudioContext(2, n, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0]], G.SR), loop=True)
        cv = c.create_convolver(pkg.AudioBuffer(ir, G.SR))
        d = c.create_delay(max_delay_time=0.05, delay_time=[0.004, 0.0113][g])
        fb = c.create_gain(0.5)
        s.connect(cv)
        cv.connect(d)
        d.connect(fb)
        fb.connect(d)
        d.connect(c.destination())
        s.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 2)
    assert float(np.abs(cpu).max()) > 1e-3
    assert maxdiff(gpu, cpu) <= TOL

    c = pkg.OfflineAudioContext(1, 128 * 4, G.SR, engine.backend)
    s = c.create_constant_source()
    cv = c.create_convolver(pkg.AudioBuffer([np.ones(4, np.float32)], G.SR))
    d = c.create_delay(max_delay_time=0.05, delay_time=0.004)
    s.connect(d)
    d.connect(cv)
    cv.connect(d)
    d.connect(c.destination())
    s.start()
    with pytest.raises(pkg.WaeError) as e:
        c.start_rendering_sync()
    assert e.value.status == 4 and "feedback" in str(e.value)


# ---- OfflineAudioContext::suspend_sync: graph mutation at render time -----------------------------------------
def test_suspend_sync_reference_case(pkg, engine):
    import test_oracle_offline as O
    O.test_suspend_sync(pkg, engine.backend)          # src/context/offline.rs:469-511 on the CUDA path
    O.test_suspend_argument_errors(pkg, engine.backend)


@pytest.mark.parametrize("chunk", [0, 256])
def test_suspend_sync_mutations_keep_node_state(pkg, engine, oracle, chunk):
    """Nodes that live across a suspend point keep their state (filter memory, delay line, oscillator phase, automation);
    nodes, connections, param events and sources added or removed in the callbacks take effect at the suspend frame."""
    n = 128 * 40

    def build(be, g):
        pcm = G.c2_source(g, n)
        c = pkg.OfflineAudioContext(2, n, G.SR, be)
        src = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        bq = c.create_biquad_filter(type_=pkg.BANDPASS, frequency=900.0, q=12.0)
        dl = c.create_delay(max_delay_time=0.05, delay_time=0.013)
        gn = c.create_gain(0.8)
        gn.gain.linear_ramp_to_value_at_time(0.2, 0.09)
        osc = c.create_oscillator(type_=pkg.SAWTOOTH, frequency=333.0)
        og = c.create_gain(0.1)
        src.connect(bq)
        bq.connect(dl)
        dl.connect(gn)
        gn.connect(c.destination())
        osc.connect(og)
        og.connect(c.destination())
        src.start()
        osc.start()

        def first(ctx):   # new branch: a second filter in parallel, an automation event on a running ramp, a new source
            hp = ctx.create_biquad_filter(type_=pkg.HIGHPASS, frequency=2500.0)
            bq.connect(hp)
            hp.connect(ctx.destination())
            gn.gain.set_value_at_time(0.9, ctx.current_time() + 0.004)
            k = ctx.create_constant_source(offset=0.05)
            k.connect(ctx.destination())
            k.start_at(ctx.current_time())
            k.stop_at(ctx.current_time() + 0.01)

        def second(ctx):  # remove the oscillator branch, re-route the delay
            og.disconnect()
            dl.disconnect()
            dl.connect(ctx.destination())

        if g == 0:
            c.suspend_sync(128 * 7 / G.SR, first)
            c.suspend_sync(128 * 19 / G.SR + 1e-4, second)   # quantised up: frame 128 * 20
        elif g == 1:
            c.suspend_sync(128 * 11 / G.SR, second)
        return c                                             # g == 2: no suspend point, same batch

    engine.set_option(pkg.OPT_CHUNK_FRAMES, chunk)
    try:
        gpu, cpu = both(pkg, engine, oracle, build, 3)
    finally:
        engine.set_option(pkg.OPT_CHUNK_FRAMES, 0)
    assert maxdiff(gpu, cpu) <= TOL


# ---- AudioBufferSource slow track ---------------------------------------------------------------
SLOW_CASES = {
    "rate_half": dict(playback_rate=0.5),
    "rate_1p7_detune": dict(playback_rate=1.7, detune=-130.0),
    "subsample_start": dict(start=0.00123),
    "offset_duration": dict(start=0.004, offset=0.0112, duration=0.0305),
    "stop": dict(stop=0.0391),
    "loop_custom_points": dict(loop=True, loop_start=0.0103, loop_end=0.0377, playback_rate=1.3),
    "loop_default_rate": dict(loop=True, playback_rate=0.77),
    "buffer_44k1_in_48k": dict(buffer_sr=44100.0),
    # renderer frame loop on the GPU (k_buffer_source_serial): what the closed-form tracks do not cover
    "reverse": dict(playback_rate=-1.0, offset=0.05, start=0.0, duration=1e308),
    "reverse_loop": dict(playback_rate=-0.8, loop=True, loop_start=0.0103, loop_end=0.0377, offset=0.03, start=0.002, duration=1e308),
    "rate_zero": dict(playback_rate=0.0, offset=0.01, start=0.0, duration=1e308),
    "tiny_loop": dict(loop=True, loop_start=0.0100, loop_end=0.01004, playback_rate=1.1),
    "rate_ramp": dict(auto="rate_ramp"),
    "detune_steps_loop": dict(auto="detune_steps", loop=True),
    "rate_lfo_krate": dict(auto="rate_lfo", loop=True, loop_start=0.005, loop_end=0.04),
    "rate_back_to_one": dict(auto="rate_back_to_one"),
}


@pytest.mark.parametrize("name", sorted(SLOW_CASES))
def test_buffer_source_slow_track(pkg, engine, oracle, name):
    o = SLOW_CASES[name]

    def build(be, g):
        rng = np.random.default_rng(50 + g)
        t = np.arange(2600) / 48000.0
        pcm = [(np.sin(2 * np.pi * (300 + 90 * g) * t + c) * 0.7 + 0.05 * rng.standard_normal(2600)).astype(np.float32) for c in range(2)]
        c = pkg.OfflineAudioContext(2, 128 * 60, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer(pcm, o.get("buffer_sr", G.SR)), detune=o.get("detune", 0.0),
                                   playback_rate=o.get("playback_rate", 1.0), loop=o.get("loop", False),
                                   loop_start=o.get("loop_start", 0.0), loop_end=o.get("loop_end", 0.0))
        s.connect(c.destination())
        auto = o.get("auto")
        if auto == "rate_ramp":
            s.playback_rate.set_value(0.5)
            s.playback_rate.linear_ramp_to_value_at_time(2.0, 0.05)
        elif auto == "detune_steps":
            s.detune.set_value_at_time(700.0, 0.01)
            s.detune.set_value_at_time(-500.0, 0.03)
            s.detune.exponential_ramp_to_value_at_time(-100.0, 0.09)
        elif auto == "rate_lfo":
            lfo = c.create_oscillator(frequency=9.0)
            amp = c.create_gain(gain=0.6)
            lfo.connect(amp)
            amp.connect(s.playback_rate)
            lfo.start()
        elif auto == "rate_back_to_one":
            s.playback_rate.set_value_at_time(1.5, 0.01)
            s.playback_rate.set_value_at_time(1.0, 0.02)
        if "offset" in o:
            s.start_at_with_offset_and_duration(o["start"], o["offset"], o["duration"])
        else:
            s.start_at(o.get("start", 0.0))
        if "stop" in o:
            s.stop_at(o["stop"])
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 2)
    assert float(np.abs(cpu).max()) > 0.3
    assert maxdiff(gpu, cpu) <= TOL


# ---- AudioParam automation and audio-rate modulation ------------------------------------------
AUTOMATIONS = {
    "linear": lambda p: (p.set_value(0.2), p.linear_ramp_to_value_at_time(1.0, 0.013)),
    "exponential": lambda p: (p.set_value(0.01), p.exponential_ramp_to_value_at_time(0.9, 0.021)),
    "set_value_at_time": lambda p: (p.set_value(0.1), p.set_value_at_time(0.7, 0.0051), p.set_value_at_time(0.3, 0.0102)),
    "set_target": lambda p: (p.set_value(0.0), p.set_target_at_time(1.0, 0.003, 0.004)),
    "set_target_then_ramp": lambda p: (p.set_value(1.0), p.set_target_at_time(0.0, 0.002, 0.003), p.linear_ramp_to_value_at_time(0.8, 0.02)),
    "value_curve": lambda p: p.set_value_curve_at_time([0.0, 1.0, 0.25, 0.75, 0.0], 0.004, 0.017),
    "cancel_and_hold": lambda p: (p.set_value(0.0), p.linear_ramp_to_value_at_time(1.0, 0.02), p.cancel_and_hold_at_time(0.011)),
    "cancel_scheduled": lambda p: (p.set_value(0.5), p.set_value_at_time(0.9, 0.004), p.linear_ramp_to_value_at_time(0.1, 0.02),
                                   p.cancel_scheduled_values(0.01)),
    "k_rate_ramp": lambda p: (p.set_automation_rate("k"), p.set_value(0.0), p.linear_ramp_to_value_at_time(1.0, 0.02)),
}


@pytest.mark.parametrize("name", sorted(AUTOMATIONS))
def test_param_automation_on_gain(pkg, engine, oracle, name):
    def build(be, g):
        c = pkg.OfflineAudioContext(1, 128 * 12 + 50, G.SR, be)
        src = c.create_oscillator(frequency=700.0)
        gn = c.create_gain()
        AUTOMATIONS[name](gn.gain)
        src.connect(gn)
        gn.connect(c.destination())
        src.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build)
    assert maxdiff(gpu, cpu) <= TOL


@pytest.mark.parametrize("name", ["linear", "set_target", "value_curve"])
def test_param_automation_on_constant_source(pkg, engine, oracle, name):
    def build(be, g):
        c = pkg.OfflineAudioContext(1, 128 * 12, G.SR, be)
        src = c.create_constant_source()
        AUTOMATIONS[name](src.offset)
        src.connect(c.destination())
        src.start_at(0.0007)
        return c

    gpu, cpu = both(pkg, engine, oracle, build)
    assert maxdiff(gpu, cpu) <= 2e-6


def test_audio_rate_param_inputs_offline_rs(pkg, engine):
    # tests/offline.rs:114-149 test_audio_param_graph on the CUDA path: intrinsic 0.5 + two audio-rate inputs
    c = pkg.OfflineAudioContext(1, 128, 48000.0, engine.backend)
    gain = c.create_gain()
    gain.gain.set_value(0.5)
    gain.connect(c.destination())
    source = c.create_constant_source()
    source.offset.set_value(0.8)
    source.connect(gain)
    for v in (0.1, 0.3):
        s2 = c.create_constant_source()
        s2.offset.set_value(v)
        s2.connect(gain.gain)
        s2.start()
    source.start()
    out = c.start_rendering_sync().get_channel_data(0)
    assert np.array_equal(out, np.full(128, np.float32(0.8) * np.float32(0.9), np.float32))


def test_fm_synthesis_audio_rate_frequency(pkg, engine, oracle):
    # an LFO / FM patch: modulator -> gain (depth) -> carrier.frequency, plus a detune ramp on the carrier
    def build(be, g):
        c = pkg.OfflineAudioContext(1, 128 * 40, G.SR, be)
        car = c.create_oscillator(type_=[pkg.SINE, pkg.SAWTOOTH][g], frequency=440.0)
        mod = c.create_oscillator(frequency=110.0)
        depth = c.create_gain(150.0)
        mod.connect(depth)
        depth.connect(car.frequency)
        car.detune.linear_ramp_to_value_at_time(700.0, 0.08)
        car.connect(c.destination())
        mod.start()
        car.start_at(0.0021)
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 2)
    assert maxdiff(gpu, cpu) <= TOL


def test_biquad_frequency_sweep(pkg, engine, oracle):
    def build(be, g):
        pcm = G.c2_source(g, 128 * 30)
        c = pkg.OfflineAudioContext(2, 128 * 30, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        f = c.create_biquad_filter(type_=[pkg.LOWPASS, pkg.PEAKING][g], frequency=200.0, q=4.0, gain=6.0)
        f.frequency.exponential_ramp_to_value_at_time(6000.0, 0.06)
        f.q.linear_ramp_to_value_at_time(0.7, 0.05)
        s.connect(f)
        f.connect(c.destination())
        s.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 2)
    assert maxdiff(gpu, cpu) <= TOL


def test_panner_and_delay_time_automation(pkg, engine, oracle):
    def build(be, g):
        pcm = G.c2_source(g, 128 * 30)
        c = pkg.OfflineAudioContext(2, 128 * 30, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]] if g else [pcm[0]], G.SR))
        sp = c.create_stereo_panner(pan=-1.0)
        sp.pan.linear_ramp_to_value_at_time(1.0, 0.07)
        d = c.create_delay(max_delay_time=0.1, delay_time=0.001)
        d.delay_time.linear_ramp_to_value_at_time(0.02, 0.06)
        s.connect(sp)
        sp.connect(d)
        d.connect(c.destination())
        s.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 2)
    assert maxdiff(gpu, cpu) <= TOL


def test_plain_c_client_renders_what_the_oracle_renders(pkg, engine, oracle, tmp_path):
    """examples/c_client.c drives the C ABI from C (own process, no Python): its per-graph RMS must be the oracle's for the same graphs."""
    import subprocess
    from test_abi_cpu import build_c_client
    r = subprocess.run([build_c_client(tmp_path), "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = [float(line.split("rms")[1]) for line in r.stdout.splitlines() if "rms" in line]
    assert len(got) == 3
    for g in range(3):
        c = pkg.OfflineAudioContext(2, 48000, 48000.0, oracle)
        osc = c.create_oscillator(type_=pkg.SAWTOOTH, frequency=110.0 * (g + 1))
        bq = c.create_biquad_filter(type_=pkg.LOWPASS, frequency=900.0, q=2.0)
        gn = c.create_gain(0.5)
        osc.connect(bq)
        bq.connect(gn)
        gn.connect(c.destination())
        gn.gain.linear_ramp_to_value_at_time(0.0, 1.0)
        osc.start()
        left = c.start_rendering_sync().get_channel_data(0).astype(np.float64)
        assert abs(got[g] - float(np.sqrt(np.mean(left * left)))) <= 2e-6


def _oneshot_mixed_batch(pkg, be, n_graphs, length, ir):
    """graphs of four kinds, interleaved: fused buffer chains, voice sums, convolvers sharing one response, a delay loop"""
    ctxs = []
    for g in range(n_graphs):
        k = g % 4
        if k == 0:
            ctxs.append(G.c2_buffer_biquad_gain(pkg, be, g, length))
        elif k == 1:
            ctxs.append(G.c3_many_voices(pkg, be, 12, length))
        elif k == 2:
            ctxs.append(G.c4_convolver(pkg, be, g, length, ir))
        else:
            c = pkg.OfflineAudioContext(2, length, G.SR, be)
            o = c.create_oscillator(type_=pkg.SAWTOOTH, frequency=110.0 + g)
            d = c.create_delay(max_delay_time=0.05, delay_time=0.011)
            fb = c.create_gain(0.4)
            o.connect(d)
            d.connect(fb)
            fb.connect(d)
            d.connect(c.destination())
            o.start()
            ctxs.append(c)
    return ctxs


@pytest.mark.parametrize("pinned_out", [False, True])
def test_one_shot_render_many_groups(pkg, engine, oracle, pinned_out):
    """wae_render_batch(HOST) on a batch large enough to be cut into graph groups: sizing and planning run on the engine's worker
    threads, the source PCM goes up from the graphs' page-locked buffers, rendered PCM comes back through the staging slots (pageable
    `out`) or straight into a page-locked `out` — same PCM as prepare / run / fetch, and as the oracle."""
    n_graphs, length = 72, 8192 * 2 + 128 * 5 + 3
    ir = G.synthetic_ir(3000, 2, decay=0.3)
    want = pkg.Batch(_oneshot_mixed_batch(pkg, engine.backend, n_graphs, length, ir))
    want.run()
    want.sync()
    ref = want.fetch()
    want.destroy()
    if pinned_out:
        import torch
        t = torch.empty(n_graphs * 2 * length, dtype=torch.float32, pin_memory=True)
        out = t.numpy().reshape(n_graphs, 2, length)
    else:
        out = np.full((n_graphs, 2, length), np.nan, np.float32)
    for _ in range(2):  # the second call draws its device memory from the engine's cache
        out[...] = np.nan
        got = pkg.render_batch_oneshot(_oneshot_mixed_batch(pkg, engine.backend, n_graphs, length, ir), out)
        assert np.array_equal(got, ref)
    cpu = G.render(pkg, _oneshot_mixed_batch(pkg, oracle, 8, length, ir))
    assert maxdiff(got[:8], cpu) <= TOL


def test_one_shot_render_reports_planner_refusals(pkg):
    """an unsupported graph anywhere in the batch fails the whole one-shot call cleanly (status + text), nothing is left running"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    length = 128 * 40
    eng2 = pkg.Engine(0)  # a fresh engine has no HRIR sphere: HRTF panners are refused
    try:
        ctxs = [G.c2_buffer_biquad_gain(pkg, eng2.backend, g, length) for g in range(70)]
        bad = pkg.OfflineAudioContext(2, length, G.SR, eng2.backend)
        p = bad.create_panner(panning_model=pkg.context.HRTF)
        s2 = bad.create_constant_source()
        s2.connect(p)
        p.connect(bad.destination())
        s2.start()
        with pytest.raises(pkg.WaeError):
            pkg.render_batch_oneshot(ctxs + [bad])
        out = pkg.render_batch_oneshot(ctxs[:3])  # the engine is still usable
        assert np.isfinite(out).all() and np.abs(out).max() > 0
    finally:
        eng2.close()


@pytest.mark.parametrize("bad", [float("nan"), float("inf")])
def test_scan_biquad_recovers_from_a_non_finite_sample(pkg, engine, oracle, bad):
    """biquad_filter.rs:881-883: `if !y.is_normal() { y = 0. }` — a NaN / Inf in the source PCM costs the reference three output samples
    and its filter state is clean again; the time-parallel scan replays the affected tile serially with the same flush (k_chain)"""
    length = 128 * 40 + 9
    def build(be, g):
        rng = np.random.default_rng(77 + g)
        pcm = rng.uniform(-0.5, 0.5, (2, length)).astype(np.float32)
        pcm[0, 1000 + 300 * g] = bad
        pcm[1, 3000] = -bad if np.isinf(bad) else bad
        c = pkg.OfflineAudioContext(2, length, G.SR, be)
        src = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        bq = c.create_biquad_filter(type_=pkg.LOWPASS, frequency=900.0 + 100 * g, q=3.0)
        b2 = c.create_biquad_filter(type_=pkg.HIGHPASS, frequency=200.0, q=1.0)
        src.connect(bq)
        bq.connect(b2)
        b2.connect(c.destination())
        src.start()
        return c
    gpu, cpu = both(pkg, engine, oracle, build, 3)
    assert np.isfinite(cpu).all() and np.isfinite(gpu).all()
    assert maxdiff(gpu, cpu) <= TOL


@pytest.mark.parametrize("chunk", [0, 2048 * 36])
def test_few_long_renders_are_cut_into_slabs_that_hand_their_state_on_before_they_render(pkg, engine, oracle, chunk):
    """A handful of (graph, channel) pairs with a long render through one biquad: k_chain<PRE> (time slabs of 32768 frames; every slab
    first runs from zero state without storing, publishes G^L s_in + that, then renders).  Against the oracle, and against the same
    render with one CTA per pair (WAE_OPT_CHAIN_PREPASS = 0); sources: buffer (streamed), oscillator, constant; with / without a
    wave-shaper behind the filter; a state carried over a chunk boundary."""
    n = 2048 * 75 + 128 * 5   # 76 tiles, the last one ragged: slabs of 16 tiles
    curve = np.tanh(np.linspace(-2, 2, 257)).astype(np.float32)

    def build(be, g):
        c = pkg.OfflineAudioContext(2, n, G.SR, be)
        if g % 3 == 0:
            pcm = G.c2_source(g, n)
            src = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        elif g % 3 == 1:
            src = c.create_oscillator(type_=pkg.SAWTOOTH, frequency=97.0 + 13 * g)
        else:
            src = c.create_constant_source(offset=0.25)
        bq = c.create_biquad_filter(type_=[pkg.LOWPASS, pkg.BANDPASS, pkg.HIGHPASS, pkg.PEAKING][g % 4], frequency=[400.0, 30.0, 2500.0, 1200.0][g % 4],
                                    q=[0.7, 25.0, 2.0, 4.0][g % 4], gain=3.0)
        gn = c.create_gain(0.6)
        src.connect(bq)
        bq.connect(gn)
        last = gn
        if g >= 4:
            last = c.create_wave_shaper(curve)
            gn.connect(last)
        last.connect(c.destination())
        src.start()
        return c

    engine.set_option(pkg.OPT_CHUNK_FRAMES, chunk)
    try:
        gpu, cpu = both(pkg, engine, oracle, build, 7)
        engine.set_option(pkg.OPT_CHAIN_PREPASS, 0)
        one = G.render(pkg, [build(engine.backend, g) for g in range(7)])
    finally:
        engine.set_option(pkg.OPT_CHAIN_PREPASS, 1)
        engine.set_option(pkg.OPT_CHUNK_FRAMES, 0)
    assert maxdiff(gpu, cpu) <= TOL
    assert maxdiff(gpu, one) <= 2e-6   # same scan, the state crosses 4 slab boundaries through G^L instead of tile by tile


@pytest.mark.parametrize("bad", [np.nan, np.inf])
def test_a_non_finite_sample_in_a_slab_that_publishes_its_state_early(pkg, engine, oracle, bad):
    """k_chain<PRE> + biquad_filter.rs:881-883: a slab that flushed a NaN / Inf ends in the same state wherever it started, so it
    publishes its zero-state end state alone (the `flushed` flag); bad samples in the first, a middle and the last slab"""
    n = 2048 * 70

    def build(be, g):
        rng = np.random.default_rng(170 + g)
        pcm = rng.uniform(-0.5, 0.5, (2, n)).astype(np.float32)
        pcm[0, 5000 + 100 * g] = bad
        pcm[1, 2048 * 16 * 2 + 777] = -bad if np.isinf(bad) else bad
        pcm[0, n - 3000] = bad
        c = pkg.OfflineAudioContext(2, n, G.SR, be)
        src = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        bq = c.create_biquad_filter(type_=pkg.LOWPASS, frequency=700.0 + 150 * g, q=6.0)
        src.connect(bq)
        bq.connect(c.destination())
        src.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 3)
    assert np.isfinite(cpu).all() and np.isfinite(gpu).all()
    assert maxdiff(gpu, cpu) <= TOL


@pytest.mark.parametrize("bad", [np.nan, -np.inf])
def test_automated_biquad_scan_recovers_from_a_non_finite_sample(pkg, engine, oracle, bad):
    """k_biquad_arate evaluates a quantum as a scan over affine maps; a NaN / Inf in the input poisons the states handed between the lanes,
    so that quantum is evaluated again by the reference's serial code with its flush (biquad_filter.rs:881-883)"""
    length = 128 * 30

    def build(be, g):
        rng = np.random.default_rng(311 + g)
        pcm = rng.uniform(-0.5, 0.5, (2, length)).astype(np.float32)
        pcm[0, 700 + 129 * g] = bad
        pcm[1, 128 * 11] = bad          # first frame of a quantum
        pcm[1, 128 * 20 + 127] = bad    # last frame of a quantum
        c = pkg.OfflineAudioContext(2, length, G.SR, be)
        src = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
        bq = c.create_biquad_filter(type_=pkg.LOWPASS, frequency=300.0, q=4.0)
        bq.frequency.exponential_ramp_to_value_at_time(6000.0, 0.07)
        bq.q.linear_ramp_to_value_at_time(0.5, 0.05)
        src.connect(bq)
        bq.connect(c.destination())
        src.start()
        return c

    gpu, cpu = both(pkg, engine, oracle, build, 3)
    assert np.isfinite(cpu).all() and np.isfinite(gpu).all()
    assert maxdiff(gpu, cpu) <= TOL


def test_one_shot_render_into_registered_caller_memory(pkg, engine):
    """wae_host_register: the caller's own buffer page-locked once, D2H of the one-shot render lands in it directly"""
    import ctypes as C
    n, length = 70, 128 * 50 + 9
    api = engine.backend.api
    out = np.full((n, 2, length), np.nan, np.float32)
    api.check(api.host_register(engine.backend.engine, out.ctypes.data_as(C.c_void_p), out.nbytes))
    try:
        got = pkg.render_batch_oneshot([G.c2_buffer_biquad_gain(pkg, engine.backend, g, length) for g in range(n)], out)
    finally:
        api.check(api.host_unregister(engine.backend.engine, out.ctypes.data_as(C.c_void_p)))
    ref = pkg.render_batch_oneshot([G.c2_buffer_biquad_gain(pkg, engine.backend, g, length) for g in range(n)])
    assert np.array_equal(got, ref)


def test_one_audio_buffer_played_by_many_sources(pkg, engine, oracle):
    # one AudioBuffer (one copy in the device slab) played by sources on every track at once: the aligned fast track (fused chain), offsets and
    # durations, a resampling playback rate, a loop with custom points, a rate ramp (serial track), a second buffer in between, two suspend-free
    # graphs per batch — each node must find its samples at the shared offset
    rng = np.random.default_rng(77)
    length = 128 * 60 + 17

    def build(be, g):
        c = pkg.OfflineAudioContext(2, length, 48000.0, be)
        pcm = rng.uniform(-1, 1, (2, 3000)).astype(np.float32) if be is engine.backend else build.pcm[g]
        build.pcm[g] = pcm
        buf = pkg.AudioBuffer(list(pcm), 48000.0)
        mono = pkg.AudioBuffer([pcm[0, :1777] * 0.5], 44100.0)
        s = c.create_buffer_source(buf)
        s.connect(c.destination())
        s.start()                                            # fast track
        for i in range(6):
            gn = c.create_gain()
            gn.gain.value = 0.3 + 0.1 * i
            gn.connect(c.destination())
            s = c.create_buffer_source(buf if i != 3 else mono, playback_rate=1.0 if i % 2 == 0 else 0.77 + 0.1 * i)
            s.connect(gn)
            s.start_at_with_offset_and_duration(0.002 * i + 0.0001 * g, 0.004 * i, 0.03 + 0.01 * i)
        lp = c.create_buffer_source(buf, loop=True, loop_start=0.01, loop_end=0.03)
        lp.connect(c.destination())
        lp.start_at(0.01)
        rp = c.create_buffer_source(buf)
        rp.playback_rate.set_value_at_time(0.5, 0.0)
        rp.playback_rate.linear_ramp_to_value_at_time(2.0, 0.1)
        rp.connect(c.destination())
        rp.start_at(0.005)
        return c
    build.pcm = {}
    gpu, cpu = both(pkg, engine, oracle, build, 2)
    assert np.abs(cpu).max() > 1.0
    assert maxdiff(gpu, cpu) <= TOL


def test_a_batch_of_few_large_graphs_is_planned_by_several_workers(pkg, engine, oracle):
    # >= 4096 nodes in a batch of few graphs without suspend points: the sizing pass and the planning pass are split over the host workers
    # (prep_begin: size_group_split, prep_plan_group: merge_builds) — the merged tables must be the ones a single planner builds: buffer
    # sources between the voice graphs (slab offsets per run), biquad chains (scan-constant indices), a convolver per voice graph (conv-input
    # indices), mixes (edge offsets).  Rendered through prepare / run / fetch AND through the one-shot call (planned on a worker there).
    length = 128 * 37 + 11
    ir = G.synthetic_ir(9000, 2, decay=0.5)

    def build(be, g):
        if g % 2 == 0:
            return G.c2_buffer_biquad_gain(pkg, be, g, length)
        return G.north_star_voices_convolver(pkg, be, 260, length, ir, seed=g)
    n = 5
    # (5 graphs, 5237 nodes in processing order: above the 4096-node threshold of the split)
    ctxs = [build(engine.backend, g) for g in range(n)]
    gpu = G.render(pkg, ctxs)
    cpu = G.render(pkg, [build(oracle, g) for g in range(n)])
    assert maxdiff(gpu, cpu) <= TOL
    host = np.zeros((n, 2, length), np.float32)
    pkg.render_batch_oneshot([build(engine.backend, g) for g in range(n)], host)
    assert np.array_equal(host, gpu)
