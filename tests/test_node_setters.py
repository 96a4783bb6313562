"""The reference's post-construction setters (one control message each): AudioBufferSourceNode::set_buffer / set_loop / set_loop_start /
set_loop_end, ConvolverNode::set_buffer / set_normalize, WaveShaperNode::set_curve / set_oversample, OscillatorNode::set_periodic_wave,
PannerNode::set_*, AnalyserNode::set_* — the way the reference's own examples and benchmarks build graphs
(`create_buffer_source()` then `set_buffer(..)`, `create_panner()` then `set_panning_model(HRTF)`).
A node configured through the setters must be THE node its options would have built: same PCM on the oracle, same plan in the library
(host only here; tests/test_gpu_reference_cases.py renders the graph-level functions on CUDA)."""
import numpy as np
import pytest

import graphs as G

RQ = 128
SR = 48000.0


def _noise(n, ch=1, seed=11):
    return list(np.random.default_rng(seed).uniform(-0.8, 0.8, (ch, n)).astype(np.float32))


def _variants(pkg, be, with_setters):
    """One context that uses every setter (with_setters) or the equivalent construction options."""
    c = pkg.OfflineAudioContext(2, RQ * 24, SR, be)
    buf = pkg.AudioBuffer(_noise(1500, 2), SR)
    ir = pkg.AudioBuffer(_noise(700, 2, seed=12), SR)
    curve = np.tanh(np.linspace(-2, 2, 33)).astype(np.float32)
    table = np.sin(np.arange(2048) / 2048.0 * 2 * np.pi * 3).astype(np.float32)
    if with_setters:
        src = c.create_buffer_source()
        src.set_buffer(buf)
        src.set_loop(True)
        src.set_loop_start(0.004)
        src.set_loop_end(0.02)
        conv = c.create_convolver()
        conv.set_normalize(False)
        conv.set_buffer(ir)
        conv.set_normalize(True)  # only the NEXT set_buffer would see it (convolver.rs:325-328)
        sh = c.create_wave_shaper()
        sh.set_curve(curve)
        sh.set_oversample(pkg.OVERSAMPLE_X2)
        osc = c.create_oscillator(frequency=300.0)
        osc.set_periodic_wave(table)
        pan = c.create_panner()
        pan.set_distance_model(pkg.context.EXPONENTIAL)
        pan.set_ref_distance(2.0)
        pan.set_max_distance(50.0)
        pan.set_rolloff_factor(1.5)
        pan.set_cone_inner_angle(40.0)
        pan.set_cone_outer_angle(100.0)
        pan.set_cone_outer_gain(0.25)
        pan.position_x.set_value(3.0)
        pan.position_z.set_value(-2.0)
        an = c.create_analyser()
        an.set_fft_size(512)
        an.set_smoothing_time_constant(0.3)
        an.set_min_decibels(-90.0)
        an.set_max_decibels(-10.0)
    else:
        src = c.create_buffer_source(buf, loop=True, loop_start=0.004, loop_end=0.02)
        conv = c.create_convolver(ir, disable_normalization=True)
        sh = c.create_wave_shaper(curve=curve, oversample=pkg.OVERSAMPLE_X2)
        osc = c.create_oscillator(frequency=300.0, periodic_wave=table)
        pan = c.create_panner(distance_model=pkg.context.EXPONENTIAL, ref_distance=2.0, max_distance=50.0, rolloff_factor=1.5,
                              cone_inner_angle=40.0, cone_outer_angle=100.0, cone_outer_gain=0.25, position=(3.0, 0.0, -2.0))
        an = c.create_analyser(fft_size=512, smoothing_time_constant=0.3, min_decibels=-90.0, max_decibels=-10.0)
    src.connect(conv)
    conv.connect(sh)
    osc.connect(sh)
    sh.connect(pan)
    pan.connect(an)
    an.connect(c.destination())
    src.start()
    osc.start()
    c._analyser = an
    return c


def test_setters_build_the_node_the_options_build(pkg, oracle):
    a = _variants(pkg, oracle, True)
    b = _variants(pkg, oracle, False)
    pa, pb = a.start_rendering_sync(), b.start_rendering_sync()
    assert np.abs(pa.get_channel_data(0)).max() > 1e-3
    for ch in range(2):
        assert np.array_equal(pa.get_channel_data(ch), pb.get_channel_data(ch))
    assert np.array_equal(a._analyser.get_float_frequency_data(), b._analyser.get_float_frequency_data())
    assert np.array_equal(a._analyser.get_byte_frequency_data(), b._analyser.get_byte_frequency_data())


def test_setters_give_the_library_the_same_plan(pkg, builder):
    if not builder.api.is_product:
        pytest.skip("plan is the library's")
    pa = pkg.context.plan_batch([_variants(pkg, builder, True)])
    pb = pkg.context.plan_batch([_variants(pkg, builder, False)])
    assert pa == pb and "k_shaper_os" in pa["kinds"] and "k_conv_mac_ifft" in pa["kinds"]


def test_setter_errors(pkg, builder):
    # audio_buffer_source.rs:283-286, waveshaper.rs:204-207 "cannot assign ... twice"; convolver.rs:520-548; panner.rs:560-640; analysis.rs:592-653
    c = pkg.OfflineAudioContext(2, RQ, SR, builder)
    buf = pkg.AudioBuffer(_noise(64), SR)
    s = c.create_buffer_source(buf)
    with pytest.raises(pkg.WaeError) as e:
        s.set_buffer(buf)
    assert "cannot assign buffer twice" in str(e.value)
    s = c.create_buffer_source()
    s.set_buffer(buf)
    with pytest.raises(pkg.WaeError):
        s.set_buffer(buf)
    sh = c.create_wave_shaper(curve=np.array([1.0], np.float32))
    with pytest.raises(pkg.WaeError) as e:  # waveshaper.rs:624-647 change_a_curve_for_another_curve_should_panic
        sh.set_curve(np.array([2.0], np.float32))
    assert "cannot assign curve twice" in str(e.value)
    sh = c.create_wave_shaper()             # waveshaper.rs:649-669 change_none_for_curve_after_build
    sh.set_curve(np.array([2.0], np.float32))
    sh.set_oversample(pkg.OVERSAMPLE_X4)
    cv = c.create_convolver()
    with pytest.raises(pkg.WaeError):
        cv.set_buffer(pkg.AudioBuffer(_noise(64), 44100.0))
    with pytest.raises(pkg.WaeError):
        cv.set_buffer(pkg.AudioBuffer(_noise(64, 3), SR))
    cv.set_buffer(buf)
    cv.set_buffer(pkg.AudioBuffer(_noise(32, 2), SR))  # a convolver may get another response
    p = c.create_panner()
    for setter, bad in [("set_ref_distance", -1.0), ("set_max_distance", 0.0), ("set_rolloff_factor", -0.5), ("set_cone_outer_gain", 1.5),
                        ("set_cone_outer_gain", -0.1), ("set_distance_model", 7), ("set_panning_model", 5)]:
        with pytest.raises(pkg.WaeError):
            getattr(p, setter)(bad)
    a = c.create_analyser()
    for setter, bad in [("set_fft_size", 13), ("set_fft_size", 16), ("set_fft_size", 65536), ("set_smoothing_time_constant", -1.0),
                        ("set_smoothing_time_constant", 2.0), ("set_min_decibels", -30.0), ("set_max_decibels", -100.0)]:
        with pytest.raises(pkg.WaeError):
            getattr(a, setter)(bad)
    a.set_min_decibels(-20.0 - 30.0)
    a.set_max_decibels(10.0)   # analysis.rs:592-597 test_set_decibels
    g = c.create_gain()
    with pytest.raises(pkg.WaeError):
        g._set_attribute(pkg._binding.ATTR_LOOP, 1.0)  # a gain has no loop


def test_buffer_source_configured_the_way_the_reference_examples_do(pkg, oracle):
    # examples/benchmarks.rs:97-105 (and every test of audio_buffer_source.rs): create_buffer_source(); set_buffer(..); set_loop(true); start()
    c = pkg.OfflineAudioContext(1, RQ * 4, SR, oracle)
    pcm = _noise(100)[0]
    src = c.create_buffer_source()
    src.set_buffer(pkg.AudioBuffer([pcm], SR))
    src.set_loop(True)
    src.connect(c.destination())
    src.start()
    out = c.start_rendering_sync().get_channel_data(0)
    assert np.array_equal(out, np.tile(pcm, 6)[:RQ * 4])


def test_hrtf_selected_with_set_panning_model(pkg, oracle):
    # benches/my_benchmark.rs:262-263: create_panner() then set_panning_model(HRTF)
    oracle.set_hrir_sphere(G.synthetic_hrir_sphere(44100, 384))  # (>= 258 taps: shorter responses vanish in the resampler)

    def build(setter):
        c = pkg.OfflineAudioContext(2, RQ * 6, SR, oracle)
        osc = c.create_oscillator()
        if setter:
            p = c.create_panner()
            p.set_panning_model(pkg.context.HRTF)
        else:
            p = c.create_panner(panning_model=pkg.context.HRTF)
        p.position_x.set_value(10.0)
        osc.connect(p)
        p.connect(c.destination())
        osc.start()
        return c.start_rendering_sync()

    a, b = build(True), build(False)
    assert np.array_equal(a.get_channel_data(0), b.get_channel_data(0)) and np.array_equal(a.get_channel_data(1), b.get_channel_data(1))
    assert np.abs(a.get_channel_data(1)).max() > 1e-3


def test_setters_at_a_suspend_point(pkg, builder):
    # a control message sent from a suspend_sync callback takes effect at that quantum.  The library plans the two segments separately;
    # two changes are not lowered (the graph keeps the CPU renderer): loop attributes of a started source, a second impulse response
    def ctx():
        c = pkg.OfflineAudioContext(2, 8192 * 3, SR, builder)
        src = c.create_buffer_source(pkg.AudioBuffer(_noise(3000, 2), SR), loop=True)
        pan = c.create_panner(position=(2.0, 0.0, -1.0))
        conv = c.create_convolver(pkg.AudioBuffer(_noise(500, 2, seed=3), SR))
        src.connect(pan)
        pan.connect(conv)
        conv.connect(c.destination())
        src.start()
        return c, src, pan, conv

    c, src, pan, conv = ctx()
    c.suspend_sync(8192 / SR, lambda _c: (pan.set_distance_model(pkg.context.LINEAR), pan.set_rolloff_factor(0.5)))
    if builder.api.is_product:
        assert pkg.context.plan_batch([c])["segments"] == 2
    else:
        out = c.start_rendering_sync()
        assert np.abs(out.get_channel_data(0)[8192 + 2000:]).max() > 1e-4
    for change in (lambda s, p, v: s.set_loop(False), lambda s, p, v: v.set_buffer(pkg.AudioBuffer(_noise(100, 2), SR))):
        c, src, pan, conv = ctx()
        c.suspend_sync(8192 / SR, lambda _c, f=change, s=src, p=pan, v=conv: f(s, p, v))
        if builder.api.is_product:
            with pytest.raises(pkg.WaeError) as e:
                pkg.context.plan_batch([c])
            assert e.value.status == 4  # WAE_UNSUPPORTED
        else:
            c.start_rendering_sync()  # the CPU renderer handles both


def test_selective_disconnect(pkg, builder):
    # src/node/audio_node.rs:304-405 over ConcreteBaseAudioContext::disconnect (concrete_base.rs:474-507)
    c = pkg.OfflineAudioContext(2, RQ, SR, builder)
    src = c.create_constant_source()
    a, b = c.create_gain(0.5), c.create_gain(0.25)
    split = c.create_channel_splitter(2)
    merge = c.create_channel_merger(2)
    lfo = c.create_oscillator()
    src.connect(a)
    src.connect(b)
    a.connect(c.destination())
    b.connect(c.destination())
    b.connect(split)
    split.connect_from_output_to_input(merge, 0, 1)
    split.connect_from_output_to_input(merge, 1, 0)
    merge.connect(c.destination())
    lfo.connect(a.gain)
    src.start()
    lfo.start()
    before = c.render_order()
    assert a.id in before and b.id in before and merge.id in before
    src.disconnect_dest(b)                                           # src -> b only
    with pytest.raises(pkg.WaeError) as e:
        src.disconnect_dest(b)                                       # not connected any more
    assert "attempting to disconnect unconnected nodes" in str(e.value)
    split.disconnect_dest_from_output_to_input(merge, 0, 1)
    with pytest.raises(pkg.WaeError):
        split.disconnect_dest_from_output_to_input(merge, 0, 1)
    with pytest.raises(pkg.WaeError):
        split.disconnect_dest_from_output(merge, 5)                  # IndexSizeError - output port 5 is out of bounds
    with pytest.raises(pkg.WaeError):
        split.disconnect_dest_from_output_to_input(merge, 1, 9)      # IndexSizeError - input port 9 is out of bounds
    split.disconnect_output(1)                                       # no destination named: nothing to complain about, ever
    split.disconnect_output(1)
    lfo.disconnect_dest(a.gain)                                      # towards an AudioParam
    with pytest.raises(pkg.WaeError):
        lfo.disconnect_dest(a.gain)
    other = pkg.OfflineAudioContext(2, RQ, SR, builder)
    with pytest.raises(pkg.WaeError):
        src.disconnect_dest(other.destination())                     # different contexts


def test_selective_disconnect_renders(pkg, oracle):
    # two constants into the destination; one is removed selectively at a suspend point
    c = pkg.OfflineAudioContext(1, RQ * 4, SR, oracle)
    k1, k2 = c.create_constant_source(offset=0.25), c.create_constant_source(offset=0.5)
    g = c.create_gain()
    k1.connect(g)
    k2.connect(g)
    k2.connect(c.destination())
    g.connect(c.destination())
    k1.start()
    k2.start()
    c.suspend_sync(2 * RQ / SR, lambda _c: k2.disconnect_dest(g))
    out = c.start_rendering_sync().get_channel_data(0)
    assert np.all(out[:2 * RQ] == 1.25) and np.all(out[2 * RQ:] == 0.75)


def _wave(api, real, imag, disable_normalization=False, n=2048):
    import ctypes as C
    fp = C.POINTER(C.c_float)
    r = None if real is None else np.ascontiguousarray(real, np.float32)
    i = None if imag is None else np.ascontiguousarray(imag, np.float32)
    length = len(r) if r is not None else (len(i) if i is not None else 0)
    out = np.zeros(n, np.float32)
    api.check(api.periodic_wave_table(None if r is None else r.ctypes.data_as(fp), None if i is None else i.ctypes.data_as(fp), length,
                                      1 if disable_normalization else 0, out.ctypes.data_as(fp), n))
    return out


def test_periodic_wave_tables(pkg, host_api):
    # src/periodic_wave.rs:278-345 wavetable_generate_sine / _2f_not_norm / _2f_norm / normalize, :221-276 the constructor panics
    i = np.arange(2048, dtype=np.float32)
    sine = np.sin(i / np.float32(2048) * np.float32(2) * np.float32(np.pi), dtype=np.float32)
    assert np.abs(_wave(host_api, [0.0, 0.0], [0.0, 1.0]) - sine).max() <= 1e-6
    assert np.abs(_wave(host_api, None, None) - sine).max() <= 1e-6              # the default is a sine
    two = (np.float32(0.5) * np.sin(i / np.float32(2048) * np.float32(2) * np.float32(np.pi)) +
           np.float32(0.5) * np.sin(np.float32(2) * i / np.float32(2048) * np.float32(2) * np.float32(np.pi))).astype(np.float32)
    assert np.abs(_wave(host_api, [0.0, 0.0, 0.0], [0.0, 0.5, 0.5], disable_normalization=True) - two).max() <= 1e-6
    assert np.abs(_wave(host_api, [0.0, 0.0, 0.0], [0.0, 0.5, 0.5]) - two / np.abs(two).max()).max() <= 1e-6
    assert np.abs(_wave(host_api, None, [0.0, 0.5, 0.5]) - two / np.abs(two).max()).max() <= 1e-6   # only `imag`
    cos = _wave(host_api, [0.0, 1.0], None)                                                       # only `real`
    assert abs(float(cos[0]) - 1.0) <= 1e-6 and abs(float(cos[512])) <= 1e-6
    for real, imag in [([0.0], None), (None, [0.0]), ([0.0], [0.0])]:                             # "length should at least 2"
        with pytest.raises(pkg.WaeError):
            _wave(host_api, real, imag)


def test_buffer_given_after_a_null_buffer_start_is_ignored(pkg, builder):
    # audio_buffer_source.rs:443-451 / test_null_buffer_start_ends_before_start_time (:1509-1534): started with no buffer and rendered ->
    # ended for good; a buffer given later (from a suspend callback) is never heard.  Given in the SAME callback as start() it plays.
    def build(start_in_callback):
        c = pkg.OfflineAudioContext(1, 48000, SR, builder)
        src = c.create_buffer_source()
        src.connect(c.destination())
        if not start_in_callback:
            src.start_at(0.75)

        def cb(ctx):
            if start_in_callback:
                src.start_at(0.75)
            src.set_buffer(pkg.AudioBuffer([np.ones(64, np.float32)], SR))
        c.suspend_sync(0.5, cb)
        return c

    if builder.api.is_product:
        dead, live = pkg.context.plan_batch([build(False)]), pkg.context.plan_batch([build(True)])
        # nothing but silence is rendered: no source kernel, no PCM upload (k_meta / k_mix_dyn carry the "silent" layout to the destination)
        assert set(dead["kinds"]) <= {"k_mix", "k_mix_dyn", "k_meta"} and dead["source_floats"] == 0
        assert any(k.startswith("k_buffer_source") or k == "k_chain" for k in live["kinds"]) and live["source_floats"] == 64
    else:
        assert not build(False).start_rendering_sync().get_channel_data(0).any()
        out = build(True).start_rendering_sync().get_channel_data(0)
        assert out[36000:36064].all() and not out[:36000].any()
