"""Pins the oracle against the reference's unit tests of its smaller node renderers: PannerNode (src/node/panner.rs:1081-1223),
ConstantSourceNode (constant_source.rs:307-372), ChannelMergerNode (channel_merger.rs:208-270), ChannelSplitterNode
(channel_splitter.rs:262-284), StereoPannerNode (stereo_panner.rs:368-553) and WaveShaperNode (waveshaper.rs:671-739).
Every function names the `#[test]` it restates and takes any backend; tests/test_gpu_reference_cases.py reruns them on CUDA."""
import numpy as np

RQ = 128


def _ones_source(pkg, c, channels, sr, n=RQ):
    src = c.create_buffer_source(pkg.AudioBuffer([np.ones(n, np.float32)] * channels, sr))
    src.start()
    return src


def _close(a, want, tol):
    a = np.asarray(a, np.float64)
    assert np.abs(a - np.asarray(want, np.float64)).max() <= tol, (a[:8], want)


def test_equal_power_mono_to_stereo(pkg, oracle):  # panner.rs:1081-1131
    sr = 44100.0
    c = pkg.OfflineAudioContext(2, RQ * 4, sr, oracle)
    src = _ones_source(pkg, c, 1, sr)
    panner = c.create_panner(panning_model=pkg.EQUALPOWER, cfg=pkg.channel_config(1, pkg.CLAMPED_MAX))
    panner.position_x.set_value(1.0)  # sound comes from the right
    src.connect(panner)
    panner.connect(c.destination())
    out = c.start_rendering_sync()
    _close(out.get_channel_data(0)[:128], 0.0, 1e-6)
    _close(out.get_channel_data(1)[:128], 1.0, 1e-6)
    _close(out.get_channel_data(0)[128:256], 0.0, 1e-6)  # no tail time
    _close(out.get_channel_data(1)[128:256], 0.0, 1e-6)


def test_equal_power_azimuth_mono_to_stereo(pkg, oracle):  # panner.rs:1133-1170
    sr = 44100.0
    c = pkg.OfflineAudioContext(2, RQ, sr, oracle)
    src = _ones_source(pkg, c, 1, sr)
    panner = c.create_panner(panning_model=pkg.EQUALPOWER)
    panner.position_y.set_value(1.0)  # sound comes from above: both ears receive equal volume
    src.connect(panner)
    panner.connect(c.destination())
    out = c.start_rendering_sync()
    _close(out.get_channel_data(0), np.sqrt(np.float32(0.5)), 1e-6)
    _close(out.get_channel_data(1), np.sqrt(np.float32(0.5)), 1e-6)


def test_equal_power_stereo_to_stereo(pkg, oracle):  # panner.rs:1172-1223
    sr = 44100.0
    c = pkg.OfflineAudioContext(2, RQ, sr, oracle)
    lis = c.listener()  # listener at (10, 0, 0), looking along +x, up = +z
    for name, v in [("position_x", 10.0), ("position_y", 0.0), ("position_z", 0.0), ("forward_x", 1.0), ("forward_y", 0.0),
                    ("forward_z", 0.0), ("up_x", 0.0), ("up_y", 0.0), ("up_z", 1.0)]:
        getattr(lis, name).set_value(v)
    src = _ones_source(pkg, c, 2, sr)
    panner = c.create_panner()
    panner.position_x.set_value(10.0)
    panner.position_y.set_value(10.0)
    panner.position_z.set_value(0.0)
    src.connect(panner)
    panner.connect(c.destination())
    out = c.start_rendering_sync()
    _close(out.get_channel_data(0), 0.2, 1e-3)  # both channels summed to the left, distance 10 -> x 0.1
    _close(out.get_channel_data(1), 0.0, 1e-3)


def test_constant_source_start_stop(pkg, oracle):  # constant_source.rs:307-338
    sr = 48000.0
    c = pkg.OfflineAudioContext(1, RQ * 4, sr, oracle)
    src = c.create_constant_source()
    src.connect(c.destination())
    src.start_at(129.0 / sr)
    src.stop_at(257.0 / sr)
    ch = c.start_rendering_sync().get_channel_data(0)
    want = np.zeros(RQ * 4, np.float32)
    want[129:257] = 1.0
    assert np.array_equal(ch, want)


def test_constant_source_start_in_the_past(pkg, oracle):  # constant_source.rs:340-357
    sr = 48000.0
    c = pkg.OfflineAudioContext(1, 2 * RQ, sr, oracle)

    def at_suspend(ctx):
        src = ctx.create_constant_source()
        src.connect(ctx.destination())
        src.start_at(0.0)

    c.suspend_sync(RQ / sr, at_suspend)
    ch = c.start_rendering_sync().get_channel_data(0)
    assert np.array_equal(ch[:RQ], np.zeros(RQ, np.float32)) and np.array_equal(ch[RQ:], np.ones(RQ, np.float32))


def test_constant_source_start_in_the_future_while_dropped(pkg, oracle):  # constant_source.rs:359-372
    sr = 48000.0
    c = pkg.OfflineAudioContext(1, 4 * RQ, sr, oracle)
    src = c.create_constant_source()
    src.connect(c.destination())
    src.start_at(258.0 / sr)
    del src  # the control handle goes away, the renderer keeps its schedule
    ch = c.start_rendering_sync().get_channel_data(0)
    assert np.array_equal(ch[:258], np.zeros(258, np.float32)) and np.array_equal(ch[258:], np.ones(254, np.float32))


def _merger_graph(pkg, c):
    merger = c.create_channel_merger(2)
    merger.connect(c.destination())
    srcs = []
    for i, v in enumerate((2.0, 3.0)):
        s = c.create_constant_source()
        s.offset.set_value(v)
        s.connect_from_output_to_input(merger, 0, i)
        s.start()
        srcs.append(s)
    return srcs


def test_channel_merger(pkg, oracle):  # channel_merger.rs:208-233 test_merge
    c = pkg.OfflineAudioContext(2, RQ, 48000.0, oracle)
    _merger_graph(pkg, c)
    out = c.start_rendering_sync()
    assert np.array_equal(out.get_channel_data(0), np.full(RQ, 2.0, np.float32))
    assert np.array_equal(out.get_channel_data(1), np.full(RQ, 3.0, np.float32))


def test_channel_merger_disconnect(pkg, oracle):  # channel_merger.rs:235-270 test_merge_disconnect
    sr, length = 48000.0, 4 * RQ
    c = pkg.OfflineAudioContext(2, length, sr, oracle)
    srcs = _merger_graph(pkg, c)
    c.suspend_sync(length / sr / 2.0, lambda ctx: srcs[1].disconnect())
    out = c.start_rendering_sync()
    assert np.array_equal(out.get_channel_data(0), np.full(length, 2.0, np.float32))
    right = out.get_channel_data(1)
    assert np.array_equal(right[:length // 2], np.full(length // 2, 3.0, np.float32))
    assert np.array_equal(right[length // 2:], np.zeros(length // 2, np.float32))


def test_channel_merger_splitter_option_errors(pkg, oracle):  # channel_merger.rs:184-206, channel_splitter.rs:236-260
    import pytest
    c = pkg.OfflineAudioContext(1, RQ, 48000.0, oracle)
    m = c.create_channel_merger(6)
    assert m.number_of_inputs() == 6
    s = c.create_channel_splitter(2)
    assert s.number_of_outputs() == 2
    with pytest.raises(pkg.WaeError):
        c.create_channel_merger(33)  # (0 means "default" at the C ABI)
    with pytest.raises(pkg.WaeError):
        c.create_channel_splitter(33)


def test_channel_splitter(pkg, oracle):  # channel_splitter.rs:262-284 test_splitter
    c = pkg.OfflineAudioContext(1, RQ, 48000.0, oracle)
    splitter = c.create_channel_splitter(2)
    splitter.connect_from_output_to_input(c.destination(), 1, 0)  # 2nd output only
    src = c.create_buffer_source(pkg.AudioBuffer([np.array([1.0], np.float32), np.array([-1.0], np.float32)], 48000.0), loop=True)
    src.start()
    src.connect(splitter)
    assert np.array_equal(c.start_rendering_sync().get_channel_data(0), np.full(RQ, -1.0, np.float32))


def _stereo_pan(pkg, be, channels, pan, cfg):
    sr = 44100.0
    c = pkg.OfflineAudioContext(2, RQ, sr, be)
    panner = c.create_stereo_panner(pan, cfg=cfg)
    panner.connect(c.destination())
    src = c.create_buffer_source(pkg.AudioBuffer([np.ones(RQ, np.float32)] * channels, sr))
    src.connect(panner)
    src.start()
    out = c.start_rendering_sync()
    return out.get_channel_data(0), out.get_channel_data(1)


def test_stereo_panner_mono_panning(pkg, oracle):  # stereo_panner.rs:368-466
    mono = pkg.channel_config(1, pkg.CLAMPED_MAX)
    left, right = _stereo_pan(pkg, oracle, 1, -1.0, mono)
    _close(left, 1.0, 0.0)
    _close(right, 0.0, 0.0)
    left, right = _stereo_pan(pkg, oracle, 1, 1.0, mono)
    _close(left, 0.0, 1e-7)
    _close(right, 1.0, 0.0)
    left, right = _stereo_pan(pkg, oracle, 1, 0.0, mono)
    _close(left * left + right * right, 1.0, 1.2e-7)


def test_stereo_panner_stereo_panning(pkg, oracle):  # stereo_panner.rs:468-553
    for pan, wl, tl, wr in [(-1.0, 2.0, 0.0, 0.0), (1.0, 0.0, 1e-7, 2.0), (0.0, 1.0, 1e-7, 1.0)]:
        left, right = _stereo_pan(pkg, oracle, 2, pan, None)
        _close(left, wl, tl)
        _close(right, wr, 0.0)


def _shape(pkg, be, data, length):
    sr = 44100.0
    c = pkg.OfflineAudioContext(1, length, sr, be)
    shaper = c.create_wave_shaper(curve=np.array([-0.5, 0.0, 0.5], np.float32))
    shaper.connect(c.destination())
    buf = np.zeros(3 * RQ, np.float32)
    buf[:len(data)] = data
    src = c.create_buffer_source(pkg.AudioBuffer([buf], sr))
    src.connect(shaper)
    src.start_at(0.0)
    return c.start_rendering_sync().get_channel_data(0)


def test_wave_shaper_boundaries(pkg, oracle):  # waveshaper.rs:671-707 test_shape_boundaries
    data = np.concatenate([np.full(RQ, -1.0), np.zeros(RQ), np.full(RQ, 1.0)]).astype(np.float32)
    assert np.array_equal(_shape(pkg, oracle, data, 3 * RQ), data * np.float32(0.5))


def test_wave_shaper_interpolation(pkg, oracle):  # waveshaper.rs:709-739 test_shape_interpolation
    data = (np.arange(RQ, dtype=np.float32) / np.float32(RQ) * np.float32(2.0) - np.float32(1.0)).astype(np.float32)
    assert np.array_equal(_shape(pkg, oracle, data, RQ), data / np.float32(2.0))


def test_up_down_mix_rules_through_a_graph(pkg, oracle):
    # src/render/quantum.rs:803-1440 test_audiobuffer_upmix_speakers / _downmix_speakers, :745-785 _mix_discrete: the same tables,
    # applied by the input mixer of an explicit-count node (src/render/graph.rs:507-521) instead of AudioRenderQuantum::mix directly
    s = np.float32(0.5) ** np.float32(0.5)
    f = np.float32
    table = [
        ([1], 2, 0, [1, 1]), ([1], 4, 0, [1, 1, 0, 0]), ([1], 6, 0, [0, 0, 1, 0, 0, 0]), ([1, 2], 4, 0, [1, 2, 0, 0]),
        ([1, 2], 6, 0, [1, 2, 0, 0, 0, 0]), ([1, 2, 3, 4], 6, 0, [1, 2, 0, 0, 3, 4]), ([1, 2], 1, 0, [1.5]), ([1, 2, 3, 4], 1, 0, [2.5]),
        ([1, 2, 3, 4, 5, 6], 1, 0, [s * (f(1) + f(2)) + f(3) + f(0.5) * (f(5) + f(6))]), ([1, 2, 3, 4], 2, 0, [2.0, 3.0]),
        ([1, 2, 3, 4, 5, 6], 2, 0, [f(1) + s * (f(3) + f(5)), f(2) + s * (f(3) + f(6))]),
        ([1, 2, 3, 4, 5, 6], 4, 0, [f(1) + s * f(3), f(2) + s * f(3), 5, 6]),
        ([1, 2, 3], 5, 1, [1, 2, 3, 0, 0]), ([1, 2, 3], 2, 1, [1, 2]), ([1, 2, 3], 4, 0, [1, 2, 3, 0]),  # 3 -> 4 has no speaker rule: discrete
    ]
    sr = 48000.0
    for vals, to, interp, want in table:
        c = pkg.OfflineAudioContext(to, RQ, sr, oracle)
        src = c.create_buffer_source(pkg.AudioBuffer([np.full(RQ, v, np.float32) for v in vals], sr))
        g = c.create_gain(1.0, cfg=pkg.channel_config(to, pkg.EXPLICIT, pkg.DISCRETE if interp else pkg.SPEAKERS))
        src.connect(g)
        g.connect(c.destination())
        src.start()
        out = c.start_rendering_sync()
        got = np.array([out.get_channel_data(i) for i in range(to)])
        assert np.all(got == got[:, :1]), (vals, to)
        assert np.abs(got[:, 0].astype(np.float64) - np.asarray(want, np.float64)).max() <= 2.4e-7 * max(1.0, max(vals)), (vals, to, got[:, 0], want)


def _mixing_case(pkg, be, dest_channels, dest_interp, count, mode, interp):
    # tests/mixing.rs:11-44 setup_with_destination_channel_config + run_with_intermediate_channel_config
    c = pkg.OfflineAudioContext(dest_channels, RQ, 44100.0, be)
    c.destination().set_channel_interpretation(dest_interp)
    constant = c.create_constant_source()
    constant.start()
    gain = c.create_gain()  # only added for mixing
    gain.set_channel_count(count)
    gain.set_channel_count_mode(mode)
    gain.set_channel_interpretation(interp)
    constant.connect(gain)
    gain.connect(c.destination())
    out = c.start_rendering_sync()
    assert out.number_of_channels() == dest_channels
    return [out.get_channel_data(i) for i in range(dest_channels)]


def test_mixing_integration_cases(pkg, oracle):  # tests/mixing.rs:48-106 (all six)
    ones, zeroes = np.ones(RQ, np.float32), np.zeros(RQ, np.float32)
    for dest, dest_interp, count, want in [
        (1, pkg.SPEAKERS, 1, [ones]),                       # test_mono_speakers
        (2, pkg.SPEAKERS, 2, [ones, ones]),                 # test_stereo_speakers
        (4, pkg.SPEAKERS, 4, [ones, ones, zeroes, zeroes]),  # test_quad_speakers
        (2, pkg.DISCRETE, 1, [ones, zeroes]),               # test_mono_to_discrete_stereo
        (2, pkg.DISCRETE, 2, [ones, zeroes]),               # test_stereo_to_discrete_stereo
        (1, pkg.DISCRETE, 2, [ones]),                       # test_stereo_to_discrete_mono
    ]:
        got = _mixing_case(pkg, oracle, dest, dest_interp, count, pkg.MAX, pkg.SPEAKERS)
        for g, w in zip(got, want):
            assert np.array_equal(g, w), (dest, dest_interp, count)


def test_channel_config_setters_take_effect_at_a_suspend_point(pkg, oracle):
    # the control message reaches the graph between two quanta (src/render/thread.rs:229-248): stereo source into a gain whose
    # count mode goes Max -> Explicit(1) half-way: [1, -0.5] before, the down-mix 0.25 on both destination channels after
    sr = 48000.0
    c = pkg.OfflineAudioContext(2, 4 * RQ, sr, oracle)
    src = c.create_buffer_source(pkg.AudioBuffer([np.full(4 * RQ, 1.0, np.float32), np.full(4 * RQ, -0.5, np.float32)], sr))
    g = c.create_gain()
    src.connect(g)
    g.connect(c.destination())
    src.start()

    def narrow(_ctx):
        g.set_channel_count(1)
        g.set_channel_count_mode(pkg.EXPLICIT)

    c.suspend_sync(2 * RQ / sr, narrow)
    out = c.start_rendering_sync()
    left, right = out.get_channel_data(0), out.get_channel_data(1)
    assert np.all(left[:2 * RQ] == 1.0) and np.all(right[:2 * RQ] == -0.5)
    assert np.all(left[2 * RQ:] == 0.25) and np.all(right[2 * RQ:] == 0.25)


def test_channel_config_constraints(pkg, oracle):
    # channel_merger.rs:195-204, channel_splitter.rs:250-259 (#[should_panic]) and the other per-node overrides of the three setters
    import pytest
    c = pkg.OfflineAudioContext(2, RQ, 48000.0, oracle)
    ok = [
        (c.create_gain(), "set_channel_count", 32), (c.create_gain(), "set_channel_count_mode", pkg.EXPLICIT),
        (c.create_channel_merger(2), "set_channel_count", 1), (c.create_channel_splitter(3), "set_channel_count", 3),
        (c.create_panner(), "set_channel_count", 1), (c.create_stereo_panner(), "set_channel_count_mode", pkg.EXPLICIT),
        (c.destination(), "set_channel_count", 2), (c.destination(), "set_channel_interpretation", pkg.DISCRETE),
        (c.create_convolver(), "set_channel_count", 1), (c.create_dynamics_compressor(), "set_channel_count_mode", pkg.EXPLICIT),
    ]
    for node, setter, v in ok:
        getattr(node, setter)(v)
    bad = [
        (c.create_gain(), "set_channel_count", 0), (c.create_gain(), "set_channel_count", 33),
        (c.create_channel_merger(2), "set_channel_count", 3), (c.create_channel_merger(2), "set_channel_count_mode", pkg.MAX),
        (c.create_channel_splitter(2), "set_channel_count", 3), (c.create_channel_splitter(2), "set_channel_count_mode", pkg.MAX),
        (c.create_channel_splitter(2), "set_channel_interpretation", pkg.SPEAKERS),
        (c.create_panner(), "set_channel_count", 3), (c.create_panner(), "set_channel_count_mode", pkg.MAX),
        (c.create_stereo_panner(), "set_channel_count", 3), (c.create_stereo_panner(), "set_channel_count_mode", pkg.MAX),
        (c.create_convolver(), "set_channel_count", 3), (c.create_convolver(), "set_channel_count_mode", pkg.MAX),
        (c.create_dynamics_compressor(), "set_channel_count", 3), (c.create_dynamics_compressor(), "set_channel_count_mode", pkg.MAX),
        (c.destination(), "set_channel_count", 1), (c.destination(), "set_channel_count_mode", pkg.MAX),
    ]
    for node, setter, v in bad:
        with pytest.raises(pkg.WaeError):
            getattr(node, setter)(v)
