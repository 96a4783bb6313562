"""Reference tests that tools/reference_test_map.py found without a restatement (tests/REFERENCE_TESTS.md), restated here: argument checks
of the AudioParam methods, the oscillator's type rules, AudioBuffer validation at the boundary, constructor values that must be audible from
the first frame, the unit tests of AudioRenderQuantum::add / mix replayed through a graph.  Graph-building checks run against BOTH libraries
(the oracle and the product's own host code, `builder`), renders against the oracle (the GPU suite reruns graph renders through
tests/test_gpu_reference_cases.py)."""
import ctypes as C

import numpy as np
import pytest

RQ = 128
SR = 48000.0


def _const(c, value, start=True):
    s = c.create_constant_source(value)
    if start:
        s.start()
    return s


def render(c):
    a = c.start_rendering_sync()
    return np.array([a.get_channel_data(i) for i in range(a.number_of_channels())])


# ---- src/param.rs:1667-1695 test_assert_strictly_positive(_fail) / test_assert_not_zero(_fail) / test_assert_sequence_length(_fail):
# the helpers guard exponential_ramp_to_value_at_time (value != 0, param.rs:490) and set_value_curve_at_time (>= 2 values, duration > 0: :618-620)
def test_audioparam_argument_asserts(pkg, builder):
    c = pkg.OfflineAudioContext(1, RQ, SR, builder)
    g = c.create_gain()
    with pytest.raises(pkg.WaeError, match="should not be equal to zero"):
        g.gain.exponential_ramp_to_value_at_time(0.0, 0.1)
    g.gain.exponential_ramp_to_value_at_time(-0.1, 0.1)  # assert_not_zero(-0.1) / (0.1) pass
    g.gain.exponential_ramp_to_value_at_time(0.1, 0.2)
    g2 = c.create_gain()
    with pytest.raises(pkg.WaeError, match="sequence length"):
        g2.gain.set_value_curve_at_time([0.0], 0.0, 0.1)
    with pytest.raises(pkg.WaeError, match="strictly positive"):
        g2.gain.set_value_curve_at_time([0.0, 1.0], 0.0, 0.0)
    with pytest.raises(pkg.WaeError):
        g2.gain.set_value_curve_at_time([0.0, 1.0], 0.0, float("inf"))  # "The provided value is non-finite"
    g2.gain.set_value_curve_at_time([0.0, 0.0], 0.0, 0.1)  # two values, duration 0.1: accepted


# ---- src/node/oscillator.rs:740-787: set_type(Custom) panics; with a periodic wave the type is Custom and set_type is ignored; it renders
def test_oscillator_type_rules(pkg, builder, oracle):
    c = pkg.OfflineAudioContext(2, 1, 44100.0, builder)
    osc = c.create_oscillator()
    with pytest.raises(pkg.WaeError, match="Custom"):
        osc.set_type(pkg.CUSTOM if hasattr(pkg, "CUSTOM") else 4)
    # PeriodicWaveOptions::default(): real = [0, 0], imag = [0, 1] -> one sine cycle; set_type(Sine) afterwards changes nothing
    table = np.sin(2 * np.pi * np.arange(2048) / 2048).astype(np.float32)
    outs = []
    for reset_type in (False, True):
        c = pkg.OfflineAudioContext(2, RQ, 44100.0, oracle)
        osc = c.create_oscillator(frequency=440.0, periodic_wave=table)
        if reset_type:
            osc.set_type(pkg.SAWTOOTH)  # ignored: the node keeps playing its wavetable
        osc.connect(c.destination())
        osc.start()
        outs.append(render(c))
    assert np.array_equal(outs[0], outs[1]) and np.abs(outs[0]).max() > 0.1
    # the product's graph half applies the same rule (plan unchanged by the ignored set_type is checked in test_node_setters.py)
    c = pkg.OfflineAudioContext(2, RQ, 44100.0, builder)
    osc = c.create_oscillator(frequency=440.0, periodic_wave=table)
    osc.set_type(pkg.SAWTOOTH)  # no error


# ---- src/buffer.rs:440-491 test_zero_channels(_from) / test_invalid_sample_rate(_from) / test_invalid_length: an AudioBuffer that the
# reference refuses to construct is refused where it enters the library (copy_buffer: channel count 1..=32, length > 0)
def test_invalid_audio_buffers_are_refused_at_the_boundary(pkg, builder):
    import importlib
    B = importlib.import_module(pkg.__name__ + "._binding")
    c = pkg.OfflineAudioContext(1, RQ, SR, builder)
    api = c._api
    one = np.zeros(8, np.float32)

    def desc(n_ch, length):
        ptrs = (C.POINTER(C.c_float) * max(n_ch, 1))(*[B.fptr(one) for _ in range(max(n_ch, 1))])
        return B.AudioBufferDesc(n_ch, length, SR, ptrs), ptrs
    for n_ch, length in [(0, 8), (33, 8), (1, 0)]:
        d, keep = desc(n_ch, length)
        o = B.BufferSourceOptions(C.pointer(d), 0.0, 1.0, 0, 0.0, 0.0)
        nid = C.c_uint32()
        st = api.create_buffer_source(c._g, C.byref(o), C.byref(nid))
        assert st != 0, (n_ch, length)
        s = c.create_buffer_source()
        with pytest.raises(pkg.WaeError):
            api.check(api.buffer_source_set_buffer(c._g, s.id, C.byref(d)))
    d, keep = desc(1, 8)   # ... and a valid one passes
    o = B.BufferSourceOptions(C.pointer(d), 0.0, 1.0, 0, 0.0, 0.0)
    nid = C.c_uint32()
    api.check(api.create_buffer_source(c._g, C.byref(o), C.byref(nid)))


# ---- src/context/offline.rs:439-460 test_sample_rate_length, render_empty_graph (render_twice_panics: the context's state machine stays on
# the Rust side of the boundary)
def test_offline_context_accessors_and_empty_graph(pkg, builder, oracle):
    c = pkg.OfflineAudioContext(1, 48000, 96000.0, builder)
    assert c.sample_rate() == 96000.0 and c.length() == 48000
    c = pkg.OfflineAudioContext(2, 555, 44100.0, oracle)
    out = render(c)
    assert out.shape == (2, 555) and not out.any() and c.length() == 555


# ---- gain.rs:209-217, delay.rs:756-764, constant_source.rs:300-306, panner.rs:1070-1079 test_audioparam_value_applies_immediately: the value
# given in the options is the param's value — restated the audible way too: it holds from the first rendered frame
def test_option_values_apply_from_the_first_frame(pkg, builder, oracle):
    for be in (builder, oracle):
        c = pkg.OfflineAudioContext(1, RQ, SR, be)
        assert c.create_gain(0.12).gain.value() == np.float32(0.12)
        assert c.create_delay(1.0, 0.12).delay_time.value() == np.float32(0.12)
        assert c.create_constant_source(12.0).offset.value() == 12.0
        p = c.create_panner(position=(1.0, 2.0, 3.0))
        assert (p.position_x.value(), p.position_y.value(), p.position_z.value()) == (1.0, 2.0, 3.0)
    c = pkg.OfflineAudioContext(1, RQ * 2, SR, oracle)
    g = c.create_gain(0.12)
    _const(c, 12.0).connect(g)
    g.connect(c.destination())
    out = render(c)[0]
    assert np.all(out == np.float32(12.0) * np.float32(0.12))
    # DelayOptions::delay_time: an impulse comes out exactly delay_time later, first quantum included
    n_delay = 37
    c = pkg.OfflineAudioContext(1, RQ * 2, SR, oracle)
    d = c.create_delay(1.0, n_delay / SR)
    src = c.create_buffer_source(pkg.AudioBuffer([np.array([1.0], np.float32)], SR))
    src.connect(d)
    d.connect(c.destination())
    src.start()
    out = render(c)[0]
    expected = np.zeros(RQ * 2, np.float32)
    expected[n_delay] = 1.0
    assert np.abs(out - expected).max() <= 1e-5


# ---- src/render/quantum.rs:703-726 test_channel_add, :1442-1474 test_audiobuffer_add, :1476-1520 test_is_silent_quantum /
# test_is_not_silent_quantum — AudioRenderQuantum::add / mix, replayed through input ports
def test_quantum_add_and_mix_through_a_graph(pkg, oracle):
    # signal + silence = signal, silence + signal = signal (an unstarted source is the silent quantum), 1 + 2 = 3
    for order in (0, 1):
        c = pkg.OfflineAudioContext(1, RQ, SR, oracle)
        nodes = [_const(c, 1.0), _const(c, 5.0, start=False)]
        for s in (nodes if order == 0 else nodes[::-1]):
            s.connect(c.destination())
        assert np.all(render(c)[0] == 1.0)
    c = pkg.OfflineAudioContext(1, RQ, SR, oracle)
    _const(c, 1.0).connect(c.destination())
    _const(c, 2.0).connect(c.destination())
    assert np.all(render(c)[0] == 3.0)
    # test_audiobuffer_add: [1, 1] (mono up-mixed as speakers) + mono 2 under (count 2, explicit, discrete) = [3, 1]
    c = pkg.OfflineAudioContext(2, RQ, SR, oracle)
    port = c.create_gain(1.0, cfg=pkg.channel_config(2, pkg.EXPLICIT, pkg.DISCRETE))
    stereo = c.create_buffer_source(pkg.AudioBuffer([np.ones(RQ, np.float32), np.ones(RQ, np.float32)], SR))
    stereo.connect(port)
    stereo.start()
    _const(c, 2.0).connect(port)
    port.connect(c.destination())
    out = render(c)
    assert np.all(out[0] == 3.0) and np.all(out[1] == 1.0)
    # test_is_not_silent_quantum: mono 1 mixed to 2 channels DISCRETE = [1, 0]; test_is_silent_quantum: silence stays silence under speakers
    c = pkg.OfflineAudioContext(2, RQ, SR, oracle)
    port = c.create_gain(1.0, cfg=pkg.channel_config(2, pkg.EXPLICIT, pkg.DISCRETE))
    _const(c, 1.0).connect(port)
    port.connect(c.destination())
    out = render(c)
    assert np.all(out[0] == 1.0) and not out[1].any()
    c = pkg.OfflineAudioContext(2, RQ, SR, oracle)
    port = c.create_gain(1.0, cfg=pkg.channel_config(2, pkg.EXPLICIT, pkg.SPEAKERS))
    _const(c, 1.0, start=False).connect(port)
    port.connect(c.destination())
    assert not render(c).any()


# ---- src/node/analyser.rs:335-344 test_construct_decibels: min -10 / max 20 is a valid pair (only min < max is required)
def test_analyser_decibel_range_above_zero(pkg, builder):
    c = pkg.OfflineAudioContext(1, RQ, 44100.0, builder)
    c.create_analyser(min_decibels=-10.0, max_decibels=20.0)


# ---- constructor tests: dynamics_compressor.rs:491-523 test_constructor_default / _non_default, oscillator.rs:691-735 assert_osc_default_build
# (_with_factory_func), waveshaper.rs:587-625 build_with_new / build_with_factory_func / test_default_options / test_user_defined_options,
# iir_filter.rs:430-449 test_constructor_and_factory, channel_splitter.rs:223-235 test_valid_constructor_options — the defaults live in the
# host mirror of the API (the option structs of include/wae.h carry every value explicitly); what they build must render
def test_constructors_and_their_defaults(pkg, builder, oracle):
    c = pkg.OfflineAudioContext(1, RQ, 44100.0, builder)
    comp = c.create_dynamics_compressor()
    got = [comp.attack.value(), comp.knee.value(), comp.ratio.value(), comp.release.value(), comp.threshold.value()]
    assert got == [float(np.float32(0.003)), 30.0, 12.0, 0.25, -24.0]
    comp = c.create_dynamics_compressor(attack=0.5, knee=12.0, ratio=1.0, release=0.75, threshold=-60.0)
    assert [comp.attack.value(), comp.knee.value(), comp.ratio.value(), comp.release.value(), comp.threshold.value()] == [0.5, 12.0, 1.0, 0.75, -60.0]
    osc = c.create_oscillator()
    assert osc.frequency.value() == 440.0 and osc.detune.value() == 0.0
    c.create_wave_shaper()                                  # curve None, oversample None
    c.create_iir_filter([1.0, 1.0, 1.0], [1.0, 1.0, 1.0])
    sp = c.create_channel_splitter(2)
    assert sp.number_of_outputs() == 2
    # "should not panic when run": default oscillator, one frame, stereo; a one-point curve with 2x over-sampling
    c = pkg.OfflineAudioContext(2, 1, 44100.0, oracle)
    osc = c.create_oscillator()
    osc.connect(c.destination())
    osc.start()
    assert render(c).shape == (2, 1)
    c = pkg.OfflineAudioContext(2, RQ, 44100.0, oracle)
    c.create_wave_shaper(curve=[1.0], oversample=pkg.OVERSAMPLE_2X if hasattr(pkg, "OVERSAMPLE_2X") else 1)
    assert not render(c).any()
    # the default sine IS type Sine at 440 Hz: same PCM as the options spelled out
    outs = []
    for kw in ({}, {"type_": pkg.SINE, "frequency": 440.0, "detune": 0.0}):
        c = pkg.OfflineAudioContext(1, RQ * 4, 44100.0, oracle)
        o = c.create_oscillator(**kw)
        o.connect(c.destination())
        o.start()
        outs.append(render(c))
    assert np.array_equal(outs[0], outs[1]) and np.abs(outs[0]).max() > 0.9


# ---- convolver.rs:551-584 test_constructor_options_buffer: a response given in the OPTIONS (not through set_buffer) convolves — identity
# response, normalised: the input scaled by the calibration 0.00125
def test_convolver_response_from_the_options(pkg, oracle):
    sr = 44100.0
    c = pkg.OfflineAudioContext(1, 10, sr, oracle)
    conv = c.create_convolver(buffer=pkg.AudioBuffer([np.array([1.0], np.float32)], sr))
    conv.connect(c.destination())
    src = c.create_buffer_source()
    src.connect(conv)
    src.set_buffer(pkg.AudioBuffer([np.array([0.0, 1.0, 0.0, -1.0, 0.0], np.float32)], sr))
    src.start()
    out = render(c)[0]
    cal = 0.00125
    assert np.abs(out - np.array([0.0, cal, 0.0, -cal, 0, 0, 0, 0, 0, 0], np.float32)).max() <= 1e-7
