"""The graph-building half of the product (csrc/wae_abi_graph.cpp: option validation, id allocation, constraint checks, processing
order) is host code and runs without a GPU: wae_graph_create accepts a NULL engine.  The argument-error tests restated from the
reference are therefore executed here against libwae_b200.so itself, on the CPU — not only against the oracle."""
import os

import numpy as np
import pytest

import test_oracle_analyser as AN
import test_oracle_kat as K
import test_oracle_nodes as N

RQ = 128


@pytest.fixture
def product(pkg):
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "web-audio-api-rs_b200", "libwae_b200.so")
    if not os.path.exists(so):
        pytest.skip("libwae_b200.so is not built (python -c 'import __graft_entry__ as g; g.build()')")
    return pkg.context.Backend(pkg.api(), None)


@pytest.mark.parametrize("case", [N.test_channel_config_constraints, N.test_channel_merger_splitter_option_errors, K.test_iir_coefficient_validation,
                                  K.test_convolver_argument_errors, AN.test_option_constraints], ids=lambda f: f.__name__)
def test_reference_argument_errors_on_the_product(pkg, product, case):
    case(pkg, product)


@pytest.mark.parametrize("which", ["product", "oracle"])
def test_node_and_param_ids_follow_the_reference(pkg, product, oracle, which):
    # src/context/concrete_base.rs:240 + SURVEY Appendix A.1: ids 0..=10 reserved, a node's id is taken before its params'
    c = pkg.OfflineAudioContext(2, RQ, 48000.0, product if which == "product" else oracle)
    osc = c.create_oscillator()       # 11, frequency 12, detune 13
    bq = c.create_biquad_filter()     # 14, q 15, detune 16, frequency 17, gain 18
    d = c.create_delay(1.0, 0.1)      # writer 19, reader 20, delayTime 21
    g = c.create_gain()               # 22, gain 23
    assert (osc.id, bq.id, d.id, g.id) == (11, 14, 19, 22)
    osc.connect(bq)
    bq.connect(d)
    d.connect(g)
    g.connect(c.destination())
    # depth-first from ascending ids (graph.rs:443-479): 0 | 11 -> 14 -> 19 -> 20 -> 22 | 12 | 13 | 15..18 | 21 | 23, post-order reversed
    assert c.render_order() == [23, 21, 18, 17, 16, 15, 13, 12, 11, 14, 19, 20, 22, 0]


def test_rendering_without_an_engine_fails_loudly(pkg, product):
    c = pkg.OfflineAudioContext(1, RQ, 48000.0, product)
    src = c.create_constant_source()
    src.connect(c.destination())
    src.start()
    with pytest.raises(pkg.WaeError):
        c.start_rendering_sync()


def test_context_option_errors(pkg, product):
    # src/context/offline.rs:78-105 / src/lib.rs:185-260: channel count, length and sample-rate ranges
    for ch, length, sr in [(0, RQ, 48000.0), (33, RQ, 48000.0), (1, 0, 48000.0), (1, RQ, 2999.0), (1, RQ, 768001.0)]:
        with pytest.raises(pkg.WaeError):
            pkg.OfflineAudioContext(ch, length, sr, product)
    pkg.OfflineAudioContext(32, 1, 3000.0, product)


def test_param_event_argument_errors(pkg, product):
    # src/param.rs:560-657 assert_* : negative times, zero exponential target, non-positive curve duration, short curves
    c = pkg.OfflineAudioContext(1, RQ, 48000.0, product)
    g = c.create_gain().gain
    g.set_value_at_time(1.0, 0.0)
    g.linear_ramp_to_value_at_time(2.0, 0.5)
    for bad in [lambda: g.set_value_at_time(1.0, -1.0), lambda: g.linear_ramp_to_value_at_time(1.0, -0.1),
                lambda: g.exponential_ramp_to_value_at_time(0.0, 1.0), lambda: g.exponential_ramp_to_value_at_time(1.0, -1.0),
                lambda: g.set_target_at_time(1.0, -1.0, 0.1), lambda: g.set_target_at_time(1.0, 1.0, -0.1),
                lambda: g.cancel_scheduled_values(-1.0), lambda: g.cancel_and_hold_at_time(-1.0),
                lambda: g.set_value_curve_at_time(np.array([1.0], np.float32), 1.0, 1.0),
                lambda: g.set_value_curve_at_time(np.array([1.0, 2.0], np.float32), -1.0, 1.0),
                lambda: g.set_value_curve_at_time(np.array([1.0, 2.0], np.float32), 1.0, 0.0)]:
        with pytest.raises(pkg.WaeError):
            bad()


@pytest.mark.parametrize("kind", ["constant", "buffer", "oscillator"])
def test_scheduled_source_state_machine(pkg, builder, kind):
    # src/node/scheduled_source.rs:267-338: start twice panics, stop before start panics, stop twice is allowed (issue #579)
    def make():
        c = pkg.OfflineAudioContext(2, 1, 44100.0, builder)
        return c, {"constant": c.create_constant_source, "buffer": c.create_buffer_source, "oscillator": c.create_oscillator}[kind]()

    c, src = make()
    src.start()
    with pytest.raises(pkg.WaeError):
        src.start()
    c, src = make()
    with pytest.raises(pkg.WaeError):
        src.stop()
    c, src = make()
    src.start()
    src.stop()
    src.stop()
    c, src = make()
    with pytest.raises(pkg.WaeError):
        src.start_at(-1.0)  # scheduled_source.rs:12-30 assert_valid_time_value / RangeError


def test_scheduling_clock_against_a_replay_of_the_reference(pkg):
    """Start / stop frames of every scheduled source come from csrc/wae_hostmath.h::SchedClock.  The reference never computes them: its
    renderers compare the start time with a clock that is `current_frame / sample_rate` at the head of each quantum (thread.rs:357-360)
    and grows by `+= dt` per frame inside it (oscillator.rs:511-557).  Replayed here literally in Python for 6000 random times."""
    import ctypes as C
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "web-audio-api-rs_b200", "libwae_b200.so")
    if not os.path.exists(so):
        pytest.skip("libwae_b200.so is not built")
    api = pkg.api()
    rng = np.random.default_rng(5)
    for i in range(6000):
        sr = float(np.float32(rng.choice([8000.0, 22050.0, 44100.0, 48000.0, 96000.0, 12345.0])))
        kind = i % 4
        if kind == 0:
            t = float(rng.uniform(0.0, 3.0))
        elif kind == 1:
            t = int(rng.integers(0, 120000)) / sr                     # exactly on a frame (as computed by a user: k / sr)
        elif kind == 2:
            t = int(rng.integers(0, 900)) * 128 / sr                  # exactly on a quantum boundary
        else:
            t = np.nextafter(int(rng.integers(1, 900)) * 128 / sr, rng.choice([0.0, 10.0]))  # one ulp off a boundary
        dt = 1.0 / sr
        # ---- literal replay
        q = max(0, int(t * sr / 128) - 2)
        want = None
        while want is None:
            now = (q * 128) / sr
            if t >= now + dt * 128:                                    # `start_time >= next_block_time`: the block is skipped
                q += 1
                continue
            for f in range(128):
                if now >= t:
                    want = (q * 128 + f, now)
                    break
                now += dt
            else:
                q += 1                                                  # not reached inside the block after all: the next block starts at/after it
                want = (q * 128, (q * 128) / sr)
        frame, ftime = C.c_int64(), C.c_double()
        api.check(api.sched_first_frame_at_or_after(sr, float(t), C.byref(frame), C.byref(ftime)))
        assert (frame.value, ftime.value) == want, (sr, t, frame.value, want)


def test_constructor_channel_configs_and_audio_buffers_are_validated(pkg, product):
    """assert_valid_number_of_channels / assert_valid_buffer_length at construction (src/lib.rs:185-228, src/buffer.rs:96-115): a channel
    count outside [1, 32], an unknown enum value, an AudioBuffer without frames or with more than 32 channels never reach the planner"""
    c = pkg.OfflineAudioContext(2, 1280, 48000.0, product)
    with pytest.raises(pkg.WaeError, match="Invalid number of channels"):
        c.create_gain(1.0, cfg=pkg.context.channel_config(33, pkg.MAX, pkg.SPEAKERS))
    with pytest.raises(pkg.WaeError, match="unknown channel count mode"):
        c.create_biquad_filter(cfg=pkg.context.channel_config(2, 7, pkg.SPEAKERS))
    with pytest.raises(pkg.WaeError, match="unknown channel interpretation"):
        c.create_delay(cfg=pkg.context.channel_config(2, pkg.MAX, 5))
    with pytest.raises(pkg.WaeError, match="Invalid length"):
        c.create_buffer_source(pkg.AudioBuffer([np.zeros(0, np.float32)], 48000.0))
    with pytest.raises(pkg.WaeError, match="Invalid number of channels"):
        c.create_buffer_source(pkg.AudioBuffer([np.zeros(8, np.float32)] * 33, 48000.0))
    c.create_gain(1.0, cfg=pkg.context.channel_config(32, pkg.EXPLICIT, pkg.DISCRETE))  # the limits themselves are fine
