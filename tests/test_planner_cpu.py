"""The planner — the largest piece of host logic of the library (csrc/wae_engine.cu: fusion, stage grouping, scheduling classes of
feedback loops, render segments of suspend points, chunk sizing, what is refused) — exercised WITHOUT a GPU through wae_batch_plan,
which runs the sizing pass wae_batch_prepare runs first (it touches no device memory) and reports what the batch would be lowered to."""
import os

import numpy as np
import pytest

import benchmark_scenarios as BS
import graphs as G

RQ = 128


@pytest.fixture
def be(pkg):
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "web-audio-api-rs_b200", "libwae_b200.so")
    if not os.path.exists(so):
        pytest.skip("libwae_b200.so is not built (python -c 'import __graft_entry__ as g; g.build()')")
    return pkg.context.Backend(pkg.api(), None)


def plan(pkg, ctxs):
    return pkg.context.plan_batch(ctxs)


def test_c2_is_one_fused_launch_for_the_whole_render(pkg, be):
    # BASELINE configs[1]: AudioBufferSource -> Biquad -> Gain -> destination: one k_chain stage, no arena, one chunk = the whole render
    p = plan(pkg, [G.c2_buffer_biquad_gain(pkg, be, g, 48000) for g in range(8)])
    assert p["kinds"] == {"k_chain": 1} and p["arena_floats_per_frame"] == 0 and p["chunks"] == 1 and p["chunk_frames"] == 48000
    assert p["source_floats"] == 8 * 2 * 48000 and not p["has_feedback"]
    # >= 512 graphs: 32 pipeline groups, each its own stage
    p = plan(pkg, [G.c2_buffer_biquad_gain(pkg, be, g, 1280) for g in range(512)])
    assert p["groups"] == 32 and p["kinds"] == {"k_chain": 32}


def test_c1_c3_voices_fuse_and_meet_in_one_mix(pkg, be):
    p = plan(pkg, [G.c1_osc_biquad(pkg, be, 48000)])
    assert p["kinds"] == {"k_chain": 1}  # the single chain writes the destination directly (mono -> stereo by copy)
    p = plan(pkg, [G.c3_many_voices(pkg, be, 64, 48000)])
    assert p["kinds"] == {"k_chain": 1, "k_mix": 1}  # 64 oscillator -> biquad chains in ONE launch, one ordered sum
    assert p["arena_floats_per_frame"] == 64           # 64 mono voices


def test_c4_north_star_c5_stage_lists(pkg, be):
    ir = G.synthetic_ir(20000, 2, decay=0.6)
    p = plan(pkg, [G.c4_convolver(pkg, be, g, 8192 * 3, ir) for g in range(4)])
    assert set(p["kinds"]) == {"k_chain", "k_conv_fft_in", "k_conv_mac_ifft", "k_mix"} and p["chunk_frames"] % 8192 == 0
    p = plan(pkg, [G.north_star_voices_convolver(pkg, be, 50, 8192 * 3, ir, seed=g) for g in range(2)])
    assert p["kinds"]["k_chain"] == 1 and "k_conv_mac_ifft" in p["kinds"]
    with pytest.raises(pkg.WaeError) as e:  # an HRTF panner needs the sphere the engine is given (wae_engine_set_hrir_sphere)
        plan(pkg, [G.c5_full_chain(pkg, be, 0, 8192, ir)])
    assert e.value.status == 4 and "HRIR sphere" in str(e.value)


@pytest.mark.parametrize("name,build", BS.SCENARIOS, ids=[n for n, _ in BS.SCENARIOS])
def test_every_reference_benchmark_scenario_is_lowered(pkg, be, name, build):
    # examples/benchmarks.rs: none of the 24 scenarios may be refused (WAE_UNSUPPORTED) by the planner
    p = plan(pkg, [build(pkg, be, 3.0) for _ in range(2)])
    assert p["stages"] >= 1 and p["segments"] == p["groups"]
    if name.startswith("Simple source test without resampling (Mono)"):
        assert p["kinds"] == {"k_chain": 1}


def test_feedback_loop_is_scheduled(pkg, be):
    def echo(with_conv):
        c = pkg.OfflineAudioContext(2, RQ * 16, 48000.0, be)
        src = c.create_constant_source()
        g = c.create_gain(0.5)
        d = c.create_delay(1.0, 0.01)
        src.connect(g)
        g.connect(d)
        if with_conv:
            conv = c.create_convolver(pkg.AudioBuffer([np.ones(8, np.float32)], 48000.0))
            d.connect(conv)
            conv.connect(g)
        else:
            d.connect(g)
        g.connect(c.destination())
        src.start()
        return c

    p = plan(pkg, [echo(False)])
    assert p["has_feedback"] and "k_delay_read" in p["kinds"] and "k_ring_write" in p["kinds"]
    # a ConvolverNode inside the loop is refused by the full planning pass only (it needs the final scheduling classes): the sizing
    # pass accepts it; tests/test_gpu_parity.py::test_unsupported_is_reported_not_faked checks the refusal on the GPU
    assert plan(pkg, [echo(True)])["has_feedback"]


def test_suspend_points_cut_the_render_into_segments(pkg, be):
    def build(cut_frames, with_conv):
        sr = 48000.0
        c = pkg.OfflineAudioContext(1, 8192 * 3, sr, be)
        src = c.create_constant_source()
        node = c.create_convolver(pkg.AudioBuffer([np.ones(4, np.float32)], sr)) if with_conv else c.create_gain(0.5)
        src.connect(node)
        node.connect(c.destination())
        src.start()
        c.suspend_sync(cut_frames / sr, lambda ctx: src.offset.set_value(0.25))
        return c

    p = plan(pkg, [build(RQ * 5, False)])
    assert p["segments"] == 2
    p = plan(pkg, [build(8192, True)])   # a convolver needs the cut on a partition boundary
    assert p["segments"] == 2
    with pytest.raises(pkg.WaeError) as e:
        plan(pkg, [build(RQ * 5, True)])
    assert e.value.status == 4
    # graphs with different suspend points are planned in different groups
    p = plan(pkg, [build(RQ * 5, False), build(RQ * 9, False)])
    assert p["groups"] == 2 and p["segments"] == 4


def test_batch_shape_errors(pkg, be):
    a = pkg.OfflineAudioContext(2, 256, 48000.0, be)
    b = pkg.OfflineAudioContext(1, 256, 48000.0, be)
    with pytest.raises(pkg.WaeError):
        plan(pkg, [a, b])  # all graphs of a batch share channels / length / sample rate
