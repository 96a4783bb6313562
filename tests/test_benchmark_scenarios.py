"""examples/benchmarks.rs, all 24 scenarios (tests/benchmark_scenarios.py): the oracle renders each one here (CPU: the graphs build,
render, and are not silent); tests/test_gpu_benchmark_scenarios.py compares the CUDA engine with the oracle on the same graphs."""
import numpy as np
import pytest

import benchmark_scenarios as BS

SECONDS = 3.0  # what the reference's DURATION = 120 s becomes here


@pytest.mark.parametrize("name,build", BS.SCENARIOS, ids=[n for n, _ in BS.SCENARIOS])
def test_scenario_renders_on_the_oracle(pkg, oracle, name, build):
    c = build(pkg, oracle, SECONDS)
    out = c.start_rendering_sync()
    pcm = np.array([out.get_channel_data(i) for i in range(out.number_of_channels())])
    assert np.isfinite(pcm).all()
    if name == "Baseline (silence)":
        assert not pcm.any()
    else:
        assert np.abs(pcm).max() > 1e-3, name


@pytest.mark.parametrize("name,build", BS.CRITERION, ids=[n for n, _ in BS.CRITERION])
def test_criterion_bench_renders_on_the_oracle(pkg, oracle, name, build):
    # benches/my_benchmark.rs: the reference's criterion / iai benchmark graphs (bench_audio_buffer_decode and the worklet one are
    # outside the path).  HRTF at 48 kHz with a 44.1 kHz sphere, as the reference's embedded one
    import graphs as G
    if "hrtf" in name:
        oracle.set_hrir_sphere(G.synthetic_hrir_sphere(44100, 256))
    c = build(pkg, oracle, 2.0)
    out = c.start_rendering_sync()
    pcm = np.array([out.get_channel_data(i) for i in range(out.number_of_channels())])
    assert pcm.shape[0] == 2 and np.isfinite(pcm).all()
    assert (np.abs(pcm).max() > 1e-3) == (name != "bench_ctor")
