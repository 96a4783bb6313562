#!/usr/bin/env python
"""Times the one-shot plugin call wae_render_batch(HOST) on the workloads whose host side dominates it: north_star (8 graphs x 1000
voices -> convolver: 24 k nodes around an 11 ms render) and the reference's `Granular synthesis` scenario (1500 grains of ONE AudioBuffer
per graph).  usage: python tools/oneshot_workloads.py [calls]   (WAE_PLAN_SPLIT=0: planning on one thread, for comparison)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import __graft_entry__ as ge
    import benchmark_scenarios as BS
    import graphs as G
    pkg = ge.build()
    calls = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    eng = pkg.Engine(0)
    ir = G.synthetic_ir(178899, 2, decay=0.6)
    cases = [("north_star 8 x 1000 voices, 10 s", 2, 480000, lambda g: G.north_star_voices_convolver(pkg, eng.backend, 1000, 480000, ir, seed=g), 8),
             ("Granular synthesis 8 graphs, 7.5 s", 1, int(7.5 * 48000), lambda g: BS.granular_synthesis(pkg, eng.backend, 120.0), 8)]
    for name, ch, length, build, n in cases:
        host = np.zeros((n, ch, length), np.float32)
        for i in range(calls):
            t0 = time.perf_counter()
            ctxs = [build(g) for g in range(n)]
            t1 = time.perf_counter()
            pkg.render_batch_oneshot(ctxs, host)
            t2 = time.perf_counter()
            print(f"{name}: call {i}: build graphs {1e3 * (t1 - t0):8.1f} ms   wae_render_batch(HOST) {1e3 * (t2 - t1):8.2f} ms   |x|max {float(np.abs(host).max()):.3f}", flush=True)
            del ctxs
    eng.close()


if __name__ == "__main__":
    main()
