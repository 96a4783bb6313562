#!/usr/bin/env python
"""Times wae_render_batch(HOST) — the one-shot plugin call — on the C2 workload: fresh graphs every call, prepare inside the call.
usage: python tools/oneshot_time.py [n_graphs] [seconds] [calls]   (WAE_PREPARE_PROFILE=1 prints nothing here: there is no prepare)"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import __graft_entry__ as ge
    import graphs as G
    pkg = ge.build()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
    calls = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    length = int(secs * G.SR)
    eng = pkg.Engine(0)
    if os.environ.get("WAE_NUMA"):
        eng.set_option(pkg.OPT_BIND_NUMA, 1)
    pcm = [G.c2_source(g, length) for g in range(n)]
    out_pageable = np.zeros((n, 2, length), np.float32)
    pinned = torch.empty(n * 2 * length, dtype=torch.float32, pin_memory=True)
    out_pinned = pinned.numpy().reshape(n, 2, length)
    for label, out in (("pageable out", out_pageable), ("pinned out", out_pinned)):
        for i in range(calls):
            t0 = time.perf_counter()
            ctxs = [G.c2_buffer_biquad_gain(pkg, eng.backend, g, length, pcm=pcm[g]) for g in range(n)]
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            pkg.render_batch_oneshot(ctxs, out)
            t3 = time.perf_counter()
            print(f"{label}: call {i}: build graphs {1e3 * (t1 - t0):8.1f} ms   wae_render_batch(HOST) {1e3 * (t3 - t2):8.2f} ms   "
                  f"({n * ((length + 127) // 128) / (t3 - t2) / 1e6:.1f} M graph-quanta/s)", flush=True)
            del ctxs
    # parity spot check against prepare / run / fetch
    ctxs = [G.c2_buffer_biquad_gain(pkg, eng.backend, g, length, pcm=pcm[g]) for g in range(min(n, 40))]
    b = pkg.Batch(ctxs)
    b.run()
    b.sync()
    ref = b.fetch()
    print("bit-equal to prepare/run/fetch on the first graphs:", bool(np.array_equal(ref, out_pinned[: len(ctxs)]) and np.array_equal(ref, out_pageable[: len(ctxs)])))
    eng.close()


if __name__ == "__main__":
    main()
