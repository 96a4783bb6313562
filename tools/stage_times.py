#!/usr/bin/env python
"""Per-stage CUDA-event times of one scenario of tests/benchmark_scenarios.py (examples/benchmarks.rs) on the engine.
    python tools/stage_times.py --scenario Envelope --graphs 16 --seconds 120"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", required=True)
    ap.add_argument("--graphs", type=int, default=16)
    ap.add_argument("--seconds", type=float, default=120.0)
    args = ap.parse_args()
    import __graft_entry__ as ge
    pkg = ge.build()
    import benchmark_scenarios as BS
    eng = pkg.Engine(0)
    for name, build in BS.SCENARIOS:
        if args.scenario.lower() not in name.lower():
            continue
        batch = pkg.Batch([build(pkg, eng.backend, args.seconds) for _ in range(args.graphs)])
        batch.set_timing(True)
        batch.run()
        batch.sync()
        batch.run()
        batch.sync()
        st = batch.stats()
        agg = {}
        for n, t, _k in batch.stage_times():
            a = agg.setdefault(n, [0.0, 0])
            a[0] += t
            a[1] += 1
        print("%s: %d graphs, %.1f ms per run, %d launches, %d chunks" % (name, args.graphs, st.last_run_ms, st.kernel_launches_per_run, st.chunks))
        for n, (t, k) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print("  %-28s %10.2f ms  (%d stage launches)" % (n, t, k))
        batch.destroy()
    eng.close()


if __name__ == "__main__":
    main()
