"""Runs one of the extra workloads a few times (for ncu captures): python tools/profile_workload.py north_star|C3|C4|C5 [graphs] [seconds]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import graphs as G  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "north_star"
    graphs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
    pkg = ge.load_package()
    eng = pkg.Engine(0)
    n = int(48000 * seconds)
    ir3 = G.synthetic_ir(144000, 2, decay=0.6)
    if name == "north_star":
        build = lambda g: G.north_star_voices_convolver(pkg, eng.backend, 1000, n, ir3, seed=g)
    elif name == "C3":
        build = lambda g: G.c3_many_voices(pkg, eng.backend, 4096, n)
    elif name == "C4":
        build = lambda g: G.c4_convolver(pkg, eng.backend, g, n, ir3)
    else:
        eng.backend.set_hrir_sphere(G.synthetic_hrir_sphere(48000, 512))
        build = lambda g: G.c5_full_chain(pkg, eng.backend, g, n, ir3)
    eng.set_option(pkg.OPT_PIPELINE_GROUPS, 1)
    import time
    t0 = time.perf_counter()
    ctxs = [build(g) for g in range(graphs)]
    t1 = time.perf_counter()
    batch = pkg.Batch(ctxs)
    t2 = time.perf_counter()
    print(f"graph building (ctypes) {1e3 * (t1 - t0):.1f} ms, wae_batch_prepare {1e3 * (t2 - t1):.1f} ms")
    batch.set_timing(True)
    for _ in range(3):
        batch.run()
        batch.sync()
    print(name, graphs, "graphs", seconds, "s:", round(batch.stats().last_run_ms, 3), "ms", batch.stage_times())
    batch.destroy()
    eng.close()


if __name__ == "__main__":
    main()
