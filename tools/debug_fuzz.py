import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as ge
import graphs as G
import test_gpu_fuzz as F
pkg = ge.load_package()
eng = pkg.Engine(0)
oracle = pkg.context.Backend(pkg.Api(ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle.so")), "wao_"))
log = []
C = pkg.OfflineAudioContext
for name in [n for n in dir(C) if n.startswith("create_")]:
    orig = getattr(C, name)
    def mk(orig, name):
        def f(self, *a, **k):
            log.append(name[7:] + "(" + ",".join(f"{kk}={vv:.3g}" if isinstance(vv, float) else f"{kk}={vv}" for kk, vv in k.items() if not hasattr(vv, "shape") and not isinstance(vv, (list, tuple)) and kk != "cfg") + ")")
            return orig(self, *a, **k)
        return f
    setattr(C, name, mk(orig, name))
seeds = [int(x) for x in sys.argv[1:]] or [1000 * s + g for s in range(24) for g in range(5)]
for seed in seeds:
    try:
        log.clear()
        a = G.render(pkg, [F.random_graph(pkg, eng.backend, seed)])[0]
        desc = " ".join(log)
        b = G.render(pkg, [F.random_graph(pkg, oracle, seed)])[0]
    except pkg.WaeError as e:
        print(seed, "ERR", str(e)[:80]); continue
    if not np.isfinite(b).all():
        print(seed, "oracle non-finite"); continue
    d = np.abs(a.astype(np.float64) - b).max(axis=0)
    bad = np.where(d > 2e-5)[0]
    if len(bad) and os.environ.get("FUZZ_TAPS"):
        for tap in range(40):
            log.clear()
            cg = F.random_graph(pkg, eng.backend, seed, tap=tap)
            if cg is None:
                break
            names = list(log)
            ta = G.render(pkg, [cg])[0]
            tb = G.render(pkg, [F.random_graph(pkg, oracle, seed, tap=tap)])[0]
            td = np.abs(ta.astype(np.float64) - tb).max(axis=0)
            tbad = np.where(td > 2e-5)[0]
            print(f"   tap {tap:2d} {chr(66)+chr(65)+chr(68) if len(tbad) else chr(111)+chr(107)} max {td.max():.2e} first {tbad[0] if len(tbad) else -1} :: {names[tap] if tap < len(names) else chr(63)}")
    if len(bad):
        print(seed, f"max {d.max():.3e} first bad frame {bad[0]} (q {bad[0] // 128}) n_bad {len(bad)} |ref| {np.abs(b).max():.3g} :: {desc}")
eng.close()
