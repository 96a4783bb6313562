import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as ge
import graphs as G
pkg = ge.load_package()
eng = pkg.Engine(0)
oracle = pkg.context.Backend(pkg.Api(ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle.so")), "wao_"))
n = 128 * 40

def build(be, variant):
    pcm = G.c2_source(0, n)
    c = pkg.OfflineAudioContext(2, n, G.SR, be)
    parts = variant.split("+")
    src = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], G.SR))
    bq = c.create_biquad_filter(type_=pkg.BANDPASS, frequency=900.0, q=12.0)
    dl = c.create_delay(max_delay_time=0.05, delay_time=0.013)
    gn = c.create_gain(0.8)
    if "ramp" in parts:
        gn.gain.linear_ramp_to_value_at_time(0.2, 0.09)
    osc = c.create_oscillator(type_=pkg.SAWTOOTH, frequency=333.0)
    og = c.create_gain(0.1)
    src.connect(bq); bq.connect(dl); dl.connect(gn); gn.connect(c.destination())
    osc.connect(og); og.connect(c.destination())
    src.start(); osc.start()
    def cb(ctx):
        if "hp" in parts:
            hp = ctx.create_biquad_filter(type_=pkg.HIGHPASS, frequency=2500.0)
            bq.connect(hp); hp.connect(ctx.destination())
        if "ev" in parts:
            gn.gain.set_value_at_time(0.9, ctx.current_time() + 0.004)
        if "k" in parts:
            k = ctx.create_constant_source(offset=0.05)
            k.connect(ctx.destination()); k.start_at(ctx.current_time()); k.stop_at(ctx.current_time() + 0.01)
        if "og" in parts:
            og.disconnect()
        if "dl" in parts:
            dl.disconnect(); dl.connect(ctx.destination())
    c.suspend_sync(128 * 7 / G.SR, cb)
    return c

for variant in ["none", "hp", "ramp", "ramp+ev", "ev", "k", "og", "dl", "ramp+hp+ev+k", "og+dl"]:
    a = G.render(pkg, [build(eng.backend, variant)])[0]
    b = G.render(pkg, [build(oracle, variant)])[0]
    d = np.abs(a.astype(np.float64) - b).max(axis=0)
    bad = np.where(d > 1e-5)[0]
    print(f"{variant:16s} max {d.max():.3e} first bad frame {bad[0] if len(bad) else -1} (quantum {bad[0] // 128 if len(bad) else -1})")
eng.close()
