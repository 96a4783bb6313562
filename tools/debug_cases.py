import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as ge
import graphs as G
pkg = ge.load_package()
eng = pkg.Engine(0)
oracle = pkg.context.Backend(pkg.Api(ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle.so")), "wao_"))
N = 128 * 30 + 57
rng = np.random.default_rng(0)
pcm1 = rng.uniform(-0.6, 0.6, (1, 1800)).astype(np.float32)
pcm2 = rng.uniform(-0.6, 0.6, (2, 1800)).astype(np.float32)

def case_delay(ch, dt, loop=False, rate=1.0, via=None):
    def build(be):
        c = pkg.OfflineAudioContext(2, N, G.SR, be)
        s = c.create_buffer_source(pkg.AudioBuffer(list(pcm1 if ch == 1 else pcm2), G.SR), loop=loop, playback_rate=rate)
        d = c.create_delay(max_delay_time=0.1, delay_time=dt)
        x = s
        if via == "analyser":
            a = c.create_analyser(fft_size=256); s.connect(a); x = a
        if via == "gain":
            a = c.create_gain(0.7); s.connect(a); x = a
        x.connect(d); d.connect(c.destination()); s.start()
        return c
    return build

def case_span(kind):
    def build(be):
        c = pkg.OfflineAudioContext(2, N, G.SR, be)
        o1 = c.create_oscillator(type_=1, frequency=328.0); o2 = c.create_oscillator(type_=2, frequency=2150.0)
        sp = c.create_stereo_panner(pan=-0.494)
        if kind == "merger":
            m = c.create_channel_merger(2); o1.connect(m); o2.connect_from_output_to_input(m, 0, 1); m.connect(sp)
        elif kind == "fanin":
            o1.connect(sp); o2.connect(sp)
        elif kind == "mono":
            o1.connect(sp)
        elif kind == "ramp":
            o1.connect(sp); sp.pan.linear_ramp_to_value_at_time(0.5, 0.05)
        elif kind == "merger_ramp":
            m = c.create_channel_merger(2); o1.connect(m); o2.connect_from_output_to_input(m, 0, 1); m.connect(sp); sp.pan.linear_ramp_to_value_at_time(0.5, 0.05)
        elif kind == "comp_merger":
            d = c.create_dynamics_compressor(); o1.connect(d); m = c.create_channel_merger(2); d.connect(m); o2.connect_from_output_to_input(m, 0, 1); m.connect(sp)
        sp.connect(c.destination()); o1.start(); o2.start()
        return c
    return build

cases = {"delay mono short 110": case_delay(1, 0.00229), "delay stereo short 110": case_delay(2, 0.00229), "delay mono 0.0181": case_delay(1, 0.0181),
         "delay stereo 0.0181": case_delay(2, 0.0181), "delay loop .73": case_delay(2, 0.00874, True, 0.73), "delay via analyser": case_delay(2, 0.00229, via="analyser"),
         "delay via gain": case_delay(2, 0.00229, via="gain"), "delay mono 0.0": case_delay(1, 0.0)}
for k in ["merger", "fanin", "mono", "ramp", "merger_ramp", "comp_merger"]:
    cases["span " + k] = case_span(k)
for name, build in cases.items():
    a = G.render(pkg, [build(eng.backend)])[0]
    b = G.render(pkg, [build(oracle)])[0]
    d = np.abs(a.astype(np.float64) - b).max(axis=0)
    bad = np.where(d > 2e-5)[0]
    print(f"{name:26s} max {d.max():.2e} first bad {bad[0] if len(bad) else -1} n_bad {len(bad)}")
eng.close()
