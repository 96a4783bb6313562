"""Graph builders shared by the parity tests, bench.py and smoke(): each function builds the SAME graph on any
backend (the CUDA engine or the CPU oracle) through the mirrored control API, so a parity test is
`render(build(gpu)) vs render(build(oracle))`.  The BASELINE.json configs C1..C5 are here."""
import numpy as np

SR = 48000.0


def c1_osc_biquad(pkg, backend, length=48000, sr=SR):
    """C1: OscillatorNode(440 Hz sine) -> BiquadFilterNode(lowpass 350 Hz, Q 1) -> destination (tests/offline.rs style)."""
    c = pkg.OfflineAudioContext(2, length, sr, backend)
    osc = c.create_oscillator()
    bq = c.create_biquad_filter()
    osc.connect(bq)
    bq.connect(c.destination())
    osc.start()
    return c


def c2_params(g):
    rng = np.random.default_rng(1000 + g)
    f0 = float(np.exp(rng.uniform(np.log(100.0), np.log(8000.0))))
    q = float(rng.uniform(0.5, 4.0))
    gain = float(rng.uniform(0.1, 0.9))
    return rng, f0, q, gain


def c2_source(g, frames):
    rng, *_ = c2_params(g)
    return rng.uniform(-1.0, 1.0, (2, frames)).astype(np.float32)


def c2_buffer_biquad_gain(pkg, backend, g, length, sr=SR, pcm=None):
    """C2: AudioBufferSource(stereo noise) -> Biquad(lowpass, seeded f0/Q) -> Gain -> destination."""
    _, f0, q, gain = c2_params(g)
    if pcm is None:
        pcm = c2_source(g, length)
    c = pkg.OfflineAudioContext(2, length, sr, backend)
    src = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], sr))
    bq = c.create_biquad_filter(type_=pkg.LOWPASS, frequency=f0, q=q)
    gn = c.create_gain(gain)
    src.connect(bq)
    bq.connect(gn)
    gn.connect(c.destination())
    src.start()
    return c


def c3_many_voices(pkg, backend, voices=4096, length=48000, sr=SR):
    """C3: `voices` x (sine osc f = 55*2^(v/512) -> bandpass f0 = 2f, Q 5), all summed at the destination."""
    c = pkg.OfflineAudioContext(2, length, sr, backend)
    for v in range(voices):
        f = 55.0 * 2.0 ** (v / 512.0)
        osc = c.create_oscillator(frequency=f)
        bq = c.create_biquad_filter(type_=pkg.BANDPASS, frequency=2.0 * f, q=5.0)
        osc.connect(bq)
        bq.connect(c.destination())
        osc.start()
    return c


def synthetic_ir(frames, channels=2, seed=99, decay=0.25, sr=SR):
    """Exponentially decaying noise (the shape of examples/benchmarks.rs:307-349 'Convolution reverb')."""
    rng = np.random.default_rng(seed)
    t = np.arange(frames) / sr
    env = np.exp(-t / decay)
    return [(rng.standard_normal(frames) * env).astype(np.float32) for _ in range(channels)]


def c4_convolver(pkg, backend, g, length, ir, sr=SR, burst=4096):
    """C4: stereo AudioBufferSource (noise burst + low-level noise) -> ConvolverNode(normalize) -> destination."""
    rng = np.random.default_rng(4000 + g)
    pcm = (rng.uniform(-1.0, 1.0, (2, length)) * 0.05).astype(np.float32)
    pcm[:, :burst] += rng.uniform(-0.5, 0.5, (2, min(burst, length))).astype(np.float32)
    c = pkg.OfflineAudioContext(2, length, sr, backend)
    src = c.create_buffer_source(pkg.AudioBuffer([pcm[0], pcm[1]], sr))
    cv = c.create_convolver(pkg.AudioBuffer(ir, sr))
    src.connect(cv)
    cv.connect(c.destination())
    src.start()
    return c


def north_star_voices_convolver(pkg, backend, voices, length, ir, sr=SR, seed=0):
    """north_star graph: `voices` x (oscillator -> biquad) summed into ONE ConvolverNode -> destination."""
    rng = np.random.default_rng(7000 + seed)
    c = pkg.OfflineAudioContext(2, length, sr, backend)
    cv = c.create_convolver(pkg.AudioBuffer(ir, sr))
    cv.connect(c.destination())
    types = [pkg.SINE, pkg.SAWTOOTH, pkg.SQUARE, pkg.TRIANGLE]
    for v in range(voices):
        f = float(55.0 * 2.0 ** rng.uniform(0.0, 6.0))
        osc = c.create_oscillator(type_=types[v % 4], frequency=f, detune=float(rng.uniform(-20, 20)))
        bq = c.create_biquad_filter(type_=pkg.LOWPASS, frequency=min(4.0 * f, 18000.0), q=float(rng.uniform(0.5, 6.0)))
        gn = c.create_gain(1.0 / voices)
        osc.connect(bq)
        bq.connect(gn)
        gn.connect(cv)
        osc.start()
    return c


def render(pkg, contexts, threads=1):
    bufs = pkg.render_batch(contexts, threads=threads)
    return np.stack([np.stack(b.channels) for b in bufs])
