"""The reference's own benchmark suite — examples/benchmarks.rs (adapted there from padenot/webaudio-benchmark) — rebuilt graph by
graph through the mirrored control API, so that each scenario can be rendered by the CUDA engine and by the oracle and compared.
Scaled: the reference renders 120 s (DURATION, benchmarks.rs:71); `seconds` here is what DURATION becomes (the per-scenario
divisions, /4, /8, /16, are kept).  The `samples/think-*.wav` assets are replaced by a seeded synthetic phrase of the same shape
(mono / stereo, at 38 kHz and 48 kHz); the two random scenarios (reverb IR, grains) use a seeded generator.
Every builder cites the lines of benchmarks.rs it restates."""
import numpy as np

SR = 48000.0


def think(sample_rate, channels, seconds=0.61):
    """Stand-in for samples/think-{mono,stereo}-{38000,48000}.wav: a short percussive phrase (decaying tones + noise)."""
    n = int(seconds * sample_rate)
    t = np.arange(n) / sample_rate
    rng = np.random.default_rng(38 + channels)
    out = []
    for c in range(channels):
        x = 0.5 * np.sin(2 * np.pi * (180.0 + 40.0 * c) * t) * np.exp(-3.0 * t) + 0.25 * np.sin(2 * np.pi * 1200.0 * t) * np.exp(-9.0 * ((t * 4.0) % 1.0))
        x = x + 0.1 * rng.standard_normal(n) * np.exp(-5.0 * t)
        out.append(x.astype(np.float32))
    return out


def _looped_source(pkg, c, rate, channels):
    src = c.create_buffer_source(pkg.AudioBuffer(think(rate, channels), rate), loop=True)
    src.start()
    return src


def _positional(pkg, c):  # benchmarks.rs:127-134 / 180-187
    p = c.create_panner()
    p.connect(c.destination())
    for name, v in [("position_x", 1.0), ("position_y", 2.0), ("position_z", 3.0), ("orientation_x", 1.0), ("orientation_y", 2.0), ("orientation_z", 3.0)]:
        getattr(p, name).set_value(v)
    return p


def baseline_silence(pkg, be, seconds):  # :86-92
    return pkg.OfflineAudioContext(1, int(seconds * SR), SR, be)


def _simple_source(ctx_channels, rate, buf_channels, positional=False):
    def build(pkg, be, seconds):  # :94-198, :200-226
        c = pkg.OfflineAudioContext(ctx_channels, int(seconds * SR), SR, be)
        src = _looped_source(pkg, c, rate, buf_channels)
        src.connect(_positional(pkg, c) if positional else c.destination())
        return c
    return build


def _simple_mixing(distinct):
    def build(pkg, be, seconds):  # :228-268 (100 looped 38 kHz sources into the stereo destination, DURATION / 4)
        c = pkg.OfflineAudioContext(2, int(seconds / 4 * SR), SR, be)
        pcm = think(38000.0, 1)
        shared = pkg.AudioBuffer(pcm, 38000.0)
        for _ in range(100):
            buf = pkg.AudioBuffer([pcm[0].copy()], 38000.0) if distinct else shared
            src = c.create_buffer_source(buf, loop=True)
            src.connect(c.destination())
            src.start()
        return c
    return build


def mixing_with_gains(pkg, be, seconds):  # :270-305
    c = pkg.OfflineAudioContext(2, int(seconds * SR), SR, be)
    gain = c.create_gain()
    gain.connect(c.destination())
    gain.gain.set_value(-1.0)
    gains_i = []
    for _ in range(4):
        g = c.create_gain()
        g.connect(gain)
        g.gain.set_value(0.25)
        gains_i.append(g)
    for _ in range(2):
        src = _looped_source(pkg, c, 38000.0, 1)
        for g in gains_i:
            gij = c.create_gain()
            gij.gain.set_value(0.5)
            gij.connect(g)
            src.connect(gij)
    return c


def convolution_reverb(pkg, be, seconds):  # :307-349 (DURATION / 8; 4 s stereo IR of decaying noise — scaled with `seconds` too)
    c = pkg.OfflineAudioContext(1, int(seconds / 8 * SR), SR, be)
    rng = np.random.default_rng(307)
    n = int(4.0 * SR * min(1.0, seconds / 120.0 * 8.0))
    env = (np.float32(1.0) - np.arange(n, dtype=np.float32) / np.float32(n)) ** np.float32(10.0)
    ir = [((rng.uniform(0.0, 2.0, n).astype(np.float32) - np.float32(1.0)) * env).astype(np.float32) for _ in range(2)]
    conv = c.create_convolver(pkg.AudioBuffer(ir, SR))
    conv.connect(c.destination())
    src = _looped_source(pkg, c, SR, 1)
    src.connect(conv)
    return c


def granular_synthesis(pkg, be, seconds):  # :351-388 (DURATION / 16, one grain every 5 ms)
    dur = seconds / 16.0
    c = pkg.OfflineAudioContext(1, int(dur * SR), SR, be)
    buf = pkg.AudioBuffer(think(SR, 1), SR)
    rng = np.random.default_rng(351)
    offset = 0.0
    while offset < dur:
        env = c.create_gain()
        env.connect(c.destination())
        src = c.create_buffer_source(buf)
        src.connect(env)
        rand_start = int(rng.integers(0, 1000)) / 1000.0 * 0.5
        rand_duration = int(rng.integers(0, 1000)) / 1000.0 * 0.999
        start = offset * rand_start
        end = start + 0.005 * rand_duration
        start_release = max(offset + end - start, 0.0)
        env.gain.set_value_at_time(0.0, offset)
        env.gain.linear_ramp_to_value_at_time(0.5, offset + 0.005)
        env.gain.set_value_at_time(0.5, start_release)
        env.gain.linear_ramp_to_value_at_time(0.0, start_release + 0.05)
        src.start_at_with_offset_and_duration(offset, start, end)
        offset += 0.005
    return c


def _synth(envelope, with_gain):
    def build(pkg, be, seconds):  # :390-467 (44.1 kHz; a one-second sawtooth note every 140 bpm sixteenth... every 0.583 s)
        sr = 44100.0
        c = pkg.OfflineAudioContext(1, int(seconds * sr), sr, be)
        offset = 0.0
        while offset < seconds:
            dst = c.destination()
            if with_gain:
                env = c.create_gain()
                env.connect(c.destination())
                dst = env
            osc = c.create_oscillator()
            osc.connect(dst)
            osc.set_type(pkg.SAWTOOTH)
            osc.frequency.set_value(110.0)
            if envelope:
                env.gain.set_value_at_time(0.0, 0.0)
                env.gain.set_value_at_time(0.5, offset)
                env.gain.set_target_at_time(0.0, offset + 0.01, 0.1)
            osc.start_at(offset)
            osc.stop_at(offset + 1.0)
            offset += 140.0 / 60.0 / 4.0
        return c
    return build


def subtractive_synth(pkg, be, seconds):  # :469-504
    sr = 44100.0
    c = pkg.OfflineAudioContext(1, int(seconds * sr), sr, be)
    filt = c.create_biquad_filter()
    filt.connect(c.destination())
    filt.frequency.set_value_at_time(0.0, 0.0)
    filt.q.set_value_at_time(20.0, 0.0)
    env = c.create_gain()
    env.connect(filt)
    env.gain.set_value_at_time(0.0, 0.0)
    osc = c.create_oscillator()
    osc.connect(env)
    osc.set_type(pkg.SAWTOOTH)
    osc.frequency.set_value(110.0)
    osc.start()
    offset = 0.0
    while offset < seconds:
        env.gain.set_value_at_time(1.0, offset)
        env.gain.set_target_at_time(0.0, offset, 0.1)
        filt.frequency.set_value_at_time(0.0, offset)
        filt.frequency.set_target_at_time(3500.0, offset, 0.03)
        offset += 140.0 / 60.0 / 16.0
    return c


def _stereo_panning(automation):
    def build(pkg, be, seconds):  # :506-543
        c = pkg.OfflineAudioContext(2, int(seconds * SR), SR, be)
        p = c.create_stereo_panner()
        p.connect(c.destination())
        if automation:
            p.pan.set_value_at_time(-1.0, 0.0)
            p.pan.set_value_at_time(0.2, 0.5)
        else:
            p.pan.set_value(0.1)
        _looped_source(pkg, c, SR, 2).connect(p)
        return c
    return build


def sawtooth_with_automation(pkg, be, seconds):  # :545-559 (the 10 s ramp is scaled with the duration)
    c = pkg.OfflineAudioContext(2, int(seconds * SR), SR, be)
    osc = c.create_oscillator()
    osc.connect(c.destination())
    osc.set_type(pkg.SAWTOOTH)
    osc.frequency.set_value(2000.0)
    osc.frequency.linear_ramp_to_value_at_time(20.0, 10.0 * seconds / 120.0 * 4.0)
    osc.start_at(0.0)
    return c


def stereo_source_with_delay(pkg, be, seconds):  # :561-578 (1 s delay: `seconds` must exceed 1)
    c = pkg.OfflineAudioContext(2, int(seconds * SR), SR, be)
    d = c.create_delay(1.0)
    d.delay_time.set_value(1.0)
    d.connect(c.destination())
    _looped_source(pkg, c, SR, 2).connect(d)
    return c


def iir_filter(pkg, be, seconds):  # :580-606 (lowpass at 200 Hz computed from the biquad)
    c = pkg.OfflineAudioContext(2, int(seconds * SR), SR, be)
    iir = c.create_iir_filter([0.0002029799640409502, 0.0004059599280819004, 0.0002029799640409502],
                              [1.0126964557853775, -1.9991880801438362, 0.9873035442146225])
    iir.connect(c.destination())
    _looped_source(pkg, c, SR, 2).connect(iir)
    return c


def biquad_filter(pkg, be, seconds):  # :608-626
    c = pkg.OfflineAudioContext(2, int(seconds * SR), SR, be)
    bq = c.create_biquad_filter()
    bq.connect(c.destination())
    bq.frequency.set_value(200.0)
    _looped_source(pkg, c, SR, 2).connect(bq)
    return c


SCENARIOS = [
    ("Baseline (silence)", baseline_silence),
    ("Simple source test without resampling (Mono)", _simple_source(1, SR, 1)),
    ("Simple source test without resampling (Stereo)", _simple_source(2, SR, 2)),
    ("Simple source test without resampling (Stereo and positional)", _simple_source(2, SR, 2, True)),
    ("Simple source test with resampling (Mono)", _simple_source(1, 38000.0, 1)),
    ("Simple source test with resampling (Stereo)", _simple_source(2, 38000.0, 2)),
    ("Simple source test with resampling (Stereo and positional)", _simple_source(2, 38000.0, 2, True)),
    ("Upmix without resampling (Mono -> Stereo)", _simple_source(2, SR, 1)),
    ("Downmix without resampling (Stereo -> Mono)", _simple_source(1, SR, 2)),
    ("Simple mixing (100x same buffer)", _simple_mixing(False)),
    ("Simple mixing (100 different buffers)", _simple_mixing(True)),
    ("Simple mixing with gains", mixing_with_gains),
    ("Convolution reverb", convolution_reverb),
    ("Granular synthesis", granular_synthesis),
    ("Synth (Sawtooth with Envelope)", _synth(True, True)),
    ("Synth (Sawtooth with gain - no automation)", _synth(False, True)),
    ("Synth (Sawtooth without gain)", _synth(False, False)),
    ("Substractive Synth", subtractive_synth),
    ("Stereo panning", _stereo_panning(False)),
    ("Stereo panning with automation", _stereo_panning(True)),
    ("Sawtooth with automation", sawtooth_with_automation),
    ("Stereo source with delay", stereo_source_with_delay),
    ("IIR filter", iir_filter),
    ("Biquad filter", biquad_filter),
]


# ---- benches/my_benchmark.rs (criterion / iai): 10 s at 48 kHz stereo (1 s for the HRTF one); `seconds` scales the 10 ------------
def _bench_ctx(pkg, be, seconds, short=False):
    return pkg.OfflineAudioContext(2, int((seconds / 10.0 if short else seconds) * SR), SR, be)


def bench_ctor(pkg, be, seconds):  # :43-46
    return _bench_ctx(pkg, be, seconds)


def bench_constant_source(pkg, be, seconds):  # :56-65 (start at 1 s, stop at 9 s of 10)
    c = _bench_ctx(pkg, be, seconds)
    src = c.create_constant_source()
    src.connect(c.destination())
    src.start_at(0.1 * seconds)
    src.stop_at(0.9 * seconds)
    return c


def bench_sine(pkg, be, seconds):  # :67-75
    c = _bench_ctx(pkg, be, seconds)
    osc = c.create_oscillator()
    osc.connect(c.destination())
    osc.start()
    return c


def bench_detuned_sine(pkg, be, seconds):  # :77-87
    c = _bench_ctx(pkg, be, seconds)
    osc = c.create_oscillator()
    osc.detune.linear_ramp_to_value_at_time(1000.0, seconds)
    osc.connect(c.destination())
    osc.start()
    return c


def bench_sine_gain(pkg, be, seconds, delay=False):  # :89-101, :103-119
    c = _bench_ctx(pkg, be, seconds)
    osc = c.create_oscillator()
    gain = c.create_gain()
    if delay:
        d = c.create_delay(0.3)
        d.delay_time.set_value(0.2)
        osc.connect(d)
        d.connect(gain)
    else:
        gain.gain.set_value(0.5)  # "avoid happy path"
        osc.connect(gain)
    gain.connect(c.destination())
    osc.start()
    return c


def _bench_buffer_src(to):
    def build(pkg, be, seconds):  # :121-196, :222-257: the looped stereo asset into one node
        c = _bench_ctx(pkg, be, seconds)
        src = c.create_buffer_source(pkg.AudioBuffer(think(SR, 2, 2.1), SR), loop=True)  # think-stereo-48000.wav is 101 129 frames
        node = to(pkg, c)
        if node is None:
            src.connect(c.destination())
        else:
            node.connect(c.destination())
            src.connect(node)
        src.start()
        return c
    return build


def _delay_02(pkg, c):
    d = c.create_delay(0.3)
    d.delay_time.set_value(0.2)
    return d


def _iir_200(pkg, c):
    return c.create_iir_filter([0.0002029799640409502, 0.0004059599280819004, 0.0002029799640409502],
                               [1.0126964557853775, -1.9991880801438362, 0.9873035442146225])


def _biquad_200(pkg, c):
    b = c.create_biquad_filter()
    b.frequency.set_value(200.0)
    return b


def _panning_automation(pkg, c):
    p = c.create_stereo_panner()
    p.pan.set_value_at_time(-1.0, 0.0)
    p.pan.set_value_at_time(0.2, 0.5)
    return p


def bench_hrtf_panners(pkg, be, seconds):  # :259-278: one oscillator into two HRTF panners at x = +-10, 1 s (the backend needs an HRIR sphere)
    c = _bench_ctx(pkg, be, seconds, short=True)
    osc = c.create_oscillator()
    for x in (10.0, -10.0):
        p = c.create_panner(panning_model=pkg.context.HRTF)
        p.position_x.set_value(x)
        p.connect(c.destination())
        osc.connect(p)
    osc.start()
    return c


CRITERION = [
    ("bench_ctor", bench_ctor),
    ("bench_constant_source", bench_constant_source),
    ("bench_sine", bench_sine),
    ("bench_detuned_sine", bench_detuned_sine),
    ("bench_sine_gain", bench_sine_gain),
    ("bench_sine_gain_delay", lambda pkg, be, seconds: bench_sine_gain(pkg, be, seconds, delay=True)),
    ("bench_buffer_src", _bench_buffer_src(lambda pkg, c: None)),
    ("bench_buffer_src_delay", _bench_buffer_src(_delay_02)),
    ("bench_buffer_src_iir", _bench_buffer_src(_iir_200)),
    ("bench_buffer_src_biquad", _bench_buffer_src(_biquad_200)),
    ("bench_stereo_positional", _bench_buffer_src(lambda pkg, c: _positional_node(pkg, c))),
    ("bench_stereo_panning_automation", _bench_buffer_src(_panning_automation)),
    ("bench_analyser_node", _bench_buffer_src(lambda pkg, c: c.create_analyser())),
    ("bench_hrtf_panners", bench_hrtf_panners),
]


def _positional_node(pkg, c):  # :198-220 (the panner of _positional, not yet connected)
    p = c.create_panner()
    for name, v in [("position_x", 1.0), ("position_y", 2.0), ("position_z", 3.0), ("orientation_x", 1.0), ("orientation_y", 2.0), ("orientation_z", 3.0)]:
        getattr(p, name).set_value(v)
    return p
