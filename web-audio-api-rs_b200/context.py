"""Host-side mirror of the reference's control API for the OfflineAudioContext path.

Same names and argument meaning as web-audio-api-rs:
  OfflineAudioContext::new / create_* / destination / start_rendering_sync   (src/context/offline.rs, base.rs)
  AudioNode::connect / connect_from_output_to_input / disconnect               (src/node/audio_node.rs:224-466)
  AudioScheduledSourceNode::start / start_at / stop / stop_at                  (src/node/scheduled_source.rs)
  AudioParam::value / set_value / *_at_time                                    (src/param.rs:336-662)
so that the parity tests read like the reference's own `#[test]`s.  Everything is forwarded through the
C ABI of include/wae.h; errors surface as WaeError carrying the reference's panic text.

One addition that the reference does not have: `render_batch([ctx, ...])` renders many independent
OfflineAudioContexts in ONE engine call — the whole point of the GPU engine.
"""
import ctypes as C
import math

import numpy as np

from . import _binding as B

F64_MAX = 1.7976931348623157e308

# enums, same order as the reference
SINE, SQUARE, SAWTOOTH, TRIANGLE, CUSTOM = range(5)
LOWPASS, HIGHPASS, BANDPASS, NOTCH, ALLPASS, PEAKING, LOWSHELF, HIGHSHELF = range(8)
MAX, CLAMPED_MAX, EXPLICIT = range(3)
SPEAKERS, DISCRETE = range(2)
EQUALPOWER, HRTF = range(2)
LINEAR, INVERSE, EXPONENTIAL = range(3)
OVERSAMPLE_NONE, OVERSAMPLE_X2, OVERSAMPLE_X4 = range(3)


class Backend:
    """An Api plus (for the product) an engine handle."""

    def __init__(self, api, engine=None):
        import weakref
        self.api = api
        self.engine = engine
        self.batches = weakref.WeakSet()  # live wae_batch handles: destroyed before their engine (Engine.close)
        self.closed = False

    def close_batches(self):
        for b in list(self.batches):
            b.destroy()
        self.closed = True

    def set_hrir_sphere(self, data):
        """load_hrtf_processor (src/node/panner.rs:39-68): hand over the bytes of the HRIR sphere (IRC_1003_C.bin format)."""
        buf = C.create_string_buffer(bytes(data), len(data))
        if self.api.is_product:
            self.api.check(self.api.engine_set_hrir_sphere(self.engine, C.cast(buf, C.c_void_p), len(data)))
        else:
            self.api.check(self.api.set_hrir_sphere(C.cast(buf, C.c_void_p), len(data)))


class AudioBuffer:
    """Planar f32 PCM (src/buffer.rs:69-72)."""

    def __init__(self, channels, sample_rate):
        self.channels = [B.as_f32(c) for c in channels]
        # AudioBuffer::from (src/buffer.rs:96-117) panics on ragged channels
        if any(len(c) != len(self.channels[0]) for c in self.channels):
            raise B.WaeError(2, "all channels of an AudioBuffer must have the same length")
        self.sample_rate = float(np.float32(sample_rate))

    @classmethod
    def zeros(cls, number_of_channels, length, sample_rate):
        return cls([np.zeros(length, np.float32) for _ in range(number_of_channels)], sample_rate)

    def number_of_channels(self):
        return len(self.channels)

    def length(self):
        return len(self.channels[0]) if self.channels else 0

    def duration(self):
        return self.length() / self.sample_rate

    def get_channel_data(self, i):
        return self.channels[i]

    def copy_to_channel(self, data, i):
        d = B.as_f32(data)
        self.channels[i][: len(d)] = d

    def _desc(self):
        return B.buffer_desc(self.channels, self.sample_rate)


def channel_config(count=0, count_mode=MAX, interpretation=SPEAKERS):
    """AudioNodeOptions; count == 0 selects the node's default config."""
    return B.ChannelConfig(count, count_mode, interpretation)


class AudioParam:
    def __init__(self, ctx, node_id, index, value, min_value=-3.4028234663852886e38, max_value=3.4028234663852886e38):
        self._ctx, self._node, self._index = ctx, node_id, index
        self._value = float(np.float32(value))
        self._min, self._max = min_value, max_value

    def _push(self, type_, value=0.0, time=0.0, aux=0.0, values=None):
        ev = B.ParamEvent(type_, float(value), float(time), float(aux), None, 0)
        keep = None
        if values is not None:
            keep = B.as_f32(values)
            ev.values = B.fptr(keep)
            ev.values_len = len(keep)
        api = self._ctx._api
        if self._node == "listener":
            api.check(api.listener_param_event_push(self._ctx._g, self._index, C.byref(ev)))
        else:
            api.check(api.param_event_push(self._ctx._g, self._node, self._index, C.byref(ev)))
        return self

    def value(self):
        return self._value

    def set_value(self, v):
        self._push(0, v)
        self._value = min(max(float(np.float32(v)), self._min), self._max)
        return self

    def set_value_at_time(self, v, start_time):
        return self._push(1, v, start_time)

    def linear_ramp_to_value_at_time(self, v, end_time):
        return self._push(2, v, end_time)

    def exponential_ramp_to_value_at_time(self, v, end_time):
        return self._push(3, v, end_time)

    def cancel_scheduled_values(self, cancel_time):
        return self._push(4, 0.0, cancel_time)

    def set_target_at_time(self, v, start_time, time_constant):
        return self._push(5, v, start_time, time_constant)

    def cancel_and_hold_at_time(self, cancel_time):
        return self._push(6, 0.0, cancel_time)

    def set_value_curve_at_time(self, values, start_time, duration):
        return self._push(7, 0.0, start_time, duration, values)

    def set_automation_rate(self, rate):
        api = self._ctx._api
        api.check(api.param_set_automation_rate(self._ctx._g, self._node, self._index, 0 if rate in ("a", "A", 0) else 1))


class AudioNode:
    def __init__(self, ctx, node_id, n_inputs=1, n_outputs=1):
        self._ctx, self.id = ctx, node_id
        self._n_inputs, self._n_outputs = n_inputs, n_outputs

    def number_of_inputs(self):
        return self._n_inputs

    def number_of_outputs(self):
        return self._n_outputs

    def _set_attribute(self, attribute, value):
        api = self._ctx._api
        api.check(api.node_set_attribute(self._ctx._g, self.id, attribute, float(value)))

    def _set_buffer(self, fn, buffer):
        d, keep = buffer._desc()
        self._ctx._keep.append((d, keep))
        api = self._ctx._api
        api.check(getattr(api, fn)(self._ctx._g, self.id, C.byref(d)))

    # AudioNode::set_channel_count / set_channel_count_mode / set_channel_interpretation (src/node/audio_node.rs:417-441)
    def set_channel_count(self, count):
        api = self._ctx._api
        api.check(api.node_set_channel_count(self._ctx._g, self.id, int(count)))

    def set_channel_count_mode(self, mode):
        api = self._ctx._api
        api.check(api.node_set_channel_count_mode(self._ctx._g, self.id, int(mode)))

    def set_channel_interpretation(self, interpretation):
        api = self._ctx._api
        api.check(api.node_set_channel_interpretation(self._ctx._g, self.id, int(interpretation)))

    def connect(self, dest):
        return self.connect_from_output_to_input(dest, 0, 0)

    def connect_from_output_to_input(self, dest, output, input):
        api = self._ctx._api
        if isinstance(dest, AudioParam):
            api.check(api.connect_param(self._ctx._g, self.id, output, 1 if dest._node == "listener" else dest._node, dest._index))
        else:
            if dest._ctx is not self._ctx:
                raise B.WaeError(1, "InvalidAccessError - Attempting to connect nodes from different contexts")
            api.check(api.connect(self._ctx._g, self.id, output, dest.id, input))
        return dest

    def disconnect(self):
        api = self._ctx._api
        api.check(api.disconnect(self._ctx._g, self.id))

    # the selective forms, src/node/audio_node.rs:304-405
    def _disconnect_from(self, output, dest, input):
        api = self._ctx._api
        if isinstance(dest, AudioParam):
            owner = 1 if dest._node == "listener" else dest._node
            api.check(api.disconnect_param(self._ctx._g, self.id, output, owner, dest._index))
        else:
            if dest is not None and dest._ctx is not self._ctx:
                raise B.WaeError(1, "InvalidAccessError - Attempting to disconnect nodes from different contexts")
            api.check(api.disconnect_from(self._ctx._g, self.id, output, 0xFFFFFFFF if dest is None else dest.id, input))

    def disconnect_dest(self, dest):
        self._disconnect_from(-1, dest, -1)

    def disconnect_output(self, output):
        self._disconnect_from(output, None, -1)

    def disconnect_dest_from_output(self, dest, output):
        self._disconnect_from(output, dest, -1)

    def disconnect_dest_from_output_to_input(self, dest, output, input):
        self._disconnect_from(output, dest, input)


class AudioScheduledSourceNode(AudioNode):
    def start(self):
        self.start_at(0.0)  # context.current_time() is 0 before an offline render

    def start_at(self, when):
        api = self._ctx._api
        api.check(api.source_start(self._ctx._g, self.id, when, 0.0, F64_MAX))

    def stop(self):
        self.stop_at(0.0)

    def stop_at(self, when):
        api = self._ctx._api
        api.check(api.source_stop(self._ctx._g, self.id, when))


class OscillatorNode(AudioScheduledSourceNode):
    def set_periodic_wave(self, table):
        """OscillatorNode::set_periodic_wave (oscillator.rs:334-337); `table` = PeriodicWave's wavetable (periodic_wave.rs:163-209)."""
        t = B.as_f32(table)
        api = self._ctx._api
        api.check(api.oscillator_set_periodic_wave(self._ctx._g, self.id, B.fptr(t), len(t)))

    def set_type(self, type_):
        api = self._ctx._api
        api.check(api.oscillator_set_type(self._ctx._g, self.id, type_))


class AudioBufferSourceNode(AudioScheduledSourceNode):
    # audio_buffer_source.rs:278-349
    def set_buffer(self, buffer):
        self._set_buffer("buffer_source_set_buffer", buffer)

    def set_loop(self, value):
        self._set_attribute(B.ATTR_LOOP, 1.0 if value else 0.0)

    def set_loop_start(self, value):
        self._set_attribute(B.ATTR_LOOP_START, value)

    def set_loop_end(self, value):
        self._set_attribute(B.ATTR_LOOP_END, value)

    def start_at_with_offset(self, start, offset):
        self.start_at_with_offset_and_duration(start, offset, F64_MAX)

    def start_at_with_offset_and_duration(self, start, offset, duration):
        api = self._ctx._api
        api.check(api.source_start(self._ctx._g, self.id, start, offset, duration))


def _response_arrays(frequency_hz):
    f = np.ascontiguousarray(frequency_hz, dtype=np.float32)
    return f, np.zeros(len(f), np.float32), np.zeros(len(f), np.float32)


class ConvolverNode(AudioNode):
    # convolver.rs:259-328
    def set_buffer(self, buffer):
        self._set_buffer("convolver_set_buffer", buffer)

    def set_normalize(self, value):
        self._set_attribute(B.ATTR_NORMALIZE, 1.0 if value else 0.0)


class WaveShaperNode(AudioNode):
    # waveshaper.rs:203-229
    def set_curve(self, curve):
        c = B.as_f32(curve)
        api = self._ctx._api
        api.check(api.wave_shaper_set_curve(self._ctx._g, self.id, B.fptr(c), len(c)))

    def set_oversample(self, oversample):
        self._set_attribute(B.ATTR_OVERSAMPLE, oversample)


class PannerNode(AudioNode):
    # panner.rs:545-657
    def set_panning_model(self, v):
        self._set_attribute(B.ATTR_PANNING_MODEL, v)

    def set_distance_model(self, v):
        self._set_attribute(B.ATTR_DISTANCE_MODEL, v)

    def set_ref_distance(self, v):
        self._set_attribute(B.ATTR_REF_DISTANCE, v)

    def set_max_distance(self, v):
        self._set_attribute(B.ATTR_MAX_DISTANCE, v)

    def set_rolloff_factor(self, v):
        self._set_attribute(B.ATTR_ROLLOFF_FACTOR, v)

    def set_cone_inner_angle(self, v):
        self._set_attribute(B.ATTR_CONE_INNER_ANGLE, v)

    def set_cone_outer_angle(self, v):
        self._set_attribute(B.ATTR_CONE_OUTER_ANGLE, v)

    def set_cone_outer_gain(self, v):
        self._set_attribute(B.ATTR_CONE_OUTER_GAIN, v)


class BiquadFilterNode(AudioNode):
    def set_type(self, type_):
        api = self._ctx._api
        api.check(api.biquad_set_type(self._ctx._g, self.id, type_))
        self.type_ = type_

    def get_frequency_response(self, frequency_hz):
        """BiquadFilterNode::get_frequency_response (src/node/biquad_filter.rs:657-735) -> (mag_response, phase_response)."""
        f, mag, phase = _response_arrays(frequency_hz)
        self._ctx._api.biquad_frequency_response(self.type_, self._ctx._sample_rate, self.frequency.value(), self.detune.value(), self.q.value(),
                                                 self.gain.value(), f.ctypes.data_as(B.c_float_p), mag.ctypes.data_as(B.c_float_p),
                                                 phase.ctypes.data_as(B.c_float_p), len(f))
        return mag, phase


class IIRFilterNode(AudioNode):
    def get_frequency_response(self, frequency_hz):
        """IIRFilterNode::get_frequency_response (src/node/iir_filter.rs:215-265) -> (mag_response, phase_response)."""
        f, mag, phase = _response_arrays(frequency_hz)
        ff, fb = self.feedforward, self.feedback
        self._ctx._api.iir_frequency_response(ff.ctypes.data_as(B.c_double_p), len(ff), fb.ctypes.data_as(B.c_double_p), len(fb), self._ctx._sample_rate,
                                              f.ctypes.data_as(B.c_float_p), mag.ctypes.data_as(B.c_float_p), phase.ctypes.data_as(B.c_float_p), len(f))
        return mag, phase


class DynamicsCompressorNode(AudioNode):
    def reduction(self):
        """DynamicsCompressorNode::reduction (src/node/dynamics_compressor.rs:204-206) after the render."""
        ctx, api = self._ctx, self._ctx._api
        out = C.c_float(0)
        if api.is_product:
            if ctx._batch is None:
                raise B.WaeError(2, "reduction is available after rendering")
            api.check(api.compressor_reduction(ctx._batch.handle, ctx._batch_index, self.id, C.byref(out)))
        else:
            api.check(api.compressor_reduction(ctx._g, self.id, C.byref(out)))
        return out.value


class AnalyserNode(AudioNode):
    def __init__(self, ctx, node_id, fft_size):
        super().__init__(ctx, node_id)
        self.fft_size = fft_size

    # analyser.rs:148-222
    def set_fft_size(self, fft_size):
        self._set_attribute(B.ATTR_FFT_SIZE, fft_size)
        self.fft_size = int(fft_size)

    def set_smoothing_time_constant(self, v):
        self._set_attribute(B.ATTR_SMOOTHING_TIME_CONSTANT, v)

    def set_min_decibels(self, v):
        self._set_attribute(B.ATTR_MIN_DECIBELS, v)

    def set_max_decibels(self, v):
        self._set_attribute(B.ATTR_MAX_DECIBELS, v)

    def frequency_bin_count(self):
        return self.fft_size // 2

    # `out`: a caller-owned array (like the reference's `&mut [f32]` / `&mut [u8]`): only the first min(len, fft_size or
    # frequency_bin_count) entries are written, the rest is left as it was
    def get_float_time_domain_data(self, n=None, out=None):
        return self._ctx._analyser_read(self, "time", n or self.fft_size, out=out)

    def get_float_frequency_data(self, n=None, out=None):
        return self._ctx._analyser_read(self, "freq", n or self.fft_size // 2, out=out)

    def get_byte_time_domain_data(self, n=None, out=None):
        return self._ctx._analyser_read(self, "time", n or self.fft_size, byte=True, out=out)

    def get_byte_frequency_data(self, n=None, out=None):
        return self._ctx._analyser_read(self, "freq", n or self.fft_size // 2, byte=True, out=out)


class AudioListener:
    def __init__(self, ctx):
        names = ["position_x", "position_y", "position_z", "forward_x", "forward_y", "forward_z", "up_x", "up_y", "up_z"]
        defaults = [0, 0, 0, 0, 0, -1, 0, 1, 0]
        for i, (n, d) in enumerate(zip(names, defaults)):
            setattr(self, n, AudioParam(ctx, "listener", i, d))


class OfflineAudioContext:
    """OfflineAudioContext::new(number_of_channels, length, sample_rate), src/context/offline.rs:78."""

    def __init__(self, number_of_channels, length, sample_rate, backend):
        self._backend = backend
        self._api = backend.api
        self._channels, self._length = int(number_of_channels), int(length)
        self._sample_rate = float(np.float32(sample_rate))
        g = C.c_void_p()
        if self._api.is_product:
            self._api.check(self._api.graph_create(backend.engine, self._channels, self._length, self._sample_rate, C.byref(g)))
        else:
            self._api.check(self._api.graph_create(self._channels, self._length, self._sample_rate, C.byref(g)))
        self._g = g
        self._keep = []
        self._dest = AudioNode(self, 0, 1, 1)
        self._batch = None
        self._batch_index = 0
        self._listener = None
        self._suspends = []
        self._current_time = 0.0

    def __del__(self):
        try:
            if self._g:
                self._api.graph_destroy(self._g)
                self._g = None
        except Exception:
            pass

    def render_order(self):
        """Ids in per-quantum processing order (Graph::order_nodes, src/render/graph.rs:331-487); host work on both libraries."""
        ids = (C.c_uint32 * 4096)()
        if self._api.is_product:
            n = C.c_uint32(0)
            self._api.check(self._api.graph_render_order(self._g, ids, 4096, C.byref(n)))
            n = n.value
        else:
            n = self._api.render_order(self._g, ids, 4096)
        return list(ids[:min(n, 4096)])

    # ---- BaseAudioContext
    def destination(self):
        return self._dest

    def sample_rate(self):
        return self._sample_rate

    def length(self):
        return self._length

    def listener(self):
        if self._listener is None:
            self._listener = AudioListener(self)
        return self._listener

    def create_buffer(self, number_of_channels, length, sample_rate):
        return AudioBuffer.zeros(number_of_channels, length, sample_rate)

    def _create(self, fn, opts):
        nid = C.c_uint32()
        self._api.check(getattr(self._api, fn)(self._g, C.byref(opts), C.byref(nid)))
        return nid.value

    def create_oscillator(self, type_=SINE, frequency=440.0, detune=0.0, periodic_wave=None):
        o = B.OscillatorOptions(type_, frequency, detune, None, 0)
        if periodic_wave is not None:
            pw = B.as_f32(periodic_wave)
            self._keep.append(pw)
            o.periodic_wave, o.periodic_wave_len, o.type = B.fptr(pw), len(pw), CUSTOM
        nid = self._create("create_oscillator", o)
        n = OscillatorNode(self, nid, 0, 1)
        nyq = self._sample_rate / 2
        n.frequency = AudioParam(self, nid, 0, frequency, -nyq, nyq)
        n.detune = AudioParam(self, nid, 1, detune, -153600.0, 153600.0)
        return n

    def create_biquad_filter(self, type_=LOWPASS, frequency=350.0, q=1.0, detune=0.0, gain=0.0, cfg=None):
        o = B.BiquadOptions(type_, q, detune, frequency, gain, cfg or channel_config())
        nid = self._create("create_biquad_filter", o)
        n = BiquadFilterNode(self, nid)
        n.type_ = type_
        n.q = AudioParam(self, nid, 0, q)
        n.detune = AudioParam(self, nid, 1, detune, -153600.0, 153600.0)
        n.frequency = AudioParam(self, nid, 2, frequency, 0.0, self._sample_rate / 2)
        n.gain = AudioParam(self, nid, 3, gain)
        return n

    def create_iir_filter(self, feedforward, feedback, cfg=None):
        ff = np.ascontiguousarray(feedforward, dtype=np.float64)
        fb = np.ascontiguousarray(feedback, dtype=np.float64)
        o = B.IirOptions(ff.ctypes.data_as(B.c_double_p), len(ff), fb.ctypes.data_as(B.c_double_p), len(fb),
                         cfg or channel_config())
        n = IIRFilterNode(self, self._create("create_iir_filter", o))
        n.feedforward, n.feedback = ff, fb
        return n

    def create_gain(self, gain=1.0, cfg=None):
        nid = self._create("create_gain", B.GainOptions(gain, cfg or channel_config()))
        n = AudioNode(self, nid)
        n.gain = AudioParam(self, nid, 0, gain)
        return n

    def create_buffer_source(self, buffer=None, detune=0.0, playback_rate=1.0, loop=False, loop_start=0.0, loop_end=0.0):
        o = B.BufferSourceOptions(None, detune, playback_rate, 1 if loop else 0, loop_start, loop_end)
        if buffer is not None:
            d, keep = buffer._desc()
            self._keep.append((d, keep))
            o.buffer = C.pointer(d)
        nid = self._create("create_buffer_source", o)
        n = AudioBufferSourceNode(self, nid, 0, 1)
        n.detune = AudioParam(self, nid, 0, detune)
        n.playback_rate = AudioParam(self, nid, 1, playback_rate)
        return n

    def create_constant_source(self, offset=1.0):
        nid = self._create("create_constant_source", B.ConstantSourceOptions(offset))
        n = AudioScheduledSourceNode(self, nid, 0, 1)
        n.offset = AudioParam(self, nid, 0, offset)
        return n

    def create_convolver(self, buffer=None, disable_normalization=False, cfg=None):
        o = B.ConvolverOptions(None, 1 if disable_normalization else 0, cfg or channel_config())
        if buffer is not None:
            d, keep = buffer._desc()
            self._keep.append((d, keep))
            o.buffer = C.pointer(d)
        return ConvolverNode(self, self._create("create_convolver", o))

    def create_wave_shaper(self, curve=None, oversample=OVERSAMPLE_NONE, cfg=None):
        o = B.WaveShaperOptions(None, 0, oversample, cfg or channel_config())
        if curve is not None:
            c = B.as_f32(curve)
            self._keep.append(c)
            o.curve, o.curve_len = B.fptr(c), len(c)
        return WaveShaperNode(self, self._create("create_wave_shaper", o))

    def create_delay(self, max_delay_time=1.0, delay_time=0.0, cfg=None):
        nid = self._create("create_delay", B.DelayOptions(max_delay_time, delay_time, cfg or channel_config()))
        n = AudioNode(self, nid)
        n.delay_time = AudioParam(self, nid, 0, delay_time, 0.0, max_delay_time)
        return n

    def create_stereo_panner(self, pan=0.0, cfg=None):
        nid = self._create("create_stereo_panner", B.StereoPannerOptions(pan, cfg or channel_config()))
        n = AudioNode(self, nid)
        n.pan = AudioParam(self, nid, 0, pan, -1.0, 1.0)
        return n

    def create_panner(self, panning_model=EQUALPOWER, distance_model=INVERSE, position=(0.0, 0.0, 0.0),
                      orientation=(1.0, 0.0, 0.0), ref_distance=1.0, max_distance=10000.0, rolloff_factor=1.0,
                      cone_inner_angle=360.0, cone_outer_angle=360.0, cone_outer_gain=0.0, cfg=None):
        o = B.PannerOptions(panning_model, distance_model, *position, *orientation, ref_distance, max_distance,
                            rolloff_factor, cone_inner_angle, cone_outer_angle, cone_outer_gain, cfg or channel_config())
        nid = self._create("create_panner", o)
        n = PannerNode(self, nid)
        for i, name in enumerate(["position_x", "position_y", "position_z", "orientation_x", "orientation_y", "orientation_z"]):
            setattr(n, name, AudioParam(self, nid, i, (list(position) + list(orientation))[i]))
        return n

    def create_analyser(self, fft_size=2048, smoothing_time_constant=0.8, min_decibels=-100.0, max_decibels=-30.0, cfg=None):
        o = B.AnalyserOptions(fft_size, smoothing_time_constant, min_decibels, max_decibels, cfg or channel_config())
        return AnalyserNode(self, self._create("create_analyser", o), fft_size)

    def create_dynamics_compressor(self, attack=0.003, knee=30.0, ratio=12.0, release=0.25, threshold=-24.0, cfg=None):
        o = B.DynamicsCompressorOptions(attack, knee, ratio, release, threshold, cfg or channel_config())
        nid = self._create("create_dynamics_compressor", o)
        n = DynamicsCompressorNode(self, nid)
        for i, (name, v) in enumerate([("attack", attack), ("knee", knee), ("ratio", ratio), ("release", release), ("threshold", threshold)]):
            setattr(n, name, AudioParam(self, nid, i, v))
        return n

    def create_channel_merger(self, number_of_inputs=6):
        return AudioNode(self, self._create("create_channel_merger", B.ChannelMergerOptions(number_of_inputs)), number_of_inputs, 1)

    def create_channel_splitter(self, number_of_outputs=6):
        return AudioNode(self, self._create("create_channel_splitter", B.ChannelSplitterOptions(number_of_outputs)), 1, number_of_outputs)

    def current_time(self):
        """BaseAudioContext::current_time: 0 before rendering, the suspend time inside a suspend_sync callback."""
        return self._current_time

    def suspend_sync(self, suspend_time, callback):
        """OfflineAudioContext::suspend_sync (src/context/offline.rs:330-387): `callback(context)` runs when the render reaches
        suspend_time (quantised up to a render quantum) and may mutate the graph."""
        self._suspends.append((float(suspend_time), callback))

    def _run_suspend_callbacks(self):
        """The control half of the render loop (src/render/thread.rs:271-290): take the suspend points in time order."""
        todo, self._suspends = sorted(self._suspends, key=lambda sc: sc[0]), []
        for t, cb in todo:
            self._api.check(self._api.graph_suspend(self._g, t))
            self._current_time = math.ceil(t * self._sample_rate / 128.0) * 128.0 / self._sample_rate
            cb(self)
        self._current_time = 0.0

    # ---- rendering
    def start_rendering_sync(self):
        """OfflineAudioContext::start_rendering_sync (src/context/offline.rs:157-185): a batch of one."""
        return render_batch([self])[0]

    def _analyser_read(self, node, kind, n, byte=False, out=None):
        if out is None:
            out = np.zeros(n, np.uint8 if byte else np.float32)
        else:
            assert out.dtype == (np.uint8 if byte else np.float32) and out.flags.c_contiguous
            n = len(out)
        ptr = out.ctypes.data_as(C.POINTER(C.c_uint8)) if byte else B.fptr(out)
        api = self._api
        name = "analyser_get_%s_%s_data" % ("byte" if byte else "float", "time_domain" if kind == "time" else "frequency")
        fn = getattr(api, name)
        if api.is_product:
            if self._batch is None:
                raise B.WaeError(2, "analyser data is available after rendering")
            api.check(fn(self._batch.handle, self._batch_index, node.id, ptr, n))
        else:
            api.check(fn(self._g, node.id, ptr, n))
        return out


class Batch:
    """wae_batch_prepare / run / fetch: a compiled batch of contexts (product only)."""

    def __init__(self, contexts):
        ctx0 = contexts[0]
        self.api = ctx0._api
        self.contexts = contexts
        self.n = len(contexts)
        self.channels, self.length = ctx0._channels, ctx0._length
        for c in contexts:
            c._run_suspend_callbacks()
        arr = (C.c_void_p * self.n)(*[c._g for c in contexts])
        h = C.c_void_p()
        self.api.check(self.api.batch_prepare(ctx0._backend.engine, arr, self.n, C.byref(h)))
        self.handle = h
        self._backend = ctx0._backend
        self._backend.batches.add(self)
        for i, c in enumerate(contexts):
            c._batch, c._batch_index = self, i

    def run(self):
        self.api.check(self.api.batch_run(self.handle))

    def upload(self):
        self.api.check(self.api.batch_upload(self.handle))

    def set_timing(self, per_stage):
        self.api.check(self.api.batch_set_timing(self.handle, 1 if per_stage else 0))

    def sync(self):
        self.api.check(self.api.batch_sync(self.handle))

    def groups(self):
        """[(first_graph, last_graph)] of the batch's graph groups (rendered one after the other)."""
        n = C.c_uint32()
        self.api.check(self.api.batch_group_count(self.handle, C.byref(n)))
        out = []
        for k in range(n.value):
            a, b = C.c_uint32(), C.c_uint32()
            self.api.check(self.api.batch_group_range(self.handle, k, C.byref(a), C.byref(b)))
            out.append((a.value, b.value))
        return out

    def run_group(self, k):
        """render of ONE graph group, asynchronous on the engine stream (call the groups in order)"""
        self.api.check(self.api.batch_run_group(self.handle, k))

    def run_pipelined(self, host_out_ptr):
        """H2D + render + D2H, overlapped per graph group; `host_out_ptr` = address of [n][ch][length] f32 (pinned)."""
        self.api.check(self.api.batch_run_pipelined(self.handle, host_out_ptr))

    def fetch(self, out=None):
        if out is None:
            out = np.empty((self.n, self.channels, self.length), np.float32)
        self.api.check(self.api.batch_fetch(self.handle, out.ctypes.data_as(C.c_void_p)))
        return out

    def device_ptr(self):
        p, n = C.c_void_p(), C.c_uint64()
        self.api.check(self.api.batch_output_device_ptr(self.handle, C.byref(p), C.byref(n)))
        return p.value, n.value

    def stats(self):
        s = B.BatchStats()
        self.api.check(self.api.batch_get_stats(self.handle, C.byref(s)))
        return s

    def stage_times(self):
        """[(kernel name, ms over the last run, instances)] — needs set_timing(True) before run()."""
        out, i = [], 0
        while True:
            name = C.create_string_buffer(64)
            ms, n = C.c_float(), C.c_uint32()
            if self.api.batch_stage_time(self.handle, i, name, C.byref(ms), C.byref(n)) != 0:
                break
            out.append((name.value.decode(), ms.value, n.value))
            i += 1
        return out

    def destroy(self):
        if self.handle:
            if not self._backend.closed:  # a batch must not outlive its engine
                self.api.batch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def plan_batch(contexts):
    """wae_batch_plan: what the contexts would be lowered to (host only, no GPU) -> dict; raises what prepare would raise."""
    api = contexts[0]._api
    assert api.is_product
    for c in contexts:
        c._run_suspend_callbacks()
    arr = (C.c_void_p * len(contexts))(*[c._g for c in contexts])
    info = B.PlanInfo()
    api.check(api.batch_plan(arr, len(contexts), C.byref(info)))
    kinds = {}
    for part in info.stage_kinds.decode().split(", "):
        if part:
            name, n = part.rsplit(" x ", 1)
            kinds[name] = int(n)
    return {"groups": info.groups, "segments": info.segments, "stages": info.stages, "has_feedback": bool(info.has_feedback),
            "chunk_frames": info.chunk_frames, "chunks": info.chunks, "arena_floats_per_frame": info.arena_floats_per_frame,
            "source_floats": info.source_floats, "kinds": kinds}


def render_batch_oneshot(contexts, out=None):
    """ONE wae_render_batch(engine, graphs, n, out, HOST) call: sizing, planning, H2D of the source PCM, render and D2H overlapped inside
    the library (csrc/wae_engine.cu: render_oneshot_host) — the call the Rust binding makes from start_rendering_sync
    (INTEGRATION.md).  `out`: optional [n][channels][length] float32 array (pageable numpy memory, or the numpy view of a pinned torch
    tensor: the library copies straight into page-locked memory).  Product only."""
    ctx0 = contexts[0]
    api = ctx0._api
    n, ch, length = len(contexts), ctx0._channels, ctx0._length
    for c in contexts:
        c._run_suspend_callbacks()
    if out is None:
        out = np.empty((n, ch, length), np.float32)
    arr = (C.c_void_p * n)(*[c._g for c in contexts])
    api.check(api.render_batch(ctx0._backend.engine, arr, n, out.ctypes.data_as(C.c_void_p), 0))
    return out


def render_batch(contexts, threads=1):
    """Render many OfflineAudioContexts. Returns a list of AudioBuffer (one per context).

    Product: one wae_render_batch call.  Oracle (tests only): wao_render per context (optionally threaded)."""
    ctx0 = contexts[0]
    api = ctx0._api
    n, ch, length = len(contexts), ctx0._channels, ctx0._length
    for c in contexts:
        c._run_suspend_callbacks()
    out = np.empty((n, ch, length), np.float32)
    arr = (C.c_void_p * n)(*[c._g for c in contexts])
    if api.is_product:
        b = Batch(contexts)
        b.run()
        b.sync()
        b.fetch(out)
    else:
        secs = C.c_double()
        api.check(api.render_many(arr, n, B.fptr(out), threads, C.byref(secs)))
    return [AudioBuffer([out[i, c] for c in range(ch)], ctx0._sample_rate) for i in range(n)]
