"""The reference's AudioParamProcessor unit tests (src/param.rs:1765-3300: handle_incoming_event + compute_intrinsic_values on
one processor, block by block) restated as data and run against BOTH param implementations of this repo on the CPU:

* "oracle"  — oracle/wao_param.cpp (the parity checker), through its wao_param_sim_* test hooks;
* "engine"  — the product's own host code: the event folding of csrc/wae_param_host.h and the state machine of
  csrc/wae_param_core.h, which is the code the planner replays at suspend points AND the code the k_param CUDA kernel
  executes (same header, __host__ __device__), through wae_param_sim_* (include/wae.h).  No GPU is involved.

Expected values that the reference computes with powf / exp are recomputed here with numpy; they are compared at 2 f32 ulp
instead of the reference's exact 0 (libm vs numpy), everything else keeps the reference's tolerance."""
import ctypes as C

import numpy as np
import pytest

A, K = 1, 0
F = np.float32


class Sim:
    def __init__(self, api, rate, default, mn, mx, pkg):
        self.api, self.pkg = api, pkg
        self.h = C.c_void_p()
        api.check(api.param_sim_create(rate, default, mn, mx, C.byref(self.h)))
        self.keep = []

    def _push(self, type_, value=0.0, time=0.0, aux=0.0, values=None):
        B = self.pkg._binding
        ev = B.ParamEvent(type_, float(value), float(time), float(aux), None, 0)
        if values is not None:
            arr = np.ascontiguousarray(values, np.float32)
            self.keep.append(arr)
            ev.values, ev.values_len = arr.ctypes.data_as(C.POINTER(C.c_float)), len(arr)
        self.api.check(self.api.param_sim_push(self.h, C.byref(ev)))

    def set_value_at_time(self, v, t):
        self._push(1, v, t)

    def linear(self, v, t):
        self._push(2, v, t)

    def exponential(self, v, t):
        self._push(3, v, t)

    def cancel(self, t):
        self._push(4, 0.0, t)

    def target(self, v, t, tc):
        self._push(5, v, t, tc)

    def cancel_and_hold(self, t):
        self._push(6, 0.0, t)

    def curve(self, values, t, duration):
        self._push(7, 0.0, t, duration, values)

    def set_rate(self, rate):
        self.api.check(self.api.param_sim_set_automation_rate(self.h, rate))

    def run(self, block_time, count=10, dt=1.0):
        out = np.zeros(count, np.float32)
        n = C.c_uint32(0)
        self.api.check(self.api.param_sim_compute(self.h, float(block_time), float(dt), count, out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n)))
        return out[:n.value].copy()

    def close(self):
        self.api.param_sim_destroy(self.h)


@pytest.fixture(params=["oracle", "engine", "engine-walk-serial", "engine-walk-record", "engine-walk-spec"])
def sim(request, pkg, oracle):
    """oracle | the library's param_compute_buffer (what the default kernel and the suspend replay run) | the sink-based walker of
    csrc/wae_param_walk.h with its serial sink, with the recording sink (k_param_parallel), and walked from predicted-then-verified states
    like the default kernel k_param_spec does."""
    import os
    if request.param == "oracle":
        api = oracle.api
    else:
        so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "web-audio-api-rs_b200", "libwae_b200.so")
        if not os.path.exists(so):
            pytest.skip("libwae_b200.so is not built (python -c 'import __graft_entry__ as g; g.build()')")
        api = pkg.api()
    walker = {"engine-walk-serial": 1, "engine-walk-record": 2, "engine-walk-spec": 3}.get(request.param, 0)
    made = []

    def make(rate, default, mn, mx):
        s = Sim(api, rate, default, mn, mx, pkg)
        if walker:
            api.check(api.param_sim_set_walker(s.h, walker))
        made.append(s)
        return s

    yield make
    for s in made:
        s.close()


def eq(got, want, tol=0.0):
    want = np.asarray(want, np.float32)
    assert got.shape == want.shape, (got, want)
    assert np.abs(got.astype(np.float64) - want).max() <= tol, (got, want)


def ulp2(want):
    return 2.4e-7 * max(1.0, float(np.abs(np.asarray(want)).max()))


def exp_ramp(start, end, n, total):
    return [F(start) * np.power(F(end) / F(start), F(t) / F(total), dtype=np.float32) for t in range(n)]


def target_curve(v0, v1, t0, tc, ts):
    return [F(v1) + (F(v0) - F(v1)) * F(np.exp(-((float(t) - t0) / tc))) for t in ts]


def test_steps_a_rate(sim):  # param.rs:1814-1872
    p = sim(A, 0.0, -10.0, 10.0)
    p.set_value_at_time(5.0, 2.0)
    p.set_value_at_time(12.0, 8.0)  # (clamped later, in mix_to_output)
    p.set_value_at_time(8.0, 10.0)
    eq(p.run(0.0), [0, 0, 5, 5, 5, 5, 5, 5, 12, 12])
    eq(p.run(10.0), [8.0])
    p = sim(A, 0.0, -10.0, 10.0)
    p.set_value_at_time(5.0, 2.0)
    p.set_value_at_time(8.0, 12.0)
    eq(p.run(0.0), [0, 0, 5, 5, 5, 5, 5, 5, 5, 5])
    eq(p.run(10.0), [5, 5, 8, 8, 8, 8, 8, 8, 8, 8])


def test_steps_k_rate(sim):  # :1874-1899
    p = sim(K, 0.0, -10.0, 10.0)
    for v, t in [(5.0, 2.0), (12.0, 8.0), (8.0, 10.0), (3.0, 14.0)]:
        p.set_value_at_time(v, t)
    eq(p.run(0.0), [0.0])
    eq(p.run(10.0), [8.0])
    eq(p.run(20.0), [3.0])


def test_linear_ramps_a_rate(sim):  # :1901-2033 linear_ramp_arate, _end_of_block, _implicit_set_value, _multiple_blocks
    p = sim(A, 0.0, -10.0, 10.0)
    p.set_value_at_time(5.0, 2.0)
    p.linear(8.0, 5.0)
    p.linear(0.0, 13.0)
    eq(p.run(0.0), [0, 0, 5, 6, 7, 8, 7, 6, 5, 4])
    p = sim(A, 0.0, -10.0, 10.0)
    p.set_value_at_time(0.0, 0.0)
    p.linear(9.0, 9.0)
    eq(p.run(0.0), [0, 1, 2, 3, 4, 5, 6, 7, 8, 9])
    p = sim(A, 0.0, -10.0, 10.0)
    eq(p.run(0.0), [0.0])
    p.linear(10.0, 20.0)  # arrives while rendering: ramps from the last event (time 0 of the implicit start) ... :1960-1992
    eq(p.run(10.0), [0, 1, 2, 3, 4, 5, 6, 7, 8, 9])
    eq(p.run(20.0), [10.0] * 10)
    p = sim(A, 0.0, -20.0, 20.0)
    p.linear(20.0, 20.0)
    eq(p.run(0.0), [0, 1, 2, 3, 4, 5, 6, 7, 8, 9])
    eq(p.run(10.0), [10, 11, 12, 13, 14, 15, 16, 17, 18, 19])
    eq(p.run(20.0), [20.0] * 10)


def test_linear_ramp_k_rate_and_start_time(sim):  # :2035-2128
    for end, last in [(20.0, 20.0), (15.0, 15.0)]:
        p = sim(K, 0.0, -20.0, 20.0)
        p.linear(end, end)
        eq(p.run(0.0), [0.0])
        eq(p.run(10.0), [10.0])
        eq(p.run(20.0), [last])
    p = sim(A, 0.0, -10.0, 10.0)
    p.set_value_at_time(1.0, 0.0)
    p.linear(-1.0, 10.0)
    eq(p.run(0.0), [1, 0.8, 0.6, 0.4, 0.2, 0, -0.2, -0.4, -0.6, -0.8], 1e-7)
    eq(p.run(10.0), [-1.0] * 10)
    p.linear(1.0, 30.0)  # starts from the previous event (t = 10), not from "now"
    eq(p.run(20.0), [0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9], 1e-7)


def test_exponential_ramps(sim):  # :2130-2400
    p = sim(A, 0.0, 0.0, 1.0)
    p.set_value_at_time(0.0001, 0.0)
    p.exponential(1.0, 10.0)
    want = exp_ramp(0.0001, 1.0, 10, 10)
    eq(p.run(0.0), want, ulp2(want))
    eq(p.run(10.0), [1.0] * 10)
    p = sim(A, 0.0, 0.0, 1.0)
    p.set_value_at_time(0.0001, 3.0)
    p.exponential(1.0, 13.0)
    res = [0.0] * 3 + exp_ramp(0.0001, 1.0, 10, 10) + [1.0] * 7
    eq(p.run(0.0), res[:10], ulp2(res))
    eq(p.run(10.0), res[10:], ulp2(res))
    # zero start value / opposite sign: the ramp degenerates into a step at its end time
    p = sim(A, 0.0, 0.0, 1.0)
    p.set_value_at_time(0.0, 0.0)
    p.exponential(1.0, 5.0)
    eq(p.run(0.0), [0, 0, 0, 0, 0, 1, 1, 1, 1, 1])
    p = sim(A, 0.0, -1.0, 1.0)
    p.set_value_at_time(-1.0, 0.0)
    p.exponential(1.0, 5.0)
    eq(p.run(0.0), [-1, -1, -1, -1, -1, 1, 1, 1, 1, 1])
    # k-rate
    p = sim(K, 0.0, 0.0, 1.0)
    p.set_value_at_time(0.0001, 3.0)
    p.exponential(1.0, 13.0)
    eq(p.run(0.0), [res[0]], ulp2(res))
    eq(p.run(10.0), [res[10]], ulp2(res))
    eq(p.run(20.0), [1.0])
    for default in (0.0, -1.0):
        p = sim(K, default, -1.0, 1.0)
        p.exponential(1.0, 5.0)
        eq(p.run(0.0), [default])
        eq(p.run(10.0), [1.0])
    # an exponential ramp pushed after a finished linear one starts from that one's end (:2362-2400)
    p = sim(A, 0.0, -10.0, 10.0)
    p.set_value_at_time(0.0, 0.0)
    p.linear(1.0, 10.0)
    eq(p.run(0.0), [0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9], 1e-7)
    eq(p.run(10.0), [1.0] * 10)
    p.exponential(0.0001, 30.0)
    eq(p.run(20.0), exp_ramp(1.0, 0.0001, 20, 20)[10:], 1e-7)


def test_exponential_ramp_to_zero_is_rejected(sim, pkg):  # :2262-2272 #[should_panic]
    p = sim(A, 1.0, 0.0, 1.0)
    with pytest.raises(pkg.WaeError):
        p.exponential(0.0, 10.0)


def test_set_target(sim):  # :2402-2775
    for pre in (True, False):  # with an explicit SetValueAtTime before it, or the implicit one
        p = sim(A, 0.0, 0.0, 1.0)
        if pre:
            p.set_value_at_time(0.0, 0.0)
        p.target(1.0, 0.0, 1.0)
        want = target_curve(0.0, 1.0, 0.0, 1.0, range(10))
        eq(p.run(0.0), want, ulp2(want))
    p = sim(A, 0.0, 0.0, 100.0)
    p.set_value_at_time(1.0, 1.0)
    p.target(42.0, 1.0, 2.1)
    want = target_curve(1.0, 42.0, 1.0, 2.1, range(10))
    want[0] = 0.0
    eq(p.run(0.0), want, ulp2(want))
    p = sim(A, 0.0, 0.0, 100.0)
    p.target(1.0, 1.0, 0.0)  # zero time constant == set_value_at_time
    eq(p.run(0.0), [0.0] + [1.0] * 9)
    # several blocks, then followed by a SetValueAtTime
    p = sim(A, 0.0, 0.0, 2.0)
    p.set_value_at_time(0.0, 0.0)
    p.target(2.0, 0.0, 1.0)
    res = target_curve(0.0, 2.0, 0.0, 1.0, range(20))
    eq(p.run(0.0), res[:10], ulp2(res))
    eq(p.run(10.0), res[10:], ulp2(res))
    p = sim(A, 0.0, 0.0, 2.0)
    p.set_value_at_time(0.0, 0.0)
    p.target(2.0, 0.0, 1.0)
    p.set_value_at_time(0.5, 15.0)
    res = target_curve(0.0, 2.0, 0.0, 1.0, range(15)) + [0.5] * 5
    eq(p.run(0.0), res[:10], ulp2(res))
    eq(p.run(10.0), res[10:], ulp2(res))
    # ends at the threshold: no subnormals, then exactly the target (:2589-2619)
    p = sim(A, 0.0, 0.0, 2.0)
    p.set_value_at_time(1.0, 0.0)
    p.target(0.0, 1.0, 0.2)
    vs = p.run(0.0, 128)
    assert np.all((vs == 0.0) | (np.abs(vs) >= np.finfo(np.float32).tiny))
    eq(p.run(10.0, 128), [0.0] * 128)
    # waits for its start time (:2621-2643)
    p = sim(A, 0.0, 0.0, 2.0)
    p.set_value_at_time(1.0, 0.0)
    p.target(0.0, 5.0, 1.0)
    assert np.all(p.run(0.0)[:6] == 1.0)
    # followed by a ramp that starts from the value the SetTarget reached (:2645-2697)
    p = sim(A, 0.0, 0.0, 10.0)
    p.set_value_at_time(0.0, 0.0)
    p.target(2.0, 0.0, 10.0)
    res = target_curve(0.0, 2.0, 0.0, 10.0, range(11))
    eq(p.run(0.0), res[:10], ulp2(res))
    v0 = res[10]
    p.linear(10.0, 20.0)
    ramp = [v0 + (F(10.0) - v0) * F(t - 10.0) / F(10.0) for t in range(10, 20)]
    eq(p.run(10.0), ramp, 1e-6)
    eq(p.run(20.0), [10.0] * 10)
    # k-rate
    p = sim(K, 0.0, 0.0, 2.0)
    p.set_value_at_time(0.0, 0.0)
    p.target(2.0, 0.0, 1.0)
    res = target_curve(0.0, 2.0, 0.0, 1.0, range(20))
    eq(p.run(0.0), [res[0]], ulp2(res))
    eq(p.run(10.0), [res[10]], ulp2(res))
    # snaps to the target once close enough (:2731-2775)
    p = sim(A, 0.0, 0.0, 1.0)
    p.set_value_at_time(1.0, 0.0)
    p.target(0.0, 0.0, 1.0)
    res = target_curve(1.0, 0.0, 0.0, 1.0, range(30))
    eq(p.run(0.0), res[:10], ulp2(res))
    eq(p.run(10.0), res[10:20], ulp2(res))
    eq(p.run(20.0), res[20:30], ulp2(res))
    eq(p.run(30.0), [0.0] * 10)


def test_cancel_scheduled_values(sim):  # :2777-2902
    p = sim(A, 0.0, 0.0, 10.0)
    for t in range(10):
        p.set_value_at_time(float(t), float(t))
    p.cancel(5.0)
    eq(p.run(0.0), [0, 1, 2, 3, 4, 4, 4, 4, 4, 4])
    p = sim(A, 0.0, 0.0, 10.0)
    p.set_value_at_time(0.0, 0.0)
    p.linear(10.0, 10.0)
    p.cancel(10.0)
    eq(p.run(0.0), [0.0] * 10)
    for explicit in (True, False):  # cancelling a ramp that is under way restores the value from before the ramp
        p = sim(A, 0.0, 0.0, 20.0)
        if explicit:
            p.set_value_at_time(0.0, 0.0)
        p.linear(20.0, 20.0)
        eq(p.run(0.0), [0, 1, 2, 3, 4, 5, 6, 7, 8, 9])
        p.cancel(10.0)
        eq(p.run(10.0), [0.0])
    p = sim(A, 0.0, 0.0, 10.0)
    p.linear(10.0, 10.0)
    p.cancel(10.0)
    eq(p.run(0.0), [0.0] * 10)


def test_cancel_and_hold(sim):  # :2904-3142
    p = sim(A, 0.0, 0.0, 10.0)
    for t in (1.0, 2.0, 3.0, 4.0):
        p.set_value_at_time(t, t)
    p.cancel_and_hold(2.5)
    eq(p.run(0.0), [0, 1, 2, 2, 2, 2, 2, 2, 2, 2])
    p = sim(A, 0.0, 0.0, 2.0)
    p.set_value_at_time(0.0, 0.0)
    p.target(2.0, 0.0, 1.0)
    p.cancel_and_hold(15.0)
    res = target_curve(0.0, 2.0, 0.0, 1.0, range(16))
    res = res[:15] + [res[15]] * 5
    eq(p.run(0.0), res[:10], ulp2(res))
    eq(p.run(10.0), res[10:], ulp2(res))
    for hold, want in [(5.0, [0, 1, 2, 3, 4, 5, 5, 5, 5, 5]), (4.5, [0, 1, 2, 3, 4, 4.5, 4.5, 4.5, 4.5, 4.5])]:
        p = sim(A, 0.0, 0.0, 10.0)
        p.linear(10.0, 10.0)
        p.cancel_and_hold(hold)
        eq(p.run(0.0), want)
    ramp = exp_ramp(0.0001, 1.0, 6, 10)
    for hold, want in [(5.0, ramp[:5] + [ramp[5]] * 5),
                       (4.5, ramp[:5] + [F(0.0001) * np.power(F(1.0) / F(0.0001), F(4.5) / F(10.0), dtype=np.float32)] * 5)]:
        p = sim(A, 0.0, 0.0, 10.0)
        p.set_value_at_time(0.0001, 0.0)
        p.exponential(1.0, 10.0)
        p.cancel_and_hold(hold)
        eq(p.run(0.0), want, ulp2(want))
    for hold, want in [(5.0, [0, 0.2, 0.4, 0.6, 0.8, 1, 1, 1, 1, 1]), (4.5, [0, 0.2, 0.4, 0.6, 0.8, 0.9, 0.9, 0.9, 0.9, 0.9])]:
        p = sim(A, 0.0, 0.0, 2.0)
        p.curve([0.0, 0.5, 1.0, 0.5, 0.0], 0.0, 10.0)
        p.cancel_and_hold(hold)
        eq(p.run(0.0), want, 1e-7)


def test_set_value_curve(sim, pkg):  # :3144-3275
    p = sim(A, 0.0, 0.0, 10.0)
    p.curve([0.0, 0.5, 1.0, 0.5, 0.0], 0.0, 10.0)
    eq(p.run(0.0), [0, 0.2, 0.4, 0.6, 0.8, 1, 0.8, 0.6, 0.4, 0.2], 1e-7)
    eq(p.run(10.0), [0.0] * 10)
    p = sim(A, 0.0, 0.0, 10.0)
    p.curve([0.0, 0.5, 1.0, 0.5, 0.0], 0.0, 20.0)
    eq(p.run(0.0), [0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9], 1e-7)
    eq(p.run(10.0), [1, 0.9, 0.8, 0.7, 0.6, 0.5, 0.4, 0.3, 0.2, 0.1], 1e-7)
    eq(p.run(20.0), [0.0] * 10)
    p = sim(A, 0.0, 0.0, 10.0)
    p.curve([0.0, 0.5, 1.0, 0.5, 0.0], 5.0, 10.0)
    eq(p.run(0.0), [0, 0, 0, 0, 0, 0, 0.2, 0.4, 0.6, 0.8])
    # overlapping another event is rejected, whichever comes first (#[should_panic] :3203-3250)
    p = sim(A, 1.0, 0.0, 1.0)
    p.set_value_at_time(0.0, 5.0)
    with pytest.raises(pkg.WaeError):
        p.curve([0.0, 0.5, 1.0, 0.5, 0.0], 0.0, 10.0)
        p.run(0.0)
    p = sim(A, 1.0, 0.0, 1.0)
    p.curve([0.0, 0.5, 1.0, 0.5, 0.0], 0.0, 10.0)
    with pytest.raises(pkg.WaeError):
        p.set_value_at_time(0.0, 5.0)
        p.run(0.0)


def test_update_automation_rate(sim):  # :3277-3314
    p = sim(A, 0.0, -10.0, 10.0)
    p.set_rate(K)
    p.set_value_at_time(2.0, 0.000001)
    eq(p.run(0.0), [0.0])
    p = sim(K, 0.0, -10.0, 10.0)
    p.set_rate(A)
    p.set_value_at_time(2.0, 0.000001)
    eq(p.run(0.0), [2.0] * 10)


def test_varying_param_size(sim):  # :3316-3392 — the block is single-valued exactly when no event touches it
    for online in (True, False):
        p = sim(A, 0.0, 0.0, 10.0)
        p.set_value_at_time(0.0, 0.0)
        p.linear(9.0, 9.0)
        if not online:
            p.set_value_at_time(1.0, 25.0)
        eq(p.run(0.0), [0, 1, 2, 3, 4, 5, 6, 7, 8, 9])
        eq(p.run(10.0), [9.0])
        if online:
            p.set_value_at_time(1.0, 25.0)
        eq(p.run(20.0), [9, 9, 9, 9, 9, 1, 1, 1, 1, 1])
        eq(p.run(30.0), [1.0])


def test_full_block_ramp_from_set_value(sim):  # :3503-3528 (the intrinsic half of test_full_render_chain)
    p = sim(A, 2.0, 2.0, 42.0)
    p._push(0, 128.0)  # SetValue: stored unclamped, clamped only in mix_to_output
    p.linear(0.0, 128.0)
    eq(p.run(0.0, 128), 128.0 - np.arange(128, dtype=np.float32))


@pytest.mark.parametrize("seed", range(60))
def test_the_walker_is_bit_identical_to_the_kernel_code_on_random_timelines(pkg, seed):
    """csrc/wae_param_walk.h (serial and recording sinks, and the speculative walk) vs csrc/wae_param_core.h::param_compute_buffer on random event timelines, 128-frame
    quanta at 48 kHz, events pushed before and during the render: every block, bit for bit (the oracle is compared at f32 resolution)."""
    import os
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "web-audio-api-rs_b200", "libwae_b200.so")
    if not os.path.exists(so):
        pytest.skip("libwae_b200.so is not built")
    api = pkg.api()
    rng = np.random.default_rng(seed)
    sims = []
    for walker in (0, 1, 2, 3):
        s = Sim(api, A if seed % 3 else K, 0.5, -10.0, 10.0, pkg)
        api.check(api.param_sim_set_walker(s.h, walker))
        sims.append(s)
    dt, t_end = 1.0 / 48000.0, 60 * 128 / 48000.0

    # identical argument streams: draw once, apply to all three
    def push_same(t_from):
        kind = int(rng.integers(7))
        t = float(rng.uniform(t_from, t_end))
        v = float(rng.uniform(0.05, 2.0))
        tc = float(rng.integers(1, 40)) * 1e-4
        curve = rng.uniform(0.0, 1.0, 5).astype(np.float32)
        dur = float(rng.uniform(1e-3, 6e-3))
        for s in sims:
            try:
                [lambda: s.set_value_at_time(v, t), lambda: s.linear(v, t), lambda: s.exponential(v, t), lambda: s.target(v, t, tc),
                 lambda: s.curve(curve, t, dur), lambda: s.cancel_and_hold(t), lambda: s.cancel(t)][kind]()
            except pkg.WaeError:
                pass

    for _ in range(int(rng.integers(3, 9))):
        push_same(0.0)
    for q in range(60):
        if rng.random() < 0.15:
            push_same(q * 128 * dt)
        outs = []
        for s in sims:
            try:
                outs.append(s.run(q * 128 * dt, 128, dt))
            except pkg.WaeError:
                outs.append(None)
        if outs[0] is None:
            assert all(o is None for o in outs[1:])
            break
        assert all(o is not None for o in outs[1:])
        assert all(np.array_equal(outs[0], o, equal_nan=True) for o in outs[1:]), (seed, q)
    for s in sims:
        s.close()


def test_speculation_predicts_the_state_inside_every_kind_of_event(pkg):
    """k_param_spec walks 32 quanta at once from PREDICTED states; a prediction that fails costs a re-walk, so inside a ramp, a set-target,
    a value curve and a stretch without events nearly all of them have to hold (the host simulator counts them: a window restarts after 32
    quanta or after a miss, so at most 31 of 32 are predictions)."""
    import os
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "web-audio-api-rs_b200", "libwae_b200.so")
    if not os.path.exists(so):
        pytest.skip("libwae_b200.so is not built")
    api = pkg.api()
    dt = 1.0 / 48000.0
    quanta = 3000  # 8 s

    def run(setup, rate=A):
        ref = Sim(api, rate, 0.5, -1000.0, 1000.0, pkg)
        spec = Sim(api, rate, 0.5, -1000.0, 1000.0, pkg)
        api.check(api.param_sim_set_walker(spec.h, 3))
        setup(ref)
        setup(spec)
        for q in range(quanta):
            a, b = ref.run(q * 128 * dt, 128, dt), spec.run(q * 128 * dt, 128, dt)
            assert np.array_equal(a, b, equal_nan=True), q
        tried, hits = C.c_uint64(0), C.c_uint64(0)
        api.check(api.param_sim_speculation(spec.h, C.byref(tried), C.byref(hits)))
        ref.close()
        spec.close()
        return tried.value, hits.value

    def envelope(p):   # examples/benchmarks.rs "Substractive Synth": a set-target restarted every 146 ms
        t = 0.0
        while t < quanta * 128 * dt:
            p.set_value_at_time(1.0, t)
            p.target(0.0, t, 0.1)
            t += 140.0 / 60.0 / 16.0

    cases = {
        "no events": lambda p: None,
        "events far away": lambda p: (p.set_value_at_time(0.2, 1.0), p.set_value_at_time(0.7, 6.5)),
        "linear ramp": lambda p: (p.set_value_at_time(100.0, 0.0), p.linear(2000.0, 7.5)),
        "exponential ramp": lambda p: (p.set_value_at_time(100.0, 0.0), p.exponential(2000.0, 7.0)),
        "set target": lambda p: p.target(3.0, 0.25, 1.5),
        "value curve": lambda p: p.curve(np.linspace(0.0, 1.0, 64).astype(np.float32), 0.1, 7.0),
        "envelope": envelope,
    }
    for name, setup in cases.items():
        for rate in (A, K):
            tried, hits = run(setup, rate)
            assert tried >= quanta * 29 // 32, (name, tried)
            floor = 0.80 if name == "envelope" else 0.97   # (55 events in 8 s: each costs one miss and a short window)
            assert hits >= floor * tried, (name, rate, tried, hits)


def test_a_pending_value_curve_is_sampled_before_its_start(sim):
    # Found by the fuzz above.  When an event ends inside a block and the next one is a value curve that starts LATER, compute_buffer
    # still visits the curve branch and leaves compute_set_value_curve_sample(next_block_time) in the intrinsic value (param.rs:1429-1498):
    # the position on the curve is negative there, Rust's `as usize` saturates it to segment 0 and the phase is the fractional part of the
    # negative position.  (A C cast would index before the curve.)
    p = sim(K, 0.0, -10.0, 10.0)
    p.set_value_at_time(1.0, 0.0)
    p.linear(2.0, 5.0)                                   # ends inside the first block of 10 frames
    values = np.array([0.25, 0.75, 0.5], np.float32)
    p.curve(values, 25.0, 10.0)                          # starts two blocks later
    eq(p.run(0.0), [0.0])                                # k-rate: the value the block started with
    position = (len(values) - 1) * (10.0 - 25.0) / 10.0  # at next_block_time = 10
    phase = F(position - np.floor(position))
    want = F(values[1] - values[0]) * phase + values[0]
    eq(p.run(10.0), [want], 1e-7)                        # the k-rate value of the second block is what the first block left behind
    eq(p.run(20.0), [F(values[1] - values[0]) * F((2 * (20.0 - 25.0) / 10.0) - np.floor(2 * (20.0 - 25.0) / 10.0)) + values[0]], 1e-7)


@pytest.mark.parametrize("seed", range(100, 160))
def test_library_and_oracle_agree_on_random_timelines(pkg, oracle, seed):
    """The two restatements of AudioParamProcessor (oracle/wao_param.cpp and the library's host folding + state machine) on random event
    timelines: events pushed before and during the render, a-rate and k-rate, 128-frame blocks at 48 kHz.  Same libm on the CPU: the
    values must agree to the last bit, and both must refuse the same pushes."""
    import os
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "web-audio-api-rs_b200", "libwae_b200.so")
    if not os.path.exists(so):
        pytest.skip("libwae_b200.so is not built")
    rng = np.random.default_rng(seed)
    rate = A if seed % 4 else K
    sims = [Sim(oracle.api, rate, 0.5, -10.0, 10.0, pkg), Sim(pkg.api(), rate, 0.5, -10.0, 10.0, pkg)]
    dt, t_end = 1.0 / 48000.0, 50 * 128 / 48000.0

    def push(t_from):
        kind = int(rng.integers(7))
        t = float(rng.uniform(t_from, t_end))
        v = float(rng.uniform(-1.5, 2.0)) if kind != 2 else float(rng.uniform(0.05, 2.0))
        tc = float(rng.integers(0, 40)) * 1e-4
        curve = rng.uniform(-1.0, 1.0, int(rng.integers(2, 7))).astype(np.float32)
        dur = float(rng.uniform(1e-3, 6e-3))
        res = []
        for s in sims:
            try:
                [lambda: s.set_value_at_time(v, t), lambda: s.linear(v, t), lambda: s.exponential(v, t), lambda: s.target(v, t, tc),
                 lambda: s.curve(curve, t, dur), lambda: s.cancel_and_hold(t), lambda: s.cancel(t)][kind]()
                res.append(True)
            except pkg.WaeError:
                res.append(False)
        return res

    def refused_by_one(res, q):
        # an event the reference panics on in the render thread (a curve overlapping another event ...): one implementation reports it
        # at the push, the other when the next block is computed — both must refuse it
        if res[0] == res[1]:
            return False
        late = sims[0] if res[0] else sims[1]
        with pytest.raises(pkg.WaeError):
            late.run(q * 128 * dt, 128, dt)
        return True

    stop = False
    for _ in range(int(rng.integers(2, 9))):
        if refused_by_one(push(0.0), 0):
            stop = True
            break
    for q in range(0 if stop else 50):
        if rng.random() < 0.2 and refused_by_one(push(q * 128 * dt), q):
            break
        outs = []
        for s in sims:
            try:
                outs.append(s.run(q * 128 * dt, 128, dt))
            except pkg.WaeError:
                outs.append(None)
        if outs[0] is None or outs[1] is None:
            assert outs[0] is None and outs[1] is None, (seed, q)
            break
        assert outs[0].shape == outs[1].shape and np.array_equal(outs[0], outs[1], equal_nan=True), (seed, q, outs[0][:4], outs[1][:4])
    for s in sims:
        s.close()
