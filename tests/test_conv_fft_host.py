"""The convolver's 8192-point shared-memory transforms, run on the HOST through wae_selftest_conv_fft: the library compiles the same
butterfly / index / twiddle functions for both sides (`__host__ __device__`), so a wrong rotation, span or bit-reversed position shows up
here, without a GPU.  Forward = decimation in frequency (natural in, bit-reversed "position" order out), inverse = decimation in time
(position order in, natural out); the real-FFT split pairs position p with the position of the mirror bin."""
import ctypes as C
import os

import numpy as np
import pytest

B = 8192
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def brev(k, bits=13):
    r = 0
    for i in range(bits):
        r |= ((k >> i) & 1) << (bits - 1 - i)
    return r


BR = np.array([brev(k) for k in range(B)])


@pytest.fixture(scope="module")
def run(pkg):
    so = os.path.join(ROOT, "web-audio-api-rs_b200", "libwae_b200.so")
    if not os.path.exists(so):
        pytest.skip("libwae_b200.so is not built")
    api = pkg.api()

    def call(data, mode):
        buf = np.ascontiguousarray(data, dtype=np.float32).copy()
        assert buf.size == 2 * B
        assert api.lib.wae_selftest_conv_fft(buf.ctypes.data_as(C.POINTER(C.c_float)), mode) == 0
        return buf
    return call


def test_complex_forward_is_the_dft_in_bit_reversed_positions(run):
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(B) + 1j * rng.standard_normal(B)).astype(np.complex64)
    got = run(x.view(np.float32), 0).view(np.complex64)
    want = np.fft.fft(x.astype(np.complex128))
    scale = np.abs(want).max()
    assert np.abs(got[BR] - want).max() <= 2e-6 * scale  # element k lives at position brev13(k)


def test_complex_inverse_takes_positions_back_to_natural_order(run):
    rng = np.random.default_rng(2)
    x = (rng.standard_normal(B) + 1j * rng.standard_normal(B)).astype(np.complex64)
    spec = np.fft.fft(x.astype(np.complex128))
    pos = np.empty(B, np.complex64)
    pos[BR] = spec.astype(np.complex64)
    got = run(pos.view(np.float32), 1).view(np.complex64)
    assert np.abs(got / B - x).max() <= 2e-6 * np.abs(x).max() * 4


def test_real_forward_gives_the_packed_half_spectrum(run):
    rng = np.random.default_rng(3)
    r = rng.uniform(-1, 1, 2 * B).astype(np.float32)
    got = run(r, 2).view(np.complex64)
    want = np.fft.rfft(r.astype(np.float64))  # B + 1 bins
    scale = np.abs(want).max()
    assert abs(got[0].real - want[0].real) <= 2e-6 * scale and abs(got[0].imag - want[B].real) <= 2e-6 * scale  # (DC, Nyquist)
    k = np.arange(1, B)
    assert np.abs(got[BR[k]] - want[k]).max() <= 2e-6 * scale


def test_real_round_trip_and_a_convolution_through_the_packed_spectra(run):
    rng = np.random.default_rng(4)
    r = rng.uniform(-1, 1, 2 * B).astype(np.float32)
    back = run(run(r, 2), 3)
    assert np.abs(back - r).max() <= 2e-6
    # overlap-save block: [previous | current] frame times [h | 0]; the second half of the product's inverse is the linear convolution
    h = np.zeros(2 * B, np.float32)
    h[:B] = (rng.uniform(-1, 1, B) * np.exp(-np.arange(B) / 900.0)).astype(np.float32)
    X = run(r, 2).view(np.complex64)
    H = run(h, 2).view(np.complex64)
    Y = X * H
    Y[0] = complex(X[0].real * H[0].real, X[0].imag * H[0].imag)  # packed bin: DC and Nyquist multiply component-wise
    y = run(Y.view(np.float32), 3)
    want = np.convolve(r.astype(np.float64), h[:B].astype(np.float64))[B:2 * B]
    assert np.abs(y[B:] - want).max() <= 1e-5 * max(1.0, np.abs(want).max())


def test_mode_is_validated(pkg):
    so = os.path.join(ROOT, "web-audio-api-rs_b200", "libwae_b200.so")
    if not os.path.exists(so):
        pytest.skip("libwae_b200.so is not built")
    api = pkg.api()
    buf = np.zeros(2 * B, np.float32)
    assert api.lib.wae_selftest_conv_fft(buf.ctypes.data_as(C.POINTER(C.c_float)), 4) != 0
    assert api.lib.wae_selftest_conv_fft(None, 0) != 0
