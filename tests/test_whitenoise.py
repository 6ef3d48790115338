"""White noise on the CPU side (no GPU needed): the oracle is pinned to the reference, and the
host build of the kernel's column routines (tests/harness) is bit-identical to the oracle.

1. oracle == golden fields written by the compiled reference (tests/golden/whitenoise_golden.npz);
2. oracle == the compiled reference live (oracle/_ref/pmesh_ref/_whitenoise), fresh cases,
   including the full-spectrum fill;
3. the reference's own tests restated (pmesh/tests/test_whitenoise.py:6-60): N-GenIC values,
   std 1/sqrt(2), partition invariance, Hermitian symmetry;
4. pmb_wnrng.h compiled for the host == oracle, bit for bit (stream and fields).
"""
import ctypes
import os
import sys

import numpy
import pytest
from numpy.testing import assert_array_equal, assert_allclose

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G  # noqa: E402


def bits(a):
    return a.view(a.real.dtype)


def half(N):
    return (N[0], N[1], N[2] // 2 + 1)


def test_oracle_against_reference_golden(oracle):
    z = numpy.load(os.path.join(HERE, "golden", "whitenoise_golden.npz"))
    for ci, (N, seed, unitary, dt, start, shape) in enumerate(G.WHITENOISE_CASES):
        want = z["wn_%d" % ci]
        got = oracle.whitenoise(numpy.zeros(want.shape, dtype=dt), start, N, seed, unitary)
        assert got.dtype == want.dtype
        assert_array_equal(bits(got), bits(want), err_msg=str(G.WHITENOISE_CASES[ci]))


@pytest.fixture(scope="module")
def refwn():
    import build_ref
    if not build_ref.have_ref():
        try:
            build_ref.build()
        except Exception:
            pass
    wn = build_ref.load_whitenoise() if build_ref.have_ref() else None
    if wn is None:
        pytest.skip("oracle/_ref not available")
    return wn


def test_oracle_against_compiled_reference(oracle, refwn):
    rng = numpy.random.default_rng(5)
    for trial in range(12):
        N = tuple(int(n) for n in rng.integers(2, 14, 3))
        if trial % 3 == 0:
            N = (N[0], N[0], N[0])
        seed = int(rng.integers(0, 2 ** 32))
        unitary = bool(trial % 2)
        dt = "complex64" if trial % 4 == 1 else "complex128"
        for full in (False, True):
            shape = N if full else half(N)
            start = [0, 0, 0]
            if trial % 5 == 2:      # a random block
                start = [int(rng.integers(0, n)) for n in shape]
                shape = tuple(int(rng.integers(1, n - s + 1)) for n, s in zip(shape, start))
            want = G.whitenoise_reference(refwn, (N, seed, unitary, dt, start, shape))
            got = oracle.whitenoise(numpy.zeros(shape, dtype=dt), start, N, seed, unitary)
            assert_array_equal(bits(got), bits(want), err_msg="%s seed=%d full=%s" % (N, seed, full))
    # strided canvas (the transposed layout): memory order (1, 2, 0)
    N = (6, 8, 10)
    store = numpy.zeros((8, 6, 6), dtype="complex128")
    view = store.transpose(2, 0, 1)
    assert view.shape == half(N)
    oracle.whitenoise(view, 0, N, 42, False)
    want = G.whitenoise_reference(refwn, (N, 42, False, "complex128", (0, 0, 0), None))
    assert_array_equal(view, want)


def test_reference_known_answers(oracle):
    # pmesh/tests/test_whitenoise.py:27-38 -- values from N-GenIC (Illustris seed)
    v = oracle.whitenoise(numpy.zeros((4, 4, 3), dtype="complex128"), 0, (4, 4, 4), 5463, False)
    assert_allclose(v[0, 1, 0], (-0.04 - 0.03j), atol=0.02)
    assert_allclose(v[1, 0, 0], (0.36 - 0.78j), atol=0.02)
    assert_allclose(v[1, 1, 0], (-0.43 + 0.33j), atol=0.02)
    assert_allclose(v[1, 1, 1], (-1.65 - 0.64j), atol=0.02)
    # :40-63 -- Hermitian: the field is the rfftn of a real field
    h = numpy.fft.rfftn(numpy.fft.irfftn(v.copy(), s=(4, 4, 4), axes=(0, 1, 2)))
    assert_array_equal(v[1, 1, 0], v[3, 3, 0].conjugate())
    assert_array_equal(v[1, 1, 2], v[3, 3, 2].conjugate())
    assert_allclose(h, v, rtol=1e-5, atol=1e-9)
    # :6-25 -- unit variance split between real and imaginary parts; a block equals the same block of the whole
    N = 64
    v = oracle.whitenoise(numpy.zeros((N, N, N // 2 + 1), dtype="complex128"), 0, (N, N, N), 1, False)
    assert_allclose(v.real.std(), 0.5 ** 0.5, rtol=2e-2)
    assert_allclose(v.imag.std(), 0.5 ** 0.5, rtol=2e-2)
    piece = oracle.whitenoise(numpy.zeros((32, 4, 4), dtype="complex128"), [2, 2, 2], (N, N, N), 1, False)
    assert_array_equal(piece, v[2:34, 2:6, 2:6])
    # :65-83 -- the full-spectrum fill is the Hermitian completion of the half fill
    N = 8
    full = oracle.whitenoise(numpy.zeros((N, N, N), dtype="complex128"), 0, (N, N, N), 1, False)
    hf = oracle.whitenoise(numpy.zeros((N, N, N // 2 + 1), dtype="complex128"), 0, (N, N, N), 1, False)
    c1 = numpy.fft.ifftn(full)
    assert_allclose(c1.imag, 0, atol=1e-9)
    assert_allclose(c1, numpy.fft.irfftn(hf, s=(N, N, N), axes=(0, 1, 2)))
    # unitary: |mode| == 1 everywhere except the mean
    u = oracle.whitenoise(numpy.zeros((N, N, N // 2 + 1), dtype="complex128"), 0, (N, N, N), 7, True)
    a = abs(u)
    assert a[0, 0, 0] == 0
    a[0, 0, 0] = 1
    assert_allclose(a, 1.0, rtol=1e-14)


def _hh_fill(h, shape, dt, start, N, seed, unitary, order=None):
    A = ctypes.c_int64 * 3
    if order is None:
        got = numpy.zeros(shape, dtype=dt)
    else:   # canvas stored with the axes permuted: got is a strided view
        store = numpy.zeros([shape[o] for o in order], dtype=dt)
        got = store.transpose(numpy.argsort(order))
        assert got.shape == tuple(shape)
    rc = h.hh_whitenoise(ctypes.c_void_p(got.ctypes.data), ctypes.c_int(got.dtype.itemsize), A(*N), A(*start), A(*shape),
                         A(*got.strides), ctypes.c_uint(seed & 0xffffffff), ctypes.c_int(int(unitary)))
    assert rc == 0
    return got


def test_host_build_of_the_kernel_routines_is_bit_exact(harness, oracle):
    for seed in (1, 2, 5463, 0, 2 ** 31 - 1, 123456789):
        got = numpy.empty(1000)
        harness.hh_wn_stream(ctypes.c_uint(seed), got.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(1000))
        want = oracle.whitenoise_stream(seed, 1000)
        assert_array_equal(got, want)
        assert (want >= 0).all() and (want < 1).all()
    for ci, (N, seed, unitary, dt, start, shape) in enumerate(G.WHITENOISE_CASES):
        if shape is None:
            shape = half(N)
        want = oracle.whitenoise(numpy.zeros(shape, dtype=dt), start, N, seed, unitary)
        got = _hh_fill(harness, shape, dt, start, N, seed, unitary)
        assert_array_equal(bits(got), bits(want), err_msg=str(G.WHITENOISE_CASES[ci]))
        got = _hh_fill(harness, shape, dt, start, N, seed, unitary, order=(1, 2, 0))
        assert_array_equal(got, want)
