"""White noise on the GPU (pmb_whitenoise through pmesh_b200.whitenoise / ParticleMesh.generate_whitenoise)
against the oracle and the reference's golden fields.

The random streams are integer-exact; what may differ from the reference is the last bit of the
device's log / sin / cos / sqrt, so float64 fields are compared to 1e-13 absolute (values are O(1);
a single wrong draw would show as an O(1) error) and float32 fields to one float32 ulp.
"""
import os
import sys

import numpy
import pytest
from numpy.testing import assert_array_equal, assert_allclose

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G  # noqa: E402


def tol(dt):
    return dict(rtol=0, atol=1e-13) if numpy.dtype(dt) == numpy.dtype("complex128") else dict(rtol=0, atol=5e-7)


def half(N):
    return (N[0], N[1], N[2] // 2 + 1)


def test_generate_against_reference_golden_and_oracle(oracle):
    from pmesh_b200.whitenoise import generate
    z = numpy.load(os.path.join(HERE, "golden", "whitenoise_golden.npz"))
    for ci, (N, seed, unitary, dt, start, shape) in enumerate(G.WHITENOISE_CASES):
        want = z["wn_%d" % ci]
        got = generate(numpy.zeros(want.shape, dtype=dt), start, N, seed, unitary)
        assert_allclose(got, want, err_msg=str(G.WHITENOISE_CASES[ci]), **tol(dt))
        assert_allclose(got, oracle.whitenoise(numpy.zeros(want.shape, dtype=dt), start, N, seed, unitary), **tol(dt))
        # exact structure: zero mean mode, real self-conjugate modes
        if tuple(start) == (0, 0, 0):
            assert got[0, 0, 0] == 0
            if N[0] % 2 == 0 and N[1] % 2 == 0 and N[2] % 2 == 0:
                assert got[N[0] // 2, N[1] // 2, N[2] // 2].imag == 0


def test_generate_device_layouts_and_blocks(oracle):
    """k-contiguous (1 rank), transposed (1, 2, 0) memory order (P ranks) and arbitrary blocks give the
    same numbers: the field does not depend on layout or partition (tests/test_whitenoise.py:6-25)"""
    from pmesh_b200.device import DeviceArray
    from pmesh_b200.whitenoise import generate
    N = (40, 24, 36)
    whole = oracle.whitenoise(numpy.zeros(half(N), dtype="complex128"), 0, N, 2024, False)
    d = DeviceArray.empty(half(N), "complex128")
    generate(d, 0, N, 2024, False)
    assert_allclose(d.to_host(), whole, **tol("complex128"))
    # transposed storage: (n1_local, nc, n0) in memory, logical (n0, n1_local, nc) view, block of j
    s1, m1 = 5, 11
    store = DeviceArray.zeros((m1, half(N)[2], N[0]), "complex128")
    es = 16
    view = DeviceArray((N[0], m1, half(N)[2]), "complex128", ptr=store.ptr,
                       strides=(es, half(N)[2] * N[0] * es, N[0] * es), base=store)
    generate(view, (0, s1, 0), N, 2024, False)
    got = store.to_host().transpose(2, 0, 1)
    assert_allclose(got, whole[:, s1:s1 + m1, :], **tol("complex128"))
    # blocks that do not start at zero along any axis, both precisions, unitary
    for dt in ("complex128", "complex64"):
        for unitary in (False, True):
            piece = generate(numpy.zeros((9, 7, 5), dtype=dt), (31, 17, 3), N, 99, unitary)
            want = oracle.whitenoise(numpy.zeros((9, 7, 5), dtype=dt), (31, 17, 3), N, 99, unitary)
            assert_allclose(piece, want, **tol(dt))
    # empty block, bad arguments
    generate(numpy.zeros((0, 4, 4), dtype="complex128"), 0, N, 1, False)
    with pytest.raises(NotImplementedError):
        generate(numpy.zeros((8, 8, 8), dtype="complex128"), 0, (8, 8, 8), 1, False)     # full spectrum
    with pytest.raises(NotImplementedError):
        generate(numpy.zeros((8, 5), dtype="complex128"), 0, (8, 8), 1, False)           # 2-D helper of the reference


def test_statistics_and_hermitian_symmetry():
    """tests/test_whitenoise.py:6-12, 40-63 at a size where every warp/tile path of the kernel is taken"""
    from pmesh_b200.whitenoise import generate
    N = 96
    v = generate(numpy.zeros((N, N, N // 2 + 1), dtype="complex128"), 0, (N, N, N), 1, False)
    assert_allclose(v.real.std(), 0.5 ** 0.5, rtol=1e-2)
    assert_allclose(v.imag.std(), 0.5 ** 0.5, rtol=1e-2)
    h = numpy.fft.rfftn(numpy.fft.irfftn(v.copy(), s=(N, N, N), axes=(0, 1, 2)))
    assert_allclose(h, v, rtol=1e-5, atol=1e-9)
    u = generate(numpy.zeros((N, N, N // 2 + 1), dtype="complex64"), 0, (N, N, N), 1, True)
    a = abs(u)
    a[0, 0, 0] = 1
    assert_allclose(a, 1.0, rtol=1e-6)


def test_particlemesh_generate_whitenoise(oracle):
    """ParticleMesh.generate_whitenoise (pm.py:1656-1696): complex and real results, mean, dtype"""
    from pmesh_b200.pm import ParticleMesh, RealField, ComplexField
    for dtype, cdt in (("f8", "complex128"), ("f4", "complex64")):
        pm = ParticleMesh(BoxSize=100.0, Nmesh=[16, 12, 20], dtype=dtype)
        N = (16, 12, 20)
        want = oracle.whitenoise(numpy.zeros(half(N), dtype=cdt), 0, N, 120577, False)
        c = pm.generate_whitenoise(120577)
        assert isinstance(c, ComplexField) and c.value.dtype == numpy.dtype(cdt)
        assert_allclose(c.value, want, **tol(cdt))
        c = pm.generate_whitenoise(120577, unitary=True, mean=2.5)
        wu = oracle.whitenoise(numpy.zeros(half(N), dtype=cdt), 0, N, 120577, True)
        wu[0, 0, 0] = 2.5
        assert_allclose(c.value, wu, **tol(cdt))
        r = pm.generate_whitenoise(120577, type="real")
        assert isinstance(r, RealField)
        wr = numpy.fft.irfftn(want.astype("complex128"), s=N, axes=(0, 1, 2)) * numpy.prod(N)
        assert_allclose(r.value, wr, rtol=0, atol=(1e-9 if dtype == "f8" else 2e-3) * abs(wr).max())
        # on a cubic mesh the noise is exactly Hermitian, so r2c of the real noise is the complex noise
        # again (tests/test_whitenoise.py:40-63; on non-cubic meshes the reference's seed spiral mixes
        # Nmesh[0] and Nmesh[1] and a few k_z = 0 / Nyquist modes lose their partner -- kept as is)
        pm = ParticleMesh(BoxSize=100.0, Nmesh=[16, 16, 16], dtype=dtype)
        want = oracle.whitenoise(numpy.zeros((16, 16, 9), dtype=cdt), 0, (16, 16, 16), 5463, False)
        r = pm.generate_whitenoise(5463, type="real")
        assert_allclose(r.r2c().value, want, rtol=0, atol=1e-10 if dtype == "f8" else 2e-5)
