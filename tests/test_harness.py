"""The device stencil code, compiled for the host (tests/harness), against the oracle.

No GPU needed: pmb_window.h / pmb_stencil.cuh compile for host and device, so the exact arithmetic
the CUDA kernels execute is checked here bit for bit (g++ -ffp-contract=off == nvcc -fmad=false).
The kernels' launch geometry, atomics and the CUB sort are what remains for the `-m gpu` tests.
"""
import ctypes

import numpy
import pytest
from numpy.testing import assert_array_equal, assert_allclose

from common import ALL_WINDOWS, random_case, resample_args


def _paint(h, *a, **kw):
    args, keep = resample_args(*a, **kw)
    rc = h.hh_paint(ctypes.byref(args))
    assert rc == 0, rc


def _readout(h, *a, **kw):
    args, keep = resample_args(*a, **kw)
    rc = h.hh_readout(ctypes.byref(args))
    assert rc == 0, rc


@pytest.mark.parametrize("name", ALL_WINDOWS)
@pytest.mark.parametrize("nd", [1, 2, 3])
def test_deterministic_paint_and_readout_bit_exact(harness, oracle, name, nd):
    rng = numpy.random.default_rng(100 + nd)
    n = 120 if nd == 3 and name in ("db20", "sym20", "db12", "sym12", "lanczos6", "lanczos5") else 300
    for dtype in ("f8", "f4"):
        for posdtype in ("f8", "f4"):
            shape, pos, mass, scale, translate, period = random_case(rng, nd, n=n, dtype=dtype, posdtype=posdtype)
            for diffdir in [None] + list(range(nd)):
                want = numpy.zeros(shape, dtype)
                oracle.paint(want, pos, name, mass=mass, diffdir=diffdir, scale=scale, translate=translate, period=period)
                got = numpy.zeros(shape, dtype)
                _paint(harness, name, -1, got, pos, mass=mass, diffdir=diffdir, scale=scale, translate=translate,
                       period=period, mode=1)
                assert_array_equal(got, want, err_msg="%s %dD %s diff=%s" % (name, nd, dtype, diffdir))
                # atomic-order emulation: float canvases add in float
                got2 = numpy.zeros(shape, dtype)
                _paint(harness, name, -1, got2, pos, mass=mass, diffdir=diffdir, scale=scale, translate=translate,
                       period=period, mode=0)
                tol = 1e-12 if dtype == "f8" else 2e-5
                assert_allclose(got2, want, rtol=tol, atol=tol * max(1.0, abs(want).max()))
                field = rng.uniform(-1, 1, shape).astype(dtype)
                w = oracle.readout(field, pos, name, diffdir=diffdir, scale=scale, translate=translate, period=period)
                for odt in ("f8", "f4"):
                    g = numpy.zeros(len(pos), odt)
                    _readout(harness, name, -1, field, pos, out=g, diffdir=diffdir, scale=scale, translate=translate, period=period)
                    assert_array_equal(g, w.astype(odt))


@pytest.mark.parametrize("name", ["cic", "tsc", "pcs", "linear", "cubic", "lanczos2", "acg3", "db6"])
def test_hsml_and_resize(harness, oracle, name):
    rng = numpy.random.default_rng(7)
    for nd in (1, 2, 3):
        shape, pos, mass, scale, translate, period = random_case(rng, nd, n=80)
        hs = rng.uniform(0.6, 2.2, len(pos))
        hs[::7] = 1.0        # exactly native support -> tuned path per particle
        for hsml in (hs, 1.0, 1.7):
            want = numpy.zeros(shape)
            oracle.paint(want, pos, name, mass=mass, hsml=hsml, scale=scale, translate=translate, period=period)
            got = numpy.zeros(shape)
            _paint(harness, name, -1, got, pos, mass=mass, hsml=hsml, scale=scale, translate=translate, period=period, mode=1)
            assert_array_equal(got, want)
            w = oracle.readout(want, pos, name, hsml=hsml, scale=scale, translate=translate, period=period)
            g = numpy.zeros(len(pos))
            _readout(harness, name, -1, want, pos, hsml=hsml, out=g, scale=scale, translate=translate, period=period)
            assert_array_equal(g, w)
        for support in (5, 8):
            want = numpy.zeros(shape)
            oracle.paint(want, pos, name, support=support, mass=mass, scale=scale, translate=translate, period=period)
            got = numpy.zeros(shape)
            _paint(harness, name, support, got, pos, mass=mass, scale=scale, translate=translate, period=period, mode=1)
            assert_array_equal(got, want)


def test_wide_support_on_the_fly(harness, oracle):
    # LANCZOS2.resize(400) in 1-D (reference tests/test_window.py:215-219): wider than the cached stencil
    for name in ("lanczos2", "lanczos3"):
        want = numpy.zeros(1000)
        oracle.paint(want, numpy.array([[500.5]]), name, support=400)
        got = numpy.zeros(1000)
        _paint(harness, name, 400, got, numpy.array([[500.5]]), mode=1)
        assert_array_equal(got, want)
        assert abs(got.sum() - 1.0) < 1e-3


def test_nonperiodic_clipping_and_strides(harness, oracle):
    rng = numpy.random.default_rng(3)
    pos = rng.uniform(-2, 8, (200, 2))
    big_w = numpy.zeros((12, 14))
    big_g = numpy.zeros((12, 14))
    for name in ("cic", "tsc", "pcs", "lanczos3"):
        want = big_w[::2, ::2]
        got = big_g[::2, ::2]
        oracle.paint(want, pos, name)                       # period 0: contributions outside are dropped
        _paint(harness, name, -1, got, pos, mode=1)
        assert_array_equal(big_g, big_w)


def test_pcs_gradient_scale_quirk(harness, oracle):
    """tuned PCS derivative lacks scale[d]; 'cubic' has it."""
    rng = numpy.random.default_rng(5)
    pos = rng.uniform(0, 8, (50, 3))
    scale = numpy.array([0.5, 2.0, 1.1])
    period = numpy.array([8, 8, 8])
    for d in range(3):
        a = numpy.zeros((8, 8, 8)); b = numpy.zeros((8, 8, 8)); c = numpy.zeros((8, 8, 8))
        _paint(harness, "pcs", -1, a, pos, diffdir=d, scale=scale, period=period, mode=1)
        _paint(harness, "cubic", -1, b, pos, diffdir=d, scale=scale, period=period, mode=1)
        _paint(harness, "pcs", -1, c, pos, diffdir=d, scale=scale, period=period, mode=1, pcsfix=1)
        assert_allclose(a * scale[d], b, rtol=1e-12, atol=1e-13)
        assert_allclose(c, b, rtol=1e-12, atol=1e-13)


# ------------------------------------------------------------------ routing arithmetic (pmb_route.h)
def _hh_decompose(h, pos, edges, P, smoothing, periodic=True, assign=None, scale=None):
    nd = len(edges)
    pos = numpy.ascontiguousarray(pos)
    sm = numpy.empty(nd); sm[:] = smoothing
    sc = numpy.ones(nd) if scale is None else numpy.asarray(scale, dtype="f8")[:nd].copy()
    e = numpy.ascontiguousarray(numpy.concatenate([numpy.asarray(x, dtype="f8") for x in edges]))
    ne = numpy.array([len(x) for x in edges], dtype="int32")
    shape = [len(x) - 1 for x in edges]
    ndom = int(numpy.prod(shape))
    if assign is None:
        if P >= ndom:
            assign = numpy.arange(ndom, dtype="int32")
        else:
            assign = numpy.empty(ndom, dtype="int32")
            for i in range(P):
                assign[i * ndom // P:(i + 1) * ndom // P] = i
    assign = numpy.ascontiguousarray(assign, dtype="int32")
    dd = numpy.zeros(shape, dtype="int16")
    for i, edge in enumerate(edges):
        edge = numpy.asarray(edge)
        d1 = (edge[1:] == edge[:-1]).reshape([-1 if ii == i else 1 for ii in range(nd)])
        dd[...] |= d1
    deg = numpy.ascontiguousarray(dd.ravel())
    cap = len(pos) * max(P, 1) + 1
    counts = numpy.zeros(P, dtype="int32")
    indices = numpy.zeros(cap, dtype="int32")
    h.hh_decompose.restype = ctypes.c_int64
    n = h.hh_decompose(ctypes.c_void_p(pos.ctypes.data), ctypes.c_int(pos.dtype.itemsize), ctypes.c_int64(len(pos)),
                       ctypes.c_int64(pos.strides[0]), ctypes.c_int64(pos.strides[1]), ctypes.c_int(nd),
                       ctypes.c_void_p(sc.ctypes.data), ctypes.c_void_p(sm.ctypes.data), ctypes.c_void_p(e.ctypes.data),
                       ctypes.c_void_p(ne.ctypes.data), ctypes.c_int(int(bool(periodic))),
                       ctypes.c_void_p(assign.ctypes.data), ctypes.c_void_p(deg.ctypes.data), ctypes.c_int(P),
                       ctypes.c_void_p(counts.ctypes.data), ctypes.c_void_p(indices.ctypes.data), ctypes.c_int64(cap))
    assert n >= 0
    return counts, indices[:n]


def test_routing_arithmetic_bit_exact(harness, oracle):
    """the rank-set arithmetic the count kernel executes (pmb_route.h, host build) == the oracle ==
    the reference's golden counts / indices, for every golden geometry and for fresh random ones"""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden as G
    z = numpy.load(os.path.join(here, "golden", "domain_golden.npz"))
    for case, (edges, P, smoothing, periodic) in enumerate(G.DOMAIN_CASES):
        pos = G.domain_inputs(case)
        for rank in (0, P - 1):
            assign = z["assign_%d_%d" % (case, rank)]
            counts, indices = _hh_decompose(harness, pos, edges, P, smoothing, periodic, assign)
            assert_array_equal(counts, z["counts_%d_%d" % (case, rank)])
            assert_array_equal(indices, z["indices_%d_%d" % (case, rank)])
    rng = numpy.random.default_rng(77)
    for trial in range(20):
        nd = int(rng.integers(1, 4))
        shape = [int(rng.integers(1, 5)) for d in range(nd)]
        box = rng.uniform(4, 64, nd)
        edges = [numpy.concatenate([[0.0], numpy.sort(rng.uniform(0, b, s - 1)), [b]]) for s, b in zip(shape, box)]
        P = int(rng.integers(1, 9))
        periodic = bool(trial % 4)
        smoothing = rng.uniform(0, 3, nd)
        scale = rng.uniform(0.5, 2.0, nd) if trial % 3 == 0 else None
        pos = rng.uniform(-0.7, 1.7, (3000, nd)) * box
        pos[:30] = 0.0
        pos[30:40] = -1e-17
        pos[40:50] = box
        if scale is not None:
            pos = pos / scale
        pos = pos.astype("f4" if trial % 5 == 1 else "f8")
        want_c, want_i = oracle.decompose(pos, edges, P, smoothing=smoothing, periodic=periodic, scale=scale)
        got_c, got_i = _hh_decompose(harness, pos, edges, P, smoothing, periodic, scale=scale)
        assert_array_equal(got_c, want_c, err_msg="trial %d" % trial)
        assert_array_equal(got_i, want_i, err_msg="trial %d" % trial)


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048, 4096])
def test_line_transform_of_the_fused_backward_pass(harness, n):
    """pmb_ifft.cuh -- the Stockham passes (radices 16, R2, R3), their shared-memory index maps (both layouts) and the
    register butterflies, run "thread" by "thread" on the host -- against numpy.fft.ifft * n"""
    rng = numpy.random.default_rng(n)
    for es, dt, tol in ((16, numpy.complex128, 2e-15), (8, numpy.complex64, 1e-6)):
        for contig in (0, 1):
            x = (rng.standard_normal((64, n)) + 1j * rng.standard_normal((64, n))).astype(dt)
            out = numpy.zeros_like(x)
            tw = numpy.exp(2j * numpy.pi * numpy.arange(n) / n).astype(dt)
            nb = harness.hh_ifft_lines(ctypes.c_int(n), ctypes.c_int(es), ctypes.c_int(contig), ctypes.c_void_p(x.ctypes.data),
                                       ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(tw.ctypes.data))
            assert 1 <= nb <= 64
            want = numpy.fft.ifft(x[:nb].astype(numpy.complex128), axis=1) * n
            assert abs(out[:nb] - want).max() <= tol * abs(want).max()
    assert harness.hh_ifft_lines(ctypes.c_int(96), ctypes.c_int(16), ctypes.c_int(0), None, None, None) == -1
