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
