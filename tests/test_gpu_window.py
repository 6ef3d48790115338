"""GPU parity of paint / readout through the public pmesh API (-> ctypes -> C ABI -> CUDA kernels).

The first block restates the reference's own known-answer tests (pmesh/tests/test_window.py) against
pmesh_b200.window; the second compares with the oracle on seeded random inputs:
bit-exact for the deterministic paint mode and for readout, 1e-6 / 1e-4 relative (f8 / f4) for the
atomic mode, as BASELINE.json's north_star states.
"""
import numpy
import pytest
from numpy.testing import assert_array_equal, assert_allclose, assert_almost_equal

from common import ALL_WINDOWS, FAST_WINDOWS, random_case

pytestmark = pytest.mark.gpu

MODES = ["deterministic", "atomic"]


@pytest.fixture(scope="module")
def W():
    from pmesh_b200 import window
    return window


# ---------------------------------------------------------------- reference known-answer tests
@pytest.mark.parametrize("mode", MODES)
def test_unweighted_weighted(W, mode):
    # pmesh/tests/test_window.py:11-41
    pos = [[0., 0.], [1., 1.], [2., 2.], [3., 3.]]
    real = numpy.zeros((4, 4))
    W.CIC.paint(real, pos, mode=mode)
    assert_array_equal(real, numpy.eye(4))
    real = numpy.zeros((4, 4))
    W.CIC.paint(real, pos, mass=[0., 1., 2., 3.], mode=mode)
    assert_array_equal(real, numpy.diag([0., 1, 2, 3]))


@pytest.mark.parametrize("mode", MODES)
def test_wide(W, mode):
    # test_window.py:43-58
    wcic = W.ResampleWindow("linear", 4)
    real = numpy.zeros((4))
    wcic.paint(real, [[1.5]], mode=mode)
    assert_almost_equal(real, [0.125, 0.375, 0.375, 0.125])
    real = numpy.zeros((4))
    wcic.paint(real, [[1.51]], mode=mode)
    assert_almost_equal(real, [0.1225, 0.3725, 0.3775, 0.1275])
    real = numpy.zeros((4))
    wcic.paint(real, [[1.5]], diffdir=0, mode=mode)
    assert_almost_equal(real, [-0.25, -0.25, 0.25, 0.25])


@pytest.mark.parametrize("mode", MODES)
def test_wrap_translate_affine_scale(W, mode):
    # test_window.py:60-116
    affine = W.Affine(ndim=2, period=2)
    for pos in ([[-.5, -.5]], [[-.5, .5]], [[-.5, 1.5]]):
        real = numpy.zeros((2, 2))
        W.CIC.paint(real, pos, transform=affine, mode=mode)
        assert_array_equal(real, [[0.25, 0.25], [0.25, 0.25]])
    real = numpy.zeros((2, 2))
    W.CIC.paint(real, [[1., 0]], transform=W.Affine(ndim=2, translate=[-1, 0]), mode=mode)
    assert_array_equal(real, [[1., 0.], [0., 0.]])
    affine = W.Affine(ndim=2)
    real = numpy.zeros((4, 4))
    W.CIC.paint(real, [[.5, .5]], transform=affine, mode=mode)
    translate = numpy.zeros((4, 4))
    W.CIC.paint(translate, [[0., 0.]], transform=affine.shift(0.5), mode=mode)
    assert_array_equal(translate, real)
    real = numpy.zeros((2, 2))
    W.CIC.paint(real, [[10., 0]], transform=W.Affine(ndim=2, translate=[-1, 0], scale=0.1), mode=mode)
    assert_almost_equal(real, [[1., 0.], [0, 0.]])


def test_scale_hsml_strides_anisotropic_diff(W):
    # test_window.py:118-186
    real = numpy.zeros(10)
    W.CIC.paint(real, [[50., 0]], hsml=1., transform=W.Affine(ndim=1, translate=[0], scale=0.1))
    assert_array_equal(real, [0., 0., 0., 0., 0., 1., 0., 0., 0., 0.])
    real = numpy.zeros(10)
    W.CIC.paint(real, [[5., 0]], hsml=None, transform=W.Affine(ndim=1, translate=[0], scale=1.))
    assert_array_equal(real, [0., 0., 0., 0., 0., 1., 0., 0., 0., 0.])
    real = numpy.zeros((20, 20))[::10, ::10]
    W.CIC.paint(real, [[1., 0]])
    assert_array_equal(real, [[0, 0], [1, 0]])
    real = numpy.zeros((2, 4))
    W.CIC.paint(real, [[0., 0], [1., 0], [0., 1], [0., 2], [0., 3]])
    assert_array_equal(real, [[1, 1, 1, 1], [1, 0, 0, 0]])
    real = numpy.zeros((2, 2))
    W.CIC.paint(real, [[0.5, 0]], diffdir=0)
    assert_array_equal(real, [[-1, 0], [1, 0]])
    real = numpy.zeros((2, 2))
    W.CIC.paint(real, [[0, 0.5]], diffdir=1)
    assert_array_equal(real, [[-1, 1], [0, 0]])


def test_nearest_lanczos_tsc_cubic_acg(W):
    # test_window.py:188-285
    real = numpy.zeros((4, 4))
    W.NEAREST.paint(real, [[1.2, 1.2]])
    want = numpy.zeros((4, 4)); want[1, 1] = 1
    assert_allclose(real, want, atol=1e-5)
    assert W.NEAREST.support == 1
    real = numpy.zeros((4, 4))
    W.LANCZOS2.paint(real, [[1.5, 1.5]])
    assert_allclose(real,
                    [[0.003977, -0.035797, -0.035797, 0.003977],
                     [-0.035797, 0.322173, 0.322173, -0.035797],
                     [-0.035797, 0.322173, 0.322173, -0.035797],
                     [0.003977, -0.035797, -0.035797, 0.003977]], atol=1e-5)
    assert W.LANCZOS2.support == 4
    a = numpy.zeros(1000)
    W.LANCZOS2.resize(400).paint(a, [[500.5]])
    b = numpy.zeros(1000)
    W.LANCZOS3.resize(400).paint(b, [[500.5]])
    real = numpy.zeros((4))
    W.TSC.paint(real, [[1.5]])
    assert_array_equal(real, [0, 0.5, 0.5, 0])
    real = numpy.zeros((4))
    W.TSC.paint(real, [[1.8]])
    assert_almost_equal(real, [0., 0.245, 0.71, 0.045])
    real = numpy.zeros((5))
    W.TSC.paint(real, [[2.]])
    assert_array_equal(real, [0, 0.125, 0.75, 0.125, 0])
    real = numpy.zeros((5))
    W.TSC.paint(real, [[0.]], transform=W.Affine(ndim=1, period=5))
    assert_array_equal(real, [0.75, 0.125, 0, 0, 0.125])
    real = numpy.zeros((6))
    W.CUBIC.paint(real, [[2.5]])
    assert_allclose(real, [0., 0.02083333, 0.47916667, 0.47916667, 0.02083333, 0.], rtol=1e-6)
    real1 = numpy.zeros((10))
    W.CUBIC.paint(real1, [[4.5]], hsml=2.0)
    real2 = numpy.zeros((10))
    W.CUBIC.resize(8).paint(real2, [[4.5]], hsml=1.0)
    assert_array_equal(real1, real2)
    real = numpy.zeros((4))
    W.ACG3.paint(real, [[2.1]], 1.0)
    assert_allclose(real, [0., 0.21347228, 0.52014034, 0.30805789])


def test_tuned_equals_generic(W):
    # test_window.py:311-360 (+ the PCS/CUBIC comparison the reference lacks)
    pos = [[1.1, 1.3, 2.5]]
    real = numpy.zeros((4, 4, 4)); real2 = numpy.zeros((4, 4, 4))
    W.CIC.paint(real, pos); W.LINEAR.paint(real2, pos)
    assert_array_equal(real, real2)
    for d in range(3):
        d1 = numpy.zeros((4, 4, 4)); d2 = numpy.zeros((4, 4, 4))
        W.CIC.paint(d1, pos, diffdir=d); W.LINEAR.paint(d2, pos, diffdir=d)
        assert_array_equal(d1, d2)
    affine = W.Affine(ndim=3, translate=[2, 1, 2], scale=[0.5, 2.0, 1.1], period=[8, 8, 8])
    numpy.random.seed(1234)
    field = numpy.random.uniform(size=(8, 8, 8))
    pos = [[1.1, 1.3, 2.9]]
    real = numpy.zeros((8, 8, 8)); real2 = numpy.zeros((8, 8, 8))
    W.TSC.paint(real, pos, transform=affine); W.QUADRATIC.paint(real2, pos, transform=affine)
    assert_array_equal(real, real2)
    assert_array_equal(W.TSC.readout(field, pos, transform=affine), W.QUADRATIC.readout(field, pos, transform=affine))
    for d in range(3):
        d1 = numpy.zeros((8, 8, 8)); d2 = numpy.zeros((8, 8, 8))
        W.TSC.paint(d1, pos, diffdir=d, transform=affine); W.QUADRATIC.paint(d2, pos, diffdir=d, transform=affine)
        assert_array_equal(d1, d2)
        assert_array_equal(W.TSC.readout(field, pos, diffdir=d, transform=affine),
                           W.QUADRATIC.readout(field, pos, diffdir=d, transform=affine))
    a = numpy.zeros((8, 8, 8)); b = numpy.zeros((8, 8, 8))
    W.PCS.paint(a, pos, transform=affine); W.CUBIC.paint(b, pos, transform=affine)
    assert_allclose(a, b, rtol=0, atol=1e-14)


def test_compensation(W):
    assert_allclose(W.CIC.get_fwindow([0, 2 * numpy.pi]), [1, 0.0], atol=1e-9)


# ---------------------------------------------------------------- oracle parity on random inputs
@pytest.mark.parametrize("name", ALL_WINDOWS)
def test_paint_readout_vs_oracle(W, oracle, name):
    rng = numpy.random.default_rng(sum(map(ord, name)))
    win = W.windows[name]
    for nd in (1, 2, 3):
        n = 150 if nd == 3 and win.support > 8 else 400
        for dtype, posdtype in (("f8", "f8"), ("f4", "f8"), ("f8", "f4"), ("f4", "f4")):
            shape, pos, mass, scale, translate, period = random_case(rng, nd, n=n, posdtype=posdtype)
            tr = W.Affine(nd, scale=scale, translate=translate, period=period)
            for diffdir in [None] + list(range(nd)):
                want = numpy.zeros(shape, dtype)
                oracle.paint(want, pos, name, mass=mass, diffdir=diffdir, scale=scale, translate=translate, period=period)
                got = numpy.zeros(shape, dtype)
                win.paint(got, pos, mass=mass, diffdir=diffdir, transform=tr, mode="deterministic")
                assert_array_equal(got, want, err_msg="det %s %dD %s/%s diff=%s" % (name, nd, dtype, posdtype, diffdir))
                got = numpy.zeros(shape, dtype)
                win.paint(got, pos, mass=mass, diffdir=diffdir, transform=tr, mode="atomic")
                tol = 1e-6 if dtype == "f8" else 1e-4      # north_star tolerances
                assert_allclose(got, want, rtol=tol, atol=tol * max(1.0, abs(want).max()))
                field = rng.uniform(-1, 1, shape).astype(dtype)
                w = oracle.readout(field, pos, name, diffdir=diffdir, scale=scale, translate=translate, period=period)
                g = win.readout(field, pos, diffdir=diffdir, transform=tr)
                assert g.dtype == numpy.dtype("f8")          # Q7: out defaults to f8
                assert_array_equal(g, w)
                g4 = win.readout(field, pos, out=numpy.zeros(len(pos), "f4"), diffdir=diffdir, transform=tr)
                assert_array_equal(g4, w.astype("f4"))


@pytest.mark.parametrize("name", ["cic", "tsc", "pcs", "cubic", "lanczos2", "db6"])
def test_hsml_vs_oracle(W, oracle, name):
    rng = numpy.random.default_rng(11)
    win = W.windows[name]
    for nd in (1, 2, 3):
        shape, pos, mass, scale, translate, period = random_case(rng, nd, n=100)
        tr = W.Affine(nd, scale=scale, translate=translate, period=period)
        hs = rng.uniform(0.6, 2.2, len(pos))
        hs[::7] = 1.0
        for hsml in (hs, hs.astype("f4"), 1.7):
            want = numpy.zeros(shape)
            oracle.paint(want, pos, name, mass=mass, hsml=hsml, scale=scale, translate=translate, period=period)
            got = numpy.zeros(shape)
            win.paint(got, pos, mass=mass, hsml=hsml, transform=tr, mode="deterministic")
            assert_array_equal(got, want)
            got = numpy.zeros(shape)
            win.paint(got, pos, mass=mass, hsml=hsml, transform=tr, mode="atomic")
            assert_allclose(got, want, rtol=1e-6, atol=1e-6 * abs(want).max())
            w = oracle.readout(want, pos, name, hsml=hsml, scale=scale, translate=translate, period=period)
            assert_array_equal(win.readout(want, pos, hsml=hsml, transform=tr), w)


def test_edge_cases(W, oracle):
    # empty particle set, canvas with a zero-length axis, particles entirely outside a non-periodic canvas
    real = numpy.zeros((4, 4))
    W.CIC.paint(real, numpy.zeros((0, 2)))
    assert (real == 0).all()
    assert W.CIC.readout(real, numpy.zeros((0, 2))).shape == (0,)
    empty = numpy.zeros((0, 4))
    W.CIC.paint(empty, [[1., 1.]])
    real = numpy.zeros((4, 4))
    W.TSC.paint(real, [[100., -50.]], mode="deterministic")
    assert (real == 0).all()
    assert W.TSC.readout(numpy.ones((4, 4)), [[100., -50.]])[0] == 0
    # strided canvas + weights, deterministic vs oracle
    rng = numpy.random.default_rng(3)
    pos = rng.uniform(-2, 8, (200, 2))
    big_w = numpy.zeros((12, 14)); big_g = numpy.zeros((12, 14))
    for name in ("cic", "pcs", "lanczos3"):
        oracle.paint(big_w[::2, ::2], pos, name)
        W.windows[name].paint(big_g[::2, ::2], pos, mode="deterministic")
        assert_array_equal(big_g, big_w)


def test_deterministic_chunking_is_invisible(W, oracle):
    """a tiny workspace forces many chunks; chunk boundaries must not change a single bit"""
    from pmesh_b200 import _lib
    rng = numpy.random.default_rng(5)
    pos = rng.uniform(0, 16, (5000, 3))
    mass = rng.uniform(0.5, 2, 5000)
    want = numpy.zeros((16, 16, 16), "f4")
    oracle.paint(want, pos, "tsc", mass=mass, period=[16, 16, 16])
    ctx = _lib.context()
    ctx.set_workspace_limit(64 << 10)       # 64 KiB -> ~75 particles per chunk
    try:
        got = numpy.zeros((16, 16, 16), "f4")
        W.TSC.paint(got, pos, mass=mass, transform=W.Affine(3, period=16), mode="deterministic")
    finally:
        ctx.set_workspace_limit(2 << 30)
    assert_array_equal(got, want)


def test_readout_grad_fused_equals_separate(W, oracle):
    from pmesh_b200.device import DeviceArray
    rng = numpy.random.default_rng(9)
    for name in ("cic", "tsc", "pcs", "lanczos2"):
        for nd in (1, 2, 3):
            shape, pos, mass, scale, translate, period = random_case(rng, nd, n=200)
            tr = W.Affine(nd, scale=scale, translate=translate, period=period)
            field = rng.uniform(-1, 1, shape)
            val, grad = W.windows[name].readout_grad(DeviceArray.from_host(field), DeviceArray.from_host(pos), transform=tr)
            assert_array_equal(val.to_host(), oracle.readout(field, pos, name, scale=scale, translate=translate, period=period))
            g = grad.to_host()
            for d in range(nd):
                assert_array_equal(g[:, d], oracle.readout(field, pos, name, diffdir=d, scale=scale, translate=translate, period=period))


def test_large_cic_properties_and_device_arrays(W, oracle):
    """cfg1 size (64^3 particles, 64^3 mesh, BoxSize 100) in full against the oracle; mass conservation"""
    from pmesh_b200.device import DeviceArray
    rng = numpy.random.default_rng(42)
    N = 64
    pos = rng.uniform(0, 100, (N ** 3, 3))
    tr = W.Affine(3, scale=N / 100.0, translate=0, period=N)
    want = numpy.zeros((N, N, N))
    oracle.paint(want, pos, "cic", scale=N / 100.0, period=[N] * 3)
    dpos = DeviceArray.from_host(pos)
    for mode, tol in (("deterministic", 0), ("atomic", 1e-6)):
        mesh = DeviceArray.zeros((N, N, N), "f8")
        W.CIC.paint(mesh, dpos, transform=tr, mode=mode)
        got = mesh.to_host()
        if tol == 0:
            assert_array_equal(got, want)
        else:
            assert_allclose(got, want, rtol=tol, atol=tol)
        assert abs(got.sum() - N ** 3) < 1e-6 * N ** 3
    r = W.CIC.readout(DeviceArray.from_host(want), dpos, transform=tr)
    assert_array_equal(r.to_host(), oracle.readout(want, pos, "cic", scale=N / 100.0, period=[N] * 3))


@pytest.mark.parametrize("name", ["nnb", "cic", "tsc", "pcs"])
@pytest.mark.parametrize("order", ["lattice", "random"])
def test_scheduled_kernels_large_inputs(W, oracle, name, order):
    """>= 2^18 particles on a 3-D mesh take the locality-scheduled kernels with warp-aggregated
    atomics (pmb_sched.cuh): full periodic canvas and a slab (size[0] < period, translated) canvas,
    lattice-ordered (merging lanes) and shuffled (no merging) particles, f8 and f4."""
    from pmesh_b200.device import DeviceArray
    N = 72
    rng = numpy.random.default_rng(17)
    q = numpy.indices((N, N, N)).reshape(3, -1).T + 0.5
    pos = (q + 2.5 * numpy.sin(2 * numpy.pi * q[:, ::-1] / N) + rng.uniform(-0.2, 0.2, q.shape)) % N
    if order == "random":
        pos = pos[rng.permutation(len(pos))]
    mass = rng.uniform(0.5, 2.0, len(pos))
    if name == "cic" and order == "random":
        pos = pos.astype("f4")          # the strided / float32 position loader of the lean CIC kernels
    dpos, dmass = DeviceArray.from_host(pos), DeviceArray.from_host(mass)
    for shape, translate in (((N, N, N), [0.0, 0.0, 0.0]), ((20, N, N), [-30.0, 0.0, 0.0])):
        tr = W.Affine(3, scale=1.0, translate=translate, period=N)
        for dtype, tol in (("f8", 1e-6), ("f4", 1e-4)):
            want = numpy.zeros(shape, dtype)
            oracle.paint(want, pos, name, mass=mass, translate=translate, period=[N] * 3)
            mesh = DeviceArray.zeros(shape, dtype)
            W.windows[name].paint(mesh, dpos, mass=dmass, transform=tr, mode="atomic")
            got = mesh.to_host()
            assert_allclose(got, want, rtol=tol, atol=tol * abs(want).max())
            field = rng.uniform(-1, 1, shape).astype(dtype)
            r = W.windows[name].readout(DeviceArray.from_host(field), dpos, transform=tr)
            assert_array_equal(r.to_host(), oracle.readout(field, pos, name, translate=translate, period=[N] * 3))
            # gradient windows (paint_jvp / the paint of a vjp) take the same scheduled / carry kernels
            for diffdir in ((1,) if dtype == "f4" else (0, 2)):
                want = numpy.zeros(shape, dtype)
                oracle.paint(want, pos, name, mass=mass, diffdir=diffdir, translate=translate, period=[N] * 3)
                mesh = DeviceArray.zeros(shape, dtype)
                W.windows[name].paint(mesh, dpos, mass=dmass, diffdir=diffdir, transform=tr, mode="atomic")
                assert_allclose(mesh.to_host(), want, rtol=tol, atol=tol * max(abs(want).max(), 1e-30))
                # ... and the gather of a gradient window, bit for bit
                r = W.windows[name].readout(DeviceArray.from_host(field), dpos, diffdir=diffdir, transform=tr)
                assert_array_equal(r.to_host(), oracle.readout(field, pos, name, diffdir=diffdir, translate=translate, period=[N] * 3))
            # value + all gradients from one sweep == the separate gathers
            val, grad = W.windows[name].readout_grad(DeviceArray.from_host(field), dpos, transform=tr)
            assert_array_equal(val.to_host(), oracle.readout(field, pos, name, translate=translate, period=[N] * 3))
            gh = grad.to_host()
            for d in range(3):
                assert_array_equal(gh[:, d], oracle.readout(field, pos, name, diffdir=d, translate=translate, period=[N] * 3))
    # anisotropic scale, non-periodic canvas (points outside are dropped), float32 positions
    tr = W.Affine(3, scale=[0.9, 1.1, 0.5], translate=[1.0, -2.0, 3.5], period=0)
    shape = (60, 70, 40)
    p4 = pos.astype("f4")
    want = numpy.zeros(shape, "f8")
    oracle.paint(want, p4, name, mass=mass, scale=[0.9, 1.1, 0.5], translate=[1.0, -2.0, 3.5], period=[0] * 3)
    mesh = DeviceArray.zeros(shape, "f8")
    W.windows[name].paint(mesh, DeviceArray.from_host(p4), mass=dmass, transform=tr, mode="atomic")
    assert_allclose(mesh.to_host(), want, rtol=1e-6, atol=1e-6 * abs(want).max())
    field = rng.uniform(-1, 1, shape)
    r = W.windows[name].readout(DeviceArray.from_host(field), DeviceArray.from_host(p4), transform=tr)
    assert_array_equal(r.to_host(), oracle.readout(field, p4, name, scale=[0.9, 1.1, 0.5], translate=[1.0, -2.0, 3.5], period=[0] * 3))


@pytest.mark.parametrize("order", ["lattice", "random"])
def test_readout_multi_equals_single_readouts(W, oracle, order):
    """pmb_readout_multi: up to three canvases read in one sweep over the particles (bulk-copy ring
    kernel on lattice order, tile-binned permutation on random order) == one readout per canvas == the
    oracle, bit for bit; full periodic canvas and a translated slab; odd particle count (tail chunk)."""
    from pmesh_b200.device import DeviceArray
    N = 72
    rng = numpy.random.default_rng(23)
    q = numpy.indices((N, N, N)).reshape(3, -1).T + 0.5
    pos = (q + 2.5 * numpy.sin(2 * numpy.pi * q[:, ::-1] / N) + rng.uniform(-0.2, 0.2, q.shape)) % N
    if order == "random":
        pos = pos[rng.permutation(len(pos))]
    pos = pos[:len(pos) - 37]
    dpos = DeviceArray.from_host(pos)
    for shape, translate in (((N, N, N), [0.0, 0.0, 0.0]), ((20, N, N), [-30.0, 0.0, 0.0])):
        tr = W.Affine(3, scale=1.0, translate=translate, period=N)
        for dtype in ("f8", "f4"):
            fields = [rng.uniform(-1, 1, shape).astype(dtype) for _ in range(3)]
            dfields = [DeviceArray.from_host(f) for f in fields]
            for nf in (1, 2, 3):
                outs = W.windows["cic"].readout_multi(dfields[:nf], dpos, transform=tr)
                for f, df, o in zip(fields, dfields, outs):
                    want = oracle.readout(f, pos, "cic", translate=translate, period=[N] * 3)
                    assert_array_equal(o.to_host(), want)
                    assert_array_equal(W.windows["cic"].readout(df, dpos, transform=tr).to_host(), want)
    # any other window falls back to one readout per canvas
    tr = W.Affine(3, scale=1.0, translate=[0.0] * 3, period=N)
    fields = [rng.uniform(-1, 1, (N, N, N)) for _ in range(2)]
    outs = W.windows["tsc"].readout_multi([DeviceArray.from_host(f) for f in fields], dpos, transform=tr)
    for f, o in zip(fields, outs):
        assert_array_equal(o.to_host(), oracle.readout(f, pos, "tsc", period=[N] * 3))


@pytest.mark.parametrize("name", ["lanczos3", "acg3", "db6", "cubic", "linear"])
def test_fused_gradient_sweep_of_generic_windows(W, oracle, name):
    """value + all gradients of a run-time-support window in ONE neighbourhood sweep (pmb_k_readout_grad_dyn)
    == the separate readout(diffdir = d) calls == the oracle, bit for bit; 3-D periodic, 2-D non-periodic
    and a translated slab canvas, f8 and f4 meshes"""
    from pmesh_b200.device import DeviceArray
    rng = numpy.random.default_rng(29)
    for shape, translate, period in (((20, 18, 22), [0.0, 0.0, 0.0], [20, 18, 22]),
                                     ((9, 18, 22), [-5.0, 0.0, 0.0], [20, 18, 22]),
                                     ((24, 30), [1.5, -2.0], [0, 0])):
        nd = len(shape)
        pos = rng.uniform(-3, 25, (3000, nd))
        dpos = DeviceArray.from_host(pos)
        tr = W.Affine(nd, scale=1.0, translate=translate, period=period)
        for dtype in ("f8", "f4"):
            field = rng.uniform(-1, 1, shape).astype(dtype)
            dfield = DeviceArray.from_host(field)
            val, grad = W.windows[name].readout_grad(dfield, dpos, transform=tr)
            assert_array_equal(val.to_host(), oracle.readout(field, pos, name, translate=translate, period=period))
            gh = grad.to_host()
            for d in range(nd):
                want = oracle.readout(field, pos, name, diffdir=d, translate=translate, period=period)
                assert_array_equal(gh[:, d], want)
                sep = W.windows[name].readout(dfield, dpos, diffdir=d, transform=tr)
                assert_array_equal(sep.to_host(), want)
