"""GPU tests of the ParticleMesh / Field API on one rank: FFT conventions, transfer kernels, the
decomposed paint, gradients (vjp/jvp) and the full PM force step against the oracle pipeline.

FFT values: the reference pins only conventions (normalisation, Hermitian layout, round trips,
SURVEY 8c); here they are compared with numpy.fft (rfftn/N, irfftn*N) to 1e-6 relative of the field
scale for f8 and 1e-4 for f4 -- the tolerances BASELINE.json's north_star states.
"""
import numpy
import pytest
from numpy.testing import assert_array_equal, assert_allclose

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    from pmesh_b200 import pm
    return pm


def rel_close(a, b, tol):
    scale = max(abs(numpy.asarray(b)).max(), 1e-300)
    assert abs(numpy.asarray(a) - numpy.asarray(b)).max() <= tol * scale, \
        (abs(numpy.asarray(a) - numpy.asarray(b)).max() / scale)


@pytest.mark.parametrize("shape", [(8,), (8, 8), (6, 10), (8, 8, 8), (4, 6, 10), (16, 8, 12), (5, 7, 9)])
@pytest.mark.parametrize("dtype", ["f8", "f4"])
def test_fft_conventions(P, oracle, shape, dtype):
    pm = P.ParticleMesh(BoxSize=8.0, Nmesh=shape, dtype=dtype)
    tol = 1e-6 if dtype == "f8" else 1e-4
    rng = numpy.random.default_rng(len(shape))
    x = rng.normal(size=shape).astype(dtype)
    real = pm.create(type="real", value=x)
    assert real.shape == tuple(shape) and tuple(real.cshape) == tuple(shape)
    cplx = real.r2c()
    cs = list(shape); cs[-1] = cs[-1] // 2 + 1
    assert tuple(cplx.cshape) == tuple(cs) and cplx.shape == tuple(cs)       # tests/test_pm.py:36-42
    rel_close(cplx.value, oracle.r2c(x.astype("f8")), tol)
    assert_array_equal(real.value, x)                                       # input preserved
    back = cplx.c2r()
    rel_close(back.value, x, tol)                                           # round trip, test_pm.py:126-141
    rel_close(cplx.value, oracle.r2c(x.astype("f8")), tol)                  # c2r preserves its input
    # in-place variants (out=Ellipsis)
    real2 = pm.create(type="real", value=x)
    c2 = real2.r2c(out=Ellipsis)
    rel_close(c2.value, oracle.r2c(x.astype("f8")), tol)
    r2 = c2.c2r(out=Ellipsis)
    rel_close(r2.value, x, tol)
    # normalisation: r2c(1)[0] == 1 (tests/test-particlemesh.py:6-12)
    one = pm.create(type="real", value=1.0)
    assert abs(one.r2c().cgetitem([0] * len(shape)) - 1.0) < 1e-6
    # c2r of numpy-made spectra (test_pm.py:421-428)
    y = oracle.r2c(x.astype("f8"))
    cf = pm.create(type="complex", value=y)
    rel_close(cf.c2r().value, x, tol)


def test_field_semantics(P):
    pm = P.ParticleMesh(BoxSize=8.0, Nmesh=[8, 8, 8], dtype="f8")
    real = pm.create(type="real")
    assert (real.value == 0).all()
    real[...] = 2.0
    real[...] *= 3.0
    assert (real.value == 6.0).all()
    assert abs(real.csum() - 6.0 * 512) < 1e-9 and abs(real.cmean() - 6.0) < 1e-12
    r2 = real + 1.0                     # numpy ufunc protocol re-wraps as a field (pm.py:189-199)
    assert isinstance(r2, P.RealField) and (r2.value == 7.0).all()
    assert (numpy.asarray(real) == 6.0).all()
    r3 = real.copy()
    r3.scale(0.5)                       # device-side scaling
    assert (r3.value == 3.0).all() and (real.value == 6.0).all()
    # wavenumbers: Nyquist negative (test_pm.py:46-53, 266-286)
    c = pm.create(type="complex")
    k0 = c.x[0].ravel()
    assert_allclose(k0, numpy.fft.fftfreq(8, 1.0 / 8) * 2 * numpy.pi / 8.0 * numpy.where(numpy.arange(8) == 4, 1, 1))
    assert k0[4] < 0
    assert c.i[2].ravel().tolist() == [0, 1, 2, 3, 4]
    assert real.slices == (slice(0, 8),) * 3
    # hermitian csetitem / cgetitem (test_pm.py:559-630, single rank)
    c.csetitem([1, 2, 3], 1 + 2j)
    assert c.cgetitem([1, 2, 3]) == 1 + 2j
    c.csetitem([0, 0, 0], 1 + 2j)
    assert c.cgetitem([0, 0, 0]) == 1.0            # self-conjugate mode keeps only the real part


@pytest.mark.parametrize("dtype", ["f8", "f4"])
def test_transfer_kernels_match_python_callables(P, oracle, dtype):
    from pmesh_b200 import transfer as T
    from pmesh_b200.window import CIC
    pm = P.ParticleMesh(BoxSize=[100.0, 80.0, 120.0], Nmesh=[8, 6, 10], dtype=dtype)
    tol = 1e-6 if dtype == "f8" else 1e-4
    rng = numpy.random.default_rng(0)
    x = rng.normal(size=(8, 6, 10)).astype(dtype)
    cplx = pm.create(type="real", value=x).r2c()
    objs = [T.GravityFD4(0), T.GravityFD4(2), T.GradientK(1), T.InverseLaplace(), T.GaussianLowpass(7.0),
            T.GradientIK(2), T.Scale(2.5)]
    for tf in objs:
        gpu = cplx.apply(tf)                                     # device kernel
        host = cplx.apply(lambda k, v, tf=tf: tf(k, v))          # the same formula through the slab path
        rel_close(gpu.value, host.value, tol)
    # against the oracle's own restatement of examples/nbody.py:162-170
    want = oracle.transfer(oracle.r2c(x.astype("f8")), [8, 6, 10], [100.0, 80.0, 120.0], "gravity_fd4", 1)
    rel_close(cplx.apply(T.GravityFD4(1)).value, want, tol)
    # compensation (kind='circular'), device vs python
    comp = CIC.get_compensation()
    rel_close(cplx.apply(comp, kind="circular").value, cplx.apply(lambda w, v: comp(w, v), kind="circular").value, tol)
    # in place
    c2 = cplx.copy()
    c2.apply(T.InverseLaplace(), out=Ellipsis)
    rel_close(c2.value, cplx.apply(T.InverseLaplace()).value, tol)


def test_paint_uniform_grid_and_decompose_single_rank(P, oracle):
    # tests/test_pm.py:826-867: a uniform grid paints to exactly 1.0, also when shifted by +-box
    pm = P.ParticleMesh(BoxSize=8.0, Nmesh=[8, 8, 8], dtype="f8")
    grid = pm.generate_uniform_particle_grid(shift=0.0)
    assert grid.dtype == numpy.dtype("f8")                       # Q8
    for shift in (0.0, 8.0, -8.0):
        for res in ("cic", "tsc", "pcs"):
            real = pm.paint(grid + shift, resampler=res, mode="deterministic")
            assert_allclose(real.value, 1.0, rtol=0, atol=1e-14)
    # decompose + paint == serial paint (test_pm.py:228-264), bit-exact on one rank
    rng = numpy.random.default_rng(3)
    pos = rng.uniform(-4, 12, (3000, 3))
    mass = rng.uniform(0.5, 2.0, 3000)
    for res in ("cic", "tsc", "db12"):
        layout = pm.decompose(pos, smoothing=res)
        got = pm.paint(pos, mass=mass, resampler=res, layout=layout, mode="deterministic")
        want = numpy.zeros((8, 8, 8))
        oracle.paint(want, pos, res, mass=mass, scale=1.0, period=[8, 8, 8])
        assert_array_equal(got.value, want)
        r = got.readout(pos, resampler=res, layout=layout)
        assert_array_equal(r, oracle.readout(want, pos, res, scale=1.0, period=[8, 8, 8]))
    # hold=True accumulates, hold=False clears
    a = pm.paint(pos, mode="deterministic")
    pm.paint(pos, hold=True, out=a, mode="deterministic")
    b = pm.paint(pos, mass=1.0, mode="deterministic")
    assert_allclose(a.value, 2 * b.value, rtol=1e-14)
    # readout dtype combinations (test_pm.py:661-677)
    for dt in ("f4", "f8"):
        pmx = P.ParticleMesh(BoxSize=8.0, Nmesh=[8, 8, 8], dtype=dt)
        f = pmx.paint(pos.astype("f4"))
        assert f.readout(pos.astype("f4")).dtype == numpy.dtype("f8")
        assert f.readout(pos, out=numpy.zeros(len(pos), "f4")).dtype == numpy.dtype("f4")


def test_2d_mesh(P, oracle):
    pm = P.ParticleMesh(BoxSize=[8.0, 4.0], Nmesh=[8, 8], dtype="f8")
    rng = numpy.random.default_rng(5)
    pos = rng.uniform(0, 8, (500, 2))
    f = pm.paint(pos, resampler="tsc", mode="deterministic")
    want = numpy.zeros((8, 8))
    oracle.paint(want, pos, "tsc", scale=[1.0, 2.0], period=[8, 8])
    assert_array_equal(f.value, want)
    rel_close(f.r2c().value, oracle.r2c(want), 1e-6)


@pytest.mark.parametrize("res", ["cic", "tsc", "pcs", "lanczos3"])
def test_gradients_vjp_jvp(P, oracle, res):
    """pmesh/tests/test_gradient.py:103-263 on a 4^3 mesh: vjp vs jvp to 1e-7, vs finite differences 1e-4;
    plus the operators against their definition through the oracle."""
    pm = P.ParticleMesh(BoxSize=4.0, Nmesh=[4, 4, 4], dtype="f8", resampler=res)
    rng = numpy.random.default_rng(7)
    pos = rng.uniform(0, 4, (30, 3))
    mass = rng.uniform(0.5, 1.5, 30)
    v = pm.create(type="real", value=rng.normal(size=(4, 4, 4)))
    vpos, vmass = rng.normal(size=(30, 3)), rng.normal(size=30)

    # paint: <v, J dx> == <J^T v, dx>
    out_pos, out_mass = pm.paint_vjp(v, pos, mass=mass)
    jvp = pm.paint_jvp(pos, mass=mass, v_pos=vpos, v_mass=vmass)
    lhs = (v.value * jvp.value).sum()
    rhs = (out_pos * vpos).sum() + (out_mass * vmass).sum()
    assert_allclose(lhs, rhs, rtol=1e-7)
    # definition through the oracle: out_pos[:, d] = readout(v, gradient=d) * mass
    for d in range(3):
        w = oracle.readout(v.value, pos, res, diffdir=d, scale=1.0, period=[4, 4, 4]) * mass
        assert_allclose(out_pos[:, d], w, rtol=1e-12, atol=1e-13)
    assert_allclose(out_mass, oracle.readout(v.value, pos, res, scale=1.0, period=[4, 4, 4]), rtol=1e-12, atol=1e-13)
    # finite difference of paint along one particle coordinate
    if res != "lanczos3":          # table windows have piecewise-constant derivatives
        eps = 1e-6
        p2 = pos.copy(); p2[3, 1] += eps
        num = ((pm.paint(p2, mass=mass, mode="deterministic").value - pm.paint(pos, mass=mass, mode="deterministic").value) * v.value).sum() / eps
        assert_allclose(out_pos[3, 1], num, rtol=1e-4, atol=1e-6)

    # readout
    field = pm.create(type="real", value=rng.normal(size=(4, 4, 4)))
    vr = rng.normal(size=30)
    out_self, out_pos = field.readout_vjp(pos, vr)
    jvp = field.readout_jvp(pos, v_self=v, v_pos=vpos)
    lhs = (vr * jvp).sum()
    rhs = (out_self.value * v.value).sum() + (out_pos * vpos).sum()
    assert_allclose(lhs, rhs, rtol=1e-7)
    with pytest.raises(ValueError):
        field.readout_vjp(pos, vr, gradient=0)
    # aliases of the older pmesh names
    assert pm.paint_gradient.__func__ is pm.paint_vjp.__func__
    # c2r_vjp / r2c_vjp (test_gradient.py:69-101): <c2r(y), v> == <y, c2r_vjp(v)> in the real-dot sense
    c = field.r2c()
    g = v.c2r_vjp()
    rel_close(g.c2r().value / 64.0, v.value, 1e-10)


@pytest.mark.parametrize("res,n", [("cic", 16), ("tsc", 16)])
@pytest.mark.parametrize("dtype", ["f8", "f4"])
def test_force_step_vs_oracle(P, oracle, res, n, dtype):
    """The PM force step of examples/nbody.py:199-218 with device-resident particles against the
    same pipeline built from the oracle pieces (C paint/readout + numpy.fft + numpy transfer)."""
    from pmesh_b200 import transfer as T
    from pmesh_b200.device import DeviceArray
    L = 100.0
    pm = P.ParticleMesh(BoxSize=L, Nmesh=[n, n, n], dtype=dtype, resampler=res)
    rng = numpy.random.default_rng(43)
    X = rng.uniform(0, L, (n ** 3, 3))
    dX = DeviceArray.from_host(X)
    layout = pm.decompose(dX, smoothing=1.0 * pm.resampler.support)
    rho = pm.create("real")
    pm.paint(dX, layout=layout, hold=False, out=rho)
    rho.scale(1.0 * pm.Nmesh.prod() / len(X))
    rhok = rho.r2c()
    F = numpy.empty_like(X)
    for d in range(3):
        F[:, d] = rhok.apply(T.GravityFD4(d)).c2r().readout(dX, layout=layout).to_host()

    mesh = numpy.zeros((n, n, n), dtype)
    oracle.paint(mesh, X, res, scale=n / L, period=[n] * 3)
    mesh = mesh * (float(n) ** 3 / len(X))
    ck = oracle.r2c(mesh.astype("f8"))
    Fw = numpy.empty_like(X)
    for d in range(3):
        fr = oracle.c2r(oracle.transfer(ck, [n] * 3, [L] * 3, "gravity_fd4", d), [n] * 3)
        Fw[:, d] = oracle.readout(fr.astype(dtype), X, res, scale=n / L, period=[n] * 3)
    tol = 1e-6 if dtype == "f8" else 2e-4
    rel_close(F, Fw, tol)


def test_conservation_properties_at_scale():
    """size-independent properties of the force step where the oracle is too slow (192^3 = 7 M
    particles; bench.py reports the same two numbers at the full 1024^3): the scatter conserves mass,
    sum(rho) == N, and the force step conserves momentum, |sum_p F(p)| << N rms(F)."""
    import ctypes
    from pmesh_b200 import _lib, nbody
    from pmesh_b200.device import DeviceArray
    from pmesh_b200.pm import ParticleMesh
    M = 192
    for window in ("cic", "tsc", "pcs"):
        pm = ParticleMesh(BoxSize=float(M), Nmesh=[M, M, M], dtype="f8", resampler=window)
        ctx = pm.ctx
        n = M ** 3
        X = DeviceArray.empty((n, 3), "f8")
        box = (ctypes.c_double * 3)(float(M), float(M), float(M))
        nn = (ctypes.c_int64 * 3)(M, M, M)
        _lib.check(ctx.lib.pmb_particles_lattice(ctx.handle, X.ptr, 8, n, 3, nn, box, 0.5, 3.0, 44, 0))
        rho = pm.paint(X)
        assert abs(rho.csum() - n) < 1e-10 * n
        F = nbody.force(pm, X)
        rms = (sum(f.dot(f) for f in F) / (3.0 * n)) ** 0.5
        assert rms > 0
        for f in F:
            assert abs(f.sum()) < 1e-9 * n * rms


@pytest.mark.parametrize("dtype", ["f8", "f4"])
@pytest.mark.parametrize("n0", [64, 128, 256, 512, 1024, 2048, 4096, 48])
def test_gradient_fields_fused_transfer_and_first_pass(P, oracle, dtype, n0):
    """pm.gradient_fields (pmb_fft_c2r_grad3: transfers folded into this library's own axis-0 inverse transform,
    pmb_ifft.cuh; cuFFT for the planes) == [rhok.apply(T_d).c2r()] and == the numpy oracle; every supported line
    length (64 .. 4096), and a length that takes the unfused calls (48)"""
    from pmesh_b200 import transfer as T
    n = (n0, 12, 10) if n0 >= 512 else (n0, 20, 18)
    box = [7.0, 9.0, 11.0]
    pm = P.ParticleMesh(BoxSize=box, Nmesh=n, dtype=dtype)
    tol = 1e-12 if dtype == "f8" else 2e-5
    x = numpy.random.default_rng(n0).normal(size=n).astype(dtype)
    rx = pm.create(type="real", value=x)
    rx.scale(1.75)                  # a pending scalar of the input rides along
    cx = rx.r2c()
    keep = cx.value.copy()
    ck = oracle.r2c(1.75 * x.astype("f8"))
    for make, name in ((T.GravityFD4, "gravity_fd4"), (T.GradientK, "gradient_k")):
        tfs = [make(d) for d in range(3)]
        fused = P.gradient_fields(cx, tfs)
        for d in range(3):
            one = cx.apply(tfs[d]).c2r()
            scale = abs(one.value).max()
            assert abs(fused[d].value - one.value).max() <= tol * scale, (name, d)
            want = oracle.c2r(oracle.transfer(ck, list(n), box, name, d), list(n))
            assert abs(fused[d].value - want).max() <= max(tol, 1e-6 if dtype == "f8" else 1e-4) * abs(want).max(), (name, d)
    assert numpy.array_equal(cx.value, keep), "the input modes are preserved"
    # into given RealFields, twice in a row (work buffers and plans are reused)
    outs = [pm.create(type="real") for d in range(3)]
    again = P.gradient_fields(cx, [T.GravityFD4(d) for d in range(3)], outs=outs)
    assert all(a is o for a, o in zip(again, outs))
    first = [o.value.copy() for o in outs]
    P.gradient_fields(cx, [T.GravityFD4(d) for d in range(3)], outs=outs)
    for d in range(3):
        assert numpy.array_equal(outs[d].value, first[d])
    # the last pass of r2c joins the kernel (pm.force_fields, pmb_fft_force3): same fields from the REAL input
    rx2 = pm.create(type="real", value=x)
    rx2.scale(1.75)
    keep_real = rx2.value.copy()
    for make in (T.GravityFD4, T.GradientK):
        tfs = [make(d) for d in range(3)]
        whole = P.force_fields(rx2, tfs)
        ref = P.gradient_fields(cx, tfs)
        for d in range(3):
            assert abs(whole[d].value - ref[d].value).max() <= tol * abs(ref[d].value).max(), (make.__name__, d)
    assert numpy.array_equal(rx2.value, keep_real), "the input field is preserved"
    # any other combination of transfers goes through apply + c2r
    mixed = P.gradient_fields(cx, [T.GravityFD4(0), T.GradientK(1)])
    assert abs(mixed[1].value - cx.apply(T.GradientK(1)).c2r().value).max() <= tol * abs(mixed[1].value).max()


@pytest.mark.parametrize("dtype", ["f8", "f4"])
def test_apply_gradients_and_device_reductions(P, oracle, dtype):
    """pm.apply_gradients (three gradient transfers in one pass over the modes) == three apply calls;
    csum / cdot / cnorm reduce on the device (pmb_field_sum / pmb_field_dot / pmb_cdot) == numpy"""
    from pmesh_b200 import transfer as T
    n = (12, 10, 16)
    pm = P.ParticleMesh(BoxSize=[7.0, 9.0, 11.0], Nmesh=n, dtype=dtype)
    tol = 1e-6 if dtype == "f8" else 1e-4
    rng = numpy.random.default_rng(3)
    x = rng.normal(size=n).astype(dtype)
    y = rng.normal(size=n).astype(dtype)
    rx, ry = pm.create(type="real", value=x), pm.create(type="real", value=y)
    cx, cy = rx.r2c(), ry.r2c()
    for make in (T.GravityFD4, T.GradientK):
        tfs = [make(d) for d in range(3)]
        fused = P.apply_gradients(cx, tfs)
        for d in range(3):
            rel_close(fused[d].value, cx.apply(tfs[d]).value, 1e-14 if dtype == "f8" else 1e-6)
    # a pending scalar of the input rides along
    rs = pm.create(type="real", value=x)
    rs.scale(3.5)
    cs = rs.r2c()
    fused = P.apply_gradients(cs, [T.GravityFD4(d) for d in range(3)])
    rel_close(fused[1].value, 3.5 * cx.apply(T.GravityFD4(1)).value, tol)
    # reductions
    rel_close(rx.csum(), x.astype("f8").sum(), tol)
    rel_close(rx.cdot(ry), (x.astype("f8") * y.astype("f8")).sum(), tol)
    rel_close(rx.cnorm(), (x.astype("f8") ** 2).sum(), tol)
    w = numpy.full(n[2] // 2 + 1, 2.0)
    w[0] = 1.0
    w[-1] = 1.0
    kx, ky = oracle.r2c(x.astype("f8")), oracle.r2c(y.astype("f8"))
    rel_close(cx.cnorm(), (abs(kx) ** 2 * w).sum(), tol)
    got = cx.cdot(cy)
    want = (kx * numpy.conj(ky) * w).sum()
    assert abs(got - want) <= tol * abs(want).max() + tol * (abs(kx) ** 2 * w).sum() ** 0.5 * (abs(ky) ** 2 * w).sum() ** 0.5
    # Parseval through the engine: cnorm of the modes == mean square of the field
    rel_close(cx.cnorm(), (x.astype("f8") ** 2).sum() / x.size, tol)
