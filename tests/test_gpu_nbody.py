"""The driver either side of the force (pmesh_b200.nbody <- examples/nbody.py): element-wise column
kernels bit-exact against numpy, force / 1-LPT / kick-drift-kick against the oracle's numpy restatement
(1e-6 relative, float64)."""
import numpy
import pytest
from numpy.testing import assert_array_equal, assert_allclose

pytestmark = pytest.mark.gpu


class Naive(object):
    """ the reference's `Naive` step factors with E(a) = 1 (examples/nbody.py:67-75) """
    @staticmethod
    def K(ai, af, ar):
        return 1.0 / (ar * ar) * (af - ai)

    @staticmethod
    def D(ai, af, ar):
        return 1.0 / (ar * ar * ar) * (af - ai)


def test_column_arithmetic_is_numpy_exact():
    from pmesh_b200.device import DeviceArray
    from pmesh_b200 import nbody
    rng = numpy.random.default_rng(3)
    for dt in ("f8", "f4"):
        n = 100003
        V = rng.normal(size=(n, 3)).astype(dt)
        S = rng.normal(size=(n, 3)).astype(dt)
        F = [rng.normal(size=n).astype(dt) for d in range(3)]
        k, dr = dt and numpy.dtype(dt).type(0.37), numpy.dtype(dt).type(-1.91)
        dV, dS = DeviceArray.from_host(V), DeviceArray.from_host(S)
        nbody.kick_drift(dV, [DeviceArray.from_host(f) for f in F], k, dS, dr)
        wV = V.copy()
        wV[...] += numpy.stack(F, axis=1) * k
        wS = S.copy()
        wS[...] += wV * dr
        assert_array_equal(dV.to_host(), wV)
        assert_array_equal(dS.to_host(), wS)
        nbody.kick_drift(dV, [DeviceArray.from_host(f) for f in F], k)          # kick only
        wV[...] += numpy.stack(F, axis=1) * k
        assert_array_equal(dV.to_host(), wV)
        assert_array_equal(dS.to_host(), wS)
        # X = S + Q, y += a x, strided column assignment
        X = DeviceArray.empty((n, 3), dt).assign_lincomb(dS, 1.0, dV, 1.0)
        assert_array_equal(X.to_host(), wS * numpy.dtype(dt).type(1) + wV * numpy.dtype(dt).type(1))
        X.iadd_scaled(dV, 0.5)
        assert_array_equal(X.to_host(), (wS + wV) + wV * numpy.dtype(dt).type(0.5))
        col = DeviceArray.from_host(F[1])
        X.column(2).assign_lincomb(col, 2.0)
        h = X.to_host()
        assert_array_equal(h[:, 2], F[1] * numpy.dtype(dt).type(2))
        assert_array_equal(h[:, 0], ((wS + wV) + wV * numpy.dtype(dt).type(0.5))[:, 0])


@pytest.mark.parametrize("window", ["cic", "tsc"])
def test_force_and_symp2_match_the_numpy_driver(oracle, window):
    from pmesh_b200 import nbody
    from pmesh_b200.pm import ParticleMesh
    n, L, Om0 = 16, 64.0, 0.3
    pm = ParticleMesh(BoxSize=L, Nmesh=[n, n, n], dtype="f8", resampler=window)
    Q = pm.generate_uniform_particle_grid(shift=0.0)
    rng = numpy.random.default_rng(12)
    S = rng.normal(scale=1.5, size=Q.shape)
    V = rng.normal(scale=0.1, size=Q.shape)
    F = nbody.force(pm, Q, S, factor=1.5 * Om0)
    want = oracle.nbody_force(Q, S, n, L, window, 1.5 * Om0)
    got = numpy.stack([f.to_host() for f in F], axis=1)
    assert_allclose(got, want, rtol=0, atol=1e-6 * abs(want).max())
    steps = numpy.linspace(0.1, 0.4, 4)
    st = nbody.symp2(pm, nbody.State(Q, S, V), steps, Naive, Om0)
    wS, wV = oracle.nbody_symp2(Q, S, V, n, L, window, steps, Naive.K, Naive.D, Om0)
    assert_allclose(st.S.to_host(), wS, rtol=0, atol=1e-6 * abs(wS).max())
    assert_allclose(st.V.to_host(), wV, rtol=0, atol=1e-6 * abs(wV).max())


def test_lpt1_from_whitenoise(oracle):
    """white noise -> linear density -> 1-LPT displacement (examples/nbody.py:245-270)"""
    from pmesh_b200 import nbody, transfer as T
    from pmesh_b200.pm import ParticleMesh
    n, L = 16, 64.0
    pm = ParticleMesh(BoxSize=L, Nmesh=[n, n, n], dtype="f8", resampler="cic")
    Q = pm.generate_uniform_particle_grid(shift=0.0)
    wn = pm.generate_whitenoise(120577, unitary=True)
    dlinear = wn.apply(T.GaussianLowpass(4.0))          # a smooth stand-in for sqrt(P(k) / V)
    want_k = oracle.transfer(oracle.whitenoise(numpy.zeros((n, n, n // 2 + 1), dtype="complex128"), 0, (n, n, n), 120577, True),
                             [n] * 3, [L] * 3, "gauss_lowpass", r=4.0)
    assert_allclose(dlinear.value, want_k, rtol=0, atol=1e-12)
    DX1 = nbody.lpt1(pm, dlinear, Q)
    want = oracle.nbody_lpt1(want_k, Q, n, L, "cic")
    assert_allclose(DX1.to_host(), want, rtol=0, atol=1e-6 * abs(want).max())
