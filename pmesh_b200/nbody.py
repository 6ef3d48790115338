"""The steps either side of the PM force: the force itself as one call, the 1-LPT displacement and
the kick-drift-kick integrator -- the reference's canonical driver, ``examples/nbody.py``, with the
particle state resident on the GPU.

The reference ships these as an example script, not as library API; the functions keep its names
and its arithmetic (file:line cited per function) so that a script written against it carries over:

    Q = pm.generate_uniform_particle_grid(shift=0)         # (N, 3), host or DeviceArray
    S, V = ...                                             # displacement, velocity, DeviceArray (N, 3)
    symp2(pm, State(Q, S, V), time_steps, factors, Om0)

Everything here is orchestration of the public pmesh API (decompose / exchange / paint / r2c / apply /
c2r / readout / gather) plus the element-wise column kernels ``pmb_kick_drift`` / ``pmb_lincomb``.
"""
import ctypes

import numpy

from . import _lib
from . import transfer as T
from .device import DeviceArray, is_device


def _dev(a):
    return a if is_device(a) else DeviceArray.from_host(numpy.ascontiguousarray(a))


def force_transfer(direction):
    """ i kfinite_d / k^2 with the 4th-order finite difference (examples/nbody.py:162-170) """
    return T.GravityFD4(direction)


def dx1_transfer(direction):
    """ i k_d / k^2 (examples/nbody.py:154-160) """
    return T.GradientK(direction)


pot_transfer = T.InverseLaplace()            # examples/nbody.py:172-175


def lowpass_transfer(r):
    """ exp(-k^2 r^2 / 2) (examples/nbody.py:177-181) """
    return T.GaussianLowpass(r)


class State(object):
    """ Q: initial (grid) positions, S: displacement, V: velocity (examples/nbody.py:78-82) """
    def __init__(self, Q, S, V):
        self.Q = _dev(Q)
        self.S = _dev(S)
        self.V = _dev(V)


def position(state, out=None):
    """ X = S + Q (examples/nbody.py:198) """
    if out is None:
        out = DeviceArray.empty(state.Q.shape, state.Q.dtype)
    return out.assign_lincomb(state.S, 1.0, state.Q, 1.0)


def force(pm, Q, S=None, factor=1.0):
    """
    The particle-mesh force at X = S + Q (examples/nbody.py:196-218): decompose, paint, x N^3/N,
    r2c, then per direction transfer -> c2r -> readout with the ghost sum.

    Returns the force column-wise: a list of ndim DeviceArray (N,) -- the layout readout / gather
    produce and the kick kernel consumes.  ``factor`` multiplies the result (the reference's
    ``1.5 * Om0``); it rides on the density's pending scale, no pass of its own.
    """
    Q = _dev(Q)
    X = Q if S is None else DeviceArray.empty(Q.shape, Q.dtype).assign_lincomb(_dev(S), 1.0, Q, 1.0)
    layout = pm.decompose(X, smoothing=1.0 * pm.resampler.support)
    # P > 1: the particles a rank keeps are painted and read where they lie (every kernel clips to the local canvas);
    # only the records that change rank travel (Layout.exchange_remote)
    tuner = pm.exchange_tuner          # which of the two is faster is measured on the first evaluations
    from .window import default_paint_mode
    if default_paint_mode() != 'atomic':
        tuner.choice = 'full'          # the deterministic paint sums in the particle order of the full exchange
    split = tuner.begin() == 'split'
    if split:
        lrem = layout.exchange_remote(X)
        rho = pm.paint(X)
        if lrem.shape[0]:
            pm.paint(lrem, out=rho, hold=True)
    else:
        lpos = layout.exchange(X)
        rho = pm.paint(lpos)
    N = pm.comm.allreduce(len(X))
    rho.scale(1.0 * pm.Nmesh.prod() / N * factor)
    # the ndim force fields are kept side by side so that ONE sweep over the particles reads them all
    # (pm.readout_fields): positions, cell indices and weights are shared by the components
    # ... and the transfers are folded into the first pass of the backward transforms (pm.gradient_fields); on one
    # rank the last pass of r2c joins them (pm.force_fields) and the density modes are never written to memory
    from .pm import force_fields, gradient_fields, readout_fields
    tfs = [force_transfer(d) for d in range(pm.ndim)]
    if pm.comm.size == 1 and pm.ndim == 3:
        f = force_fields(rho, tfs)
    else:
        f = gradient_fields(rho.r2c(out=Ellipsis), tfs)
    if split:
        F = readout_fields(f, X, remote=(layout, lrem))
    else:
        F = readout_fields(f, lpos, gather=layout)
    tuner.end()
    return F


def lpt1(pm, dlinear, Q):
    """ 1-LPT displacement DX1[:, d] = readout(c2r(i k_d / k^2 dlinear), Q) (examples/nbody.py:262-270);
        returns a DeviceArray (N, ndim) """
    Q = _dev(Q)
    layout = pm.decompose(Q)
    lpos = layout.exchange(Q)
    DX1 = DeviceArray.empty(Q.shape, Q.dtype)
    tmp = pm.create('complex')
    for d in range(pm.ndim):
        f = dlinear.apply(dx1_transfer(d), out=tmp).c2r(out=Ellipsis)
        col = layout.gather(f.readout(lpos))
        DX1.column(d).assign_lincomb(col, 1.0)
    return DX1


def kick_drift(V, F, kick, S=None, drift=0.0):
    """ V += F * kick; S += V * drift in one pass over the particles (S None: kick only) """
    assert V.is_contiguous and (S is None or (S.is_contiguous and S.shape == V.shape))
    ncol = V.shape[1] if V.ndim == 2 else 1
    assert len(F) == ncol
    cols = (ctypes.c_void_p * ncol)(*[f.ptr for f in F])
    for f in F:
        assert f.is_contiguous and f.shape == (V.shape[0],) and f.dtype == V.dtype
    _lib.check(V.ctx.lib.pmb_kick_drift(V.ctx.handle, V.ptr, None if S is None else S.ptr, cols, ncol,
                                        float(kick), float(drift), V.dtype.itemsize, V.shape[0]))


def symp2(pm, state, time_steps, factors, Om0, callback=None):
    """
    second-order kick-drift-kick (examples/nbody.py:84-102):

        F = force(Q + S)
        for ai, af:  ac = sqrt(ai af)
            V += F K(ai, ac, ai);  S += V D(ai, af, ac);  F = force(Q + S);  V += F K(ac, af, af)

    factors : object with K(ai, af, ar) and D(ai, af, ar) (the reference's FastPM / Quinn / Naive)
    """
    K, D = factors.K, factors.D
    F = force(pm, state.Q, state.S, factor=1.5 * Om0)
    for ai, af in zip(time_steps[:-1], time_steps[1:]):
        ac = (ai * af) ** 0.5
        kick_drift(state.V, F, K(ai, ac, ai), state.S, D(ai, af, ac))
        F = force(pm, state.Q, state.S, factor=1.5 * Om0)
        kick_drift(state.V, F, K(ac, af, af))
        if callback is not None:
            callback(af, state)
    return state
