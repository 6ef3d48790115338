"""CPU ORACLE for the particle-mesh hot path -- TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  ``pmesh_b200`` never does.

Parity status: PINNED.  ``tests/test_oracle.py`` checks every function here
against (a) the reference's own compiled C/Cython (``oracle/_ref``, built by
``oracle/build_ref.py`` from ``/root/reference``), bit for bit, and (b) the
known-answer vectors of the reference's test-suite (``pmesh/tests/test_window.py``,
``test_domain.py``) restated in ``tests/golden``, and (c) whole-pipeline
vectors computed by the reference's own ``pmesh/pm.py`` running on single-rank
stand-ins for pfft / mpsort / mpi4py (``tests/golden/reference_pm.py``,
``tests/test_reference_pipeline.py``): force step, k grids, white noise -> LPT,
vjp / jvp.  Only the FFT arithmetic itself is "parity unpinned by the
reference" (PFFT cannot be built here): ``numpy.fft.rfftn(x)/N`` and
``irfftn(y)*N`` stand in for it on both sides.

Layers
------
* windows / paint / readout : ``oracle/pm_oracle.c`` (plain C, -ffp-contract=off)
* gridnd_fill               : ``oracle/pm_oracle.c``
* GridND.decompose          : numpy, restating pmesh/domain.py:561-652
* Layout.exchange / gather  : numpy simulation of the Alltoallv over all ranks
* r2c / c2r / transfer      : numpy.fft, restating pmesh/pm.py:655-694, 987-1019,
                              1202-1226 and examples/nbody.py:154-181
* white noise               : ``oracle/whitenoise_oracle.c`` (ranlxd1 in integer arithmetic +
                              the N-GenIC column scheme), pinned to the compiled reference
"""
import ctypes
import os
import subprocess

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "libpm_oracle.so")
_TABLES = os.path.join(HERE, "..", "pmesh_b200", "data", "window_tables.npz")

KINDS = dict(
    nearest=0, linear=1, cubic=2, quadratic=3,
    lanczos2=4, lanczos3=5, lanczos4=6, lanczos5=7, lanczos6=8,
    acg2=9, acg3=10, acg4=11, acg5=12, acg6=13,
    db6=14, db12=15, db20=16, sym6=17, sym12=18, sym20=19,
    tunednnb=20, tunedcic=21, tunedtsc=22, tunedpcs=23,
)
# public names of pmesh.window (window.py:230-255) -> kind strings
NAMES = dict(nnb="tunednnb", cic="tunedcic", tsc="tunedtsc", pcs="tunedpcs")
for _k in list(KINDS):
    if not _k.startswith("tuned"):
        NAMES[_k] = _k

MAXDIM = 8


def build(force=False):
    """gcc -O2 -ffp-contract=off pm_oracle.c whitenoise_oracle.c -> libpm_oracle.so"""
    srcs = [os.path.join(HERE, "pm_oracle.c"), os.path.join(HERE, "whitenoise_oracle.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(x) for x in srcs):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO] + srcs + ["-lm"])
    return _SO


class _Painter(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int), ("support", ctypes.c_int), ("ndim", ctypes.c_int),
        ("order", ctypes.c_int * MAXDIM),
        ("scale", ctypes.c_double * MAXDIM), ("translate", ctypes.c_double * MAXDIM),
        ("period", ctypes.c_ssize_t * MAXDIM),
        ("canvas", ctypes.c_void_p), ("elsize", ctypes.c_int),
        ("size", ctypes.c_ssize_t * MAXDIM), ("strides", ctypes.c_ssize_t * MAXDIM),
        ("table", ctypes.c_void_p), ("tablesize", ctypes.c_int),
        ("step", ctypes.c_double), ("hsupport", ctypes.c_double),
        ("pcs_scale_fix", ctypes.c_int),
    ]


_lib = None
_tables = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ora_paint.argtypes = [ctypes.POINTER(_Painter), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_ssize_t]
        _lib.ora_readout.argtypes = [ctypes.POINTER(_Painter), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_ssize_t]
        _lib.ora_paint.restype = None
        _lib.ora_readout.restype = None
    return _lib


def tables():
    global _tables
    if _tables is None:
        z = numpy.load(_TABLES)
        _tables = {}
        for name in z.files:
            if name.endswith("_meta"):
                continue
            step, support, hs = z[name + "_meta"]
            _tables[name] = (numpy.ascontiguousarray(z[name], dtype="f8"), float(step), float(support), float(hs))
    return _tables


def kind_of(name):
    """'cic' / 'CIC' / 'tunedcic' / int -> enum value"""
    if isinstance(name, (int, numpy.integer)):
        return int(name)
    n = name.lower()
    return KINDS[NAMES.get(n, n)]


def window_support(name, support=-1):
    s, ns = ctypes.c_int(), ctypes.c_int()
    lib().ora_window_support(kind_of(name), int(support), ctypes.byref(s), ctypes.byref(ns))
    return s.value, ns.value


def _painter(real, kind, support, order, scale, translate, period, pcs_scale_fix=0):
    assert real.dtype.kind == "f" and real.dtype.itemsize in (4, 8)
    nd = real.ndim
    p = _Painter()
    p.kind = kind_of(kind)
    p.support = int(support)
    p.ndim = nd
    for d in range(nd):
        p.order[d] = int(order[d])
        p.scale[d] = float(scale[d])
        p.translate[d] = float(translate[d])
        p.period[d] = int(period[d])
        p.size[d] = real.shape[d]
        p.strides[d] = real.strides[d]
    p.canvas = real.ctypes.data
    p.elsize = real.dtype.itemsize
    keep = None
    for tname, k in KINDS.items():
        if k == p.kind and tname in tables():
            vals, step, sup, hs = tables()[tname]
            p.table = vals.ctypes.data
            p.tablesize = len(vals)
            p.step = step
            p.hsupport = hs
            keep = vals
    p.pcs_scale_fix = int(pcs_scale_fix)
    return p, keep


def _affine(ndim, scale, translate, period):
    def arr(v, default, dt):
        a = numpy.empty(ndim, dtype=dt)
        a[...] = default if v is None else v
        return a
    return arr(scale, 1.0, "f8"), arr(translate, 0.0, "f8"), arr(period, 0, "intp")


def _pos(pos, ndim):
    pos = numpy.asarray(pos)
    return numpy.ascontiguousarray(pos[:, :ndim], dtype="f8")   # f4 -> f8 per element, as _window.pyx:159


def _col(v, n):
    if v is None:
        return None
    a = numpy.empty(n, dtype="f8")
    a[...] = v
    return a


def paint(real, pos, kind="cic", support=-1, mass=None, hsml=None, diffdir=None,
          scale=None, translate=None, period=None, pcs_scale_fix=0):
    """In-place scatter-add into ``real`` (any strides, f4/f8) -- ResampleWindow.paint, window.py:106-163."""
    nd = real.ndim
    order = numpy.zeros(nd, dtype=int)
    if diffdir is not None:
        order[diffdir] = 1
    scale, translate, period = _affine(nd, scale, translate, period)
    p, keep = _painter(real, kind, support, order, scale, translate, period, pcs_scale_fix)
    x = _pos(pos, nd)
    m, h = _col(mass, len(x)), _col(hsml, len(x))
    lib().ora_paint(ctypes.byref(p), x.ctypes.data, None if m is None else m.ctypes.data,
                    None if h is None else h.ctypes.data, len(x))
    return real


def readout(real, pos, kind="cic", support=-1, hsml=None, diffdir=None, out=None,
            scale=None, translate=None, period=None, pcs_scale_fix=0):
    """Gather from ``real`` -- ResampleWindow.readout, window.py:165-221 (out defaults to f8)."""
    nd = real.ndim
    order = numpy.zeros(nd, dtype=int)
    if diffdir is not None:
        order[diffdir] = 1
    scale, translate, period = _affine(nd, scale, translate, period)
    p, keep = _painter(real, kind, support, order, scale, translate, period, pcs_scale_fix)
    x = _pos(pos, nd)
    h = _col(hsml, len(x))
    res = numpy.zeros(len(x), dtype="f8")
    lib().ora_readout(ctypes.byref(p), x.ctypes.data, None if h is None else h.ctypes.data, res.ctypes.data, len(x))
    if out is None:
        return res
    out[...] = res      # cast to out dtype like `out[i] = value` in _window.pyx:205
    return out


# --------------------------------------------------------------------------- domain
def _pymod(a, b):
    return numpy.remainder(a, b)


def decompose_patches(pos, edges, smoothing, periodic=True, scale=None):
    """(sil, sir) int16 (ndim, N) of GridND.decompose, pmesh/domain.py:601-630.
    ``scale`` restates the transform of ParticleMesh.decompose (pm.py:1786-1790): x -> scale * x."""
    pos = numpy.asarray(pos)
    ndim = len(edges)
    n = len(pos)
    sm = numpy.empty(ndim, dtype="f8")
    sm[:] = smoothing
    sc = numpy.empty(ndim, dtype="f8")
    sc[:] = 1.0 if scale is None else scale
    sil = numpy.empty((ndim, n), dtype="i2")
    sir = numpy.empty((ndim, n), dtype="i2")
    for j in range(ndim):
        e = numpy.asarray(edges[j])
        shape_j = len(e) - 1
        x = pos[:, j] if scale is None else sc[j] * pos[:, j]
        if n == 0:
            continue
        if periodic:
            box = e[-1]
            c = _pymod(x, box)
            l = numpy.digitize(_pymod(c - sm[j], box), e, right=False)
            r = numpy.digitize(_pymod(c + sm[j], box), e, right=False)
            p = numpy.digitize(c, e, right=False)
            sil[j] = p - (p - l) % shape_j - 1
            sir[j] = p + (r - p) % shape_j
        else:
            l = numpy.digitize(x - sm[j], e, right=False)
            r = numpy.digitize(x + sm[j], e, right=False)
            sil[j] = (l - 1).clip(0, shape_j)
            sir[j] = r.clip(0, shape_j)
    return sil, sir


def default_assign(ndomains, nranks):
    """GridND.__init__ default DomainAssign, pmesh/domain.py:384-392"""
    if nranks >= ndomains:
        return numpy.arange(ndomains, dtype="int32")
    a = numpy.empty(ndomains, dtype="int32")
    for i in range(nranks):
        a[i * ndomains // nranks:(i + 1) * ndomains // nranks] = i
    return a


def degenerate_domains(edges):
    """GridND.__init__ DomainDegenerate, pmesh/domain.py:396-405"""
    shape = [len(e) - 1 for e in edges]
    dd = numpy.zeros(shape, dtype="int16")
    for i, e in enumerate(edges):
        e = numpy.asarray(e)
        d1 = (e[1:] == e[:-1]).reshape([-1 if ii == i else 1 for ii in range(len(shape))])
        dd[...] |= d1
    return dd.ravel()


def gridnd_fill(sil, sir, shape, nranks, periodic, degenerate, assign):
    """(counts int32[nranks], indices int32[sum]) -- both passes of _domain.pyx:9-122"""
    L = lib()
    ndim, n = sil.shape
    sil = numpy.ascontiguousarray(sil, dtype="i2")
    sir = numpy.ascontiguousarray(sir, dtype="i2")
    dims = numpy.ascontiguousarray(shape, dtype="int32")
    deg = numpy.ascontiguousarray(degenerate, dtype="i2")
    asg = numpy.ascontiguousarray(assign, dtype="int32")
    counts = numpy.zeros(nranks, dtype="int32")
    args = (dims.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(ndim), sil.ctypes.data_as(ctypes.c_void_p),
            sir.ctypes.data_as(ctypes.c_void_p), ctypes.c_ssize_t(n), ctypes.c_int(int(bool(periodic))),
            deg.ctypes.data_as(ctypes.c_void_p), asg.ctypes.data_as(ctypes.c_void_p))
    L.ora_gridnd_fill(ctypes.c_int(0), counts.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(nranks), *args, None)
    indices = numpy.empty(int(counts.sum()), dtype="int32")
    L.ora_gridnd_fill(ctypes.c_int(1), counts.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(nranks), *args,
                      indices.ctypes.data_as(ctypes.c_void_p))
    return counts, indices


def decompose(pos, edges, nranks, smoothing=0, periodic=True, assign=None, scale=None):
    """sendcounts, indices of GridND(edges, comm of size nranks).decompose(pos, smoothing)"""
    shape = numpy.array([len(e) - 1 for e in edges], dtype="int32")
    if assign is None:
        assign = default_assign(int(numpy.prod(shape)), nranks)
    pos = numpy.asarray(pos)
    if len(pos) == 0:
        return numpy.zeros(nranks, dtype="int32"), numpy.empty(0, dtype="int32")
    sil, sir = decompose_patches(pos, edges, smoothing, periodic, scale)
    return gridnd_fill(sil, sir, shape, nranks, periodic, degenerate_domains(edges), assign)


def exchange_all(datas, layouts):
    """Simulated Alltoallv of Layout._exchange (domain.py:173-206) for ALL ranks at once.
    datas[r] is rank r's array, layouts[r] = (sendcounts, indices). Returns the list of received arrays."""
    P = len(datas)
    sent = []
    for r in range(P):
        counts, ind = layouts[r]
        buf = numpy.asarray(datas[r]).take(ind, axis=0)
        off = numpy.concatenate([[0], numpy.cumsum(counts)])
        sent.append([buf[off[q]:off[q + 1]] for q in range(P)])
    return [numpy.concatenate([sent[r][q] for r in range(P)], axis=0) for q in range(P)]


def gather_all(localdatas, layouts, sendlengths):
    """Simulated Layout.gather(mode='sum') (domain.py:208-318) for all ranks at once:
    reverse Alltoallv, then bincount in `indices` order (fp64 accumulate, cast back)."""
    P = len(localdatas)
    outs = []
    for r in range(P):
        counts, ind = layouts[r]
        pieces = []
        for q in range(P):
            recvcounts_q = [layouts[rr][0][q] for rr in range(P)]    # what q received from each rank
            off = numpy.concatenate([[0], numpy.cumsum(recvcounts_q)])
            pieces.append(numpy.asarray(localdatas[q])[off[r]:off[r + 1]])
        back = numpy.concatenate(pieces, axis=0)
        outs.append(bincount_sum(ind, back, sendlengths[r]))
    return outs


def bincount_sum(indices, values, minlength, dtype=None):
    """bincountv of domain.py:26-48 for 1-D or (N, k) values"""
    values = numpy.asarray(values)
    dtype = values.dtype if dtype is None else dtype
    out = numpy.empty((minlength,) + values.shape[1:], dtype=dtype)
    for index in numpy.ndindex(*values.shape[1:]):
        sl = (Ellipsis,) + index
        out[sl] = numpy.bincount(indices, values[sl], minlength=minlength)
    return out


# --------------------------------------------------------------------------- FFT / transfer
def r2c(real):
    """RealField.r2c: rfftn / prod(Nmesh)   (pm.py:689-692)"""
    return numpy.fft.rfftn(real) / numpy.prod(real.shape)


def c2r(cplx, nmesh):
    """ComplexField.c2r: unnormalised inverse (pm.py:1017)"""
    return numpy.fft.irfftn(cplx, s=tuple(nmesh), axes=tuple(range(len(nmesh)))) * numpy.prod(nmesh)


def wavenumbers(nmesh, boxsize, dtype="f8"):
    """k arrays of _init_o_coords (pm.py:1202-1226) for the full half-complex grid, broadcastable"""
    nd = len(nmesh)
    ks = []
    for d in range(nd):
        n = nmesh[d]
        m = n // 2 + 1 if d == nd - 1 else n
        w = numpy.arange(m, dtype=dtype)
        w[w >= n // 2] -= n
        w *= (2 * numpy.pi / n)
        k = (w * n / boxsize[d]).astype(dtype)
        s = [1] * nd
        s[d] = m
        ks.append(k.reshape(s))
    return ks


def transfer(cplx, nmesh, boxsize, kind, direction=0, r=None):
    """The force-step transfer functions of examples/nbody.py:154-181 on a full complex grid."""
    k = wavenumbers(nmesh, boxsize)
    k2 = sum(ki ** 2 for ki in k)
    if kind == "gravity_fd4":
        k2 = k2 + 0 * cplx.real
        k2[k2 == 0] = 1.0
        C = boxsize[direction] / nmesh[direction]
        w = k[direction] * C
        kfinite = 1.0 / C * 1 / 6.0 * (8 * numpy.sin(w) - numpy.sin(2 * w))
        return 1j * kfinite / k2 * cplx
    if kind == "gradient_k":
        k2 = k2 + 0 * cplx.real
        k2[k2 == 0] = 1.0
        return 1j * k[direction] / k2 * cplx
    if kind == "inv_laplace":
        k2 = k2 + 0 * cplx.real
        k2[k2 == 0] = 1.0
        return -1. / k2 * cplx
    if kind == "gauss_lowpass":
        return numpy.exp(-0.5 * k2 * r ** 2) * cplx
    if kind == "ik":
        return 1j * k[direction] * cplx
    raise ValueError(kind)


# ------------------------------------------------------------------------------------------ white noise
def whitenoise_stream(seed, n):
    """the first n uniforms of the ranlxd1 stream seeded with `seed` (pmesh/gsl/ranlxd.c)"""
    L = lib()
    out = numpy.empty(n, dtype='f8')
    L.wn_oracle_stream(ctypes.c_ulong(seed), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n))
    return out


def whitenoise(complex, start, Nmesh, seed, unitary=False):
    """fill the local block `complex` (complex64 / complex128, any strides) of the Nmesh^3 Fourier mesh
    starting at `start`; restates pmesh/whitenoise.py:4-24 -> _whitenoise.pyx:25-45 (3-D only)"""
    assert complex.ndim == 3 and complex.dtype.kind == 'c'
    L = lib()
    A = ctypes.c_ssize_t * 3
    st = numpy.empty(3, dtype='intp'); st[:] = start
    nm = numpy.empty(3, dtype='intp'); nm[:] = Nmesh
    rc = L.wn_oracle_fill(ctypes.c_void_p(complex.ctypes.data), ctypes.c_int(complex.dtype.itemsize),
                            A(*nm), A(*st), A(*complex.shape), A(*complex.strides),
                            ctypes.c_uint(seed), ctypes.c_int(bool(unitary)))
    assert rc == 0
    return complex


# ------------------------------------------------------------------------------------------ the driver
def nbody_force(Q, S, nmesh, boxsize, window, factor=1.0):
    """examples/nbody.py:196-218 on one rank with numpy: the force at X = S + Q, shape (N, 3)"""
    n, L = nmesh, boxsize
    X = S + Q
    rho = numpy.zeros((n, n, n))
    paint(rho, X, window, scale=n / L, period=[n] * 3)
    rho[...] *= 1.0 * n ** 3 / len(X)
    rhok = r2c(rho)
    F = numpy.empty_like(Q)
    for d in range(3):
        fr = c2r(transfer(rhok, [n] * 3, [L] * 3, "gravity_fd4", d), [n] * 3)
        F[..., d] = readout(fr, X, window, scale=n / L, period=[n] * 3)
    return factor * F


def nbody_symp2(Q, S, V, nmesh, boxsize, window, time_steps, K, D, Om0):
    """examples/nbody.py:84-102 (symp2) with numpy in-place arithmetic; returns (S, V)"""
    S, V = S.copy(), V.copy()
    F = nbody_force(Q, S, nmesh, boxsize, window, 1.5 * Om0)
    for ai, af in zip(time_steps[:-1], time_steps[1:]):
        ac = (ai * af) ** 0.5
        V[...] += F * K(ai, ac, ai)
        S[...] += V * D(ai, af, ac)
        F[...] = nbody_force(Q, S, nmesh, boxsize, window, 1.5 * Om0)
        V[...] += F * K(ac, af, af)
    return S, V


def nbody_lpt1(dlinear, Q, nmesh, boxsize, window):
    """examples/nbody.py:262-270: DX1[:, d] = readout(c2r(i k_d / k^2 dlinear), Q)"""
    n, L = nmesh, boxsize
    DX1 = numpy.zeros_like(Q)
    for d in range(3):
        fr = c2r(transfer(dlinear, [n] * 3, [L] * 3, "gradient_k", d), [n] * 3)
        DX1[..., d] = readout(fr, Q, window, scale=n / L, period=[n] * 3)
    return DX1
