"""Resampling windows -- the pmesh.window API (reference pmesh/window.py) on B200 kernels.

Same names and call signatures as the reference: ``Affine``, ``ResampleWindow``
(``paint``, ``readout``, ``resize``, ``get_fwindow``, ``get_compensation``),
``FindResampler``, the ``windows`` / ``methods`` registry with its 24 names in
upper and lower case and as module globals (window.py:230-262).

Arrays may be numpy (host; copied to the GPU and back around each call, for
drop-in compatibility with the reference's tests) or ``DeviceArray`` (resident).
All arithmetic happens in libpmesh_b200.so; nothing here falls back to the CPU.

Extra keyword (not in the reference): ``mode`` = ``'atomic'`` (default,
red.global.add; order of additions undefined, results within 1e-6 / 1e-4 rel.
of the reference for f8 / f4 meshes) or ``'deterministic'`` (sort by cell and
sequential per-cell sums: bit-equal to the reference).  The default can be set
with the environment variable ``PMESH_B200_PAINT_MODE``.
"""
import ctypes
import os

import numpy
from numpy.lib.stride_tricks import as_strided

from . import _lib
from .device import DeviceArray, is_device

# kind string -> enum; numbering of the reference's C header (pmesh/_window_imp.h:4-28)
KINDS = dict(
    nearest=0, linear=1, cubic=2, quadratic=3,
    lanczos2=4, lanczos3=5, lanczos4=6, lanczos5=7, lanczos6=8,
    acg2=9, acg3=10, acg4=11, acg5=12, acg6=13,
    db6=14, db12=15, db20=16, sym6=17, sym12=18, sym20=19,
    tunednnb=20, tunedcic=21, tunedtsc=22, tunedpcs=23,
)

_MODES = {"atomic": _lib.PMB_MODE_ATOMIC, "deterministic": _lib.PMB_MODE_DETERMINISTIC}


def default_paint_mode():
    return os.environ.get("PMESH_B200_PAINT_MODE", "atomic")


# opt-in fix of the tuned-PCS derivative: the reference omits scale[d]
PCS_GRADIENT_SCALE_FIX = False


def _mkarr(var, shape, dtype):
    var = numpy.asarray(var, dtype=dtype)
    if numpy.isscalar(shape):
        shape = (int(shape),)
    if len(var.shape) == 0:
        return as_strided(var, shape=shape, strides=[0] * len(shape))
    r = numpy.empty(shape, dtype)
    r[...] = var
    return r


class Affine(object):
    """ Defines an affine Transformation, used by ResampleWindow (reference window.py:18-55).

        Parameters
        ----------
            translate : array_like, in integer mesh units.
            period : array_like in integer mesh units.
            scale : factor that multiples on position to obtain mesh units.
    """
    def __init__(self, ndim, scale=None, translate=None, period=None):
        if scale is None:
            scale = 1.0
        if translate is None:
            translate = 0
        if period is None:
            period = 0
        self.scale = _mkarr(scale, ndim, 'f8')
        self.period = _mkarr(period, ndim, 'intp')
        self.translate = _mkarr(translate, ndim, 'f8')
        self.ndim = ndim

    def rescale(self, amount):
        """ Returns a new Affine where the scale is multipled by amount. """
        return Affine(self.ndim, self.scale * amount, self.translate, self.period)

    def shift(self, amount):
        """ Returns a new Affine where the translate is shifted by amount (integer mesh units). """
        return Affine(self.ndim, self.scale, self.translate + amount, self.period)


def _as_float_array(a):
    """host array -> f4/f8 array (anything else is promoted to f8, like numpy.asarray + the
    reference's fused float/double memoryviews would require)"""
    a = numpy.asarray(a)
    if a.dtype not in (numpy.dtype('f4'), numpy.dtype('f8')):
        a = a.astype('f8')
    return a


class _Column(object):
    """a per-particle scalar column (mass / hsml / out) resolved to (ptr, elsize, stride, scalar)"""
    def __init__(self, value, n, default):
        self.keep = None
        self.ptr, self.elsize, self.stride, self.scalar = None, 8, 0, default
        if value is None:
            return
        if is_device(value):
            assert value.ndim == 1 and value.shape[0] == n, "column length mismatch"
            assert value.dtype in (numpy.dtype('f4'), numpy.dtype('f8'))
            self.ptr, self.elsize, self.stride = value.ptr, value.dtype.itemsize, value.strides[0]
            self.keep = value
            return
        v = _as_float_array(value)
        if v.ndim == 0:
            self.scalar = float(v)
            return
        v = numpy.ascontiguousarray(numpy.broadcast_to(v, (n,)))
        d = DeviceArray.from_host(v)
        self.ptr, self.elsize, self.stride, self.keep = d.ptr, d.dtype.itemsize, d.strides[0], d


class ResampleWindow(object):
    """A resampling window (reference window.py:57-221 over the Cython class _window.pyx:67-205).

    ``kind`` is one of the kind strings ('tunedcic', 'lanczos3', ...) or the integer enum of the
    reference's C header; ``support`` (-1: native) rescales the window.
    """
    def __init__(self, kind, support=-1):
        lib = _lib.load()
        self.kind = kind
        if kind in KINDS:
            self._kind = KINDS[kind]
        else:
            self._kind = int(kind)
        s, ns = ctypes.c_int(), ctypes.c_int()
        _lib.check(lib.pmb_window_query(self._kind, int(support), ctypes.byref(s), ctypes.byref(ns)))
        self.support = s.value
        self.nativesupport = ns.value

    def resize(self, support):
        """ Change the support of the window, returning a new window. """
        return ResampleWindow(self.kind, support)

    def get_compensation(self):
        """ Return a function that compensates the resampling window by deconvolving in Fourier
            space; use with ComplexField.apply(kind='circular') (reference window.py:65-80).
            The returned callable carries ``.transfer`` so that ``apply`` runs it on the GPU. """
        def function(w, v):
            tf = 1.0
            for wi in w:
                tf = tf * self.get_fwindow(wi)
            return v / tf
        from .transfer import Compensate
        function.transfer = Compensate(self)
        return function

    def get_fwindow(self, w):
        """ 1d fourier space window function T(w) of the resample window, w = circular frequency;
            1 for windows without an analytic transform (reference window.py:82-104). """
        w1d = numpy.ascontiguousarray(numpy.reshape(w, -1), dtype='float64')
        out = numpy.zeros_like(w1d)
        _lib.check(_lib.load().pmb_window_fwindow(self._kind, self.support, w1d.ctypes.data, out.ctypes.data, len(w1d)))
        return out.reshape(numpy.shape(w))

    # ------------------------------------------------------------------ marshalling
    def _args(self, mesh, pos, hsml, diffdir, transform):
        """common part of paint/readout: geometry + particle columns -> ResampleArgs"""
        ndim = mesh.ndim
        if not 1 <= ndim <= 3:
            raise _lib.PmbError("meshes of %d dimensions are not supported by the GPU kernels (1..3)" % ndim)
        if transform is None:
            transform = Affine(ndim)
        assert isinstance(transform, Affine)
        a = _lib.ResampleArgs()
        a.kind = self._kind
        a.support = self.support
        a.ndim = ndim
        for d in range(ndim):
            a.order[d] = 1 if (diffdir is not None and diffdir % ndim == d) else 0
            a.scale[d] = transform.scale[d]
            a.translate[d] = transform.translate[d]
            a.period[d] = transform.period[d]
            a.size[d] = mesh.shape[d]
            a.strides[d] = mesh.strides[d]
        a.mesh = mesh.ptr
        a.mesh_elsize = mesh.dtype.itemsize
        keep = []
        if is_device(pos):
            dpos = pos
            assert dpos.dtype in (numpy.dtype('f4'), numpy.dtype('f8'))
        else:
            dpos = DeviceArray.from_host(_as_float_array(pos))
        assert dpos.ndim == 2 and dpos.shape[1] >= ndim, "pos must be (N, >=ndim)"
        keep.append(dpos)
        n = dpos.shape[0]
        a.pos = dpos.ptr
        a.pos_elsize = dpos.dtype.itemsize
        a.npart = n
        a.pos_stride0, a.pos_stride1 = dpos.strides
        h = _Column(hsml, n, 1.0)
        keep.append(h)
        a.hsml, a.hsml_elsize, a.hsml_stride, a.hsml_scalar = h.ptr, h.elsize, h.stride, h.scalar
        a.pcs_gradient_scale_fix = int(PCS_GRADIENT_SCALE_FIX)
        return a, keep, n

    @staticmethod
    def _device_mesh(real, need_values=True):
        """host canvas -> contiguous device copy (returns DeviceArray, host_target or None)"""
        if is_device(real):
            if real.dtype.kind == 'c':
                # the real part of a complex canvas (window.py:161-162)
                ft = numpy.dtype('f%d' % (real.dtype.itemsize // 2))
                real = DeviceArray(real.shape, ft, ptr=real.ptr, strides=real.strides, base=real, ctx=real.ctx)
            assert real.dtype.kind == 'f'
            return real, None
        host = real
        if numpy.iscomplexobj(host):
            host = host.real
        assert host.dtype.kind == 'f'
        if host.dtype.itemsize not in (4, 8):
            raise _lib.PmbError("canvas dtype %s is not supported (float32 / float64)" % host.dtype)
        return DeviceArray.from_host(host), host

    def paint(self, real, pos, hsml=None, mass=None, diffdir=None, transform=None, mode=None):
        """
            paint to a field; original values are preserved (added to).

            real : ndarray or DeviceArray canvas, float32/float64, 1..3 dimensions, any strides
            pos : (N, ndim) positions;  mass : (N,) weights, scalar, or None for 1
            hsml : (N,) dimensionless scaling of the window support, scalar, or None
            diffdir : axis of differentiation or None;  transform : Affine (position -> grid units)
            mode : 'atomic' | 'deterministic' | None (module default)
        """
        mesh, host = self._device_mesh(real)
        a, keep, n = self._args(mesh, pos, hsml, diffdir, transform)
        m = _Column(1.0 if mass is None else mass, n, 1.0)
        a.mass, a.mass_elsize, a.mass_stride, a.mass_scalar = m.ptr, m.elsize, m.stride, m.scalar
        a.mode = _MODES[mode or default_paint_mode()]
        ctx = mesh.ctx
        ctx.ensure_tables()
        _lib.check(ctx.lib.pmb_paint(ctx.handle, ctypes.byref(a)))
        if host is not None:
            host[...] = mesh.to_host()
        del keep, m

    def readout(self, real, pos, hsml=None, out=None, diffdir=None, transform=None):
        """
            readout from a field at positions ``pos``; returns ``out`` (float64 zeros if None,
            reference window.py:201-202).  A DeviceArray ``pos`` with ``out=None`` gives a DeviceArray.
        """
        mesh, _ = self._device_mesh(real)
        a, keep, n = self._args(mesh, pos, hsml, diffdir, transform)
        host_out = None
        if out is None:
            dout = DeviceArray.empty((n,), 'f8')
            if not is_device(pos):
                host_out = numpy.zeros(numpy.shape(pos)[:-1], dtype='f8')
        elif is_device(out):
            dout = out
        else:
            host_out = out
            odt = out.dtype if out.dtype in (numpy.dtype('f4'), numpy.dtype('f8')) else numpy.dtype('f8')
            dout = DeviceArray.empty((n,), odt)
        assert dout.ndim == 1 and dout.shape[0] == n
        a.out, a.out_elsize, a.out_stride = dout.ptr, dout.dtype.itemsize, dout.strides[0]
        ctx = mesh.ctx
        ctx.ensure_tables()
        _lib.check(ctx.lib.pmb_readout(ctx.handle, ctypes.byref(a)))
        del keep
        if host_out is not None:
            host_out[...] = dout.to_host().reshape(host_out.shape)
            return host_out
        return dout

    def readout_multi(self, reals, pos, hsml=None, outs=None, transform=None):
        """readout of up to three canvases of identical shape / strides / dtype at the same positions in one
        sweep over the particles (pmb_readout_multi): cell indices and weights are computed once.  Each
        result equals ``readout(real, pos)`` bit for bit.  Returns a list of DeviceArray."""
        meshes = [self._device_mesh(r)[0] for r in reals]
        m0 = meshes[0]
        nf = len(meshes)
        assert 1 <= nf <= 3
        for m in meshes[1:]:
            assert m.shape == m0.shape and m.strides == m0.strides and m.dtype == m0.dtype, "canvases must share one geometry"
        a, keep, n = self._args(m0, pos, hsml, None, transform)
        if outs is None:
            outs = [DeviceArray.empty((n,), 'f8') for _ in meshes]
        for o in outs:
            assert is_device(o) and o.ndim == 1 and o.shape[0] == n and o.dtype == outs[0].dtype
        a.out_elsize = outs[0].dtype.itemsize
        mp = (ctypes.c_void_p * nf)(*[m.ptr for m in meshes])
        op = (ctypes.c_void_p * nf)(*[o.ptr for o in outs])
        os_ = (ctypes.c_int64 * nf)(*[o.strides[0] for o in outs])
        ctx = m0.ctx
        ctx.ensure_tables()
        _lib.check(ctx.lib.pmb_readout_multi(ctx.handle, ctypes.byref(a), nf, mp, op, os_))
        del keep
        return list(outs)

    def readout_multi_gather(self, reals, pos, ghost_outs, own_outs, own_index_ptr, own_begin, own_count, transform=None):
        """readout_multi fused with the routing of Layout.gather('sum') (pmb_readout_multi_gather): results of
        the rank's own particles [own_begin, own_begin + own_count) go to row own_index[k] of ``own_outs``,
        those of the ghosts to the compact columns ``ghost_outs``.  Returns False when no fused kernel exists
        for this window / geometry (nothing is written then)."""
        meshes = [self._device_mesh(r)[0] for r in reals]
        m0 = meshes[0]
        nf = len(meshes)
        for m in meshes[1:]:
            assert m.shape == m0.shape and m.strides == m0.strides and m.dtype == m0.dtype, "canvases must share one geometry"
        a, keep, n = self._args(m0, pos, None, None, transform)
        a.out_elsize = 8
        mp = (ctypes.c_void_p * nf)(*[m.ptr for m in meshes])
        gp = (ctypes.c_void_p * nf)(*[o.ptr for o in ghost_outs])
        op = (ctypes.c_void_p * nf)(*[o.ptr for o in own_outs])
        ctx = m0.ctx
        ctx.ensure_tables()
        rc = ctx.lib.pmb_readout_multi_gather(ctx.handle, ctypes.byref(a), nf, mp, gp, op, ctypes.c_void_p(own_index_ptr),
                                              int(own_begin), int(own_count))
        del keep
        if rc == -5:          # PMB_EUNSUPPORTED
            return False
        _lib.check(rc)
        return True

    def readout_grad(self, real, pos, hsml=None, transform=None, want_value=True):
        """value and all ndim gradients in one neighbour sweep (device arrays): (value | None, grad (N, ndim)).
        Each column equals readout(diffdir=d) bit for bit; this is the paint_vjp / readout_vjp helper."""
        mesh, _ = self._device_mesh(real)
        a, keep, n = self._args(mesh, pos, hsml, None, transform)
        val = DeviceArray.empty((n,), 'f8') if want_value else None
        grad = DeviceArray.empty((n, mesh.ndim), 'f8')
        if val is not None:
            a.out, a.out_stride = val.ptr, val.strides[0]
        a.out_elsize = 8
        ctx = mesh.ctx
        ctx.ensure_tables()
        _lib.check(ctx.lib.pmb_readout_grad(ctx.handle, ctypes.byref(a), grad.ptr, grad.strides[0], grad.strides[1]))
        del keep
        return val, grad


def FindResampler(window):
    if isinstance(window, str) and window in windows:
        window = windows[window]
    if not isinstance(window, ResampleWindow):
        raise TypeError("argument is not a ResampleWindow name or a ResampleWindow object")
    return window


windows = dict(
    NEAREST=ResampleWindow(kind="nearest"),
    LINEAR=ResampleWindow(kind="linear"),
    NNB=ResampleWindow(kind="tunednnb"),
    CIC=ResampleWindow(kind="tunedcic"),
    TSC=ResampleWindow(kind="tunedtsc"),
    PCS=ResampleWindow(kind="tunedpcs"),
    QUADRATIC=ResampleWindow(kind="quadratic"),
    CUBIC=ResampleWindow(kind="cubic"),
    LANCZOS2=ResampleWindow(kind="lanczos2"),
    LANCZOS3=ResampleWindow(kind="lanczos3"),
    LANCZOS4=ResampleWindow(kind="lanczos4"),
    LANCZOS5=ResampleWindow(kind="lanczos5"),
    LANCZOS6=ResampleWindow(kind="lanczos6"),
    ACG2=ResampleWindow(kind="acg2"),
    ACG3=ResampleWindow(kind="acg3"),
    ACG4=ResampleWindow(kind="acg4"),
    ACG5=ResampleWindow(kind="acg5"),
    ACG6=ResampleWindow(kind="acg6"),
    DB6=ResampleWindow(kind="db6"),
    DB12=ResampleWindow(kind="db12"),
    DB20=ResampleWindow(kind="db20"),
    SYM6=ResampleWindow(kind="sym6"),
    SYM12=ResampleWindow(kind="sym12"),
    SYM20=ResampleWindow(kind="sym20"),
)

for m, p in list(windows.items()):
    windows[m.lower()] = p
    globals()[m] = p

# compatible.
methods = windows
del m, p
