"""ParticleMesh / RealField / ComplexField -- the pmesh.pm API (reference pmesh/pm.py) on one B200
per process.

Names, positional order and defaults follow the reference.  Fields own DEVICE
memory; ``Field.value`` (and ``field[...]``, numpy ufuncs, ``numpy.asarray(field)``) materialise a
host mirror for drop-in compatibility and mark it authoritative, so the next device operation
uploads it again.  The force-step path -- ``paint`` -> ``r2c`` -> ``apply(Transfer)`` -> ``c2r`` ->
``readout`` with ``DeviceArray`` particles -- never leaves the GPU.

Differences from the reference that a user can see (documented in DESIGN.md):
  * the process mesh is a slab decomposition ``np=[P]`` (the reference defaults to a 2-D pencil
    mesh for 3-D fields); with one process they coincide;
  * dtype 'f8' / 'f4' only (no c2c transforms);
  * ``apply`` runs recognised ``pmesh_b200.transfer`` objects on the GPU; any other callable is
    evaluated on the host slab by slab like the reference does (compatibility path);
  * ``ravel/unravel`` of a transposed complex field and the Fourier-space ``resample`` need the
    reference's distributed sort on more than one rank and raise NotImplementedError there; on one
    rank, and for real fields on any number of ranks, they work.
"""
import ctypes
import functools
import numbers
import operator
import warnings

import numpy
from numpy.lib.mixins import NDArrayOperatorsMixin as NDArrayLike

from . import _lib
from . import comm as _comm
from . import domain
from .device import DeviceArray, is_device
from .transfer import find_transfer
from .window import FindResampler, Affine

_gettype = type


def is_inplace(out):
    return out is Ellipsis


class slab(numpy.ndarray):
    pass


class xslab(list):
    def normp(self, p=2, zeromode=None):
        """ returns the p-norm of the vector, matching the broadcast shape (reference pm.py:122-137) """
        kk = (sum([abs(ki) ** p for ki in self]))
        if zeromode is not None:
            kk[kk == 0] = zeromode
        return kk


class slabiter(object):
    """iterate a host array slab by slab along the axis of the largest stride (reference pm.py:87-120)"""
    def __init__(self, field, value):
        if field.ndim == 2:
            axis = 2
            self.optimized_view = value[None, ...]
            self.nslabs = 1
            self.optx = [xx[None, ...] for xx in field.x]
            self.opti = [ii[None, ...] for ii in field.i]
        else:
            axissort = numpy.argsort(field._layout_strides)[::-1]
            axis = axissort[0]
            self.optimized_view = value.transpose(axissort)
            self.nslabs = field.shape[axis]
            self.optx = [xx.transpose(axissort) for xx in field.x]
            self.opti = [ii.transpose(axissort) for ii in field.i]
        self.axis = axis
        self.Nmesh = field.Nmesh
        self.BoxSize = field.BoxSize
        self.x = xslabiter(self, axis, self.nslabs, self.optx)
        self.i = xslabiter(self, axis, self.nslabs, self.opti)

    def __iter__(self):
        for irow in range(self.nslabs):
            s = self.optimized_view[irow].view(type=slab)
            s.x = [x[0] if d != self.axis else x[irow] for d, x in enumerate(self.optx)]
            s.i = [x[0] if d != self.axis else x[irow] for d, x in enumerate(self.opti)]
            s.BoxSize = self.BoxSize
            s.Nmesh = self.Nmesh
            yield s


class xslabiter(slabiter):
    """ iterating will yield the sparse coordinates of a list of slabs """
    def __init__(self, slabiter, axis, nslabs, optx):
        self.axis = axis
        self.BoxSize = slabiter.BoxSize
        self.Nmesh = slabiter.Nmesh
        self.nslabs = nslabs
        self.optx = optx

    def __iter__(self):
        for irow in range(self.nslabs):
            kk = [x[0] if d != self.axis else x[irow] for d, x in enumerate(self.optx)]
            s = xslab(kk)
            s.BoxSize = self.BoxSize
            s.Nmesh = self.Nmesh
            yield s


class _Base(object):
    """The physical (device) memory of a field; shared by in-place r2c / c2r partners.
    Stands in for pfft.LocalBuffer (reference pm.py:226)."""
    def __init__(self, pm):
        self.dev = DeviceArray.zeros((pm._alloc_elems,), pm.dtype)
        # State of the MEMORY, shared by every Field that views it (a real field and its in-place
        # complex partner, or two handles on the same base):
        # pending -- a scalar factor not yet multiplied into the device values (value == pending *
        #   memory).  scale() and the 1/prod(Nmesh) of r2c only update it; the linear consumers that
        #   multiply anyway (r2c, c2r, the transfer kernels) fold it in, everything else multiplies
        #   it in first;
        # version -- bumped by every write to the device values, so that a view can tell that its host
        #   mirror is older than the memory even if another view did the writing.
        self.pending = 1.0
        self.version = 0

    def __contains__(self, other):
        return other is self


class Field(NDArrayLike):
    """ Base class for RealField and ComplexField (reference pm.py:156-651). """
    def __repr__(self):
        return '%s:' % self.__class__.__name__ + repr(self.value)

    _HANDLED_TYPES = (numpy.ndarray, numbers.Number)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        out = kwargs.get('out', ())
        for x in inputs + out:
            if not isinstance(x, self._HANDLED_TYPES + (Field,)):
                return NotImplemented
        inputs = tuple(x.value if isinstance(x, Field) else x for x in inputs)
        if out:
            kwargs['out'] = tuple(x.value if isinstance(x, Field) else x for x in out)
        result = getattr(ufunc, method)(*inputs, **kwargs)

        def cast(result):
            if result.dtype == '?':
                return result
            if result.shape != self.shape:
                return result
            return self.pm.create(_gettype(self), value=result)
        if type(result) is tuple:
            return tuple(cast(x) for x in result)
        elif method == 'at':
            return None
        else:
            return cast(result)

    def _check_compatible(self, other):
        if isinstance(other, Field):
            if not isinstance(other, _gettype(self)):
                raise TypeError("type of two operands of cdot must be the same type")
        else:
            assert all(numpy.shape(other) == self.shape)

    def copy(self):
        r = self.pm.create(_gettype(self))
        if self._dev_valid:
            self.pm.ctx.d2d(r._base.dev.ptr, self._base.dev.ptr, self._base.dev.nbytes)
            r._mark_device_written(pending=self._pending)
        else:
            r._host[...] = self._host
            r._dev_valid = False
            r._host_valid = True
            r._host_version = r._base.version
        return r

    def __init__(self, pm, base=None):
        if base is None:
            base = _Base(pm)
        self._base = base
        self.pm = pm
        self.BoxSize = pm.BoxSize
        self.Nmesh = pm.Nmesh
        self.ndim = len(pm.Nmesh)
        L = pm._layout
        if isinstance(self, RealField):
            shape, start, estrides = L['i_shape'], L['i_start'], L['i_strides']
            self._dtype = pm.dtype
            self.cshape = numpy.array(pm.Nmesh, dtype='intp')
        elif isinstance(self, (TransposedComplexField, UntransposedComplexField)):
            if isinstance(self, UntransposedComplexField) and pm.comm.size > 1:
                raise NotImplementedError("the untransposed complex layout is only available on one rank")
            shape, start, estrides = L['o_shape'], L['o_start'], L['o_strides']
            self._dtype = numpy.dtype('c%d' % (2 * pm.dtype.itemsize))
            cs = numpy.array(pm.Nmesh, dtype='intp')
            cs[-1] = cs[-1] // 2 + 1
            self.cshape = cs
        else:
            raise TypeError("Only RealField and ComplexField. No more subclassing")
        self.shape = tuple(int(s) for s in shape)
        self.start = numpy.array(start, dtype='intp')
        self._layout_strides = tuple(int(s) * self._dtype.itemsize for s in estrides)
        self.size = int(numpy.prod(self.shape, dtype='i8'))
        # number of reals the padded real layout spans: shape[0] * (stride of axis 0)
        self._padded_reals = int(self.shape[0]) * int(estrides[0]) if len(self.shape) > 1 else int(2 * (pm.Nmesh[-1] // 2 + 1))
        self._dev = DeviceArray(self.shape, self._dtype, ptr=base.dev.ptr, strides=self._layout_strides,
                                base=base.dev, ctx=pm.ctx)
        # host mirror (lazily allocated); a fresh field is all zeros on the device
        self._host_arr = None
        self._host_valid = False
        self._host_version = -1
        self._dev_valid = True

        self.x = pm.create_coords(type(self), return_indices=False)
        self.i = pm.create_coords(type(self), return_indices=True)

        self.slices = tuple([slice(s, s + n) for s, n in zip(self.start, self.shape)])
        self.csize = functools.reduce(operator.mul, self.cshape, 1)

    # ------------------------------------------------------------------ host <-> device coherence
    @property
    def dtype(self):
        return self._dtype

    @property
    def _pending(self):
        return self._base.pending

    @_pending.setter
    def _pending(self, v):
        self._base.pending = float(v)

    @property
    def _host(self):
        if self._host_arr is None:
            self._host_arr = numpy.zeros(self.shape, dtype=self._dtype)
        return self._host_arr

    def _materialize(self):
        """multiply a pending scalar factor into the device values"""
        if self._pending != 1.0:
            factor, self._pending = self._pending, 1.0
            ctx = self.pm.ctx
            es = self.pm.dtype.itemsize
            # flat over the dense storage (for real fields this includes the r2c padding, harmlessly)
            nreal = self._padded_reals if isinstance(self, RealField) else 2 * self.size
            n = (ctypes.c_int64 * 3)(nreal)
            st = (ctypes.c_int64 * 3)(es)
            _lib.check(ctx.lib.pmb_field_scale(ctx.handle, self._dev.ptr, es, 0, 1, n, st, float(factor)))

    def _host_is_stale(self):
        """this view handed its host array out (it is host-authoritative), but ANOTHER view of the same memory
        (an in-place r2c / c2r partner, a field created with base=) has written the device since: the memory
        is newer than the mirror.  The device wins -- in the reference `.value` IS the one shared buffer, so
        the later writer wins there too; what is lost is a write made through a retained `.value` array after
        the other view's device operation, which the reference would have kept."""
        if not self._dev_valid and self._host_version != self._base.version:
            self._dev_valid = True
            self._host_valid = False
            return True
        return False

    def _sync_host(self):
        self._host_is_stale()
        # the device copy is authoritative unless this view handed its host array out (_dev_valid False)
        if self._dev_valid and (not self._host_valid or self._host_version != self._base.version):
            self._materialize()
            self._host[...] = self._dev.to_host()
            self._host_valid = True
            self._host_version = self._base.version

    @property
    def value(self):
        """host view of the local values.  Handing it out makes the host copy authoritative: the next
        device operation uploads it (the caller may have written through the returned array)."""
        self._sync_host()
        self._dev_valid = False
        self._host_version = self._base.version
        return self._host

    @value.setter
    def value(self, v):
        self._host[...] = v
        self._host_valid = True
        self._dev_valid = False
        self._host_version = self._base.version       # the mirror is newer than anything written so far

    def readonly_value(self):
        """host copy of the values without invalidating the device copy"""
        self._sync_host()
        r = self._host.view()
        r.flags.writeable = False
        return r

    def _device(self, absorb=False):
        """DeviceArray view of the field, uploading the host mirror if it is authoritative.
        absorb=True: the caller folds ``self._pending`` into its own arithmetic; otherwise the pending
        factor is multiplied in first."""
        self._host_is_stale()
        if not self._dev_valid:
            h = self._host
            if self.size:
                extent = sum((n - 1) * s for n, s in zip(self.shape, self._layout_strides)) + self._dtype.itemsize
                hull = numpy.zeros(extent, dtype='u1')
                numpy.ndarray(self.shape, self._dtype, buffer=hull, strides=self._layout_strides)[...] = h
                self.pm.ctx.h2d(self._dev.ptr, hull, extent)
            # the memory now holds exactly the host values
            self._base.pending = 1.0
            self._base.version += 1
            self._host_version = self._base.version
            self._dev_valid = True
        if not absorb:
            self._materialize()
        return self._dev

    def _mark_device_written(self, pending=1.0):
        self._dev_valid = True
        self._host_valid = False
        self._base.pending = float(pending)
        self._base.version += 1

    @property
    def flat(self):
        return self.value.flat

    def __getitem__(self, index):
        return self.value.__getitem__(index)

    def __setitem__(self, index, value):
        return self.value.__setitem__(index, value)

    def __array__(self, dtype=None, copy=None):
        return self.value if dtype is None else self.value.astype(dtype)

    # ------------------------------------------------------------------ device-side elementwise helpers
    def fill(self, value=0.0):
        """set every value (device)"""
        ctx = self.pm.ctx
        if value == 0:
            # the whole allocation, padding included: one memset
            ctx.memset(self._base.dev.ptr, 0, self._base.dev.nbytes)
        elif isinstance(self, RealField):
            sz = (ctypes.c_int64 * 3)(*self.shape)
            st = (ctypes.c_int64 * 3)(*self._layout_strides)
            _lib.check(ctx.lib.pmb_field_fill(ctx.handle, self._dev.ptr, self.pm.dtype.itemsize, self.ndim, sz, st, float(value)))
        else:
            raise NotImplementedError
        self._mark_device_written()
        return self

    def scale(self, factor):
        """value[...] *= factor on the device (the `rho[...] *= fac` step of the force, nbody.py:205-207)"""
        # lazy: the factor is carried along and folded into the next r2c / transfer kernel (or
        # multiplied in by one streaming pass as soon as anything else looks at the values)
        self._device(absorb=True)
        self._mark_device_written(pending=self._pending * float(factor))
        return self

    # ------------------------------------------------------------------ collective indexing (host-side, not hot)
    def _ctol(self, index):
        index = numpy.array(index, copy=True)
        if len(index) == self.ndim + 1:
            value = self.plain
            index1 = index[:-1]
        elif len(index) == self.ndim:
            value = self.value
            index1 = index
        else:
            raise IndexError("Only vector index in global indexing is supported. for complex append 0 or 1 for real and imag")
        index1[index1 < 0] += self.Nmesh[index1 < 0]
        if all(index1 >= self.start) and all(index1 < self.start + self.shape):
            return value, tuple(list(index1 - self.start) + list(index[self.ndim:]))
        else:
            return value, None

    def cgetitem(self, index):
        """ get a value from absolute index collectively (reference pm.py:287-296). """
        value, localindex = self._ctol(index)
        ret = value[localindex] if localindex is not None else 0
        return self.pm.comm.allreduce(ret)

    def csetitem(self, index, y):
        """ set a value at an absolute index collectively, maintaining Hermitian conjugation;
            returns the value actually set (reference pm.py:298-345). """
        index = numpy.array(index, copy=True)
        value, localindex = self._ctol(index)
        if isinstance(self, BaseComplexField):
            dualindex = numpy.negative(index)
            if len(dualindex) == self.ndim + 1:
                dualindex[-1] *= -1
            dualindex[:self.ndim] += self.Nmesh
            dualindex[:self.ndim] %= self.Nmesh
            unused, duallocalindex = self._ctol(dualindex)
        else:
            duallocalindex = None
        dualy = y
        if localindex is None:
            y = 0
        if duallocalindex is None:
            dualy = 0
        if len(index) == self.ndim + 1 and index[-1] == 1:
            dualy = -dualy
            if localindex is not None and duallocalindex is not None:
                if localindex == duallocalindex:
                    y = 0
                    dualy = 0
        elif len(index) == self.ndim:
            dualy = numpy.conjugate(dualy)
            if localindex is not None and duallocalindex is not None:
                if localindex == duallocalindex:
                    dualy = dualy.real
                    y = y.real
        if localindex is not None:
            value[localindex] = y
        if duallocalindex is not None:
            value[duallocalindex] = dualy
        return self.pm.comm.allreduce(y)

    @property
    def compressed(self):
        """ Whether the field is stored in the half-space (Hermitian compressed) format. """
        if self.Nmesh[-1] == self.cshape[-1]:
            return False
        elif self.Nmesh[-1] // 2 + 1 == self.cshape[-1]:
            return True
        else:
            raise ValueError("The mesh shape (%s) and the complex field shape (%s) are inconsistent." %
                             (str(self.Nmesh), str(self.cshape)))

    @property
    def slabs(self):
        return slabiter(self, self.value)

    def sort(self, out=None):
        warnings.warn("Use ravel instead of sort", DeprecationWarning, stacklevel=2)
        return self.ravel(out)

    def unsort(self, flatiter):
        warnings.warn("Use pm.unravel instead of unsort", DeprecationWarning, stacklevel=2)
        return self.unravel(flatiter)

    # the field resampling / I/O-order operations are compositions of the hot-path operators and live
    # in pmesh_b200/resample.py
    def ravel(self, out=None):
        """ Ravel the field to 'C'-order, partitioned by ranks (reference pm.py:389-424). """
        from . import resample as _rs
        return _rs.ravel(self, out)

    def unravel(self, flatiter):
        """ Unsort c-ordered field values to the field; self is updated (reference pm.py:426-448). """
        from . import resample as _rs
        return _rs.unravel(self, flatiter)

    def resample(self, out):
        """ Resample by filling 0 or truncating Fourier modes, onto the mesh of ``out`` (reference pm.py:479-547). """
        from . import resample as _rs
        return _rs.fourier_resample(self, out)

    def preview(self, Nmesh=None, axes=None, resampler=None, method=None):
        """ the mesh as a numpy array on every rank, optionally resampled / projected (reference pm.py:549-615). """
        from . import resample as _rs
        return _rs.preview(self, Nmesh, axes, resampler, method)

    def cast(self, type=None, out=None):
        """ cast the field object to the given type, maintaining the meaning of the field: real <->
            complex through r2c / c2r (reference pm.py:450-477) """
        type = _typestr_to_type(type) if type is not None else _gettype(self)
        if out is None:
            out = self.pm.create(type=type)
        elif not isinstance(out, type):
            out = self.pm.create(type=type, base=out._base)
        if isinstance(self, RealField) and isinstance(out, BaseComplexField):
            return self.r2c(out)
        if isinstance(self, BaseComplexField) and isinstance(out, RealField):
            return self.c2r(out)
        # same kind: the transposed and untransposed complex layouts coincide on one rank, the only
        # place this engine offers both
        if out is not self:
            out.value = self.value
        return out

    def apply(self, func, kind, out):
        """ implements all kinds of apply operations (reference pm.py:617-648) """
        if out is None:
            out = self.pm.create(type=_gettype(self))
        if is_inplace(out):
            out = self

        tf = find_transfer(func)
        if tf is not None and isinstance(self, BaseComplexField) and kind != tf.apply_kind:
            raise ValueError("transfer %s expects apply(kind=%r), got kind=%r" % (type(tf).__name__, tf.apply_kind, kind))
        if tf is not None and isinstance(self, BaseComplexField) and isinstance(out, BaseComplexField):
            ctx = self.pm.ctx
            src = self._device(absorb=True)
            params = (ctypes.c_double * 4)(*tf.params())
            box = (ctypes.c_double * 3)(*[float(b) for b in self.BoxSize])
            _lib.check(ctx.lib.pmb_transfer_scaled(self.pm._plan, tf.kind, int(tf.direction), params, box,
                                                   float(self._pending), src.ptr, out._dev.ptr))
            out._mark_device_written()
            return out

        # compatibility path: arbitrary python callable, evaluated on host slabs like the reference
        if isinstance(out, numpy.ndarray):
            assert out.shape == self.shape
            outvalue = out
        else:
            assert isinstance(out, _gettype(self))
            assert out.shape == self.shape
            outvalue = out.value
        myvalue = self.value
        outslabs = slabiter(self, outvalue)
        myslabs = slabiter(self, myvalue)
        for x, i, islab, oslab in zip(myslabs.x, myslabs.i, myslabs, outslabs):
            if kind == 'relative':
                oslab[...] = func(x, islab)
            elif kind == 'index':
                oslab[...] = func(i, islab)
            elif kind == 'absolute':
                oslab[...] = func(x, islab)
            elif kind == 'wavenumber':
                oslab[...] = func(x, islab)
            elif kind == 'circular':
                w = [ki * L / N for ki, L, N in zip(x, self.BoxSize, self.Nmesh)]
                oslab[...] = func(w, islab)
            else:
                raise ValueError("unknown kind of apply function.")
        return out


class RealField(Field):
    def __init__(self, pm, base=None):
        Field.__init__(self, pm, base)

    def r2c(self, out=None):
        """ Perform real to complex transformation: rfftn / prod(Nmesh) (reference pm.py:655-694). """
        if out is None:
            out = TransposedComplexField(self.pm)
        if is_inplace(out):
            out = self
        if out is self:
            out = TransposedComplexField(self.pm, base=self._base)
        assert isinstance(out, (BaseComplexField,))
        ctx = self.pm.ctx
        src = self._device(absorb=True)
        # PFFT normalization, same as FastPM (pm.py:692) -- carried as the pending factor of the
        # result together with whatever was pending on the input (the transform is linear)
        scale = float(numpy.prod(self.Nmesh ** -1.0))
        pending = self._pending * scale
        _lib.check(ctx.lib.pmb_fft_r2c(self.pm._plan, src.ptr, out._dev.ptr, 1.0))
        out._mark_device_written(pending=pending)
        if out._base is self._base:
            self._host_valid = False     # the real values are gone
        return out

    def ctranspose(self, axes):
        """ Collectively transpose a RealField onto a ParticleMesh with permuted axes (reference pm.py:696-723). """
        from . import resample as _rs
        return _rs.ctranspose(self, axes)

    def csum(self, dtype=None):
        """ Collective sum of the entire mesh (reference pm.py:725-739). """
        if dtype is None:
            dtype = self.dtype
        if numpy.dtype(dtype) == self.dtype and self.ndim <= 3:
            # device reduction (float64 accumulation), then the scalar allreduce
            a = self._device(absorb=True)
            out = ctypes.c_double(0.0)
            sz = (ctypes.c_int64 * 3)(*a.shape)
            st = (ctypes.c_int64 * 3)(*a.strides)
            _lib.check(self.pm.ctx.lib.pmb_field_sum(self.pm.ctx.handle, a.ptr, a.dtype.itemsize, len(a.shape), sz, st, ctypes.byref(out)))
            return self.pm.comm.allreduce(self.dtype.type(out.value * self._pending))
        v = self.readonly_value()
        arg = numpy.argsort(self._layout_strides)
        sum1 = v.transpose(arg[::-1])
        for d in range(self.ndim):
            sum1 = sum1.sum(axis=-1, dtype=dtype)
        return self.pm.comm.allreduce(sum1)

    def cmean(self, dtype=None):
        """ Collective mean of the entire mesh. """
        return self.csum(dtype=dtype) / self.csize

    def readout(self, pos, hsml=None, out=None, resampler=None, transform=None, gradient=None, layout=None):
        """
        Read out from real field at positions (reference pm.py:745-791).

        pos : (N, ndim) positions in simulation units, numpy or DeviceArray
        hsml : per-particle scaling of the resampling window, or None
        gradient : None or the direction of the window derivative
        layout : Layout from pm.decompose; positions are routed to the owning ranks and the results
                 reduced back (ghost sum)
        """
        if not transform:
            transform = self.pm.affine
        if resampler is None:
            resampler = self.pm.resampler
        resampler = FindResampler(resampler)
        if layout is None:
            return resampler.readout(self._device(), pos, hsml=hsml, out=out, transform=transform, diffdir=gradient)
        else:
            localpos = layout.exchange(pos)
            localhsml = exchange(layout, hsml)
            localresult = self.readout(localpos, hsml=localhsml, resampler=resampler,
                                       transform=transform, gradient=gradient, out=None, layout=None)
            return layout.gather(localresult, out=out)

    def readout_vjp(self, pos, v, resampler=None, transform=None, gradient=None,
                    out_self=None, out_pos=None, layout=None):
        """ back-propagate the gradient of readout; returns (out_self, out_pos) (reference pm.py:793-846) """
        if out_pos is not False:
            if gradient is not None:
                raise ValueError("gradient of gradient is not yet supported")
            if out_pos is None:
                out_pos = numpy.zeros_like(numpy.asarray(pos))
            if is_inplace(out_pos):
                out_pos = pos
            if out_pos is pos:
                pos = pos.copy()
            for d in range(pos.shape[1]):
                self.readout(pos, out=out_pos[:, d], resampler=resampler, transform=transform, gradient=d, layout=layout)
                out_pos[:, d] *= v
        if out_self is not False:
            if out_self is None:
                out_self = RealField(self.pm)
            if is_inplace(out_self):
                out_self = self
            self.pm.paint(pos, mass=v, resampler=resampler, transform=transform, gradient=gradient, hold=False,
                          layout=layout, out=out_self)
        return out_self, out_pos

    readout_gradient = readout_vjp   # older pmesh spelling (north_star wording)

    def readout_jvp(self, pos, v_self=None, v_pos=None, resampler=None, transform=None, gradient=None, layout=None):
        """ f_i = W_qi A_q (reference pm.py:848-859) """
        jvp = numpy.zeros(len(pos))
        if v_pos is not None:
            for d in range(self.ndim):
                jvp[...] += self.readout(pos, resampler=resampler, transform=transform, gradient=d, layout=layout) * v_pos[..., d]
        if v_self is not None:
            jvp[...] += v_self.readout(pos, resampler=resampler, transform=transform, gradient=None, layout=layout)
        return jvp

    def paint(self, pos, mass=1.0, resampler=None, transform=None, hold=False, gradient=None, layout=None):
        warnings.warn("Use ParticleMesh.paint instead", DeprecationWarning, stacklevel=2)
        self.pm.paint(pos, mass=mass, resampler=resampler, transform=transform, hold=hold, gradient=gradient, layout=layout, out=self)

    def c2r_vjp(v, out=None):
        """ Back-propagate the gradient of c2r from self to out """
        out = v.r2c(out)
        out.scale(float(numpy.prod(out.pm.Nmesh ** 1.0)))
        return out

    def apply(self, func, kind="relative", out=None):
        """ apply func(r, y) to the field; kind 'relative' | 'index' | 'absolute' (reference pm.py:872-895) """
        assert kind in ['relative', 'index', 'absolute']
        return Field.apply(self, func, kind, out)

    def cdot(self, other):
        self._check_compatible(other)
        if isinstance(other, RealField) and other.pm is self.pm and self.ndim <= 3:
            # sum(self * other) on the device (pmb_field_dot), then the scalar allreduce (pm.py:897-902)
            a, b = self._device(), other._device()
            if a.strides == b.strides and a.shape == b.shape:
                out = ctypes.c_double(0.0)
                sz = (ctypes.c_int64 * 3)(*a.shape)
                st = (ctypes.c_int64 * 3)(*a.strides)
                _lib.check(self.pm.ctx.lib.pmb_field_dot(self.pm.ctx.handle, a.ptr, b.ptr, a.dtype.itemsize, len(a.shape), sz, st, ctypes.byref(out)))
                return self.pm.comm.allreduce(self.dtype.type(out.value * self._pending * other._pending))
        return self.pm.comm.allreduce(numpy.sum(self[...] * other[...]))

    def cnorm(self):
        return self.cdot(self)


class BaseComplexField(Field):
    def __init__(self, pm, base=None):
        Field.__init__(self, pm, base)

    @property
    def real(self):
        return self.value.real

    @property
    def imag(self):
        return self.value.imag

    @property
    def plain(self):
        return self.value.view(dtype=(self.value.real.dtype, 2))

    def _expand_hermitian(self, i, y):
        if not self.compressed:
            return y
        y = y.copy()
        mask = (i[-1] != 0) & (i[-1] != self.Nmesh[-1] // 2)
        y += mask * y
        return y

    def _device_cdot(self, other):
        """ rank-local sum over the stored modes of conj(other) * self, conjugate modes counted (pmb_cdot) """
        a, b = self._device(), other._device()
        res = (ctypes.c_double * 2)(0.0, 0.0)
        _lib.check(self.pm.ctx.lib.pmb_cdot(self.pm._plan, a.ptr, b.ptr, res))
        f = self._pending * other._pending
        return self.dtype.type(complex(res[0] * f, res[1] * f))

    _default_norm = None

    def cnorm(self, metric=None, norm=None):
        """ compute the norm collectively; the conjugates are added too (reference pm.py:920-943) """
        if norm is None:
            if metric is None and self.compressed:
                # |y|^2 summed on the device (float64 accumulation)
                return self.pm.comm.allreduce(self._device_cdot(self).real)
            norm = lambda x: x.real ** 2 + x.imag ** 2
        def filter2(k, y):
            y = norm(y)
            if metric is not None:
                k = k.normp(p=2) ** 0.5
                y *= metric(k)
            return y
        return self.pm.comm.allreduce(self.apply(filter2)
                                      .apply(self._expand_hermitian, kind='index', out=Ellipsis)
                                      .value.sum())

    def cdot(self, other, metric=None):
        """ Collective inner product between the independent modes of two Complex Fields (pm.py:945-974) """
        if isinstance(other, Field):
            if not isinstance(other, _gettype(self)):
                raise TypeError("type of two operands of cdot must be the same type")
        if metric is None and isinstance(other, BaseComplexField) and other.pm is self.pm and self.compressed:
            return self.pm.comm.allreduce(self._device_cdot(other))
        r = self.pm.create(type=_gettype(self), value=other)
        r.value[...] = numpy.conj(r.value[...])
        r.value[...] *= self.value
        r.apply(self._expand_hermitian, kind='index', out=Ellipsis)
        if metric is not None:
            r.apply(lambda k, y: y * metric(k.normp() ** 0.5), out=Ellipsis)
        return self.pm.comm.allreduce(r.value.sum())

    def cdot_vjp(self, v, metric=None):
        """ backtrace gradient of cdot against other (partial gradient; correct for cdot().real) """
        r = self * v
        if metric is not None:
            r.apply(lambda k, y: y * metric(k.normp() ** 0.5), out=Ellipsis)
        return r

    def c2r(self, out=None):
        """ complex to real: unnormalised inverse transform (reference pm.py:987-1019) """
        if out is None:
            out = RealField(self.pm)
        if is_inplace(out):
            out = self
        if out is self:
            out = RealField(self.pm, self._base)
        assert isinstance(out, RealField)
        ctx = self.pm.ctx
        src = self._device(absorb=True)
        pending = self._pending
        _lib.check(ctx.lib.pmb_fft_c2r(self.pm._plan, src.ptr, out._dev.ptr))
        out._mark_device_written(pending=pending)
        if out._base is self._base:
            self._host_valid = False
        return out

    def r2c_vjp(v, out=None):
        """ Back-propagate the gradient of r2c to self. """
        out = v.c2r(out)
        out.scale(float(numpy.prod(out.pm.Nmesh ** -1.0)))
        return out

    def decompress_vjp(v, out=None):
        """ Back-propagate the gradient of decompress from self to out (reference pm.py:1028-1045). """
        if out is None:
            out = v.pm.create(type=_gettype(v))
        if is_inplace(out):
            out = v
        for i, a, b in zip(out.slabs.i, out.slabs, v.slabs):
            mask = numpy.ones(a.shape, '?')
            for ii, n in zip(i, out.Nmesh):
                mask &= (n - ii) % n == ii
            a[~mask] = 2 * b[~mask]
            a[mask] = b[mask]
        return out

    def apply(self, func, kind="wavenumber", out=None):
        """ apply func(k, y) to the field; kind 'wavenumber' | 'circular' | 'index' (reference pm.py:1047-1070).
            ``pmesh_b200.transfer`` objects run on the GPU. """
        assert kind in ['wavenumber', 'circular', 'index']
        return Field.apply(self, func, kind, out)


class UntransposedComplexField(BaseComplexField):
    """ A complex field with untransposed representation (single rank only in this engine). """
    def __init__(self, pm, base=None):
        Field.__init__(self, pm, base)


class TransposedComplexField(BaseComplexField):
    """ A complex field with transposed representation: with P > 1 ranks it is distributed along
        axis 1 and stored in memory order (1, 2, 0). """
    def __init__(self, pm, base=None):
        Field.__init__(self, pm, base)


# backward-compatbility, alias TranposedComplexField to ComplexField
ComplexField = TransposedComplexField


def c2r_fields(fields, outs=None):
    """
    ``[f.c2r(out=o) for f, o in zip(fields, outs)]`` (``outs`` None: new RealFields; Ellipsis entries: in place)
    with the global transposes of the transforms overlapped with each other's local FFTs on P > 1 ranks
    (engine extension, pmb_fft_c2r_multi; values equal the separate calls).
    """
    pm = fields[0].pm
    if outs is None:
        outs = [None] * len(fields)
    res, cptr, rptr, pend = [], [], [], []
    for f, o in zip(fields, outs):
        assert isinstance(f, BaseComplexField) and f.pm is pm
        if o is None:
            o = RealField(pm)
        if is_inplace(o) or o is f:
            o = RealField(pm, f._base)
        assert isinstance(o, RealField)
        src = f._device(absorb=True)
        cptr.append(src.ptr)
        rptr.append(o._dev.ptr)
        pend.append(f._pending)
        res.append(o)
    for i in range(0, len(fields), 4):
        n = len(cptr[i:i + 4])
        ca = (ctypes.c_void_p * n)(*cptr[i:i + 4])
        ra = (ctypes.c_void_p * n)(*rptr[i:i + 4])
        _lib.check(pm.ctx.lib.pmb_fft_c2r_multi(pm._plan, n, ca, ra))
    for f, o, p in zip(fields, res, pend):
        o._mark_device_written(pending=p)
        if o._base is f._base:
            f._host_valid = False
    return res


def apply_gradients(field, transfers, outs=None):
    """
    ``[field.apply(t, out=o) for t, o in zip(transfers, outs)]`` for the three gradient transfers of a force
    or displacement evaluation (``GravityFD4(0..2)`` or ``GradientK(0..2)``, examples/nbody.py:154-170) in ONE
    pass over the modes of ``field`` (engine extension: the modes are read once, 1 / k^2 is formed once).
    Any other combination is applied one transfer at a time.
    """
    pm = field.pm
    if outs is None:
        outs = [pm.create(type=_gettype(field)) for _ in transfers]
    tfs = [find_transfer(t) for t in transfers]
    same = (len(tfs) == 3 and pm.ndim == 3 and all(t is not None for t in tfs)
            and tfs[0].kind in (_lib.TF_GRAVITY_FD4, _lib.TF_GRADIENT_K) and all(t.kind == tfs[0].kind for t in tfs)
            and [int(t.direction) for t in tfs] == [0, 1, 2] and isinstance(field, BaseComplexField)
            and all(isinstance(o, BaseComplexField) and o is not field for o in outs))
    if not same:
        return [field.apply(t, out=o) for t, o in zip(transfers, outs)]
    src = field._device(absorb=True)
    box = (ctypes.c_double * 3)(*[float(b) for b in field.BoxSize])
    ptrs = (ctypes.c_void_p * 3)(*[o._dev.ptr for o in outs])
    _lib.check(pm.ctx.lib.pmb_transfer_grad3(pm._plan, tfs[0].kind, box, float(field._pending), src.ptr, ptrs))
    for o in outs:
        o._mark_device_written()
    return list(outs)


def gradient_fields(field, transfers, outs=None):
    """
    ``[field.apply(t).c2r(out=o) for t, o in zip(transfers, outs)]`` for the three gradient transfers of a force or
    displacement evaluation (``GravityFD4(0..2)`` or ``GradientK(0..2)``; examples/nbody.py:154-170, 211-213) with
    the transfers FUSED into the first pass of the backward transforms (engine extension, pmb_fft_c2r_grad3: the
    density modes are multiplied as the axis-0 transform loads them; no pass over the modes for the transfer, no
    intermediate complex fields).  ``outs``: RealFields (None: new ones).  Equals
    ``c2r_fields(apply_gradients(field, transfers))`` to rounding, and runs exactly that for any other combination.
    """
    pm = field.pm
    n = len(transfers)
    if outs is None:
        outs = [None] * n
    outs = [RealField(pm) if o is None else o for o in outs]
    tfs = [find_transfer(t) for t in transfers]
    same = (n == 3 and pm.ndim == 3 and all(t is not None for t in tfs)
            and tfs[0].kind in (_lib.TF_GRAVITY_FD4, _lib.TF_GRADIENT_K) and all(t.kind == tfs[0].kind for t in tfs)
            and [int(t.direction) for t in tfs] == [0, 1, 2] and isinstance(field, BaseComplexField)
            and all(isinstance(o, RealField) and o.pm is pm and o._base is not field._base for o in outs)
            and len(set(id(o._base) for o in outs)) == 3)
    if not same:
        return c2r_fields(apply_gradients(field, transfers), outs=outs)
    src = field._device(absorb=True)
    box = (ctypes.c_double * 3)(*[float(b) for b in field.BoxSize])
    ptrs = (ctypes.c_void_p * 3)(*[o._dev.ptr for o in outs])
    _lib.check(pm.ctx.lib.pmb_fft_c2r_grad3(pm._plan, tfs[0].kind, box, float(field._pending), src.ptr, ptrs))
    for o in outs:
        o._mark_device_written()
    return list(outs)


def force_fields(rho, transfers, outs=None):
    """
    ``[rho.r2c().apply(t).c2r(out=o) for t, o in zip(transfers, outs)]`` -- the Fourier part of a force evaluation
    (examples/nbody.py:205-213) -- for the three gradient transfers.  On one rank the last pass of r2c, the transfers
    and the first pass of the three c2r are ONE kernel (engine extension, pmb_fft_force3): the density modes are never
    written to memory.  Anything else runs ``gradient_fields(rho.r2c(), transfers, outs)``.  ``rho`` is preserved.
    """
    pm = rho.pm
    n = len(transfers)
    tfs = [find_transfer(t) for t in transfers]
    same = (n == 3 and pm.ndim == 3 and pm.comm.size == 1 and isinstance(rho, RealField) and all(t is not None for t in tfs)
            and tfs[0].kind in (_lib.TF_GRAVITY_FD4, _lib.TF_GRADIENT_K) and all(t.kind == tfs[0].kind for t in tfs)
            and [int(t.direction) for t in tfs] == [0, 1, 2])
    if same:
        if outs is None:
            outs = [None] * n
        outs = [RealField(pm) if o is None else o for o in outs]
        same = (all(isinstance(o, RealField) and o.pm is pm and o._base is not rho._base for o in outs)
                and len(set(id(o._base) for o in outs)) == 3)
    if same:
        src = rho._device(absorb=True)
        box = (ctypes.c_double * 3)(*[float(b) for b in rho.BoxSize])
        ptrs = (ctypes.c_void_p * 3)(*[o._dev.ptr for o in outs])
        pre = float(rho._pending) * float(numpy.prod(pm.Nmesh.astype('f8') ** -1.0))
        rc = pm.ctx.lib.pmb_fft_force3(pm._plan, tfs[0].kind, box, pre, src.ptr, ptrs)
        if rc == 0:
            for o in outs:
                o._mark_device_written()
            return list(outs)
        if rc != -5:          # anything but PMB_EUNSUPPORTED
            _lib.check(rc)
    return gradient_fields(rho.r2c(), transfers, outs=outs)


def readout_fields(fields, pos, resampler=None, transform=None, layout=None, gather=None, remote=None):
    """
    Read several RealFields of one ParticleMesh at the same positions in ONE sweep over the particles
    (engine extension; the force step's three components, examples/nbody.py:211-216, share the pass
    over the positions).  Equals ``[f.readout(pos, layout=layout) for f in fields]`` value for value.

    pos : DeviceArray (N, ndim);  returns a list of DeviceArray (N,)
    layout : exchange ``pos`` first and reduce the results back (like RealField.readout)
    gather : ``pos`` is ALREADY the result of ``gather.exchange(...)``; the results are reduced back to the
             original particles, ``[gather.gather(c) for c in columns]``, without the gather's own pass over
             the particles: the kernel writes the results of the rank's own particles straight into the
             gathered columns and only the ghosts travel (sums agree with Layout.gather to rounding: the
             own value is added first instead of in rank order)
    remote : ``(layout, layout.exchange_remote(pos))`` with ``pos`` the ORIGINAL particles: nothing of the rank's own
             block is moved -- every particle is read where it lies (the kernels clip to the local canvas), the ghosts
             received from other ranks are read separately, their partial sums travel back and are added in rank order
             (same sums as ``gather=``; Layout.exchange_remote)
    """
    pm = fields[0].pm
    if not transform:
        transform = pm.affine
    resampler = FindResampler(pm.resampler if resampler is None else resampler)
    if layout is not None:
        pos = layout.exchange(pos)
        gather = layout
    if remote is not None:
        # pos are the ORIGINAL particles (never moved), remote = (layout, layout.exchange_remote(pos)): every particle is
        # read where it lies -- the kernels clip to the local canvas, so a particle gets the partial sum of its points
        # on this rank -- the ghosts received from other ranks are read separately and their partial sums travel back
        lay, rpos = remote
        own = resampler.readout_multi([f._device() for f in fields], pos, transform=transform)
        if lay.comm.size == 1:
            return own
        ghosts = (resampler.readout_multi([f._device() for f in fields], rpos, transform=transform) if rpos.shape[0]
                  else [DeviceArray.empty((1,), 'f8') for _ in fields])
        return lay.gather_add_ghosts(ghosts, own)
    plan = gather.fused_gather_plan() if (gather is not None and is_device(pos)) else None
    # the fused kernel exists for the CIC window on 3-D meshes with at least 2^18 particles (pmb_readout_multi_gather)
    if plan is not None and len(fields) <= 3 and resampler.kind == 'tunedcic' and pm.ndim == 3 and pos.shape[0] >= (1 << 18):
        own_begin, own_count, own_index = plan
        nghost = int(pos.shape[0]) - own_count
        own = [DeviceArray.zeros((int(gather.sendlength),), 'f8') for _ in fields]
        ghosts = [DeviceArray.empty((max(nghost, 1),), 'f8') for _ in fields]
        if resampler.readout_multi_gather([f._device() for f in fields], pos, ghosts, own, own_index, own_begin, own_count,
                                          transform=transform):
            return gather.gather_add_ghosts(ghosts, own)
        del own, ghosts
    out = []
    for i in range(0, len(fields), 3):
        out += resampler.readout_multi([f._device() for f in fields[i:i + 3]], pos, transform=transform)
    if gather is not None:
        out = [gather.gather(o) for o in out]
    return out


def reindex(Nsrc, Ndest):
    """ index table between frequency axes of different lengths (reference pm.py:1128-1144) """
    from .resample import mode_table
    return mode_table(Nsrc, Ndest)


def exchange(layout, value):
    """ exchange per-particle columns; scalars are not exchanged (reference pm.py:1146-1157) """
    if value is None:
        return None
    if is_device(value):
        return layout.exchange(value)
    if numpy.isscalar(value):
        value = numpy.array(value)
    value = numpy.asarray(value)
    if value.ndim != 0:
        localvalue = layout.exchange(value)
    else:
        localvalue = value
    return localvalue


def _typestr_to_type(typestr):
    if not isinstance(typestr, type):
        if typestr == 'real':
            typestr = RealField
        elif typestr == 'complex':
            typestr = ComplexField
        elif typestr == 'transposedcomplex':
            typestr = TransposedComplexField
        elif typestr == 'untransposedcomplex':
            typestr = UntransposedComplexField
        else:
            raise ValueError('mode must be real or complex, or ')
    if not issubclass(typestr, Field):
        raise TypeError("mode must be a subclass of %s" % str(Field))
    return typestr


def _init_i_coords(layout, Nmesh, BoxSize, dtype):
    """ real-space coordinates x in [-L/2, L/2) and integer indices (reference pm.py:1178-1200) """
    x = []
    i_ind = []
    ndim = len(Nmesh)
    for d in range(ndim):
        t = numpy.ones(ndim, dtype='intp')
        t[d] = layout['i_shape'][d]
        i_indi = numpy.arange(t[d], dtype='intp') + layout['i_start'][d]
        ri = numpy.arange(t[d], dtype=dtype) + layout['i_start'][d]
        ri[ri >= Nmesh[d] // 2] -= Nmesh[d]
        xi = ri * BoxSize[d] / Nmesh[d]
        i_ind.append(i_indi.reshape(t))
        x.append(xi.reshape(t))
    return x, i_ind


def _init_o_coords(layout, Nmesh, BoxSize, dtype):
    """ wavenumbers k (Nyquist negative) and integer indices (reference pm.py:1202-1226) """
    k = []
    o_ind = []
    ndim = len(Nmesh)
    for d in range(ndim):
        s = numpy.ones(ndim, dtype='intp')
        s[d] = layout['o_shape'][d]
        o_indi = numpy.arange(s[d], dtype='intp') + layout['o_start'][d]
        wi = numpy.arange(s[d], dtype=dtype) + layout['o_start'][d]
        wi[wi >= Nmesh[d] // 2] -= Nmesh[d]
        wi *= (2 * numpy.pi / Nmesh[d])
        ki = wi * Nmesh[d] / BoxSize[d]
        ki_type = ki.astype(dtype)
        o_ind.append(o_indi.reshape(s))
        k.append(ki_type.reshape(s))
    return k, o_ind


class ParticleMesh(object):
    """
    ParticleMesh provides an interface to solve for forces with the particle mesh method
    (reference pm.py:1245-1488).

    Attributes
    ----------
    np      : process mesh; this engine uses slabs, np = [comm.size]
    comm    : communicator (default: world)
    Nmesh   : array of int, number of mesh points per side; the length is the dimension
    dtype   : 'f8' or 'f4'
    BoxSize : array of float
    domain  : :py:class:`pmesh_b200.domain.GridND` of the real-space slabs
    affine  : position -> local grid units ; affine_grid : global grid -> local grid units
    """
    def __init__(self, Nmesh, BoxSize=1.0, comm=None, np=None, dtype='f8',
                 plan_method='estimate', resampler='cic'):
        if comm is None:
            comm = _comm.world()
        self.comm = comm

        if len(Nmesh) == 1 and self.comm.size != 1:
            raise ValueError("Running 1d transforms on multiple ranks is not supported")
        if np is None:
            # the reference defaults to pfft.split_size_2d for 3-D meshes (pm.py:1319-1327); on one NVSwitch
            # node (<= 8 GPUs) slabs need one global transpose per transform instead of two, so they are the
            # default here.  np=[P0, P1] selects pencils.
            np = [] if len(Nmesh) == 1 else [self.comm.size]
        np = [int(p) for p in np]
        if len(np) > len(Nmesh) - 1 and len(Nmesh) > 1:
            raise ValueError("process mesh of %d dimensions for a %d-dimensional mesh" % (len(np), len(Nmesh)))
        if len(np) > 2:
            raise NotImplementedError("process meshes of more than 2 dimensions are not implemented")
        if int(numpy.prod(np, dtype='i8')) != self.comm.size and not (len(np) == 0 and self.comm.size == 1):
            raise ValueError("np must multiply to the communicator size")
        # a trailing (or leading) 1 makes it a slab decomposition
        if len(np) == 2 and np[1] == 1:
            np = [np[0]]
        self.np = np
        self._procmesh = (np[0], np[1]) if len(np) == 2 else (self.comm.size, 1)
        if self._procmesh[1] > 1 and len(Nmesh) != 3:
            raise NotImplementedError("pencil decompositions are implemented for 3-D meshes")
        self._use_padded = True

        dtype = numpy.dtype(dtype)
        if dtype not in (numpy.dtype('f8'), numpy.dtype('f4')):
            raise ValueError("dtype must be f8 or f4 (c2c transforms are not implemented)")
        if plan_method not in ("estimate", "measure", "exhaustive"):
            raise KeyError(plan_method)

        self.Nmesh = numpy.array(Nmesh, dtype='i8')
        self.ndim = len(self.Nmesh)
        self.BoxSize = numpy.empty(len(Nmesh), dtype='f8')
        self.BoxSize[:] = BoxSize
        self.dtype = dtype

        self.ctx = _lib.context()
        self.comm.ensure_device_comm(self.ctx)
        nm = (ctypes.c_int64 * 3)(*[int(n) for n in self.Nmesh])
        plan = ctypes.c_void_p()
        npm = (ctypes.c_int * 2)(*self._procmesh)
        _lib.check(self.ctx.lib.pmb_fft_create_np(self.ctx.handle, self.ndim, nm, dtype.itemsize, npm, ctypes.byref(plan)))
        self._plan = plan
        arrs = [(ctypes.c_int64 * 3)() for _ in range(6)]
        ra, ca = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self.ctx.lib.pmb_fft_layout(plan, *arrs, ctypes.byref(ra), ctypes.byref(ca)))
        names = ['i_start', 'i_shape', 'i_strides', 'o_start', 'o_shape', 'o_strides']
        self._layout = dict((n, numpy.array(list(a)[:self.ndim], dtype='intp')) for n, a in zip(names, arrs))
        self._alloc_elems = int(ra.value)

        # domain decomposition of real space = the FFT slabs (reference pm.py:1443-1461)
        edges = []
        for d in range(self.ndim):
            n = int(self.Nmesh[d])
            parts = self._procmesh[d] if d < 2 else 1
            if parts > 1:
                blk = (n + parts - 1) // parts          # FFTW / PFFT default block: ceil(n / P)
                edges.append(numpy.array([min(r * blk, n) for r in range(parts + 1)], dtype='intp'))
            else:
                edges.append(numpy.array([0, n], dtype='intp'))
        self._i_edges = edges
        shape = numpy.array([len(g) - 1 for g in edges], dtype='int32')
        size = numpy.prod(shape)
        ilocal = tuple(numpy.flatnonzero(iedges == istart)[0] for istart, iedges in zip(self._layout['i_start'], edges))
        ilocal = numpy.ravel_multi_index(ilocal, shape, mode='raise', order='C')
        ilocals = numpy.array(self.comm.allgather(int(ilocal)))
        DomainAssign = numpy.empty(size, dtype='int32')
        for irank, il in enumerate(ilocals):
            start = il * size // self.comm.size
            end = (il + 1) * size // self.comm.size
            DomainAssign[start:end] = irank
        self.domain = domain.GridND(edges, comm=self.comm, DomainAssign=DomainAssign)

        # Transform from simulation unit to local grid unit.
        self.affine = Affine(self.ndim,
                             translate=-self._layout['i_start'],
                             scale=1.0 * self.Nmesh / self.BoxSize,
                             period=self.Nmesh)
        # Transform from global grid unit to local grid unit.
        self.affine_grid = Affine(self.ndim,
                                  translate=-self._layout['i_start'],
                                  scale=1.0,
                                  period=self.Nmesh)
        self.resampler = FindResampler(resampler)
        self._coords = {}

    def __del__(self):
        try:
            if getattr(self, '_plan', None):
                self.ctx.lib.pmb_fft_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    # ------------------------------------------------------------------ small API
    @property
    def partition(self):
        """ dict view of the real / complex partition (stands in for pfft.Partition) """
        return self._layout

    def fft_library_ms(self, reset=False):
        """ milliseconds spent inside cuFFT exec calls (library time, reported separately) """
        ms = ctypes.c_float()
        _lib.check(self.ctx.lib.pmb_fft_library_ms(self._plan, ctypes.byref(ms), int(reset)))
        return ms.value

    @property
    def exchange_tuner(self):
        """ measured choice between the split and the full particle exchange of a force evaluation (domain.ExchangeTuner) """
        if getattr(self, "_exchange_tuner", None) is None:
            from .domain import ExchangeTuner
            self._exchange_tuner = ExchangeTuner(self.comm, self.ctx)
        return self._exchange_tuner

    def fft_fused_stats(self, reset=False):
        """ (ms, launches): time inside the fused transfer + axis-0 transform kernels (pmb_ifft.cuh) since the last reset """
        ms = ctypes.c_float()
        nl = ctypes.c_int64()
        _lib.check(self.ctx.lib.pmb_fft_fused_stats(self._plan, ctypes.byref(ms), ctypes.byref(nl), int(reset)))
        return ms.value, nl.value

    def fft_transpose_stats(self, reset=False):
        """ (ms, bytes): time inside the transpose kernels of the distributed transforms and the bytes they
            stored into other ranks' memory over NVLink, since the last reset """
        ms = ctypes.c_float()
        nb = ctypes.c_double()
        _lib.check(self.ctx.lib.pmb_fft_transpose_stats(self._plan, ctypes.byref(ms), ctypes.byref(nb), int(reset)))
        return ms.value, nb.value

    def create_coords(self, field_type, return_indices=False):
        """ coordinate arrays (or integer indices) broadcastable to the field (reference pm.py:1505-1531) """
        field_type = _typestr_to_type(field_type)
        key = RealField if issubclass(field_type, RealField) else BaseComplexField
        if key not in self._coords:
            if key is RealField:
                self._coords[key] = _init_i_coords(self._layout, self.Nmesh, self.BoxSize, self.dtype)
            else:
                self._coords[key] = _init_o_coords(self._layout, self.Nmesh, self.BoxSize, self.dtype)
        x, i = self._coords[key]
        if return_indices:
            return [ii.copy() for ii in i]
        return [xx.copy() for xx in x]

    def resize(self, Nmesh):
        warnings.warn("ParticleMesh.resize method is deprecated. Use reshape method with full Nmesh as a tuple.", DeprecationWarning, stacklevel=2)
        return self.reshape(Nmesh=Nmesh)

    def reshape(self, Nmesh=None, BoxSize=None, dtype=None, resampler=None):
        """ a new ParticleMesh with some parameters replaced (reference pm.py:1541-1600) """
        if Nmesh is None:
            Nmesh = self.Nmesh
        if numpy.isscalar(Nmesh):
            Nmesh = [Nmesh] * self.ndim
        return ParticleMesh(Nmesh=Nmesh,
                            BoxSize=self.BoxSize if BoxSize is None else BoxSize,
                            comm=self.comm, np=self.np,
                            dtype=self.dtype if dtype is None else dtype,
                            resampler=self.resampler if resampler is None else resampler)

    def create(self, type=None, base=None, value=None, mode=None):
        """
            Create a field object (reference pm.py:1602-1634).

            type: 'real', 'complex', 'untransposedcomplex', or the Field classes
            base : reuse the physical memory of an existing field (`obj._base`)
            value : initialise the field with the values
        """
        if mode is not None:
            warnings.warn("argument mode is deprecated. use type=%s instead" % mode, DeprecationWarning, stacklevel=2)
            if type is None:
                type = mode
            else:
                raise ValueError("both mode and type are specified, possiblity arguments are arranged in wrong order")
        type = _typestr_to_type(type)
        r = type(self, base=base)
        if value is not None:
            r[...] = value
        return r

    def unravel(self, type, flatiter):
        """ Unravel c-ordered field values into a new field of the given type (reference pm.py:1636-1654). """
        r = self.create(type=type)
        r.unravel(flatiter)
        return r

    def generate_whitenoise(self, seed, unitary=False, mean=0, type=TransposedComplexField, mode=None, base=None):
        """ Generate white noise to the field with the given seed (reference pm.py:1656-1696).

            The scheme is compatible with Gadget / N-GenIC for three-dimensional meshes and does not
            depend on the number of ranks.

            seed : int
            mean : float, the mean of the field (the k = 0 mode)
            unitary : True for a unitary white noise (amplitude fixed to 1, only the phase is random)
            type : the field to return; a RealField is the c2r of the complex noise
        """
        from .whitenoise import generate
        if mode is not None:
            warnings.warn("mode argument is deprecated, use type", DeprecationWarning, stacklevel=2)
            type = mode
        type = _typestr_to_type(type)
        complex_type = TransposedComplexField if type is RealField else type
        complex = self.create(type=complex_type, base=base)
        generate(complex._dev, complex.start, complex.Nmesh, seed, bool(unitary))
        complex._mark_device_written()
        if mean != 0:
            # the generator leaves the k = 0 mode at zero (pm.py:1685-1691 sets it to `mean`)
            complex.csetitem([0] * self.ndim, mean)
        if type is RealField:
            return complex.c2r(out=Ellipsis)
        return complex

    def mesh_coordinates(self, dtype=None):
        coord = numpy.indices(tuple(self._layout['i_shape']), dtype).reshape(self.ndim, -1).T
        return coord + self._layout['i_start']

    def generate_uniform_particle_grid(self, shift=None, dtype=None, return_id=False):
        """
            uniform grid of particles, one per (local) mesh point, in BoxSize coordinates
            (reference pm.py:1705-1752; the default dtype is numpy's, quirk Q8).
        """
        if shift is None:
            warnings.warn(
                "calling generate_uniform_particle_grid without a shift argument is deprecated."
                "use shift=0.5 for the previous default behavior. ", DeprecationWarning, 2)
            shift = 0.5
        shift = numpy.broadcast_to(shift, self.ndim)
        source = self.mesh_coordinates(dtype)
        source[...] += shift
        source[...] *= self.BoxSize / self.Nmesh
        source.flags.writeable = False
        if not return_id:
            return source
        isource = self.mesh_coordinates('i4')
        id = numpy.int64(isource[:, 0])
        for i in range(1, self.ndim):
            id[...] *= self.Nmesh[i]
            id[...] += isource[:, i]
        return source, id

    # ------------------------------------------------------------------ the hot path
    def decompose(self, pos, smoothing=None, transform=None):
        """
        Create a domain decompose layout for particles at given coordinates (reference pm.py:1754-1793).

        smoothing : None (use self.resampler), a window / window name (0.5 * support), or a number /
                    array in mesh cells: the size of the buffer region around a domain.
        """
        if smoothing is None:
            smoothing = self.resampler
        try:
            smoothing = FindResampler(smoothing)
            smoothing = smoothing.support * 0.5
        except TypeError:
            pass
        if transform is None:
            transform = self.affine
        # Transform from simulation unit to global grid unit (the shift is local per rank: not used)
        return self.domain.decompose(pos, smoothing=smoothing, transform=domain.ScaleTransform(transform.scale))

    def paint(self, pos, hsml=None, mass=1.0, resampler=None, transform=None, hold=False, gradient=None,
              layout=None, out=None, mode=None):
        """
        Paint particles into a RealField (reference pm.py:1795-1869).

        pos : (N, ndim) positions in simulation units (numpy or DeviceArray)
        mass : scalar or (N,) weights;  hsml : per-particle window scaling or None
        hold : if True, add to ``out`` instead of clearing it first
        gradient : None or the direction of the window derivative
        layout : Layout from decompose(); particles are first exchanged to the owning ranks
        mode : 'atomic' | 'deterministic' | None (engine extension, see pmesh_b200.window)
        """
        if not transform:
            transform = self.affine
        if resampler is None:
            resampler = self.resampler
        resampler = FindResampler(resampler)
        if out is None:
            out = self.create(type=RealField)
        if not hold:
            out.fill(0.0)
        if layout is None:
            mesh = out._device()
            resampler.paint(mesh, pos, hsml=hsml, mass=mass, transform=transform, diffdir=gradient, mode=mode)
            out._mark_device_written()
            return out
        else:
            localpos = layout.exchange(pos)
            localmass = exchange(layout, mass)
            localhsml = exchange(layout, hsml)
            return self.paint(localpos, mass=localmass, hsml=localhsml, resampler=resampler,
                              transform=transform, hold=True, gradient=gradient, layout=None, out=out, mode=mode)

    def paint_jvp(self, pos, mass=1.0, v_pos=None, v_mass=None, resampler=None, transform=None, gradient=None, layout=None, out=None):
        """ A_q = W_qi M_i (reference pm.py:1872-1888) """
        assert gradient is None   # second order is not supported yet
        if out is None:
            out = self.create(type=RealField)
        out.fill(0.0)
        if v_pos is not None:
            for d in range(pos.shape[1]):
                self.paint(pos, mass=v_pos[..., d] * mass,
                           resampler=resampler, transform=transform, gradient=d, hold=True, layout=layout, out=out)
        if v_mass is not None:
            self.paint(pos, mass=v_mass,
                       resampler=resampler, transform=transform, gradient=None, hold=True, layout=layout, out=out)
        return out

    def paint_vjp(self, v, pos, mass=1.0, resampler=None, transform=None, gradient=None,
                  out_pos=None, out_mass=None, layout=None):
        """ back-propagate the gradient of paint from v; returns (out_pos, out_mass) (reference pm.py:1890-1935) """
        if out_pos is not False:
            if gradient is not None:
                raise ValueError("gradient of gradient is not yet supported")
            if out_pos is None:
                out_pos = numpy.zeros_like(numpy.asarray(pos))
            if is_inplace(out_pos):
                out_pos = pos
            if out_pos is pos:
                pos = pos.copy()
            for d in range(pos.shape[1]):
                v.readout(pos, out=out_pos[:, d], resampler=resampler, transform=transform, gradient=d, layout=layout)
                out_pos[..., d] *= mass
        if out_mass is not False:
            if out_mass is None:
                out_mass = numpy.zeros(len(pos))
            if is_inplace(out_mass):
                out_mass = mass
            v.readout(pos, out=out_mass, resampler=resampler, transform=transform, gradient=gradient, layout=layout)
        return out_pos, out_mass

    paint_gradient = paint_vjp   # older pmesh spelling (north_star wording)

    def upsample(self, source, resampler=None, keep_mean=False):
        """ Resample an image by reading the source out at the pixel positions of this pm
            (reference pm.py:1937-1989).  keep_mean: conserve the mean rather than the total mass. """
        from . import resample as _rs
        return _rs.upsample(self, source, resampler, keep_mean)

    def downsample(self, source, resampler=None, keep_mean=False):
        """ Resample an image by painting the pixels of the source onto this pm (reference pm.py:1991-2027). """
        from . import resample as _rs
        return _rs.downsample(self, source, resampler, keep_mean)
