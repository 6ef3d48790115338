"""Device-resident arrays.

``DeviceArray`` is the handle the pmesh-compatible API accepts wherever the
reference takes a numpy array of particle data (positions, masses, readout
results): pass numpy and every call pays a host<->device copy; pass a
DeviceArray and the data stays in HBM between calls.
"""
import numpy

from . import _lib


def _c_strides(shape, itemsize):
    st = []
    acc = itemsize
    for n in reversed(shape):
        st.append(acc)
        acc *= max(int(n), 1)
    return tuple(reversed(st))


class DeviceArray(object):
    """A C-contiguous (or strided view of a) block of device memory with numpy-like metadata."""

    def __init__(self, shape, dtype, ptr=None, strides=None, base=None, ctx=None):
        self.ctx = ctx or _lib.context()
        self.shape = tuple(int(s) for s in (shape if numpy.ndim(shape) else (shape,)))
        self.dtype = numpy.dtype(dtype)
        self.strides = tuple(strides) if strides is not None else _c_strides(self.shape, self.dtype.itemsize)
        self.base = base
        if ptr is None:
            self._owned = True
            self.nbytes_alloc = max(int(numpy.prod(self.shape, dtype="i8")) * self.dtype.itemsize, 16)
            self.ptr = self.ctx.malloc(self.nbytes_alloc)
        else:
            self._owned = False
            self.ptr = int(ptr)
            self.nbytes_alloc = 0

    def __del__(self):
        try:
            if getattr(self, "_owned", False) and self.ptr:
                self.ctx.free(self.ptr)
                self.ptr = 0
        except Exception:
            pass

    # ---- numpy-like metadata ----------------------------------------------------
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(numpy.prod(self.shape, dtype="i8"))

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    def __len__(self):
        return self.shape[0]

    @property
    def is_contiguous(self):
        return self.strides == _c_strides(self.shape, self.dtype.itemsize)

    # ---- construction / transfer ---------------------------------------------------
    @classmethod
    def empty(cls, shape, dtype="f8", ctx=None):
        return cls(shape, dtype, ctx=ctx)

    @classmethod
    def zeros(cls, shape, dtype="f8", ctx=None):
        a = cls(shape, dtype, ctx=ctx)
        a.ctx.memset(a.ptr, 0, a.nbytes)
        return a

    @classmethod
    def from_host(cls, array, dtype=None, ctx=None):
        h = numpy.ascontiguousarray(array, dtype=dtype)
        a = cls(h.shape, h.dtype, ctx=ctx)
        if h.nbytes:
            a.ctx.h2d(a.ptr, h, h.nbytes)
        return a

    def to_host(self, out=None):
        """blocking copy to a numpy array (contiguous arrays and views with a contiguous hull)"""
        if self.is_contiguous:
            h = numpy.empty(self.shape, dtype=self.dtype) if out is None else out
            assert h.flags.c_contiguous and h.nbytes == self.nbytes
            if self.nbytes:
                self.ctx.d2h(h, self.ptr, self.nbytes)
            return h
        # strided view: copy the enclosing hull, then view it
        if self.size == 0:
            r = numpy.empty(self.shape, self.dtype)
        else:
            assert all(s >= 0 for s in self.strides)
            extent = sum((n - 1) * s for n, s in zip(self.shape, self.strides)) + self.dtype.itemsize
            hull = numpy.empty(extent, dtype="u1")
            self.ctx.d2h(hull, self.ptr, extent)
            r = numpy.array(numpy.ndarray(self.shape, self.dtype, buffer=hull, strides=self.strides))
        if out is not None:
            out[...] = r
            return out
        return r

    def set(self, array):
        h = numpy.ascontiguousarray(array, dtype=self.dtype)
        assert self.is_contiguous and h.shape == self.shape, (h.shape, self.shape)
        if h.nbytes:
            self.ctx.h2d(self.ptr, h, h.nbytes)
        return self

    def copy(self):
        assert self.is_contiguous
        a = DeviceArray(self.shape, self.dtype, ctx=self.ctx)
        self.ctx.d2d(a.ptr, self.ptr, self.nbytes)
        return a

    def fill_zero(self):
        assert self.is_contiguous
        self.ctx.memset(self.ptr, 0, self.nbytes)
        return self

    # ---- element-wise column arithmetic (pmb_axpy / pmb_lincomb) --------------------------------
    def _flat(self):
        """(ptr, byte stride, n) of the array seen as a 1-D strided column"""
        if self.is_contiguous:
            return self.ptr, self.dtype.itemsize, self.size
        if self.ndim == 1:
            return self.ptr, self.strides[0], self.shape[0]
        raise ValueError("element-wise device arithmetic needs a contiguous array or a 1-D strided column")

    def iadd_scaled(self, x, a=1.0):
        """self += a * x   (numpy: ``self[...] += x * a``)"""
        assert x.shape == self.shape and x.dtype == self.dtype and self.dtype.kind == 'f'
        py, sy, n = self._flat()
        px, sx, _ = x._flat()
        _lib.check(self.ctx.lib.pmb_axpy(self.ctx.handle, py, sy, px, sx, float(a), self.dtype.itemsize, n))
        return self

    def assign_lincomb(self, x, a=1.0, y=None, b=1.0):
        """self = a * x + b * y   (y None: self = a * x)"""
        assert x.shape == self.shape and x.dtype == self.dtype and self.dtype.kind == 'f'
        po, so, n = self._flat()
        px, sx, _ = x._flat()
        py, sy = (None, 0)
        if y is not None:
            assert y.shape == self.shape and y.dtype == self.dtype
            py, sy, _ = y._flat()
        _lib.check(self.ctx.lib.pmb_lincomb(self.ctx.handle, po, so, px, sx, float(a), py, sy, float(b),
                                            self.dtype.itemsize, n))
        return self

    def imod(self, period):
        """self %= period (numpy's floored modulo), in place"""
        assert self.dtype.kind == 'f'
        px, sx, n = self._flat()
        _lib.check(self.ctx.lib.pmb_column_mod(self.ctx.handle, px, sx, float(period), self.dtype.itemsize, n))
        return self

    def sum(self):
        """sum of all elements in float64 (device reduction; float32 / float64 arrays)"""
        import ctypes
        assert self.dtype.kind == 'f' and self.dtype.itemsize in (4, 8)
        ptr, stride, n = self._flat()
        out = ctypes.c_double(0.0)
        sz = (ctypes.c_int64 * 3)(n)
        st = (ctypes.c_int64 * 3)(stride)
        _lib.check(self.ctx.lib.pmb_field_sum(self.ctx.handle, ptr, self.dtype.itemsize, 1, sz, st, ctypes.byref(out)))
        return out.value

    def dot(self, other):
        """sum(self * other) in float64 (device reduction)"""
        import ctypes
        assert other.shape == self.shape and other.dtype == self.dtype and self.dtype.kind == 'f'
        px, sx, n = self._flat()
        py, sy, _ = other._flat()
        out = ctypes.c_double(0.0)
        _lib.check(self.ctx.lib.pmb_dot(self.ctx.handle, px, sx, py, sy, self.dtype.itemsize, n, ctypes.byref(out)))
        return out.value

    def column(self, d):
        """view of column d of an (N, k) array"""
        assert self.ndim == 2
        return DeviceArray((self.shape[0],), self.dtype, ptr=self.ptr + d * self.strides[1],
                           strides=(self.strides[0],), base=self, ctx=self.ctx)

    def __array__(self, dtype=None, copy=None):
        h = self.to_host()
        return h if dtype is None else h.astype(dtype)

    def __repr__(self):
        return "DeviceArray(shape=%s, dtype=%s, ptr=0x%x)" % (self.shape, self.dtype, self.ptr)


def is_device(x):
    return isinstance(x, DeviceArray)


class PinnedArray(object):
    """page-locked host memory exposed as a numpy array (``.array``); used for fast H2D / D2H staging"""
    def __init__(self, shape, dtype, ctx=None):
        import ctypes
        self.ctx = ctx or _lib.context()
        self.shape = tuple(int(s) for s in (shape if numpy.ndim(shape) else (shape,)))
        self.dtype = numpy.dtype(dtype)
        nbytes = max(int(numpy.prod(self.shape, dtype="i8")) * self.dtype.itemsize, 16)
        self.ptr = self.ctx.malloc_host(nbytes)
        buf = (ctypes.c_char * nbytes).from_address(self.ptr)
        self.array = numpy.frombuffer(buf, dtype=self.dtype, count=int(numpy.prod(self.shape, dtype="i8"))).reshape(self.shape)

    def __del__(self):
        try:
            if self.ptr:
                self.array = None
                self.ctx.free_host(self.ptr)
                self.ptr = 0
        except Exception:
            pass
