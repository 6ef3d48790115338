"""ctypes binding of libpmesh_b200.so (C ABI declared in include/pmesh_b200.h).

There is no fallback: if the shared library is missing or a call fails, an
exception is raised.  One context (one CUDA stream) per process/GPU.
"""
import ctypes
import os

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpmesh_b200.so")

c_i64 = ctypes.c_int64
c_i64_3 = ctypes.c_int64 * 3
c_int_3 = ctypes.c_int * 3
c_dbl_3 = ctypes.c_double * 3

PMB_MODE_ATOMIC = 0
PMB_MODE_DETERMINISTIC = 1

TF_SCALE, TF_GRAVITY_FD4, TF_GRADIENT_K, TF_INV_LAPLACE, TF_GAUSS_LOWPASS, TF_COMPENSATE, TF_IK, TF_POWERLAW = range(8)


class PmbError(RuntimeError):
    pass


class ResampleArgs(ctypes.Structure):
    """struct pmb_resample_args"""
    _fields_ = [
        ("kind", ctypes.c_int), ("support", ctypes.c_int), ("ndim", ctypes.c_int),
        ("order", c_int_3), ("scale", c_dbl_3), ("translate", c_dbl_3), ("period", c_i64_3),
        ("mesh", ctypes.c_void_p), ("mesh_elsize", ctypes.c_int),
        ("size", c_i64_3), ("strides", c_i64_3),
        ("pos", ctypes.c_void_p), ("pos_elsize", ctypes.c_int), ("npart", c_i64),
        ("pos_stride0", c_i64), ("pos_stride1", c_i64),
        ("mass", ctypes.c_void_p), ("mass_elsize", ctypes.c_int), ("mass_stride", c_i64),
        ("mass_scalar", ctypes.c_double),
        ("hsml", ctypes.c_void_p), ("hsml_elsize", ctypes.c_int), ("hsml_stride", c_i64),
        ("hsml_scalar", ctypes.c_double),
        ("out", ctypes.c_void_p), ("out_elsize", ctypes.c_int), ("out_stride", c_i64),
        ("mode", ctypes.c_int), ("pcs_gradient_scale_fix", ctypes.c_int),
    ]


class DecomposeArgs(ctypes.Structure):
    """struct pmb_decompose_args"""
    _fields_ = [
        ("pos", ctypes.c_void_p), ("pos_elsize", ctypes.c_int), ("npart", c_i64),
        ("pos_stride0", c_i64), ("pos_stride1", c_i64),
        ("ndim", ctypes.c_int), ("scale", c_dbl_3), ("smoothing", c_dbl_3),
        ("edges_h", ctypes.c_void_p), ("nedges", c_int_3), ("periodic", ctypes.c_int),
        ("domain_assign_h", ctypes.c_void_p), ("domain_degenerate_h", ctypes.c_void_p),
        ("nranks", ctypes.c_int),
    ]


# every exported symbol of include/pmesh_b200.h (checked by tests/test_abi.py)
SYMBOLS = [
    "pmb_last_error", "pmb_version", "pmb_device_count", "pmb_ctx_create", "pmb_ctx_destroy", "pmb_ctx_sync",
    "pmb_malloc", "pmb_free", "pmb_malloc_host", "pmb_free_host", "pmb_memcpy_h2d", "pmb_memcpy_d2h",
    "pmb_memcpy_d2d", "pmb_memset", "pmb_memcpy_h2d_async", "pmb_memcpy_d2h_async", "pmb_stream_record", "pmb_stream_wait", "pmb_stream_sync", "pmb_mem_info", "pmb_timer_start", "pmb_timer_stop", "pmb_launch_count",
    "pmb_flush_l2", "pmb_set_workspace_limit", "pmb_window_set_table", "pmb_window_query", "pmb_window_fwindow",
    "pmb_paint", "pmb_readout", "pmb_readout_multi", "pmb_readout_multi_gather", "pmb_readout_grad", "pmb_bin_stats", "pmb_bin_release", "pmb_bin_invalidate", "pmb_set_trim_callback", "pmb_ctx_trim", "pmb_field_fill", "pmb_field_scale", "pmb_field_sum", "pmb_field_dot",
    "pmb_axpy", "pmb_lincomb", "pmb_column_mod", "pmb_kick_drift", "pmb_dot",
    "pmb_particles_uniform", "pmb_particles_lattice", "pmb_particles_replicate",
    "pmb_decompose_count", "pmb_decompose_fill", "pmb_decompose_identity", "pmb_take", "pmb_gather_sum", "pmb_gather_sum_segments", "pmb_gather_add_segments",
    "pmb_comm_unique_id", "pmb_comm_init_rank", "pmb_comm_destroy", "pmb_comm_rank", "pmb_alltoallv",
    "pmb_allreduce_f64", "pmb_allgather_bytes", "pmb_barrier",
    "pmb_fft_create", "pmb_fft_create_np", "pmb_fft_destroy", "pmb_fft_layout", "pmb_fft_r2c", "pmb_fft_c2r", "pmb_fft_c2r_multi", "pmb_fft_library_ms", "pmb_fft_transpose_stats",
    "pmb_transfer", "pmb_transfer_scaled", "pmb_transfer_grad3", "pmb_fft_c2r_grad3", "pmb_fft_force3", "pmb_fft_fused_stats", "pmb_cdot", "pmb_whitenoise",
]

_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_int64
_Z = ctypes.c_size_t
_D = ctypes.c_double
_ARGTYPES = {
    "pmb_device_count": [_P],
    "pmb_ctx_create": [_I, _P], "pmb_ctx_destroy": [_P], "pmb_ctx_sync": [_P],
    "pmb_malloc": [_P, _Z, _P], "pmb_free": [_P, _P], "pmb_malloc_host": [_P, _Z, _P], "pmb_free_host": [_P, _P],
    "pmb_memcpy_h2d": [_P, _P, _P, _Z], "pmb_memcpy_d2h": [_P, _P, _P, _Z], "pmb_memcpy_d2d": [_P, _P, _P, _Z],
    "pmb_memset": [_P, _P, _I, _Z],
    "pmb_memcpy_h2d_async": [_P, _P, _P, _Z], "pmb_memcpy_d2h_async": [_P, _P, _P, _Z],
    "pmb_stream_record": [_P, _I, _I], "pmb_stream_wait": [_P, _I, _I], "pmb_stream_sync": [_P, _I], "pmb_mem_info": [_P, _P, _P],
    "pmb_timer_start": [_P, _I], "pmb_timer_stop": [_P, _I, _P], "pmb_launch_count": [_P, _P, _I],
    "pmb_flush_l2": [_P], "pmb_set_workspace_limit": [_P, _Z],
    "pmb_window_set_table": [_P, _I, _P, _I, _D, _D, _D],
    "pmb_window_query": [_I, _I, _P, _P], "pmb_window_fwindow": [_I, _I, _P, _P, _L],
    "pmb_paint": [_P, _P], "pmb_readout": [_P, _P], "pmb_readout_grad": [_P, _P, _P, _L, _L], "pmb_bin_stats": [_P, _P, _P], "pmb_bin_release": [_P], "pmb_bin_invalidate": [_P], "pmb_ctx_trim": [_P], "pmb_set_trim_callback": [_P, _P, _P],
    "pmb_readout_multi": [_P, _P, _I, _P, _P, _P],
    "pmb_readout_multi_gather": [_P, _P, _I, _P, _P, _P, _P, _L, _L],
    "pmb_field_fill": [_P, _P, _I, _I, _P, _P, _D], "pmb_field_scale": [_P, _P, _I, _I, _I, _P, _P, _D],
    "pmb_field_sum": [_P, _P, _I, _I, _P, _P, _P],
    "pmb_field_dot": [_P, _P, _P, _I, _I, _P, _P, _P],
    "pmb_axpy": [_P, _P, _L, _P, _L, _D, _I, _L],
    "pmb_lincomb": [_P, _P, _L, _P, _L, _D, _P, _L, _D, _I, _L],
    "pmb_column_mod": [_P, _P, _L, _D, _I, _L],
    "pmb_kick_drift": [_P, _P, _P, _P, _I, _D, _D, _I, _L],
    "pmb_dot": [_P, _P, _L, _P, _L, _I, _L, _P],
    "pmb_particles_uniform": [_P, _P, _I, _L, _I, _P, ctypes.c_uint64, _L],
    "pmb_particles_lattice": [_P, _P, _I, _L, _I, _P, _P, _D, _D, ctypes.c_uint64, _L],
    "pmb_particles_replicate": [_P, _P, _I, _P, _L, _I, _P, _P, _L, _L],
    "pmb_decompose_count": [_P, _P, _P, _P], "pmb_decompose_fill": [_P, _P, _P],
    "pmb_decompose_identity": [_P, _P],
    "pmb_take": [_P, _P, _L, _P, _L, _P],
    "pmb_gather_sum": [_P, _P, _I, _I, _P, _P, _I, _L, _P, _I],
    "pmb_gather_sum_segments": [_P, _P, _I, _I, _P, _P, _I, _L, _P, _I],
    "pmb_gather_add_segments": [_P, _P, _I, _I, _P, _P, _I, _I, _L, _P],
    "pmb_comm_unique_id": [_P], "pmb_comm_init_rank": [_P, _P, _I, _I], "pmb_comm_destroy": [_P],
    "pmb_comm_rank": [_P, _P, _P],
    "pmb_alltoallv": [_P, _P, _P, _P, _P, _P, _P, _L],
    "pmb_allreduce_f64": [_P, _P, _L, _I], "pmb_allgather_bytes": [_P, _P, _P, _L], "pmb_barrier": [_P],
    "pmb_fft_create": [_P, _I, _P, _I, _P], "pmb_fft_create_np": [_P, _I, _P, _I, _P, _P], "pmb_fft_destroy": [_P],
    "pmb_fft_layout": [_P, _P, _P, _P, _P, _P, _P, _P, _P],
    "pmb_fft_r2c": [_P, _P, _P, _D], "pmb_fft_c2r": [_P, _P, _P], "pmb_fft_c2r_multi": [_P, _I, _P, _P], "pmb_fft_library_ms": [_P, _P, _I], "pmb_fft_transpose_stats": [_P, _P, _P, _I],
    "pmb_transfer": [_P, _I, _I, _P, _P, _P, _P],
    "pmb_transfer_scaled": [_P, _I, _I, _P, _P, _D, _P, _P],
    "pmb_cdot": [_P, _P, _P, _P],
    "pmb_transfer_grad3": [_P, _I, _P, _D, _P, _P],
    "pmb_fft_c2r_grad3": [_P, _I, _P, _D, _P, _P], "pmb_fft_force3": [_P, _I, _P, _D, _P, _P], "pmb_fft_fused_stats": [_P, _P, _P, _I],
    "pmb_whitenoise": [_P, _P, _I, _P, _P, _P, _P, ctypes.c_uint, _I],
}

_lib = None


def load():
    """Load the shared library (no GPU needed for loading). Raises if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PmbError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C pmesh_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.pmb_last_error.restype = ctypes.c_char_p
        for name in SYMBOLS:
            fn = getattr(lib, name)
            if name != "pmb_last_error":
                fn.restype = ctypes.c_int
            if name in _ARGTYPES:
                fn.argtypes = _ARGTYPES[name]
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise PmbError("libpmesh_b200: %s (code %d)" % (load().pmb_last_error().decode(), rc))


def _vp(x):
    """device pointer / host array -> c_void_p"""
    if x is None:
        return None
    if isinstance(x, numpy.ndarray):
        return ctypes.c_void_p(x.ctypes.data)
    return ctypes.c_void_p(int(x))


class Context(object):
    """pmb_ctx wrapper; the process-wide instance is ``context()``."""

    def __init__(self, device=None):
        lib = load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
            n = ctypes.c_int(0)
            rc = lib.pmb_device_count(ctypes.byref(n))
            check(rc)
            if n.value <= 0:
                raise PmbError("no CUDA device visible; pmesh_b200 has no CPU fallback")
            device = device % n.value
        self.lib = lib
        self.device = device
        h = ctypes.c_void_p()
        check(lib.pmb_ctx_create(ctypes.c_int(device), ctypes.byref(h)))
        self.handle = h
        self._tables_loaded = False
        # the library asks for the cached blocks back when its own work space does not fit (pmb_set_trim_callback)
        import weakref
        me = weakref.ref(self)

        def _trim(_arg):
            c = me()
            if c is not None:
                c.empty_cache()
        self._trim_cb = ctypes.CFUNCTYPE(None, ctypes.c_void_p)(_trim)
        check(lib.pmb_set_trim_callback(h, ctypes.cast(self._trim_cb, ctypes.c_void_p), None))

    # -- memory ---------------------------------------------------------------
    # Device memory goes through a small caching allocator: freed blocks are kept by size class and
    # handed out again without cudaMalloc/cudaFree (and their implicit synchronisation).  Reuse is
    # safe because every kernel and copy of this context is ordered on its single stream.
    @staticmethod
    def _size_class(nbytes):
        nbytes = max(int(nbytes), 512)
        step = 1 << max(nbytes.bit_length() - 5, 9)      # <= 3.2 % rounding waste, >= 512 B granularity
        return (nbytes + step - 1) // step * step

    def malloc(self, nbytes):
        size = self._size_class(nbytes)
        pool = self.__dict__.setdefault("_pool", {})
        sizes = self.__dict__.setdefault("_sizes", {})
        blocks = pool.get(size)
        if blocks:
            ptr = blocks.pop()
            self._pooled -= size
            return ptr
        p = ctypes.c_void_p()
        rc = self.lib.pmb_malloc(self.handle, ctypes.c_size_t(size), ctypes.byref(p))
        if rc != 0 and pool:
            self.empty_cache()
            rc = self.lib.pmb_malloc(self.handle, ctypes.c_size_t(size), ctypes.byref(p))
        if rc != 0:
            # the library's own caches (scratch, sorted particle copies) go last: the calls that use them fall back
            # to slower paths instead of failing
            self.lib.pmb_ctx_trim(self.handle)
            rc = self.lib.pmb_malloc(self.handle, ctypes.c_size_t(size), ctypes.byref(p))
        check(rc)
        sizes[p.value] = size
        return p.value or 0

    _pooled = 0

    def free(self, ptr):
        if not ptr:
            return
        size = self.__dict__.setdefault("_sizes", {}).get(ptr)
        if size is None:
            check(self.lib.pmb_free(self.handle, ctypes.c_void_p(ptr)))
            return
        self.__dict__.setdefault("_pool", {}).setdefault(size, []).append(ptr)
        self._pooled += size
        if self._pooled > self._pool_limit():
            self.trim_cache(self._pool_limit())

    def _pool_limit(self):
        # blocks kept for reuse: up to 45 % of the device memory.  A step of the multi-rank force path frees
        # and re-allocates tens of GB of particle buffers (exchange, readout, ghost sum); letting the pool
        # overflow turned every one of them into cudaFree + cudaMalloc (measured: 198 ms instead of ~70 ms
        # per step at 2 GPUs).  cudaMalloc failures elsewhere empty the pool and retry.
        if "_limit" not in self.__dict__:
            self._limit = int(0.45 * self.mem_info()[1])
        return self._limit

    def trim_cache(self, target):
        """release pooled blocks, largest size classes first, until at most `target` bytes stay pooled"""
        pool = self.__dict__.setdefault("_pool", {})
        sizes = self.__dict__.setdefault("_sizes", {})
        for size in sorted(pool.keys(), reverse=True):
            blocks = pool[size]
            while blocks and self._pooled > target:
                ptr = blocks.pop()
                sizes.pop(ptr, None)
                self._pooled -= size
                check(self.lib.pmb_free(self.handle, ctypes.c_void_p(ptr)))
            if self._pooled <= target:
                break

    def empty_cache(self):
        """release every pooled block back to the driver"""
        pool = self.__dict__.setdefault("_pool", {})
        sizes = self.__dict__.setdefault("_sizes", {})
        for size, blocks in pool.items():
            for ptr in blocks:
                sizes.pop(ptr, None)
                check(self.lib.pmb_free(self.handle, ctypes.c_void_p(ptr)))
        pool.clear()
        self._pooled = 0

    def malloc_host(self, nbytes):
        p = ctypes.c_void_p()
        check(self.lib.pmb_malloc_host(self.handle, ctypes.c_size_t(int(nbytes)), ctypes.byref(p)))
        return p.value or 0

    def free_host(self, ptr):
        if ptr:
            check(self.lib.pmb_free_host(self.handle, ctypes.c_void_p(ptr)))

    def h2d(self, dst, src_h, nbytes):
        check(self.lib.pmb_memcpy_h2d(self.handle, ctypes.c_void_p(dst), _vp(src_h), ctypes.c_size_t(int(nbytes))))

    def d2h(self, dst_h, src, nbytes):
        check(self.lib.pmb_memcpy_d2h(self.handle, _vp(dst_h), ctypes.c_void_p(src), ctypes.c_size_t(int(nbytes))))

    def d2d(self, dst, src, nbytes):
        check(self.lib.pmb_memcpy_d2d(self.handle, ctypes.c_void_p(dst), ctypes.c_void_p(src), ctypes.c_size_t(int(nbytes))))

    # copy streams: 1 = host -> device, 2 = device -> host; 0 is the compute stream (include/pmesh_b200.h)
    def h2d_async(self, dst, src_h, nbytes):
        check(self.lib.pmb_memcpy_h2d_async(self.handle, ctypes.c_void_p(dst), _vp(src_h), ctypes.c_size_t(int(nbytes))))

    def d2h_async(self, dst_h, src, nbytes):
        check(self.lib.pmb_memcpy_d2h_async(self.handle, _vp(dst_h), ctypes.c_void_p(src), ctypes.c_size_t(int(nbytes))))

    def stream_record(self, stream, event):
        check(self.lib.pmb_stream_record(self.handle, int(stream), int(event)))

    def stream_wait(self, stream, event):
        check(self.lib.pmb_stream_wait(self.handle, int(stream), int(event)))

    def stream_sync(self, stream):
        check(self.lib.pmb_stream_sync(self.handle, int(stream)))

    def memset(self, dst, byte, nbytes):
        check(self.lib.pmb_memset(self.handle, ctypes.c_void_p(dst), ctypes.c_int(byte), ctypes.c_size_t(int(nbytes))))

    def mem_info(self):
        f, t = ctypes.c_size_t(), ctypes.c_size_t()
        check(self.lib.pmb_mem_info(self.handle, ctypes.byref(f), ctypes.byref(t)))
        return f.value, t.value

    def sync(self):
        check(self.lib.pmb_ctx_sync(self.handle))

    # -- timing ---------------------------------------------------------------
    def timer_start(self, slot=0):
        check(self.lib.pmb_timer_start(self.handle, ctypes.c_int(slot)))

    def timer_stop(self, slot=0):
        ms = ctypes.c_float()
        check(self.lib.pmb_timer_stop(self.handle, ctypes.c_int(slot), ctypes.byref(ms)))
        return ms.value

    def launch_count(self, reset=False):
        n = c_i64()
        check(self.lib.pmb_launch_count(self.handle, ctypes.byref(n), ctypes.c_int(int(reset))))
        return n.value

    def flush_l2(self):
        check(self.lib.pmb_flush_l2(self.handle))

    def set_workspace_limit(self, nbytes):
        check(self.lib.pmb_set_workspace_limit(self.handle, int(nbytes)))

    # -- window tables --------------------------------------------------------
    def ensure_tables(self):
        """upload the lanczos/acg/db/sym lookup tables once (pmesh_b200/data/window_tables.npz)"""
        if self._tables_loaded:
            return
        from .window import KINDS
        z = numpy.load(os.path.join(_HERE, "data", "window_tables.npz"))
        for name in z.files:
            if name.endswith("_meta"):
                continue
            vals = numpy.ascontiguousarray(z[name], dtype="f8")
            step, support, hs = [float(v) for v in z[name + "_meta"]]
            check(self.lib.pmb_window_set_table(self.handle, ctypes.c_int(KINDS[name]), _vp(vals), ctypes.c_int(len(vals)),
                                                ctypes.c_double(step), ctypes.c_double(support), ctypes.c_double(hs)))
        self._tables_loaded = True


_ctx = None


def context():
    """The process-wide context on GPU ``LOCAL_RANK`` (default 0)."""
    global _ctx
    if _ctx is None:
        _ctx = Context()
    return _ctx
