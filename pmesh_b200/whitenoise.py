"""White noise in Fourier space -- the pmesh.whitenoise API (reference pmesh/whitenoise.py:4-40).

``generate(complex, start, Nmesh, seed, unitary)`` fills the local block ``complex`` of the Hermitian
half-spectrum with the N-GenIC / Gadget scheme on the RANLUX ``ranlxd1`` generator
(pmesh/_whitenoise_generics.h:29-232, pmesh/gsl/ranlxd.c): one GPU thread per Fourier column
(``pmb_whitenoise``).  The random streams are bit-identical to the reference's, so the field does
not depend on how the mesh is partitioned over ranks.

``complex`` may be a numpy array (filled through a device buffer and copied back, like the window
functions accept host canvases) or a ``DeviceArray`` (filled in place, any strides).
"""
import ctypes

import numpy

from . import _lib
from .device import DeviceArray, is_device


def generate(complex, start, Nmesh, seed, unitary):
    """
        The result is always hermitian.

        complex : (n0, n1, n2) complex64 / complex128 block of the compressed (k_z <= N/2) Fourier mesh
        start   : index of its first element in the global mesh
        Nmesh   : the global mesh
        unitary : True for a unitary field (amplitude fixed to 1), False for a true Gaussian field
    """
    ndim = len(complex.shape)
    _start = numpy.empty(ndim, dtype='int64')
    _Nmesh = numpy.empty(ndim, dtype='int64')
    _start[:] = start
    _Nmesh[:] = Nmesh
    if ndim != 3:
        # the reference's 1-D / 2-D branch is a numpy RandomState + fftn test helper
        # (pmesh/whitenoise.py:25-38): not a kernel, not part of this engine
        raise NotImplementedError("white noise is generated for 3-D meshes")
    if numpy.dtype(complex.dtype).kind != 'c':
        raise TypeError("the canvas must be complex")
    if _start[2] + complex.shape[2] > _Nmesh[2] // 2 + 1:
        raise NotImplementedError("only the compressed (k_z <= N/2) half of the Fourier mesh is generated; "
                                  "the full-spectrum fill of the reference serves c2c meshes")
    if is_device(complex):
        dev = complex
    else:
        dev = DeviceArray.empty(complex.shape, complex.dtype)
    ctx = dev.ctx
    A = ctypes.c_int64 * 3
    _lib.check(ctx.lib.pmb_whitenoise(ctx.handle, dev.ptr, dev.dtype.itemsize, A(*_Nmesh), A(*_start),
                                      A(*dev.shape), A(*dev.strides), int(seed) & 0xffffffff, int(bool(unitary))))
    if dev is not complex:
        complex[...] = dev.to_host()
    return complex
