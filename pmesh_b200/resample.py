"""Field resampling and I/O-order operations -- compositions of the hot-path operators.

Everything here is orchestration: `decompose` / `exchange` / `paint` / `readout` / `r2c` / `c2r` do the
work on the GPU.  The methods of ``Field`` / ``ParticleMesh`` of the same names forward to these
functions; the behaviour follows the reference (file:line per function).
"""
import functools

import numpy

from .window import Affine


def _density_ratio(src_pm, dst_pm):
    """cells per unit volume of the source mesh over that of the destination mesh"""
    return (src_pm.Nmesh.prod() / src_pm.BoxSize.prod()) / (dst_pm.Nmesh.prod() / dst_pm.BoxSize.prod())


# ---------------------------------------------------------------------------------------------- C order
def c_order_is_local(field):
    """True when the global C-order ravel of the field, cut into pieces of the local sizes, is every
    rank's own values in local C order: one rank, or a field distributed along axis 0 only (the
    real-space slabs of this engine).  The reference sorts with mpsort in general (pm.py:418-420);
    here the one layout that would need the exchange is the transposed complex field on P > 1 ranks."""
    from .pm import RealField
    return field.pm.comm.size == 1 or isinstance(field, RealField)


def _flat_target(field, out):
    from .pm import is_inplace
    if out is None:
        out = numpy.empty_like(field.value)
    if is_inplace(out):
        out = field.value
    if not isinstance(out, numpy.flatiter):
        out = out.flat
    assert isinstance(out, numpy.flatiter)
    assert len(out) == field.size
    return out


def ravel(field, out=None):
    """reference pm.py:389-424; out: a flatiter / array (its .flat is used) or Ellipsis for in place"""
    out = _flat_target(field, out)
    if not c_order_is_local(field):
        raise NotImplementedError("ravel of a transposed complex field on more than one rank needs the "
                                  "distributed sort (mpsort); not built")
    out[...] = numpy.array(field.value.flat)
    return out


def unravel(field, flatiter):
    """reference pm.py:426-448"""
    if not isinstance(flatiter, numpy.flatiter):
        flatiter = flatiter.flat
    assert isinstance(flatiter, numpy.flatiter)
    assert field.pm.comm.allreduce(len(flatiter)) == field.csize
    if not c_order_is_local(field):
        raise NotImplementedError("unravel of a transposed complex field on more than one rank needs the "
                                  "distributed sort (mpsort); not built")
    field.value.flat[...] = numpy.array(flatiter)


# ------------------------------------------------------------------------------------- Fourier resample
def mode_table(Nsrc, Ndest):
    """for every index of a length-Ndest frequency axis, the index of the same frequency on a
    length-Nsrc axis, -1 where the source does not carry it (`reindex`, reference pm.py:1128-1144):
    mode_table(8, 4) -> [0, 1, 2, 7];  mode_table(4, 8) -> [0, 1, 2, -1, -1, -1, -1, 3]"""
    t = numpy.arange(Ndest)
    t[Ndest // 2 + 1:] = numpy.arange(Nsrc - Ndest // 2 + 1, Nsrc, 1)
    t[Nsrc // 2 + 1: Ndest - Nsrc // 2 + 1] = -1
    return t


def fourier_resample(field, out):
    """zero-fill or truncate Fourier modes onto the mesh of `out`; converts between real and complex
    as needed (reference pm.py:479-547)"""
    from .pm import Field, RealField, TransposedComplexField, _gettype
    assert isinstance(out, Field)
    if all(out.Nmesh == field.Nmesh):
        # same mesh: only the representation changes
        field.cast(type=_gettype(out), out=out)
    if field.pm.comm.size > 1:
        raise NotImplementedError("Fourier-space resample on more than one rank needs the distributed take "
                                  "(mpsort); not built")
    ndim = field.ndim
    src = field.cast(type=TransposedComplexField)
    dest = out.pm.create(type=TransposedComplexField, base=out._base, value=0)
    sv = src.value
    dv = numpy.zeros(dest.shape, dtype=dest.dtype)
    carried = numpy.ones(dest.shape, dtype='?')
    where = []
    for d in range(ndim):
        t = mode_table(src.Nmesh[d], dest.Nmesh[d])[numpy.r_[dest.slices[d]]]
        ok = (t >= 0) & (t < src.cshape[d])
        shp = [-1 if dd == d else 1 for dd in range(ndim)]
        carried &= ok.reshape(shp)
        where.append(numpy.where(ok, t, 0).reshape(shp))
    dv[carried] = sv[tuple(numpy.broadcast_arrays(*where))][carried]
    # keep the result the transform of a real field, and drop the Nyquist planes of both meshes
    # (pm.py:520-542: "the nyquist is messy due to hermitian constraints")
    i = dest.i
    selfconj = functools.reduce(numpy.bitwise_and, [(n - ii) % n == ii for ii, n in zip(i, dest.Nmesh)])
    dv.imag[numpy.broadcast_to(selfconj, dest.shape)] = 0
    for mesh in (dest.Nmesh, src.Nmesh):
        nyq = functools.reduce(numpy.bitwise_or, [ii == n // 2 for ii, n in zip(i, mesh)])
        dv[numpy.broadcast_to(nyq, dest.shape)] = 0
    dest.value = dv
    if isinstance(out, RealField):
        dest.c2r(out)
    elif out is not dest:
        out.value = dest.value
    return out


# --------------------------------------------------------------------------------- real-space resample
def upsample(pm, source, resampler=None, keep_mean=False):
    """read `source` out at the pixel positions of `pm` (reference pm.py:1937-1989)"""
    from .pm import RealField
    assert isinstance(source, RealField)
    pixels = pm.mesh_coordinates(dtype=pm.dtype)
    to_source = Affine(pm.ndim, translate=-source.start, scale=1.0 * source.Nmesh / pm.Nmesh, period=source.Nmesh)
    # the reference builds the layout twice; the second, with its fixed 1.6 cells of smoothing, is the
    # one used (pm.py:1971-1972)
    layout = source.pm.decompose(pixels, smoothing=1.6, transform=to_source)
    values = source.readout(pixels, resampler=resampler, layout=layout, transform=to_source)
    if not keep_mean:
        values *= _density_ratio(source.pm, pm)
    # every pixel is a mesh point of pm and already on its rank: nearest-point paint, no exchange
    return pm.paint(pixels, mass=values, resampler='nnb', transform=pm.affine_grid)


def downsample(pm, source, resampler=None, keep_mean=False):
    """paint the pixels of `source` onto `pm` (reference pm.py:1991-2027)"""
    from .pm import RealField
    assert isinstance(source, RealField)
    pixels = source.pm.mesh_coordinates(dtype=pm.dtype)
    values = source.readout(pixels, resampler='nnb', transform=source.pm.affine_grid)
    to_me = pm.affine_grid.rescale(1.0 * pm.Nmesh / source.Nmesh)
    if keep_mean:
        values /= _density_ratio(source.pm, pm)
    layout = pm.decompose(pixels, smoothing=resampler, transform=to_me)
    return pm.paint(pixels, mass=values, layout=layout, resampler=resampler, transform=to_me)


def ctranspose(field, axes):
    """permute the coordinates of a RealField onto a ParticleMesh with permuted BoxSize / Nmesh; like
    the reference (pm.py:696-723) with a nearest-point readout and paint"""
    assert len(numpy.unique(axes)) == field.ndim
    assert numpy.max(axes) == field.ndim - 1
    axes = numpy.array(axes, dtype='intp')
    pm = field.pm.reshape(BoxSize=field.BoxSize[axes], Nmesh=field.Nmesh[axes])
    points = field.pm.generate_uniform_particle_grid(shift=0)
    values = field.readout(points, resampler='nnb')
    points = points[..., axes]
    layout = pm.decompose(points, smoothing='nnb')
    return pm.paint(points, mass=values, resampler='nnb', layout=layout)


def preview(field, Nmesh=None, axes=None, resampler=None, method=None):
    """the mesh as one numpy array on every rank, optionally at another resolution and summed over the
    axes that are not listed (reference pm.py:549-615)"""
    from .pm import BaseComplexField
    if axes is None:
        axes = list(range(field.ndim))
    elif not hasattr(axes, '__iter__'):
        axes = [axes]
    else:
        axes = list(axes)
    if isinstance(field, BaseComplexField):
        field = field.c2r()
    if Nmesh is not None and all(Nmesh == field.Nmesh):
        Nmesh = None
    image = field
    if Nmesh is not None:
        pm = field.pm.reshape(Nmesh)
        if method is None:
            method = 'downsample' if any(pm.Nmesh < field.Nmesh) else 'upsample'
        if method == 'downsample':
            image = pm.downsample(field, resampler=resampler, keep_mean=True)
        elif method == 'upsample':
            image = pm.upsample(field, resampler=resampler, keep_mean=True)
        else:
            raise ValueError("method can only be downsample or upsample")
    result = numpy.zeros([image.cshape[i] for i in axes], dtype=image.dtype)
    mine = tuple(image.slices[i] for i in axes)
    local = image[...]
    projected = [d for d in range(field.ndim) if d not in axes]
    local = local.transpose(axes + projected)
    if projected:
        local = local.sum(axis=tuple(range(len(axes), field.ndim)))
    result[mine] += local
    return field.pm.comm.Allreduce_inplace(result)
