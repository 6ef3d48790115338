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
# The reference moves values between the field layout and the global C order with mpsort (a distributed
# sort: pm.py:418-420, 444-446, 513).  The keys here are global C indices, so no sort is needed: the owner of
# every index is known from the chunk sizes, and the values travel in one exchange.  These are host-side
# I/O-order operations (they work on the numpy mirror of a field, like the reference); the exchange goes
# through the communicator's object collectives.
def c_order_is_local(field):
    """True when the global C-order ravel of the field, cut into pieces of the local sizes, is every
    rank's own values in local C order: one rank, or a field distributed along axis 0 only (real-space
    slabs)."""
    if field.pm.comm.size == 1:
        return True
    return all(int(field.shape[d]) == int(field.cshape[d]) for d in range(1, field.ndim))


def _alltoall(comm, pieces):
    """pieces[q] goes to rank q; returns what every rank sent to this one, in rank order"""
    if comm.size == 1:
        return [pieces[0]]
    everything = comm.allgather(pieces)
    return [everything[src][comm.rank] for src in range(comm.size)]


def _global_c_index(field):
    """global C-order index of every local element (local logical C order), int64"""
    idx = numpy.zeros(tuple(int(n) for n in field.shape), dtype='i8')
    stride = 1
    for d in reversed(range(field.ndim)):
        r = numpy.arange(int(field.start[d]), int(field.start[d]) + int(field.shape[d]), dtype='i8') * stride
        shp = [1] * field.ndim
        shp[d] = -1
        idx = idx + r.reshape(shp)
        stride *= int(field.cshape[d])
    return idx.ravel()


def _chunk_offsets(comm, n):
    sizes = numpy.array(comm.allgather(int(n)), dtype='i8')
    return numpy.concatenate([[0], numpy.cumsum(sizes)])


def dist_put(comm, values, gindex, outlen):
    """out[g - first] = v for every (g, v) of every rank, where the global C order is cut into chunks of
    `outlen` elements per rank (mpsort.sort by a key that is the global index)"""
    offs = _chunk_offsets(comm, outlen)
    dest = numpy.searchsorted(offs[1:], gindex, side='right')
    pieces = []
    for q in range(comm.size):
        m = dest == q
        pieces.append((gindex[m] - offs[q], values[m]))
    out = numpy.empty(int(outlen), dtype=values.dtype)
    for li, vv in _alltoall(comm, pieces):
        out[li] = vv
    return out


def dist_take(comm, chunk, gindex):
    """chunks[...][gindex] for a 1-D array distributed in rank order as `chunk` (mpsort.take / permute)"""
    chunk = numpy.asarray(chunk)
    offs = _chunk_offsets(comm, len(chunk))
    gindex = numpy.asarray(gindex, dtype='i8')
    owner = numpy.searchsorted(offs[1:], gindex, side='right')
    where = [numpy.flatnonzero(owner == q) for q in range(comm.size)]
    asked = _alltoall(comm, [gindex[w] - offs[q] for q, w in enumerate(where)])
    answers = _alltoall(comm, [chunk[a] for a in asked])
    out = numpy.empty(len(gindex), dtype=chunk.dtype)
    for w, v in zip(where, answers):
        out[w] = v
    return out


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
    if c_order_is_local(field):
        out[...] = numpy.array(field.value.flat)
    else:
        out[...] = dist_put(field.pm.comm, numpy.array(field.value.flat), _global_c_index(field), int(field.size))
    return out


def unravel(field, flatiter):
    """reference pm.py:426-448"""
    if not isinstance(flatiter, numpy.flatiter):
        flatiter = flatiter.flat
    assert isinstance(flatiter, numpy.flatiter)
    assert field.pm.comm.allreduce(len(flatiter)) == field.csize
    if field.pm.comm.size == 1:
        field.value.flat[...] = numpy.array(flatiter)
    else:
        v = field.value
        v.flat[...] = dist_take(field.pm.comm, numpy.array(flatiter), _global_c_index(field))
        field.value = v


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
    ndim = field.ndim
    src = field.cast(type=TransposedComplexField)
    dest = out.pm.create(type=TransposedComplexField, base=out._base, value=0)
    if field.pm.comm.size > 1:
        return _fourier_resample_distributed(src, dest, out)
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


def _fourier_resample_distributed(src, dest, out):
    """the same on P > 1 ranks: the source modes are fetched from their owners with a distributed take on
    the C-order ravel of the source (reference pm.py:493-516)"""
    from .pm import RealField
    comm = src.pm.comm
    ndim = src.ndim
    flat = numpy.empty(int(src.size), dtype=src.dtype)
    ravel(src, out=flat)
    carried = numpy.ones(tuple(int(n) for n in dest.shape), dtype='?')
    gidx = numpy.zeros(carried.shape, dtype='i8')
    stride = 1
    for d in reversed(range(ndim)):
        t = mode_table(src.Nmesh[d], dest.Nmesh[d])[numpy.r_[dest.slices[d]]]
        ok = (t >= 0) & (t < src.cshape[d])
        shp = [-1 if dd == d else 1 for dd in range(ndim)]
        carried &= ok.reshape(shp)
        gidx = gidx + (numpy.where(ok, t, 0).astype('i8') * stride).reshape(shp)
        stride *= int(src.cshape[d])
    dv = numpy.zeros(carried.shape, dtype=dest.dtype)
    dv[carried] = dist_take(comm, flat, gidx[carried])
    i = dest.i
    selfconj = functools.reduce(numpy.bitwise_and, [(n - ii) % n == ii for ii, n in zip(i, dest.Nmesh)])
    dv.imag[numpy.broadcast_to(selfconj, dv.shape)] = 0
    for mesh in (dest.Nmesh, src.Nmesh):
        nyq = functools.reduce(numpy.bitwise_or, [ii == n // 2 for ii, n in zip(i, mesh)])
        dv[numpy.broadcast_to(nyq, dv.shape)] = 0
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
