"""helpers shared by the test modules"""
import ctypes

import numpy

from pmesh_b200 import _lib
from pmesh_b200.window import KINDS

# public window names -> kind strings (reference window.py:230-255)
PUBLIC = dict(nnb="tunednnb", cic="tunedcic", tsc="tunedtsc", pcs="tunedpcs")
for _k in KINDS:
    if not _k.startswith("tuned"):
        PUBLIC[_k] = _k

ALL_WINDOWS = sorted(PUBLIC)
FAST_WINDOWS = ["nnb", "cic", "tsc", "pcs", "linear", "quadratic", "cubic", "lanczos2", "lanczos3", "acg3", "db12", "sym6"]


def resample_args(kind, support, mesh, pos, mass=None, hsml=None, out=None, diffdir=None,
                  scale=None, translate=None, period=None, mode=0, pcsfix=0):
    """pmb_resample_args over HOST numpy arrays (for the host harness). Returns (args, keepalive)."""
    nd = mesh.ndim
    a = _lib.ResampleArgs()
    a.kind = KINDS[PUBLIC.get(kind, kind)] if isinstance(kind, str) else int(kind)
    a.support = int(support)
    a.ndim = nd
    sc = numpy.empty(nd); sc[:] = 1.0 if scale is None else scale
    tr = numpy.empty(nd); tr[:] = 0.0 if translate is None else translate
    pe = numpy.empty(nd, dtype="i8"); pe[:] = 0 if period is None else period
    for d in range(nd):
        a.order[d] = 1 if diffdir == d else 0
        a.scale[d] = sc[d]
        a.translate[d] = tr[d]
        a.period[d] = pe[d]
        a.size[d] = mesh.shape[d]
        a.strides[d] = mesh.strides[d]
    a.mesh = mesh.ctypes.data
    a.mesh_elsize = mesh.dtype.itemsize
    keep = [mesh, pos]
    a.pos = pos.ctypes.data
    a.pos_elsize = pos.dtype.itemsize
    a.npart = len(pos)
    a.pos_stride0, a.pos_stride1 = pos.strides
    a.mass_scalar = 1.0
    a.hsml_scalar = 1.0
    for name, v in (("mass", mass), ("hsml", hsml)):
        if v is None:
            continue
        if numpy.ndim(v) == 0:
            setattr(a, name + "_scalar", float(v))
            continue
        v = numpy.asarray(v)
        keep.append(v)
        setattr(a, name, v.ctypes.data)
        setattr(a, name + "_elsize", v.dtype.itemsize)
        setattr(a, name + "_stride", v.strides[0])
    if out is not None:
        keep.append(out)
        a.out = out.ctypes.data
        a.out_elsize = out.dtype.itemsize
        a.out_stride = out.strides[0]
    a.mode = mode
    a.pcs_gradient_scale_fix = pcsfix
    return a, keep


def random_case(rng, nd, n=300, dtype="f8", posdtype="f8", lo=-3.0, hi=9.0):
    shape = (7, 6, 5)[:nd]
    pos = rng.uniform(lo, hi, (n, nd)).astype(posdtype)
    mass = rng.uniform(0.5, 2.0, n)
    scale = numpy.array([0.5, 2.0, 1.1])[:nd]
    translate = numpy.array([2.0, -1.0, 0.3])[:nd]
    period = numpy.array([7, 6, 5])[:nd]
    return shape, pos, mass, scale, translate, period
