"""Generate the committed golden vectors from the REFERENCE ITSELF (run in the build container only).

Inputs are seeded; outputs come from the reference's own compiled C/Cython (oracle/_ref, built by
oracle/build_ref.py from /root/reference) and from the reference's own pure-python
``pmesh/domain.py`` imported from /root/reference with a fake single-process ``mpi4py``.
The GPU box has no /root/reference: tests there read only the .npz written here.

    python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import build_ref  # noqa: E402

REF = "/root/reference"

WINDOW_CASES = []
for name, kind in [("nnb", "tunednnb"), ("cic", "tunedcic"), ("tsc", "tunedtsc"), ("pcs", "tunedpcs"),
                   ("nearest", "nearest"), ("linear", "linear"), ("quadratic", "quadratic"), ("cubic", "cubic"),
                   ("lanczos2", "lanczos2"), ("lanczos3", "lanczos3"), ("acg3", "acg3"), ("acg6", "acg6"),
                   ("db6", "db6"), ("db12", "db12"), ("sym20", "sym20")]:
    for nd in (1, 2, 3):
        for dtype in ("f8", "f4"):
            WINDOW_CASES.append((name, kind, nd, dtype))


def window_inputs(seed, nd, n):
    rng = numpy.random.default_rng(seed)
    shape = (7, 6, 5)[:nd]
    pos = rng.uniform(-3.0, 9.0, (n, nd))
    mass = rng.uniform(0.5, 2.0, n)
    scale = numpy.array([0.5, 2.0, 1.1])[:nd]
    translate = numpy.array([2.0, -1.0, 0.3])[:nd]
    period = numpy.array([7, 6, 5])[:nd]
    field = rng.uniform(-1, 1, shape)
    return shape, pos, mass, scale, translate, period, field


def fake_mpi():
    """~40 lines of mpi4py.MPI good enough to import the reference's domain.py in one process"""
    class Dt(object):
        def Create_contiguous(self, n): return self
        def Commit(self): pass
        def Free(self): pass

    class FakeComm(object):
        def __init__(self, rank=0, size=1): self.rank, self.size = rank, size
        def Barrier(self): pass
        def Alltoall(self, s, r): r[...] = s
        def allgather(self, x): return [x] * self.size
        def bcast(self, x, root=0): return x
        def allreduce(self, x, op=None): return x

    MPI = types.ModuleType("mpi4py.MPI")
    MPI.COMM_WORLD = FakeComm()
    MPI.BYTE = Dt()
    MPI.SUM = "sum"
    MPI.FakeComm = FakeComm
    pkg = types.ModuleType("mpi4py")
    pkg.MPI = MPI
    sys.modules["mpi4py"] = pkg
    sys.modules["mpi4py.MPI"] = MPI
    return MPI


def load_reference_domain():
    """import /root/reference/pmesh/domain.py as pmesh_ref.domain (its `from ._domain import` resolves
    to the compiled reference extension)"""
    import importlib.util
    MPI = fake_mpi()
    build_ref.load()
    spec = importlib.util.spec_from_file_location("pmesh_ref.domain", os.path.join(REF, "pmesh", "domain.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["pmesh_ref.domain"] = mod
    spec.loader.exec_module(mod)
    return mod, MPI


DOMAIN_CASES = [
    # (edges, P, smoothing, periodic)
    ([numpy.linspace(0, 4, 5)], 4, 1, True),
    ([numpy.linspace(0, 4, 3), numpy.linspace(0, 4, 3)], 4, 0.5, True),
    ([numpy.linspace(0, 64, 3), numpy.linspace(0, 64, 3), numpy.linspace(0, 64, 2)], 4, 1.0, True),
    ([numpy.linspace(0, 64, 9), numpy.array([0, 64.]), numpy.array([0, 64.])], 8, 1.5, True),
    ([numpy.linspace(0, 10, 4), numpy.linspace(0, 10, 3)], 6, 0.7, False),
    ([numpy.array([0, 0, 2, 4., 4.]), numpy.array([0, 2., 4.])], 8, 0.3, True),
    ([numpy.linspace(0, 8, 3), numpy.linspace(0, 8, 3), numpy.linspace(0, 8, 3)], 3, 5.0, True),
]


def domain_inputs(case):
    edges, P, smoothing, periodic = DOMAIN_CASES[case]
    nd = len(edges)
    rng = numpy.random.default_rng(1000 + case)
    box = numpy.array([e[-1] for e in edges])
    pos = rng.uniform(-0.5, 1.5, (2000, nd)) * box
    pos[:20] = 0.0
    pos[20:30] = -1e-17
    pos[30:40] = box
    return pos


# (Nmesh, seed, unitary, complex dtype, block start, block shape or None for the whole half-spectrum)
WHITENOISE_CASES = [
    ((4, 4, 4), 5463, False, "complex128", (0, 0, 0), None),       # the N-GenIC case of tests/test_whitenoise.py:27-38
    ((8, 8, 8), 1, False, "complex128", (0, 0, 0), None),
    ((8, 8, 8), 1, True, "complex128", (0, 0, 0), None),
    ((16, 16, 16), 120577, False, "complex64", (0, 0, 0), None),   # the seed of examples/nbody.py:338
    ((12, 12, 12), 3, False, "complex128", (2, 5, 1), (7, 4, 5)),  # a block: partition invariance
    ((8, 6, 10), 9, False, "complex128", (0, 0, 0), None),         # non-cubic: the spiral's mixed indices
    ((7, 7, 7), 3, True, "complex64", (0, 0, 0), None),            # odd mesh
    ((8, 8, 8), 2 ** 31 + 5, False, "complex128", (0, 0, 0), None),  # seed with bit 31 set (int sign extension)
    ((8, 8, 8), 0, False, "complex128", (0, 0, 0), None),          # seed 0 -> 1
]


def whitenoise_reference(wn, case):
    N, seed, unitary, dt, start, shape = case
    if shape is None:
        shape = (N[0], N[1], N[2] // 2 + 1)
    v = numpy.zeros(shape, dtype=dt)
    wn.generate(v, numpy.array(start, dtype="intp"), numpy.array(N, dtype="intp"), seed, int(unitary))
    return v


def main():
    assert build_ref.build(), "needs /root/reference"
    w, d = build_ref.load()

    wn = build_ref.load_whitenoise()
    out = {}
    for ci, case in enumerate(WHITENOISE_CASES):
        out["wn_%d" % ci] = whitenoise_reference(wn, case)
    # the raw ranlxd1 stream, recovered from a unitary field: phase = u * 2 pi is not invertible
    # bit-exactly, so the stream itself is pinned through the fields above and the live comparison
    numpy.savez_compressed(os.path.join(HERE, "whitenoise_golden.npz"), **out)
    print("whitenoise_golden.npz:", len(out), "arrays")

    class RW(w.ResampleWindow):
        pass

    out = {}
    for ci, (name, kind, nd, dtype) in enumerate(WINDOW_CASES):
        shape, pos, mass, scale, translate, period, field = window_inputs(ci, nd, 60)
        for diffdir in [None] + list(range(nd)):
            order = numpy.zeros(nd, dtype=int)
            if diffdir is not None:
                order[diffdir] = 1
            real = numpy.zeros(shape, dtype)
            RW(kind).paint(real, pos, None, mass, order, scale, translate, period.astype(numpy.intp))
            f = field.astype(dtype)
            ro = numpy.zeros(len(pos))
            RW(kind).readout(f, pos, None, ro, order, scale, translate, period.astype(numpy.intp))
            key = "%s_%d_%s_%s" % (name, nd, dtype, "v" if diffdir is None else "d%d" % diffdir)
            out["paint_" + key] = real
            out["readout_" + key] = ro
    numpy.savez_compressed(os.path.join(HERE, "window_golden.npz"), **out)
    print("window_golden.npz:", len(out), "arrays")

    dom, MPI = load_reference_domain()
    out = {}
    for case, (edges, P, smoothing, periodic) in enumerate(DOMAIN_CASES):
        pos = domain_inputs(case)
        for rank in (0, P - 1):
            g = dom.GridND(edges, comm=MPI.FakeComm(rank, P), periodic=periodic)
            layout = g.decompose(pos, smoothing=smoothing)
            out["counts_%d_%d" % (case, rank)] = numpy.asarray(layout.sendcounts)
            out["indices_%d_%d" % (case, rank)] = numpy.asarray(layout.indices)
            out["assign_%d_%d" % (case, rank)] = numpy.asarray(g.DomainAssign)
    numpy.savez_compressed(os.path.join(HERE, "domain_golden.npz"), **out)
    print("domain_golden.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
