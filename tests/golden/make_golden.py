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


if __name__ == "__main__" and "--pipeline" not in sys.argv:
    main()


# ------------------------------------------------------------------------------------------------
# whole-pipeline vectors from the reference's own pmesh/pm.py (run on the stand-ins of
# tests/golden/reference_pm.py: numpy.fft behind pfft's interface, one rank)
PIPELINE_CASES = [
    # (window, Nmesh, BoxSize, dtype, number of particles, seed)
    ("cic", 8, 100.0, "f8", 300, 1),
    ("tsc", 12, 64.0, "f8", 500, 2),
    ("pcs", 8, 8.0, "f8", 300, 3),
    ("cic", 8, 100.0, "f4", 300, 4),
]
VJP_CASES = [("pcs", 8, 10.0, 200, 11), ("lanczos3", 8, 10.0, 120, 12), ("cic", 6, 6.0, 150, 13)]


def pipeline_inputs(case):
    window, n, L, dt, npart, seed = case
    rng = numpy.random.default_rng(500 + seed)
    return rng.uniform(-0.2 * L, 1.2 * L, (npart, 3))


def vjp_inputs(case):
    window, n, L, npart, seed = case
    rng = numpy.random.default_rng(600 + seed)
    pos = rng.uniform(0, L, (npart, 3))
    mass = rng.uniform(0.5, 2.0, npart)
    field = rng.uniform(-1, 1, (n, n, n))
    v = rng.uniform(-1, 1, npart)
    return pos, mass, field, v


def fd4_force_transfer(direction):
    """the force kernel of examples/nbody.py:162-170 as a python callable for the reference's apply()"""
    def filt(k, v):
        k2 = sum(ki ** 2 for ki in k)
        k2[k2 == 0] = 1.0
        C = (v.BoxSize / v.Nmesh)[direction]
        w = k[direction] * C
        kfinite = 1.0 / C * 1 / 6.0 * (8 * numpy.sin(w) - numpy.sin(2 * w))
        return 1j * kfinite / k2 * v
    return filt


def make_pipeline_golden():
    sys.path.insert(0, HERE)
    import reference_pm
    ns = reference_pm.load()
    assert ns is not None, "needs /root/reference"
    passed, failed = reference_pm.selftest(verbose=False)
    assert not failed, failed            # the stand-ins carry the reference's own test-suite
    PM = ns.pm
    out = {"selftest_passed": numpy.array(len(passed))}
    for ci, case in enumerate(PIPELINE_CASES):
        window, n, L, dt, npart, seed = case
        pos = pipeline_inputs(case)
        pm = PM.ParticleMesh(BoxSize=L, Nmesh=[n, n, n], dtype=dt, resampler=window)
        layout = pm.decompose(pos, smoothing=1.0 * pm.resampler.support)
        rho = pm.create("real")
        rho.paint(pos, layout=layout, hold=False)
        out["rho_%d" % ci] = numpy.array(rho.value)
        rho[...] *= 1.0 * pm.Nmesh.prod() / len(pos)
        rhok = rho.r2c()
        out["rhok_%d" % ci] = numpy.array(rhok.value)
        F = numpy.empty((len(pos), 3))
        for d in range(3):
            F[:, d] = rhok.apply(fd4_force_transfer(d)).c2r().readout(pos, layout=layout)
        out["force_%d" % ci] = F
    # coordinates of a non-cubic mesh (pm.py:1178-1226)
    pm = PM.ParticleMesh(BoxSize=[8.0, 12.0, 5.0], Nmesh=[8, 6, 10], dtype="f8")
    c, r = pm.create("complex"), pm.create("real")
    for d in range(3):
        out["kx_%d" % d] = numpy.array(c.x[d]).ravel()
        out["rx_%d" % d] = numpy.array(r.x[d]).ravel()
    # white noise -> linear field -> 1-LPT displacement (examples/nbody.py:245-270)
    pm = PM.ParticleMesh(BoxSize=64.0, Nmesh=[16, 16, 16], dtype="f8", resampler="cic")
    wn = pm.generate_whitenoise(120577, unitary=True)
    out["wn_unitary"] = numpy.array(wn.value)
    dlinear = wn.apply(lambda k, v: numpy.exp(-0.5 * sum(ki ** 2 for ki in k) * 4.0 ** 2) * v)
    out["dlinear"] = numpy.array(dlinear.value)
    Q = pm.generate_uniform_particle_grid(shift=0.0)
    layout = pm.decompose(Q)
    DX1 = numpy.zeros_like(Q)

    def dx1(direction):
        def filt(k, v):
            k2 = sum(ki ** 2 for ki in k)
            k2[k2 == 0] = 1.0
            return 1j * k[direction] / k2 * v
        return filt
    for d in range(3):
        DX1[:, d] = dlinear.apply(dx1(d)).c2r().readout(Q, layout=layout)
    out["dx1"] = DX1
    out["wn_real_mean2"] = numpy.array(pm.generate_whitenoise(7, type="real", mean=2.0).value)
    # back-propagation operators (pm.py:793-859, 1872-1935), BASELINE configs[3] windows
    for ci, case in enumerate(VJP_CASES):
        window, n, L, npart, seed = case
        pos, mass, field, v = vjp_inputs(case)
        pm = PM.ParticleMesh(BoxSize=L, Nmesh=[n, n, n], dtype="f8", resampler=window)
        vf = pm.create("real", value=2 * field)
        gpos, gmass = pm.paint_vjp(vf, pos, mass=mass)
        out["paint_vjp_pos_%d" % ci], out["paint_vjp_mass_%d" % ci] = gpos, gmass
        f = pm.create("real", value=field)
        gself, gpos = f.readout_vjp(pos, v=2 * v)
        out["readout_vjp_self_%d" % ci], out["readout_vjp_pos_%d" % ci] = numpy.array(gself.value), gpos
        out["paint_jvp_%d" % ci] = numpy.array(pm.paint_jvp(pos, mass=mass, v_pos=numpy.ones_like(pos) * [0.1, -0.2, 0.3], v_mass=v).value)
        out["readout_jvp_%d" % ci] = f.readout_jvp(pos, v_self=vf, v_pos=numpy.ones_like(pos) * [0.1, -0.2, 0.3])
    numpy.savez_compressed(os.path.join(HERE, "pipeline_golden.npz"), **out)
    print("pipeline_golden.npz:", len(out), "arrays; reference tests passed on the stand-ins:", len(passed))


if __name__ == "__main__" and "--pipeline" in sys.argv:
    make_pipeline_golden()
