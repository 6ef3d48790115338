"""Run the REFERENCE's own pmesh/pm.py in this container (test infrastructure, build container only).

pmesh.pm needs pfft-python, mpsort and mpi4py, none of which can be installed here (no MPI, no FFTW).
For ONE rank they reduce to very little, so this module provides stand-ins and imports the reference's
unmodified Python from /root/reference on top of the reference's compiled extensions (oracle/_ref):

  pfft    -> single-process partitions / buffers / plans whose `execute` is numpy.fft (unnormalised
             forward, unnormalised backward: FFTW conventions), padded in-place layout included;
  mpsort  -> sort / permute / take on local arrays;
  mpi4py  -> a size-1 communicator.

Everything except the FFT arithmetic itself is then the reference's real code: ParticleMesh set-up,
Field views, r2c / c2r normalisation (pm.py:689-692), apply() and the k / x coordinates
(pm.py:1178-1226), cgetitem / csetitem, paint / readout orchestration, resample, whitenoise ...
`make_golden.py` uses it to write whole-pipeline golden vectors; the reference's own test-suite runs
on it too (see `python tests/golden/reference_pm.py --selftest`).
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PMESH_REFERENCE", "/root/reference")


# ------------------------------------------------------------------------------------------ mpi4py
def fake_mpi4py():
    class Dt(object):
        def Create_contiguous(self, n):
            return self

        def Commit(self):
            return self

        def Free(self):
            pass

    class Comm(object):
        rank, size = 0, 1

        def Barrier(self):
            pass

        barrier = Barrier

        def Alltoall(self, s, r):
            r[...] = s

        def Alltoallv(self, send, recv):
            sbuf, (scounts, soffs), _ = send
            rbuf, (rcounts, roffs), _ = recv
            n = int(scounts[0])
            rbuf[int(roffs[0]):int(roffs[0]) + n] = sbuf[int(soffs[0]):int(soffs[0]) + n]

        def allgather(self, x):
            return [x]

        def bcast(self, x, root=0):
            return x

        def allreduce(self, x, op=None):
            return x

        def Allreduce(self, send, recv, op=None):
            if send is not MPI.IN_PLACE:
                recv[...] = send

    MPI = types.ModuleType("mpi4py.MPI")
    MPI.Comm = Comm
    MPI.COMM_WORLD = Comm()
    MPI.COMM_SELF = Comm()
    MPI.BYTE = Dt()
    MPI.SUM = "sum"
    MPI.IN_PLACE = object()
    MPI._addressof = id
    pkg = types.ModuleType("mpi4py")
    pkg.MPI = MPI
    return pkg, MPI


# ------------------------------------------------------------------------------------------ mpsort
def fake_mpsort():
    m = types.ModuleType("mpsort")

    def sort(data, orderby=None, comm=None, out=None, **kw):
        data = numpy.array(data)
        order = numpy.argsort(numpy.array(orderby), kind="stable")
        r = data[order]
        if out is None:
            return r
        out[...] = r
        return out

    def permute(data, argindex, comm=None, out=None):
        data = numpy.array(data)
        r = numpy.empty_like(data)
        r[numpy.array(argindex)] = data
        if out is None:
            return r
        out[...] = r
        return out

    def take(data, argindex, comm=None, out=None):
        r = numpy.array(data)[numpy.array(argindex)]
        if out is None:
            return r
        out[...] = r
        return out
    m.sort, m.permute, m.take = sort, permute, take
    return m


# ------------------------------------------------------------------------------------------ pfft
def fake_pfft():
    m = types.ModuleType("pfft")

    class Flags(object):
        PFFT_DESTROY_INPUT = 1
        PFFT_PRESERVE_INPUT = 2
        PFFT_PADDED_R2C = 4
        PFFT_PADDED_C2R = 8
        PFFT_ESTIMATE = 16
        PFFT_MEASURE = 32
        PFFT_EXHAUSTIVE = 64
        PFFT_TRANSPOSED_OUT = 128
        PFFT_TRANSPOSED_IN = 256

    class Type(object):
        PFFT_R2C, PFFT_C2R, PFFTF_R2C, PFFTF_C2R, PFFT_C2C, PFFTF_C2C = "r2c8", "c2r8", "r2c4", "c2r4", "c2c8", "c2c4"

    class Direction(object):
        PFFT_FORWARD, PFFT_BACKWARD = -1, 1

    def split_size_2d(s):
        a = int(s ** 0.5) + 1
        while a > 1 and s % a:
            a -= 1
        return (a, s // a)

    class ProcMesh(object):
        def __init__(self, np, comm=None):
            assert all(int(n) == 1 for n in np), "the stand-in is single-rank"
            self.np = np
            self.comm = comm

    class Partition(object):
        def __init__(self, type, n, procmesh, flags):
            n = numpy.array(n, dtype="intp")
            self.type, self.n, self.flags = type, n, flags
            self.ndim = len(n)
            self.is_c2c = type in (Type.PFFT_C2C, Type.PFFTF_C2C)
            self.rdtype = numpy.dtype("f4" if type.endswith("4") else "f8")
            self.cdtype = numpy.dtype("c8" if type.endswith("4") else "c16")
            self.padded = bool(flags & Flags.PFFT_PADDED_R2C)
            no = n.copy()
            if not self.is_c2c:
                no[-1] = n[-1] // 2 + 1
            self.local_i_start = numpy.zeros(self.ndim, dtype="intp")
            self.local_o_start = numpy.zeros(self.ndim, dtype="intp")
            self.local_i_shape = n.copy()
            self.local_o_shape = no
            self.i_edges = [numpy.array([0, s]) for s in n]
            self.o_edges = [numpy.array([0, s]) for s in no]
            self.local_ni = self.local_i_shape.copy()
            self.local_no = self.local_o_shape.copy()
            # number of real words of a buffer that can hold either representation
            self.alloc = int(2 * numpy.prod(no)) if not self.is_c2c else int(2 * numpy.prod(n))

    class LocalBuffer(object):
        def __init__(self, partition, base=None):
            self.partition = partition
            if base is None:
                self.store = numpy.zeros(partition.alloc, dtype=partition.rdtype)
            else:
                self.store = base.store
                assert len(self.store) >= partition.alloc

        def __contains__(self, other):
            return other.store is self.store

        def view_raw(self):
            return self.store

        def view_input(self):
            p = self.partition
            if p.is_c2c:
                return self.store[:2 * int(numpy.prod(p.n))].view(p.cdtype).reshape(tuple(p.n))
            if p.padded:
                shp = tuple(p.n[:-1]) + (2 * (int(p.n[-1]) // 2 + 1),)
                full = self.store[:int(numpy.prod(shp))].reshape(shp)
                return full[..., :int(p.n[-1])]
            return self.store[:int(numpy.prod(p.n))].reshape(tuple(p.n))

        def view_output(self):
            p = self.partition
            no = tuple(p.local_o_shape)
            return self.store[:2 * int(numpy.prod(no))].view(p.cdtype).reshape(no)

    class Plan(object):
        def __init__(self, partition, direction, bufferin, bufferout, type=None, flags=0):
            self.partition, self.direction = partition, direction

        def execute(self, i, o):
            p = self.partition
            axes = tuple(range(p.ndim))
            if p.is_c2c:
                if self.direction == Direction.PFFT_FORWARD:
                    o.view_output()[...] = numpy.fft.fftn(i.view_input().copy(), axes=axes)
                else:
                    o.view_input()[...] = numpy.fft.ifftn(i.view_output().copy(), axes=axes) * numpy.prod(p.n)
            elif self.direction == Direction.PFFT_FORWARD:
                o.view_output()[...] = numpy.fft.rfftn(i.view_input().astype("f8"), axes=axes)
            else:
                y = i.view_output().astype("c16")
                o.view_input()[...] = numpy.fft.irfftn(y, s=tuple(int(x) for x in p.n), axes=axes) * numpy.prod(p.n)

    m.Flags, m.Type, m.Direction = Flags, Type, Direction
    m.split_size_2d, m.ProcMesh, m.Partition, m.LocalBuffer, m.Plan = split_size_2d, ProcMesh, Partition, LocalBuffer, Plan
    return m


_loaded = None


def load():
    """returns the namespace of the reference package: .pm, .window, .domain, .whitenoise (None when
    /root/reference or the compiled extensions are missing)"""
    global _loaded
    if _loaded is not None:
        return _loaded
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    if not os.path.isdir(os.path.join(REF, "pmesh")) or not build_ref.build():
        return None
    build_ref.load()
    build_ref.load_whitenoise()
    pkg, MPI = fake_mpi4py()
    sys.modules["mpi4py"] = pkg
    sys.modules["mpi4py.MPI"] = MPI
    sys.modules["pfft"] = fake_pfft()
    sys.modules["mpsort"] = fake_mpsort()
    ns = types.SimpleNamespace()
    for name in ("window", "domain", "whitenoise", "pm"):
        spec = importlib.util.spec_from_file_location("pmesh_ref." + name, os.path.join(REF, "pmesh", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["pmesh_ref." + name] = mod
        spec.loader.exec_module(mod)
        setattr(ns, name, mod)
        setattr(sys.modules["pmesh_ref"], name, mod)
    _loaded = ns
    return ns


SKIP = ("c2c", "2d_2d", "reshape", "respawn", "leak", "test_1d")     # need c2c / process meshes / pfft internals


def selftest(verbose=True):
    """run the reference's own tests (test_pm.py, test_whitenoise.py, test_gradient.py) on the stand-ins;
    returns (passed, failed) lists of test names"""
    ns = load()
    assert ns is not None, "needs /root/reference"
    sys.modules["pmesh"] = sys.modules["pmesh_ref"]
    for name in ("pm", "window", "domain", "whitenoise"):
        sys.modules["pmesh." + name] = getattr(ns, name)
    import warnings
    MPI = sys.modules["mpi4py.MPI"]
    passed, failed = [], []
    for fname in ("test_pm.py", "test_whitenoise.py", "test_gradient.py"):
        spec = importlib.util.spec_from_file_location("ref_" + fname[:-3], os.path.join(REF, "pmesh", "tests", fname))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        for tname in sorted(n for n in dir(mod) if n.startswith("test_")):
            if any(k in tname for k in SKIP):
                continue
            fn = getattr(mod, tname)
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    if fn.__code__.co_argcount:
                        fn(MPI.COMM_WORLD)
                    else:
                        fn()
                passed.append(fname + "::" + tname)
            except Exception as e:      # report, keep going
                failed.append(fname + "::" + tname + "  " + type(e).__name__ + ": " + str(e)[:120].replace("\n", " "))
    if verbose:
        print("%d passed, %d failed" % (len(passed), len(failed)))
        for f in failed:
            print("FAILED", f)
    return passed, failed


if __name__ == "__main__":
    if "--selftest" in sys.argv:
        sys.exit(1 if selftest()[1] else 0)
    ns = load()
    print("reference pm importable:", ns is not None)
