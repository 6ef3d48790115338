"""Time GridND.decompose (routing kernels) in the geometry of a P-rank slab job on ONE GPU.

    python tools/bench_route.py [--n 268435456] [--nmesh 1024]

The communicator is a stand-in that only reports (rank, size): the routing kernels see exactly the
edges / DomainAssign / rank count of the real job, no NCCL is involved.
"""
import argparse
import ctypes
import json
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class OneRankOf(object):
    def __init__(self, rank, size):
        self.rank, self.size = rank, size

    def Barrier(self):
        pass

    def Alltoall(self, send, recv):
        recv[...] = send

    def allgather(self, x):
        return [x] * self.size

    def bcast(self, x, root=0):
        return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 28)
    ap.add_argument("--nmesh", type=int, default=1024)
    a = ap.parse_args()
    from pmesh_b200 import _lib, domain
    from pmesh_b200.device import DeviceArray
    M = a.nmesh
    X = DeviceArray.empty((a.n, 3), "f8")
    ctx = X.ctx
    box = (ctypes.c_double * 3)(float(M), float(M), float(M))
    nn = (ctypes.c_int64 * 3)(M, M, M)
    _lib.check(ctx.lib.pmb_particles_lattice(ctx.handle, X.ptr, 8, a.n, 3, nn, box, 0.5, 3.0, 44, 0))
    ctx.sync()
    out = {}
    for P, shape in ((2, (2, 1, 1)), (8, (8, 1, 1)), (8, (2, 4, 1)), (32, (4, 8, 1))):
        edges = [numpy.linspace(0, M, s + 1) for s in shape]
        g = domain.GridND(edges, comm=OneRankOf(0, P))
        for _ in range(2):
            lay = g.decompose(X, smoothing=1.0)
        ctx.sync()
        ctx.timer_start(3)
        for _ in range(5):
            lay = g.decompose(X, smoothing=1.0)
        ms = ctx.timer_stop(3) / 5
        out["P=%d %s" % (P, "x".join(map(str, shape)))] = {
            "ms": round(ms, 3), "Gparticles_per_s": round(a.n / ms / 1e6, 2),
            "GB_per_s_algorithmic(28B)": round(a.n * 28 / ms / 1e6, 1), "nsend": int(lay.sendcounts.sum())}
    print(json.dumps({"n": a.n, "routing": out}))


if __name__ == "__main__":
    main()
