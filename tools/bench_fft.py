#!/usr/bin/env python
"""r2c / c2r / transfer alone at benchmark size: ms per call (CUDA events on the library's stream) and the
fraction of the measured HBM copy bandwidth on the algorithmic bytes (3 passes x read + write of the mesh
for a 3-D transform = 48 B / cell in f8; 32 B / cell for one transfer).

    python tools/bench_fft.py --nmesh 1024 [--dtype f8]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nmesh", type=int, default=1024)
    ap.add_argument("--dtype", default="f8")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    from pmesh_b200 import comm as C, transfer as T
    from pmesh_b200.pm import ParticleMesh
    comm = C.world()
    M = a.nmesh
    pm = ParticleMesh(BoxSize=float(M), Nmesh=[M, M, M], dtype=a.dtype, comm=comm)
    ctx = pm.ctx
    peak, _ = bench.peaks()
    es = pm.dtype.itemsize
    ncell = M ** 3 / comm.size
    rho = pm.generate_whitenoise(1, type="real")
    rhok = pm.create("complex")
    tmp = pm.create("complex")
    tmp2 = pm.create("complex")
    real2 = pm.create("real")

    def timeit(fn):
        for _ in range(2):
            fn()
        ctx.timer_start(3)
        for _ in range(a.reps):
            fn()
        return ctx.timer_stop(3) / a.reps

    rows = {}
    rows["r2c_out_of_place"] = timeit(lambda: rho.r2c(out=rhok))
    rows["transfer"] = timeit(lambda: rhok.apply(T.GravityFD4(0), out=tmp))
    rows["transfer_inplace"] = timeit(lambda: tmp2.apply(T.GravityFD4(0), out=Ellipsis))

    def c2r_ip():
        rhok.apply(T.Scale(1.0), out=tmp)
        tmp.c2r(out=Ellipsis)

    def c2r_oop():
        rhok.apply(T.Scale(1.0), out=tmp)
        tmp.c2r(out=real2)
    t_copy = timeit(lambda: rhok.apply(T.Scale(1.0), out=tmp))
    rows["c2r_in_place"] = timeit(c2r_ip) - t_copy
    rows["c2r_out_of_place"] = timeit(c2r_oop) - t_copy
    rows["r2c_in_place"] = timeit(lambda: real2.r2c(out=Ellipsis).c2r(out=Ellipsis)) - rows["c2r_in_place"]
    out = {"nmesh": M, "dtype": a.dtype, "n_gpus": comm.size}
    for k, v in rows.items():
        v = comm.allreduce(v, op=C.MAX)
        ab = (4 * es if k.startswith("transfer") else 6 * es) * ncell
        out[k + "_ms"] = round(v, 3)
        out[k + "_frac"] = round(ab / (v * 1e-3) / 1e9 / peak, 3)
    if comm.rank == 0:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
