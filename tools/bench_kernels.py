#!/usr/bin/env python
"""A/B timing of the paint / readout kernels alone at benchmark size under different run-time switches
(environment variables read by the library per call), on the three bench inputs.

    python tools/bench_kernels.py --nmesh 1024 --inputs lattice,zeldovich,uniform \
        --env "PMB_RING=0" --env "PMB_RING=1" --env "PMB_RING=1 PMB_RING_READOUT3_MINB=2"

One JSON line per (input, env) with ms per launch (CUDA events on the library's stream, 5 launches after 2
warm-ups) and the fraction of the measured HBM copy bandwidth on the algorithmic bytes."""
import argparse
import json
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nmesh", type=int, default=1024)
    ap.add_argument("--window", default="cic")
    ap.add_argument("--dtype", default="f8")
    ap.add_argument("--inputs", default="lattice,zeldovich,uniform")
    ap.add_argument("--env", action="append", default=[])
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    from pmesh_b200 import comm as C
    from pmesh_b200.device import DeviceArray
    from pmesh_b200.pm import ParticleMesh
    comm = C.world()
    M = a.nmesh
    pm = ParticleMesh(BoxSize=float(M), Nmesh=[M, M, M], dtype=a.dtype, resampler=a.window, comm=comm)
    ctx = pm.ctx
    peak, _ = bench.peaks()
    a.particles = "zeldovich"
    a.paint_mode = "atomic"
    envs = a.env or [""]
    fields = [pm.create("real") for _ in range(3)]
    for f in fields:
        f.fill(1.0)
    es = pm.dtype.itemsize
    ncell = int(numpy.prod(pm._layout['i_shape']))
    for kind in a.inputs.split(","):
        X = bench.make_particles(pm, a, comm, kind)[0]
        n = X.shape[0]
        outs = [DeviceArray.empty((n,), "f8") for _ in range(3)]
        for env in envs:
            saved = {}
            for kv in env.split():
                k, v = kv.split("=")
                saved[k] = os.environ.get(k)
                os.environ[k] = v
            row = {"input": bench.INPUT_LABEL[kind], "env": env, "nmesh": M, "window": a.window}
            mesh = [f._device() for f in fields]

            def timeit(fn):
                for _ in range(2):
                    fn()
                ctx.timer_start(3)
                for _ in range(a.reps):
                    fn()
                return ctx.timer_stop(3) / a.reps
            tp = timeit(lambda: pm.resampler.paint(mesh[0], X, transform=pm.affine, mode="atomic"))
            tr = timeit(lambda: pm.resampler.readout(mesh[0], X, out=outs[0], transform=pm.affine))
            t3 = timeit(lambda: pm.resampler.readout_multi(mesh, X, outs=outs, transform=pm.affine))
            row["paint_ms"] = round(tp, 3)
            row["readout_ms"] = round(tr, 3)
            row["readout3_ms"] = round(t3, 3)
            row["paint_frac"] = round((n * 24.0 + ncell * es) / (tp * 1e-3) / 1e9 / peak, 4)
            row["readout_frac"] = round((n * 32.0 + ncell * es) / (tr * 1e-3) / 1e9 / peak, 4)
            row["readout3_frac"] = round((n * 48.0 + 3 * ncell * es) / (t3 * 1e-3) / 1e9 / peak, 4)
            print(json.dumps(row))
            sys.stdout.flush()
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
            fields[0].fill(1.0)
        del X, outs


if __name__ == "__main__":
    main()
