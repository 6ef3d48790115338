#!/usr/bin/env python
"""Timing of paint / readout on a particle array WITHOUT spatial order (uniform random), at benchmark size:
the reorder (pmb_bin.cuh) built from scratch, re-validated (content hash) and reused, against the permutation
walk (PMB_BIN=0).

    python tools/bench_bin.py --nmesh 1024 [--env "PMB_BIN_PLAIN=1"] [--perm]

One JSON line per env: ms per call (CUDA events on the library's stream).
  *_build_ms : the call right after pmb_bin_release (probe + count + scan + scatter + the op itself)
  *_ms       : the call with a valid cached copy (hash pass + the op itself [+ return to the caller's order])"""
import argparse
import ctypes
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
    ap.add_argument("--input", default="uniform")
    ap.add_argument("--env", action="append", default=[])
    ap.add_argument("--perm", action="store_true", help="also time the permutation walk (PMB_BIN=0)")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    from pmesh_b200 import _lib
    from pmesh_b200 import comm as C
    from pmesh_b200.device import DeviceArray
    from pmesh_b200.pm import ParticleMesh
    comm = C.world()
    M = a.nmesh
    pm = ParticleMesh(BoxSize=float(M), Nmesh=[M, M, M], dtype="f8", resampler=a.window, comm=comm)
    ctx = pm.ctx
    peak, _ = bench.peaks()
    a.particles = a.input
    a.paint_mode = "atomic"
    fields = [pm.create("real") for _ in range(3)]
    for f in fields:
        f.fill(1.0)
    mesh = [f._device() for f in fields]
    X = bench.make_particles(pm, a, comm, a.input)[0]
    n = X.shape[0]
    outs = [DeviceArray.empty((n,), "f8") for _ in range(3)]
    envs = (a.env or [""]) + (["PMB_BIN=0"] if a.perm else [])

    def release():
        _lib.check(ctx.lib.pmb_bin_release(ctx.handle))

    def once(fn):
        ctx.timer_start(3)
        fn()
        return ctx.timer_stop(3)

    def cached(fn):
        fn()
        ctx.timer_start(3)
        for _ in range(a.reps):
            fn()
        return ctx.timer_stop(3) / a.reps

    ops = {
        "paint": lambda: pm.resampler.paint(mesh[0], X, transform=pm.affine, mode="atomic"),
        "readout": lambda: pm.resampler.readout(mesh[1], X, out=outs[0], transform=pm.affine),
        "readout3": lambda: pm.resampler.readout_multi(mesh, X, outs=outs, transform=pm.affine),
    }
    bytes_alg = {"paint": 32.0, "readout": 40.0, "readout3": 72.0}
    for env in envs:
        saved = {}
        for kv in env.split():
            k, v = kv.split("=")
            saved[k] = os.environ.get(k)
            os.environ[k] = v
        row = {"input": bench.INPUT_LABEL[a.input], "env": env, "nmesh": M, "window": a.window, "particles": n}
        for name, fn in ops.items():
            fn()                      # warm: scratch, schedules
            release()
            row[name + "_build_ms"] = round(once(fn), 3)
            t = cached(fn)
            row[name + "_ms"] = round(t, 3)
            row[name + "_frac"] = round(n * bytes_alg[name] / (t * 1e-3) / 1e9 / peak, 4)
        release()
        # one force evaluation's worth: paint builds, the three-field gather reuses
        t = once(lambda: (ops["paint"](), ops["readout3"]()))
        row["paint+readout3_ms"] = round(t, 3)
        row["paint+readout3_frac"] = round(n * (32.0 + 72.0) / (t * 1e-3) / 1e9 / peak, 4)
        b, held = ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.check(ctx.lib.pmb_bin_stats(ctx.handle, ctypes.byref(b), ctypes.byref(held)))
        row["reorders"] = b.value
        row["bytes_held"] = held.value
        print(json.dumps(row))
        sys.stdout.flush()
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        fields[0].fill(1.0)


if __name__ == "__main__":
    main()
