"""Paint / readout / gradient throughput per resampling window on ONE GPU (BASELINE configs[1] and [3]
shapes: TSC at 256^3, PCS and lanczos3 with the vjp operators at 512^3, CIC at 512^3 for reference).

    python tools/bench_windows.py [--n 512] [--windows cic,tsc,pcs,lanczos3]

Per window: paint (atomic and deterministic), readout, readout_grad (value + 3 gradients in one
sweep = the back-propagation kernel of paint_vjp / readout_vjp), in ms and Mparticles/s, with the
algorithmic HBM bytes of SURVEY 8(d) as a fraction of the measured copy bandwidth.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--windows", default="cic,tsc,pcs,lanczos3")
    ap.add_argument("--dtype", default="f8")
    a = ap.parse_args()
    from pmesh_b200 import _lib
    from pmesh_b200.device import DeviceArray
    from pmesh_b200.pm import ParticleMesh
    from pmesh_b200.window import FindResampler
    peak = 6549.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    M = a.n
    npart = M ** 3
    pm = ParticleMesh(BoxSize=float(M), Nmesh=[M, M, M], dtype=a.dtype)
    ctx = pm.ctx
    X = DeviceArray.empty((npart, 3), "f8")
    box = (ctypes.c_double * 3)(float(M), float(M), float(M))
    nn = (ctypes.c_int64 * 3)(M, M, M)
    # cfg4: particles offset 0.1 - 0.9 cell from the lattice: lattice + smooth displacement
    _lib.check(ctx.lib.pmb_particles_lattice(ctx.handle, X.ptr, 8, npart, 3, nn, box, 0.5, 0.4, 46, 0))
    ctx.sync()
    es = pm.dtype.itemsize
    ncell = M ** 3
    res = {}
    for name in a.windows.split(","):
        w = FindResampler(name)
        rho = pm.create("real")
        out = DeviceArray.empty((npart,), "f8")
        r = {"support": int(w.support), "points": int(w.support) ** 3}

        def timeit(fn, reps):
            fn()
            ctx.sync()
            ctx.timer_start(4)
            for _ in range(reps):
                fn()
            return ctx.timer_stop(4) / reps
        reps = 3 if w.support <= 4 else 1
        t = timeit(lambda: w.paint(rho._device(), X, transform=pm.affine, mode="atomic"), reps)
        ab = npart * 24.0 + ncell * es
        r["paint_atomic"] = {"ms": round(t, 3), "Mp_s": round(npart / t / 1e3, 1), "frac_hbm": round(ab / t / 1e6 / peak, 4)}
        if w.support <= 4:
            t = timeit(lambda: w.paint(rho._device(), X, transform=pm.affine, mode="deterministic"), 1)
            r["paint_deterministic"] = {"ms": round(t, 3), "Mp_s": round(npart / t / 1e3, 1)}
        t = timeit(lambda: w.readout(rho._device(), X, out=out, transform=pm.affine), reps)
        ab = npart * 32.0 + ncell * es
        r["readout"] = {"ms": round(t, 3), "Mp_s": round(npart / t / 1e3, 1), "frac_hbm": round(ab / t / 1e6 / peak, 4)}
        t = timeit(lambda: w.readout_grad(rho._device(), X, transform=pm.affine), reps)
        ab = npart * (24.0 + 8.0 * 4) + ncell * es
        r["readout_value_and_3_gradients"] = {"ms": round(t, 3), "Mp_s": round(npart / t / 1e3, 1), "frac_hbm": round(ab / t / 1e6 / peak, 4)}
        t = timeit(lambda: w.paint(rho._device(), X, transform=pm.affine, diffdir=0, mode="atomic"), reps)
        r["paint_gradient_dir0"] = {"ms": round(t, 3), "Mp_s": round(npart / t / 1e3, 1)}
        res[name] = r
        del rho, out
    print(json.dumps({"nmesh": M, "nparticles": npart, "dtype": a.dtype, "hbm_peak_gbs": peak, "windows": res}))


if __name__ == "__main__":
    main()
