#!/usr/bin/env python
"""The three backward transforms of a force evaluation at benchmark size: the fused path (pm.gradient_fields:
transfers folded into this library's axis-0 inverse transform, pmb_ifft.cuh, + cuFFT 2-D c2r over the planes)
against the unfused one (pm.apply_gradients + pm.c2r_fields), ms per evaluation from CUDA events on the library's
stream, and the fused kernel's own time / fraction of the measured HBM copy bandwidth on its algorithmic bytes.

    python tools/bench_ifft.py --nmesh 1024 [--dtype f8]         (environment: PMB_IFFT_NARROW, PMB_IFFT_CTAS)
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
    ap.add_argument("--no-unfused", action="store_true")
    a = ap.parse_args()
    from pmesh_b200 import comm as C, transfer as T
    from pmesh_b200.pm import ParticleMesh, RealField, apply_gradients, c2r_fields, gradient_fields
    comm = C.world()
    M = a.nmesh
    pm = ParticleMesh(BoxSize=float(M), Nmesh=[M, M, M], dtype=a.dtype, comm=comm)
    ctx = pm.ctx
    peak, _ = bench.peaks()
    rhok = pm.generate_whitenoise(1, type="complex")
    tmp = [pm.create("complex") for d in range(3)]
    treal = [RealField(pm, t._base) for t in tmp]
    tf = [T.GravityFD4(d) for d in range(3)]

    def timeit(fn):
        for _ in range(2):
            fn()
        ctx.timer_start(3)
        for _ in range(a.reps):
            fn()
        return comm.allreduce(ctx.timer_stop(3) / a.reps, op=C.MAX)

    out = {"nmesh": M, "dtype": a.dtype, "n_gpus": comm.size,
           "env": dict((k, v) for k, v in os.environ.items() if k.startswith("PMB_"))}
    pm.fft_fused_stats(reset=True)
    pm.fft_library_ms(reset=True)
    out["fused_ms"] = round(timeit(lambda: gradient_fields(rhok, tf, outs=treal)), 3)
    ms, n = pm.fft_fused_stats(reset=True)
    out["fused_cufft_ms"] = round(pm.fft_library_ms(reset=True) / (a.reps + 2), 3)
    if n:
        ncell = 1
        for s in pm._layout['o_shape']:
            ncell *= int(s)
        per_eval = n / float(a.reps + 2)
        nbytes = ncell * 2 * pm.dtype.itemsize * (1 + 3.0 / per_eval)
        out["kernel"] = {"name": "pmb_k_ifft_grad", "launches_per_evaluation": per_eval, "ms_per_launch": round(ms / n, 3),
                         "algorithmic_gb_per_launch": round(nbytes / 1e9, 3),
                         "achieved_gbs": round(nbytes / (ms / n * 1e-3) / 1e9, 1),
                         "frac_of_hbm_peak": round(nbytes / (ms / n * 1e-3) / 1e9 / peak, 4)}
    if comm.size == 1:
        from pmesh_b200.pm import force_fields
        rho = pm.generate_whitenoise(2, type="real")
        rk = pm.create("complex")
        pm.fft_fused_stats(reset=True)
        pm.fft_library_ms(reset=True)
        out["r2c_then_fused_ms"] = round(timeit(lambda: gradient_fields(rho.r2c(out=rk), tf, outs=treal)), 3)
        pm.fft_fused_stats(reset=True)
        pm.fft_library_ms(reset=True)
        out["whole_ms"] = round(timeit(lambda: force_fields(rho, tf, outs=treal)), 3)
        ms, n = pm.fft_fused_stats(reset=True)
        out["whole_cufft_ms"] = round(pm.fft_library_ms(reset=True) / (a.reps + 2), 3)
        out["whole_kernel_ms_per_launch"] = round(ms / max(n, 1), 3)
        del rho, rk
    if not a.no_unfused:
        out["unfused_ms"] = round(timeit(lambda: c2r_fields(apply_gradients(rhok, tf, outs=tmp), outs=[Ellipsis] * 3)), 3)
        out["unfused_cufft_ms"] = round(pm.fft_library_ms(reset=True) / (a.reps + 2), 3)
    if comm.rank == 0:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
