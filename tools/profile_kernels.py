"""One launch of each kernel that has no line in the bench's ncu capture, at sizes that keep an
`ncu --set full` session short:

    ncu --set full --clock-control none --import-source on \\
        -k regex:"pmb_k_paint_carry32|pmb_k_route_count|pmb_k_route_fill|pmb_k_whitenoise|pmb_k_kick_drift|pmb_k_transfer|pmb_k_take|pmb_k_gather_pass" \\
        -c 12 -o gpurun_out/r1_other_kernels python tools/profile_kernels.py
"""
import ctypes
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    from bench_route import OneRankOf
    from pmesh_b200 import _lib, domain, nbody, transfer as T
    from pmesh_b200.device import DeviceArray
    from pmesh_b200.pm import ParticleMesh
    from pmesh_b200.window import FindResampler
    M = 256
    n = M ** 3
    pm = ParticleMesh(BoxSize=float(M), Nmesh=[M, M, M], dtype="f8")
    ctx = pm.ctx
    X = DeviceArray.empty((n, 3), "f8")
    box = (ctypes.c_double * 3)(float(M), float(M), float(M))
    nn = (ctypes.c_int64 * 3)(M, M, M)
    _lib.check(ctx.lib.pmb_particles_lattice(ctx.handle, X.ptr, 8, n, 3, nn, box, 0.5, 0.4, 46, 0))
    rho = pm.create("real")
    for name in ("tsc", "pcs"):                                   # pmb_k_paint_carry32<3>, <4>
        FindResampler(name).paint(rho._device(), X, transform=pm.affine, mode="atomic")
    g = domain.GridND([numpy.linspace(0, M, 9), numpy.array([0.0, M]), numpy.array([0.0, M])], comm=OneRankOf(0, 8))
    lay = g.decompose(X, smoothing=1.0)                           # pmb_k_route_count / fill, 8 slabs
    send = DeviceArray.empty((int(lay.sendcounts.sum()), 24), "u1")
    _lib.check(ctx.lib.pmb_take(ctx.handle, X.ptr, 24, lay.indices_device.ptr, int(lay.sendcounts.sum()), send.ptr))
    vals = DeviceArray.empty((int(lay.sendcounts.sum()),), "f8")
    out = DeviceArray.empty((n,), "f8")
    offs = numpy.zeros(9, dtype="i8")
    offs[1:] = numpy.cumsum(lay.sendcounts)
    _lib.check(ctx.lib.pmb_gather_sum(ctx.handle, vals.ptr, 8, 1, lay.indices_device.ptr, offs.ctypes.data, 8, n, out.ptr, 8))
    wn = pm.generate_whitenoise(1)                                # pmb_k_whitenoise<double, true>
    wn.apply(T.GravityFD4(0), out=Ellipsis)                       # pmb_k_transfer
    V = DeviceArray.zeros((n, 3), "f8")
    S = DeviceArray.zeros((n, 3), "f8")
    F = [DeviceArray.zeros((n,), "f8") for d in range(3)]
    nbody.kick_drift(V, F, 0.1, S, 0.2)                           # pmb_k_kick_drift
    ctx.sync()
    print("done")


if __name__ == "__main__":
    main()
