"""BASELINE configs[4] in miniature (or in full on 8 GPUs): white-noise initial conditions, 1-LPT
displacements, then KDK particle-mesh steps with a force mesh `boost` times finer than the particle
grid -- the flow of the reference's examples/nbody.py (simulate(), :245-288) on device-resident state.

    python tools/bench_kdk.py --npart 256 --boost 2 --steps 10
    python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_kdk.py --npart 1024 --boost 2

Prints one JSON line: ms per KDK step (two kicks, one drift, one force evaluation), ms for the IC.
"""
import argparse
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Naive(object):
    """ the reference's `Naive` step factors for an Einstein-de Sitter E(a) = a^-1.5 """
    @staticmethod
    def K(ai, af, ar):
        return 1.0 / (ar * ar * ar ** -1.5) * (af - ai)

    @staticmethod
    def D(ai, af, ar):
        return 1.0 / (ar * ar * ar * ar ** -1.5) * (af - ai)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--npart", type=int, default=256, help="particles per side")
    ap.add_argument("--boost", type=int, default=2, help="force mesh = boost x particle grid")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--seed", type=int, default=120577)      # examples/nbody.py:338
    a = ap.parse_args()
    from pmesh_b200 import comm as C, nbody, transfer as T
    from pmesh_b200.device import DeviceArray
    from pmesh_b200.pm import ParticleMesh
    comm = C.world()
    L = float(a.npart)
    pm = ParticleMesh(BoxSize=L, Nmesh=[a.npart] * 3, dtype="f8", comm=comm)
    ctx = pm.ctx
    t0 = time.perf_counter()
    wn = pm.generate_whitenoise(a.seed, unitary=True)
    # a smooth, red spectrum: sqrt(P) ~ exp(-k^2 r^2 / 2) / k on the mesh scale (stand-in for a CDM P(k))
    dlinear = wn.apply(T.GaussianLowpass(2.0)).apply(T.InverseLaplace()).apply(T.Scale(-0.05))
    Q = DeviceArray.from_host(pm.generate_uniform_particle_grid(shift=0.0))
    DX1 = nbody.lpt1(pm, dlinear, Q)
    a0 = 0.1
    # normalise the displacement field to 0.3 cells rms at the start (it grows ~ 10x to a = 1)
    rms = (comm.allreduce(DX1.dot(DX1), op=C.SUM) / (3.0 * comm.allreduce(Q.shape[0], op=C.SUM))) ** 0.5
    S = DeviceArray.empty(Q.shape, "f8").assign_lincomb(DX1, 0.3 / rms)
    V = DeviceArray.empty(Q.shape, "f8").assign_lincomb(S, a0 ** 2 * a0 ** -1.5)
    ctx.sync()
    t_ic = (time.perf_counter() - t0) * 1e3
    del wn, dlinear, DX1
    fpm = ParticleMesh(BoxSize=L, Nmesh=[a.npart * a.boost] * 3, dtype="f8", resampler="cic", comm=comm)
    state = nbody.State(Q, S, V)
    steps = numpy.linspace(a0, 1.0, a.steps + 1)
    nbody.symp2(fpm, state, steps[:2], Naive, 0.3)           # warm-up step (plans, schedules, allocator)
    comm.Barrier()
    ctx.sync()
    ctx.timer_start(5)
    nbody.symp2(fpm, state, steps, Naive, 0.3)
    ms = ctx.timer_stop(5)
    ms = comm.allreduce(ms, op=C.MAX)
    disp = comm.allreduce(state.S.dot(state.S), op=C.SUM)
    ntot = comm.allreduce(Q.shape[0], op=C.SUM)
    if comm.rank == 0:
        print(json.dumps({"workload": "white-noise IC + 1-LPT + %d KDK steps, %d^3 particles, %d^3 force mesh" % (a.steps, a.npart, a.npart * a.boost),
                          "n_gpus": comm.size, "ic_ms": round(t_ic, 1), "kdk_ms_per_step": round(ms / a.steps, 3),
                          "force_evaluations": a.steps + 1, "rms_displacement_cells": round((disp / (3.0 * ntot)) ** 0.5, 4)}))


if __name__ == "__main__":
    main()
