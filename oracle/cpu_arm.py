"""The reference's CPU implementation of the PM force step on all host cores -- TEST / BENCH
INFRASTRUCTURE, not product code (only bench.py's ``--impl reference`` and ``cpu_baseline`` legs use it).

The reference is SPMD: one serial process per MPI rank, the mesh split into slabs, PFFT for the
transforms (pmesh/pm.py).  MPI and PFFT do not exist in this image, so the same program is run as K
forked worker processes over shared memory, one slab of the real mesh per worker:

  per worker (its own lattice block of particles, reference code, one core each)
    decompose   : oracle.decompose -- the numpy part of GridND.decompose (pmesh/domain.py:561-652) +
                  gridnd_fill, on the K-slab grid
    exchange    : ndarray.take(indices) of the self block (domain.py:188); the ghosts a rank RECEIVES from
                  its neighbours were routed once during set-up and are appended (the MPI Alltoallv itself
                  is not modelled: communication is free for the reference)
    paint       : the reference's own compiled C (oracle/_ref: pmesh/_window_imp.c through _window.pyx)
                  into the worker's slab of the shared mesh, translate = -slab start (pm.py:1466-1469)
    transfer    : numpy on the worker's slab of k space (examples/nbody.py:162-170)
    readout x 3 : the reference's compiled C on the slab (+ one ghost plane), then bincount gather
  parent
    r2c, c2r x3 : scipy.fft (pocketfft) with workers=K threads on the whole shared mesh -- the labelled
                  stand-in for PFFT

``force_step(n, window, cores, steps)`` returns the wall-clock seconds per step.  Every array is first
touched before the timed region; the timed region is bracketed by barriers.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))


def _kernels():
    """(paint, readout, kind): the compiled reference when oracle/_ref is present, else the oracle port"""
    sys.path.insert(0, HERE)
    import build_ref
    import oracle
    oracle.build()
    mods = build_ref.load()
    if mods is not None:
        w = mods[0]

        class RW(w.ResampleWindow):      # the cdef class stores attributes: it needs a Python subclass (as pmesh/window.py has)
            pass

        def paint(real, pos, window, translate, period):
            nd = real.ndim
            RW(oracle.NAMES[window]).paint(
                real, pos, None, numpy.ones(len(pos)), numpy.zeros(nd, dtype=int), numpy.ones(nd),
                numpy.asarray(translate, dtype="f8"), numpy.asarray(period, dtype=numpy.intp))

        def readout(real, pos, window, translate, period):
            nd = real.ndim
            out = numpy.zeros(len(pos))
            RW(oracle.NAMES[window]).readout(
                real, pos, None, out, numpy.zeros(nd, dtype=int), numpy.ones(nd),
                numpy.asarray(translate, dtype="f8"), numpy.asarray(period, dtype=numpy.intp))
            return out
        return paint, readout, "reference", oracle

    def paint(real, pos, window, translate, period):
        oracle.paint(real, pos, window, scale=1.0, translate=translate, period=period)

    def readout(real, pos, window, translate, period):
        return oracle.readout(real, pos, window, scale=1.0, translate=translate, period=period)
    return paint, readout, "port", oracle


def _lattice_block(n, p0, p1):
    """particles of lattice planes [p0, p1) (indices taken mod n): cell centres displaced by one sine mode per
    axis, amplitude 3 cells -- the closed form of bench.py's `lattice` input (BoxSize = n, cell units)"""
    planes = numpy.arange(p0, p1) % n
    q = numpy.stack(numpy.meshgrid(planes, numpy.arange(n), numpy.arange(n), indexing="ij"), axis=-1).reshape(-1, 3) + 0.5
    x = (q + 3.0 * numpy.sin(2 * numpy.pi * 4 * q[:, ::-1] / n)) % n
    return x


def _shared(shape, dtype):
    nbytes = int(numpy.prod(shape)) * numpy.dtype(dtype).itemsize
    raw = mp.RawArray("b", nbytes)
    return numpy.frombuffer(raw, dtype=dtype).reshape(shape)


def _worker(r, K, n, window, steps, barrier, rho, ck, tk, fr, support, fsq):
    try:
        _worker_body(r, K, n, window, steps, barrier, rho, ck, tk, fr, support, fsq)
    except BaseException:
        barrier.abort()                             # let everybody fail instead of waiting for ever
        raise


def _worker_body(r, K, n, window, steps, barrier, rho, ck, tk, fr, support, fsq):
    paint, readout, kind, oracle = _kernels()
    edges = [numpy.linspace(0, n, K + 1), numpy.array([0.0, n]), numpy.array([0.0, n])]
    s0, s1 = r * n // K, (r + 1) * n // K
    smoothing = 1.0 * support                       # bench.py / nbody.py: smoothing = 1.0 * resampler.support
    # ---- set-up (untimed): my lattice block; the ghosts my neighbours would send me ----
    X = _lattice_block(n, s0, s1)
    ghosts = []
    if K > 1:
        halo = 4 + int(numpy.ceil(smoothing)) + 1    # displacement amplitude 3(+) cells + smoothing, in planes
        for a, b in ((s0 - halo, s0), (s1, s1 + halo)):
            cand = _lattice_block(n, a, b)
            counts, ind = oracle.decompose(cand, edges, K, smoothing=smoothing)
            off = numpy.concatenate([[0], numpy.cumsum(counts)])
            ghosts.append(cand.take(ind[off[r]:off[r + 1]], axis=0))
    # wavenumbers of this worker's slab of the half-complex grid; Nyquist negative (pm.py:1213-1219)
    def wn(m):
        w = numpy.arange(m, dtype="f8")
        w[w >= n // 2] -= n
        return w * (2 * numpy.pi / n)
    kx = [wn(n)[s0:s1, None, None], wn(n)[None, :, None], wn(n // 2 + 1)[None, None, :]]
    translate = numpy.array([-float(s0), 0.0, 0.0])
    period = [n, n, n]
    for _ in range(steps):
        barrier.wait()                              # step start
        counts, ind = oracle.decompose(X, edges, K, smoothing=smoothing)
        off = numpy.concatenate([[0], numpy.cumsum(counts)])
        mine = ind[off[r]:off[r + 1]]
        lpos = numpy.concatenate([X.take(mine, axis=0)] + ghosts, axis=0)
        rho[s0:s1] = 0.0                            # pm.paint(hold=False)
        paint(rho[s0:s1], lpos, window, translate, period)
        barrier.wait()                              # -> parent: x Nmesh^3 / Np, r2c
        for d in range(3):
            barrier.wait()                          # parent finished r2c (d == 0) / the previous readout round
            k2 = kx[0] ** 2 + kx[1] ** 2 + kx[2] ** 2
            k2[k2 == 0] = 1.0
            wd = kx[d]                               # C = BoxSize / Nmesh = 1
            tk[s0:s1] = 1j * (1.0 / 6.0 * (8 * numpy.sin(wd) - numpy.sin(2 * wd))) / k2 * ck[s0:s1]
            barrier.wait()                          # -> parent: c2r
            barrier.wait()                          # parent finished c2r
            f = readout(fr[s0:s1], lpos, window, translate, period)
            F = oracle.bincount_sum(mine, f[:len(mine)], len(X))  # Layout.gather('sum') of the self block
            # diagnostic only (tests): the ghosts' share would come back through the reverse Alltoallv
            fsq[r, d] = float((F * F).sum())
        barrier.wait()                              # step end


def force_step(n=512, window="cic", cores=None, steps=1, warmup=0):
    """seconds per force step (mean over `steps`) of the reference's CPU path on `cores` processes"""
    paint, readout, kind, oracle = _kernels()
    support = oracle.window_support(window)[0]
    K = int(cores or os.cpu_count() or 1)
    while n % K:
        K -= 1
    try:
        import scipy.fft as sfft
        rfftn = lambda a: sfft.rfftn(a, workers=K)
        irfftn = lambda a, out: numpy.copyto(out, sfft.irfftn(a, s=(n, n, n), workers=K))
        fft_name = "scipy.fft workers=%d" % K
    except Exception:
        rfftn = numpy.fft.rfftn
        irfftn = lambda a, out: numpy.copyto(out, numpy.fft.irfftn(a, s=(n, n, n)))
        fft_name = "numpy.fft"
    ctx = mp.get_context("fork")
    rho = _shared((n, n, n), "f8")
    fr = _shared((n, n, n), "f8")
    ck = _shared((n, n, n // 2 + 1), "c16")
    tk = _shared((n, n, n // 2 + 1), "c16")
    fsq = _shared((K, 3), "f8")
    for a in (rho, fr, ck, tk):
        a[...] = 0                                   # first touch outside the timed region
    total = steps + warmup
    barrier = ctx.Barrier(K + 1)
    procs = [ctx.Process(target=_worker, args=(r, K, n, window, total, barrier, rho, ck, tk, fr, support, fsq)) for r in range(K)]
    for p in procs:
        p.start()
    times = []
    try:
        for it in range(total):
            barrier.wait(timeout=3600)               # step start (after the workers' set-up on the first)
            t0 = time.perf_counter()
            barrier.wait(timeout=3600)               # paint done
            rho *= 1.0                               # rho *= Nmesh^3 / Np (nbody.py:207); Np = Nmesh^3 here
            ck[...] = rfftn(rho)
            ck *= 1.0 / float(n) ** 3
            for d in range(3):
                barrier.wait(timeout=3600)           # workers: transfer
                barrier.wait(timeout=3600)           # transfer done
                irfftn(tk, fr)
                fr *= float(n) ** 3
                barrier.wait(timeout=3600)           # workers: readout
            barrier.wait(timeout=3600)               # step end
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.terminate()
    return {"seconds_per_step": float(numpy.mean(times)), "kind": kind, "cores": K, "n": n, "fft": fft_name,
            "steps": steps, "rho_sum": float(rho.sum()), "force_sq_self_blocks": fsq.sum(axis=0).tolist()}


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    print(force_step(n=n, steps=1, warmup=0, cores=int(sys.argv[2]) if len(sys.argv) > 2 else None))
