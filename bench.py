#!/usr/bin/env python
"""bench.py -- the PM force step of BASELINE.json on N B200s of one node.

A "step" is one full CIC particle-mesh force evaluation (examples/nbody.py:199-218 of the
reference): decompose -> exchange -> paint -> x N^3/Np -> r2c -> {gravity transfer -> c2r ->
readout -> ghost sum} x 3, for `--nmesh`^3 particles on a `--nmesh`^3 mesh (default 1024: the
configuration BASELINE.json's metric "PM force step ms (1024^3)" is quoted on; it fits one GPU).
Total work is fixed as N grows (strong scaling); each rank starts with the particles of its own
lattice slab, displaced Zel'dovich-style.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # the reference's CPU kernels on the host cores

Prints ONE JSON line (rank 0).  Timing: CUDA events on the library's stream, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WINDOW_NAMES = {"cic": "CIC", "tsc": "TSC", "pcs": "PCS", "nnb": "NNB"}
INPUTS = ("zeldovich", "uniform", "lattice")
INPUT_LABEL = {"zeldovich": "zeldovich_gaussian", "uniform": "uniform_random", "lattice": "lattice_sine"}


def metric_name(args):
    """BASELINE.json's metric, spelled with the mesh and window actually run"""
    return ("PM force step ms (%d^3 %s: decompose, paint, r2c, gravity transfer, c2r x3, readout x3)"
            % (args.nmesh, WINDOW_NAMES.get(args.window, args.window)))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nmesh", type=int, default=1024)
    ap.add_argument("--window", default="cic")
    ap.add_argument("--dtype", default="f8")
    ap.add_argument("--particles", default="zeldovich", choices=list(INPUTS),
                    help="input of the timed force step: zeldovich = lattice displaced by a Gaussian random field "
                         "(P(k) ~ k^-2, rms 3 cells; SURVEY cfg3), uniform = uniform random (no order at all), "
                         "lattice = lattice displaced by one sine mode per axis (order fully preserved)")
    ap.add_argument("--inputs", default="zeldovich,uniform,lattice",
                    help="inputs for which paint / readout are timed alone (comma separated)")
    ap.add_argument("--no-verify", action="store_true", help="skip the full-size parity check against the oracle")
    ap.add_argument("--np", default="", help="process mesh, e.g. 2,4 for pencils (default: slabs, np = [gpus])")
    ap.add_argument("--paint-mode", default="atomic", choices=["atomic", "deterministic"])
    ap.add_argument("--breakdown", action="store_true", help="also time every stage separately (stderr)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--split", action="store_true", help="P > 1: always the split exchange (default: measured choice)")
    ap.add_argument("--no-split", action="store_true",
                    help="P > 1: move every particle through take + alltoallv (Layout.exchange) instead of painting / reading the "
                         "rank's own particles where they lie (Layout.exchange_remote)")
    ap.add_argument("--no-forward-fusion", action="store_true",
                    help="one rank: full r2c by cuFFT, then the fused transfer + axis-0 inverse kernel (pm.gradient_fields)")
    ap.add_argument("--unfused", action="store_true",
                    help="transfer pass + cuFFT for all three axes of the backward transforms (the path before pmb_ifft.cuh)")
    ap.add_argument("--e2e-double", action="store_true", help="e2e: upload || compute || download with two position buffers")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=256,
                    help="side of the sample problem of the cpu_baseline leg of the default arm (a few seconds on all cores)")
    ap.add_argument("--cpu-sample-reference", type=int, default=512,
                    help="side of the sample problem each step of `--impl reference` measures (~10 s per step on 16 cores)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.p = None
        self.device = device
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(numpy.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ ours
def _slab_rows(M, comm):
    """rows [first, first + n) of the C-order M^3 lattice owned by this rank: whole planes"""
    planes = (M + comm.size - 1) // comm.size
    first = min(comm.rank * planes, M) * M * M
    n = min((comm.rank + 1) * planes, M) * M * M - first
    return first, n


def make_particles(pm, args, comm, kind=None):
    """device-resident positions of this rank (rank r starts with the particles of lattice slab r):

    zeldovich : the lattice (shift 0.5) displaced by a Gaussian random displacement field generated with
                the engine's own operators, the reference's 1-LPT recipe (examples/nbody.py:253-274):
                white noise (seed 44) -> |k|^-1 (P(k) ~ k^-2) -> i k_d / k^2 -> c2r -> readout at the
                lattice, normalised to rms |psi| = 3 cells; x = (q + psi) mod L.
    uniform   : uniform random in the whole box (counter-based generator, seed 45): no order.
    lattice   : the lattice displaced by one sine mode per axis, amplitude 3 cells: lattice order kept."""
    import ctypes
    from pmesh_b200 import _lib, comm as C
    from pmesh_b200 import transfer as T
    from pmesh_b200.device import DeviceArray
    kind = kind or args.particles
    M = args.nmesh
    ntot = M ** 3
    first, n = _slab_rows(M, comm)
    X = DeviceArray.empty((n, 3), "f8")
    ctx = X.ctx
    box = (ctypes.c_double * 3)(*[float(b) for b in pm.BoxSize])
    nn = (ctypes.c_int64 * 3)(M, M, M)
    if kind == "uniform":
        _lib.check(ctx.lib.pmb_particles_uniform(ctx.handle, X.ptr, 8, n, 3, box, 45, first))
    elif kind == "lattice":
        _lib.check(ctx.lib.pmb_particles_lattice(ctx.handle, X.ptr, 8, n, 3, nn, box, 0.5, 3.0, 44, first))
    else:
        _lib.check(ctx.lib.pmb_particles_lattice(ctx.handle, X.ptr, 8, n, 3, nn, box, 0.5, 0.0, 44, first))   # q
        delta = pm.generate_whitenoise(44, type="complex")
        delta = delta.apply(T.PowerLaw(-1.0), out=Ellipsis)
        layout = pm.decompose(X)
        lq = layout.exchange(X)
        tmp = pm.create("complex")
        psi = []
        for d in range(3):
            f = delta.apply(T.GradientK(d), out=tmp).c2r(out=Ellipsis)
            psi.append(layout.gather(f.readout(lq)))
            del f
        del tmp, delta, lq, layout
        sq = comm.allreduce(sum(p.dot(p) for p in psi), op=C.SUM)
        cell = float(pm.BoxSize[0]) / M
        fac = 3.0 * cell / max((sq / ntot) ** 0.5, 1e-300)
        for d in range(3):
            X.column(d).iadd_scaled(psi[d], fac)
            X.column(d).imod(float(pm.BoxSize[d]))
        del psi
    ctx.sync()
    return X, ntot


class ForceStep(object):
    """The force step on device-resident particles, written against the public pmesh API."""
    def __init__(self, pm, args):
        from pmesh_b200 import transfer as T
        self.pm, self.args = pm, args
        self.rho = pm.create("real")
        self._rhok = None         # density modes: only the paths that store them allocate the field
        self.tmp = [pm.create("complex") for d in range(3)]
        from pmesh_b200.pm import RealField
        self.treal = [RealField(pm, t._base) for t in self.tmp]
        self.tf = [T.GravityFD4(d) for d in range(3)]
        self.stage = {}

    @property
    def rhok(self):
        if self._rhok is None:
            self._rhok = self.pm.create("complex")
        return self._rhok

    def _t(self, name, fn):
        if not self.args.breakdown:
            return fn()
        ctx = self.pm.ctx
        ctx.timer_start(1)
        r = fn()
        self.stage[name] = self.stage.get(name, 0.0) + ctx.timer_stop(1)
        return r

    def __call__(self, X, ntot, F):
        pm = self.pm
        # the positions of a real step are new: a tile-sorted copy cached from the previous evaluation (particle
        # arrays without spatial order, pmb_bin.cuh) must not be reused just because the bench repeats its input
        from pmesh_b200 import _lib
        _lib.check(pm.ctx.lib.pmb_bin_invalidate(pm.ctx.handle))
        for d in range(3):
            F[d] = None           # the previous evaluation's columns go back to the allocator before new ones are made
        layout = self._t("decompose", lambda: pm.decompose(X, smoothing=1.0 * pm.resampler.support))
        # P > 1: split or full exchange, whichever the first (warm-up) evaluations measured faster (domain.ExchangeTuner)
        tuner = pm.exchange_tuner
        if self.args.no_split or self.args.paint_mode != "atomic":
            tuner.choice = 'full'
        elif self.args.split:
            tuner.choice = 'split' if pm.comm.size > 1 else 'full'
        split = tuner.begin() == 'split'
        if split:
            # the particles a rank keeps are painted and read WHERE THEY LIE (the kernels clip to the local slab):
            # only the records that change rank are packed and travel (Layout.exchange_remote)
            lrem = self._t("exchange", lambda: layout.exchange_remote(X))
            lpos = X

            def paint2():
                pm.paint(X, out=self.rho, mode=self.args.paint_mode)
                if lrem.shape[0]:
                    pm.paint(lrem, out=self.rho, mode=self.args.paint_mode, hold=True)
            self._t("paint", paint2)
        else:
            lpos = self._t("exchange", lambda: layout.exchange(X))
            self._t("paint", lambda: pm.paint(lpos, out=self.rho, mode=self.args.paint_mode))
        self._t("scale", lambda: self.rho.scale(1.0 * pm.Nmesh.prod() / ntot))
        from pmesh_b200.pm import apply_gradients, c2r_fields, force_fields, gradient_fields, readout_fields
        whole = pm.comm.size == 1 and not self.args.unfused and not self.args.no_forward_fusion
        if whole:
            # one rank: the last pass of r2c, the gravity transfers and the first pass of the three c2r are ONE kernel of
            # this library (pmb_ifft.cuh, pmb_fft_force3); cuFFT transforms the planes; the density modes never reach HBM
            real = self._t("r2c+transfer+c2r", lambda: force_fields(self.rho, self.tf, outs=self.treal))
        else:
            self._t("r2c", lambda: self.rho.r2c(out=self.rhok))
        if whole:
            pass
        elif self.args.unfused:
            # the three gravity transfers read the density modes once (pm.apply_gradients) ...
            self._t("transfer", lambda: apply_gradients(self.rhok, self.tf, outs=self.tmp))
            # ... the three backward transforms overlap their NVLink transposes with each other's local FFTs
            real = self._t("c2r", lambda: c2r_fields(self.tmp, outs=[Ellipsis] * 3))
        else:
            # the gravity transfers are folded into the first (axis-0) pass of the three backward transforms, this
            # library's own kernel (pmb_ifft.cuh); cuFFT does the remaining 2-D c2r over the planes
            real = self._t("transfer+c2r", lambda: gradient_fields(self.rhok, self.tf, outs=self.treal))
        # the three force fields are read in ONE sweep over the particles (shared positions / weights)
        # ... and the ghost sum is fused into it: F[d] = layout.gather(real[d].readout(lpos)), nbody.py:214-216
        if split:
            Fn = self._t("readout+gather", lambda: readout_fields(real, X, remote=(layout, lrem)))
        else:
            Fn = self._t("readout+gather", lambda: readout_fields(real, lpos, gather=layout))
        tuner.end()
        for d in range(3):
            F[d] = Fn[d]
        return F


def _strided_sum(ctx, a):
    """float64 sum of a (strided, up to 3-D) device view (pmb_field_sum)"""
    import ctypes
    from pmesh_b200 import _lib
    out = ctypes.c_double(0.0)
    sz = (ctypes.c_int64 * 3)(*a.shape)
    st = (ctypes.c_int64 * 3)(*a.strides)
    _lib.check(ctx.lib.pmb_field_sum(ctx.handle, a.ptr, a.dtype.itemsize, len(a.shape), sz, st, ctypes.byref(out)))
    return out.value


def fused_block(pm, ms, launches, args, peak):
    """the transfer + axis-0 inverse transform kernel of this library (pmb_k_ifft_grad): CUDA-event time per launch and
    its algorithmic bytes -- the local complex cells are read once (16 B at f8) and written once per direction"""
    if not launches:
        return None
    lay = pm._layout
    ncell = int(numpy.prod(lay['o_shape']))
    csz = 2 * pm.dtype.itemsize
    per_step = launches / float(args.steps)
    ndir = 3.0 / per_step                      # directions served by one launch (3 on one rank, 1 on slabs)
    nbytes = ncell * csz * (1 + ndir)
    ms_launch = ms / launches
    gbs = nbytes / (ms_launch * 1e-3) / 1e9
    return {"kernel": "pmb_k_ifft_grad", "launches_per_step": per_step, "ms_per_launch": round(ms_launch, 3),
            "algorithmic_gb_per_launch": round(nbytes / 1e9, 3), "achieved_gbs": round(gbs, 1),
            "frac_of_hbm_peak": round(gbs / peak, 4),
            "replaces": ("the axis-0 pass of cuFFT's r2c (32 B per complex cell: the kernel takes every line forward first, "
                         "pm.force_fields) + " if pm.comm.size == 1 and not args.no_forward_fusion else "") +
                        "pmb_k_transfer_grad3 (64 B per complex cell) + the axis-0 pass of cuFFT in each of the three c2r (96 B)"}


def kernel_names(args, nl):
    """names of the paint / readout kernels the dispatch of pmb_resample.cu picks (for the roofline block)"""
    big = nl >= (1 << 18)
    if args.window == "cic" and big:
        return {"paint": "pmb_k_paint_cic_carry32", "readout": "pmb_k_readout_cic32_ring"}
    if args.window in ("nnb", "cic", "tsc", "pcs"):
        return {"paint": ("pmb_k_paint_carry32" if args.window in ("tsc", "pcs") else "pmb_k_paint_sched") if big else "pmb_k_paint_tuned",
                "readout": "pmb_k_readout_sched" if big else "pmb_k_readout_tuned"}
    return {"paint": "pmb_k_paint_dyn", "readout": "pmb_k_readout_dyn"}


def time_paint_readout(pm, args, comm, X, peak, R=5):
    """paint and readout alone on the local (exchanged) particles of input X: ms per launch from CUDA
    events on the library's stream, algorithmic bytes and the roofline fractions"""
    from pmesh_b200 import comm as C
    from pmesh_b200.device import DeviceArray
    ctx = pm.ctx
    es = pm.dtype.itemsize
    layout = pm.decompose(X, smoothing=1.0 * pm.resampler.support)
    lpos = layout.exchange(X)
    nl = lpos.shape[0]
    ncell_local = int(numpy.prod(pm._layout['i_shape']))
    rho = pm.create("real")
    import ctypes
    from pmesh_b200 import _lib

    def reorders():
        b = ctypes.c_int64(0)
        _lib.check(ctx.lib.pmb_bin_stats(ctx.handle, ctypes.byref(b), None))
        return b.value
    b0 = reorders()
    for _ in range(2):
        pm.resampler.paint(rho._device(), lpos, transform=pm.affine, mode=args.paint_mode)
    ctx.timer_start(2)
    for _ in range(R):
        pm.resampler.paint(rho._device(), lpos, transform=pm.affine, mode=args.paint_mode)
    t_paint = ctx.timer_stop(2) / R
    # particle arrays without spatial order are painted / read through a tile-sorted copy kept by the library
    # (pmb_bin.cuh): the calls above reuse it (content hash + kernel); a force evaluation pays the reorder ONCE,
    # in its paint.  t_first = a paint that has to build the copy.
    reordered = reorders() > b0
    t_first = t_paint
    if reordered:
        _lib.check(ctx.lib.pmb_bin_invalidate(ctx.handle))      # the copy's memory stays, its content is forgotten
        ctx.timer_start(2)
        pm.resampler.paint(rho._device(), lpos, transform=pm.affine, mode=args.paint_mode)
        t_first = ctx.timer_stop(2)
    out = DeviceArray.empty((nl,), "f8")
    for _ in range(2):
        pm.resampler.readout(rho._device(), lpos, out=out, transform=pm.affine)
    ctx.timer_start(2)
    for _ in range(R):
        pm.resampler.readout(rho._device(), lpos, out=out, transform=pm.affine)
    t_read = ctx.timer_stop(2) / R
    if reordered:
        _lib.check(ctx.lib.pmb_bin_release(ctx.handle))     # the copy's memory goes back before the next input
    ab_paint = nl * 24.0 + ncell_local * es          # pos (3 x f8) read + one mesh write pass
    ab_read = nl * (24.0 + 8.0) + ncell_local * es   # pos read + f8 result write + one mesh read pass
    nsum = comm.allreduce(nl, op=C.SUM)
    tp, tr = comm.allreduce(t_paint, op=C.MAX), comm.allreduce(t_read, op=C.MAX)
    tf = comm.allreduce(t_first, op=C.MAX)
    row = {"paint_ms": round(tp, 4), "readout_ms": round(tr, 4),
           "paint_frac": round(ab_paint / (t_paint * 1e-3) / 1e9 / peak, 4),
           "readout_frac": round(ab_read / (t_read * 1e-3) / 1e9 / peak, 4),
           "paint_readout_gparticles_per_s": round(nsum / ((tf + tr) * 1e-3) / 1e9, 3),
           "paint_readout_frac": round((ab_paint + ab_read) / ((t_first + t_read) * 1e-3) / 1e9 / peak, 4),
           "local_particles": int(nl)}
    if reordered:
        row["reordered"] = ("tile-sorted copy of the records: paint_ms / readout_ms reuse it (content hash + kernel "
                            "[+ return to the caller's order]); paint_first_ms builds it; paint_readout_* count paint_first_ms + readout_ms")
        row["paint_first_ms"] = round(tf, 4)
        row["reorder_ms"] = round(max(tf - tp, 0.0), 4)
    return row, (t_paint, ab_paint, t_read, ab_read, nl)


def replica_parity(pm, args, comm, step):
    """Parity of the WHOLE pipeline at the full benchmark size against the CPU oracle.

    The benchmark-size problem is built as rep^3 periodic replicas of a small problem (ns^3 uniform
    random particles on an ns^3 mesh, same cell size).  Density and force of the big problem are then
    the periodic repetition of the small one's, which the CPU oracle (oracle/: the reference's paint /
    readout arithmetic + numpy FFT) computes in seconds.  Coordinates are multiples of 2^-20 cells, so
    the stencil weights of a replica and of its original are bit-identical; what differs is the order
    of additions and the size of the FFT.  Returns relative errors (max over ranks):
    paint: max |rho - rho_oracle| / max |rho_oracle| on a sub-block of every rank's slab;
    force: max |F - F_oracle| / max |F_oracle| over ~1e5 sampled particles."""
    import ctypes
    from pmesh_b200 import _lib, comm as C
    from pmesh_b200.device import DeviceArray
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    oracle.build()
    M = args.nmesh
    ns = 128 if M % 128 == 0 and M >= 256 else (M // 2 if M % 2 == 0 else 0)
    if ns < 8 or pm.BoxSize[0] != float(M):
        return {"skipped": "mesh %d has no periodic sub-division" % M}
    rep = M // ns
    nblocks = rep ** 3
    b_first = comm.rank * nblocks // comm.size
    b_n = (comm.rank + 1) * nblocks // comm.size - b_first
    nsm = ns ** 3
    rng = numpy.random.default_rng(46)
    xs = rng.integers(0, ns << 20, size=(nsm, 3)).astype("f8") / float(1 << 20)      # cell units = box units
    ctx = pm.ctx
    dxs = DeviceArray.from_host(xs)
    X = DeviceArray.empty((b_n * nsm, 3), "f8")
    nrep = (ctypes.c_int64 * 3)(rep, rep, rep)
    per = (ctypes.c_double * 3)(float(ns), float(ns), float(ns))
    _lib.check(ctx.lib.pmb_particles_replicate(ctx.handle, X.ptr, 8, dxs.ptr, nsm, 3, nrep, per, b_first, b_n))
    del dxs
    n = X.shape[0]
    ntot = nblocks * nsm
    F = [None] * 3
    step(X, ntot, F)
    # ---- oracle of the small problem ----
    t0 = time.perf_counter()
    rho_s = numpy.zeros((ns, ns, ns))
    oracle.paint(rho_s, xs, args.window, scale=1.0, period=[ns] * 3)
    ck = oracle.r2c(rho_s * (float(ns) ** 3 / nsm))
    # ---- force of sampled particles ----
    nsamp = max(1, min(n, 100000 // comm.size))
    pick = numpy.sort(numpy.random.default_rng(47 + comm.rank).choice(n, size=nsamp, replace=False)).astype("i4") if n else numpy.zeros(0, "i4")
    dpick = DeviceArray.from_host(pick)
    ferr, fmax = 0.0, 0.0
    for d in range(3):
        fr = oracle.c2r(oracle.transfer(ck, [ns] * 3, [float(ns)] * 3, "gravity_fd4", d), [ns] * 3)
        ref = oracle.readout(fr, xs[pick % nsm], args.window, scale=1.0, period=[ns] * 3)
        got = DeviceArray.empty((len(pick),), "f8")
        if len(pick):
            _lib.check(ctx.lib.pmb_take(ctx.handle, F[d].ptr, 8, dpick.ptr, len(pick), got.ptr))
            got = got.to_host()
            ferr = max(ferr, float(abs(got - ref).max()))
            fmax = max(fmax, float(abs(ref).max()))
    ferr = comm.allreduce(ferr, op=C.MAX) / max(comm.allreduce(fmax, op=C.MAX), 1e-300)
    # ---- density on a sub-block of this rank's slab ----
    layout = pm.decompose(X, smoothing=1.0 * pm.resampler.support)
    lpos = layout.exchange(X)
    rho = pm.paint(lpos, mode=args.paint_mode)
    mesh = rho._device()
    i0 = int(pm._layout['i_start'][0])
    nb = min(32, int(mesh.shape[0]))
    perr = 0.0
    if nb > 0:
        sub = DeviceArray((nb, ns, ns), mesh.dtype, ptr=mesh.ptr, strides=mesh.strides, base=mesh, ctx=ctx).to_host()
        want = rho_s[(numpy.arange(nb) + i0) % ns]
        perr = float(abs(sub - want).max() / abs(rho_s).max())
    perr = comm.allreduce(perr, op=C.MAX)
    return {"paint_rel_err": perr, "force_rel_err": ferr, "replicas": "%d^3 of a %d^3 problem" % (rep, ns),
            "sampled_particles": int(comm.allreduce(len(pick), op=C.SUM)),
            "oracle_seconds": round(time.perf_counter() - t0, 2)}


def run_ours(args):
    from pmesh_b200 import _lib, comm as C
    from pmesh_b200.device import DeviceArray, PinnedArray
    from pmesh_b200.pm import ParticleMesh

    comm = C.world()
    if comm.size != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE is %d (launch with torch.distributed.run)" % (args.gpus, comm.size))
    M = args.nmesh
    npm = [int(v) for v in args.np.split(",")] if args.np else None
    pm = ParticleMesh(BoxSize=float(M), Nmesh=[M, M, M], dtype=args.dtype, resampler=args.window, comm=comm, np=npm)
    ctx = pm.ctx
    X, ntot = make_particles(pm, args, comm)
    n = X.shape[0]
    F = [None] * 3                # force components (SoA): every step returns new columns
    step = ForceStep(pm, args)

    for _ in range(args.warmup):
        step(X, ntot, F)
    if args.breakdown:
        step.stage = {}
    comm.Barrier()
    ctx.sync()
    clocks = ClockSampler(ctx.device)
    ctx.launch_count(reset=True)
    pm.fft_library_ms(reset=True)
    pm.fft_transpose_stats(reset=True)
    pm.fft_fused_stats(reset=True)
    ctx.timer_start(0)
    for _ in range(args.steps):
        step(X, ntot, F)
    ms = ctx.timer_stop(0)
    comm.Barrier()
    stage_ms = dict((k, round(v / args.steps, 3)) for k, v in step.stage.items())
    launches = ctx.launch_count()
    fft_ms = pm.fft_library_ms()
    xp_ms, xp_bytes = pm.fft_transpose_stats()
    fu_ms, fu_n = pm.fft_fused_stats()
    clk = clocks.stop()
    ms_step = comm.allreduce(ms / args.steps, op=C.MAX)
    fft_step = comm.allreduce(fft_ms / args.steps, op=C.MAX)

    # ---- size-independent properties of the timed input at full size (outside every timed region) ----
    # mass conservation of the scatter: sum(rho) == number of particles; momentum conservation of
    # the whole force step (same window for paint and readout, antisymmetric transfer):
    # |sum_p F_d(p)| << N * rms(F)
    es = pm.dtype.itemsize
    fsum = [comm.allreduce(F[d].sum(), op=C.SUM) for d in range(3)]
    fsq = comm.allreduce(sum(F[d].dot(F[d]) for d in range(3)), op=C.SUM)
    frms = (fsq / (3.0 * ntot)) ** 0.5
    verify = {"net_force_over_n_rms_force": max(abs(f) for f in fsum) / (ntot * max(frms, 1e-300))}

    # ---- dominant kernels alone: paint and readout on the local particles, per input (roofline) ----
    nb, fb = X.nbytes, F[0].nbytes
    F = None                      # the force columns are not needed any more; the input rows need the room
    _lib.check(ctx.lib.pmb_bin_release(ctx.handle))
    ctx.empty_cache()
    peak, peak_src = peaks()
    inputs = {}
    dom_stats = None
    todo = [args.particles] + [k for k in args.inputs.split(",") if k and k != args.particles]
    for kind in todo:
        Xk = X if kind == args.particles else make_particles(pm, args, comm, kind)[0]
        row, stats = time_paint_readout(pm, args, comm, Xk, peak)
        inputs[INPUT_LABEL[kind]] = row
        if kind == args.particles:
            dom_stats = stats
        if Xk is not X:
            del Xk
    t_paint, ab_paint, t_read, ab_read, nl = dom_stats
    main = inputs[INPUT_LABEL[args.particles]]
    dom = ("paint", t_paint, ab_paint) if t_paint >= t_read else ("readout", t_read, ab_read)
    achieved = dom[2] / (dom[1] * 1e-3) / 1e9
    kname = kernel_names(args, nl)[dom[0]]
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get("%s:%s:%d:%d:%s" % (kname, args.window, M, comm.size, args.particles))
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kname, "input": INPUT_LABEL[args.particles], "achieved": round(achieved, 1),
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "ms_per_launch": round(dom[1], 4),
                "algorithmic_bytes_per_launch": dom[2],
                "paint_ms": main["paint_ms"], "readout_ms": main["readout_ms"],
                "paint_frac": main["paint_frac"], "readout_frac": main["readout_frac"]}

    # mass conservation on the timed input
    layout = pm.decompose(X, smoothing=1.0 * pm.resampler.support)
    lpos = layout.exchange(X)
    rho = pm.paint(lpos, mode=args.paint_mode)
    mass = comm.allreduce(_strided_sum(ctx, rho._device()), op=C.SUM)
    verify["mass_conservation_rel_err"] = abs(mass - ntot) / ntot
    del lpos, layout, rho

    # ---- end to end through the public API with host buffers (pinned), H2D + D2H inside the timing ----
    # Every step uploads its positions from pinned host memory, runs the force step and downloads the three
    # force columns.  `serial_ms`: one step with upload -> compute -> download strictly one after the other.
    # `value`: a stream of independent force evaluations (an ensemble / parameter sweep): the upload of step
    # k + 1 runs on the copy stream while step k's forces are downloaded on the other (PCIe is full duplex) --
    # every byte of every step still crosses PCIe inside the timed region; single device buffers for X and F
    # (compute k + 1 waits for upload k + 1 and for download k).
    e2e = None
    if not args.no_e2e:
        Xh = PinnedArray((n, 3), "f8")
        Fh = PinnedArray((3, n), "f8")
        ctx.d2h(Xh.array, X.ptr, X.nbytes)
        _lib.check(ctx.lib.pmb_bin_release(ctx.handle))
        ke = max(1, min(args.steps, 8))

        def stream_single():
            """upload k + 1 || download k, compute between them; single device buffers"""
            Xd = X
            ctx.h2d_async(Xd.ptr, Xh.array, nb)
            ctx.stream_record(1, 0)                   # event 0: upload of the coming step is complete
            Fk = None
            for it in range(ke):
                ctx.stream_wait(0, 0)                 # compute waits for its positions ...
                if it > 0:
                    ctx.stream_wait(0, 2)             # ... and for the previous download: its F buffers are recycled now
                Fk = None
                Fk = step(Xd, ntot, [None] * 3)
                ctx.stream_record(0, 1)               # event 1: forces of this step are complete
                ctx.stream_wait(2, 1)
                for d in range(3):
                    ctx.d2h_async(Fh.array[d], Fk[d].ptr, fb)
                ctx.stream_record(2, 2)               # event 2: download complete
                if it + 1 < ke:
                    ctx.stream_wait(1, 1)             # the next upload overwrites Xd: after the compute that reads it
                    ctx.h2d_async(Xd.ptr, Xh.array, nb)
                    ctx.stream_record(1, 0)
            ctx.stream_sync(2)
            ctx.stream_sync(1)
            ctx.sync()
            return Fk

        def stream_double(X2):
            """upload k + 1 || compute k || download k - 1: two position buffers, the force columns of a step stay
            alive until their download has completed (events: 0 / 1 upload into buffer b complete, 2 / 3 compute of
            a step on buffer b complete, 4 / 5 its download complete)"""
            Xd = [X, X2]
            ctx.h2d_async(Xd[0].ptr, Xh.array, nb)
            ctx.stream_record(1, 0)
            hold = {}
            Fk = None
            for it in range(ke):
                b = it & 1
                if it + 1 < ke:
                    if it >= 1:
                        ctx.stream_wait(1, 2 + (1 - b))   # buffer 1 - b was read by the compute of step it - 1
                    ctx.h2d_async(Xd[1 - b].ptr, Xh.array, nb)
                    ctx.stream_record(1, 1 - b)
                ctx.stream_wait(0, b)
                if it >= 2:
                    ctx.stream_wait(0, 4 + b)             # the columns of step it - 2 go back to the pool only now
                    hold.pop(it - 2)
                Fk = step(Xd[b], ntot, [None] * 3)
                ctx.stream_record(0, 2 + b)
                ctx.stream_wait(2, 2 + b)
                for d in range(3):
                    ctx.d2h_async(Fh.array[d], Fk[d].ptr, fb)
                ctx.stream_record(2, 4 + b)
                hold[it] = Fk
            ctx.stream_sync(2)
            ctx.stream_sync(1)
            ctx.sync()
            hold.clear()
            return Fk

        overlap = "upload of step k+1 || download of step k (independent evaluations); compute between them"
        Fk = None
        X2 = None
        # measured at 1 GPU, 1024^3 (ms per step): single buffers 727 - 753; three-way overlap 773 -- the copies and the
        # HBM-bound kernels slow each other down; --e2e-double selects it
        try:
            free_b = ctx.mem_info()[0] + ctx._pooled
            # two more sets of force columns and one more position buffer have to fit beside the step's own fields
            if args.e2e_double and comm.allreduce(1 if free_b > 2.6 * nb + (8 << 30) else 0, op=C.MIN):
                X2 = DeviceArray.empty(X.shape, "f8")
        except Exception:
            X2 = None
        comm.Barrier()
        ctx.sync()
        t0 = time.perf_counter()
        if X2 is not None:
            try:
                Fk = stream_double(X2)
                overlap = "upload of step k+1 || compute of step k || download of step k-1 (independent evaluations, two position buffers)"
            except Exception as e:      # out of device memory on a single large rank: fall back, time again
                sys.stderr.write("e2e: double-buffered stream failed (%s); single buffers\n" % e)
                Fk = None
                X2 = None
                ctx.stream_sync(2)
                ctx.stream_sync(1)
                ctx.sync()
                ctx.empty_cache()
                t0 = time.perf_counter()
        if X2 is None:
            Fk = stream_single()
        comm.Barrier()
        e2e_ms = comm.allreduce((time.perf_counter() - t0) * 1e3 / ke, op=C.MAX)
        del X2
        Xd = X
        # one strictly serial step for comparison
        comm.Barrier()
        t0 = time.perf_counter()
        ctx.h2d(Xd.ptr, Xh.array, nb)
        Fk = None
        Fk = step(Xd, ntot, [None] * 3)
        for d in range(3):
            ctx.d2h(Fh.array[d], Fk[d].ptr, fb)
        ctx.sync()
        comm.Barrier()
        serial_ms = comm.allreduce((time.perf_counter() - t0) * 1e3, op=C.MAX)
        e2e = {"value": round(e2e_ms, 3), "unit": "ms", "h2d_bytes_per_step": int(nb),
               "d2h_bytes_per_step": int(3 * fb), "steps": ke, "serial_ms": round(serial_ms, 3),
               "overlap": overlap}
        del Xh, Fh, Xd, Fk

    # ---- parity at full size against the oracle (periodic replicas of a small problem) ----
    del X
    if not args.no_verify:
        par = replica_parity(pm, args, comm, step)
        verify["parity"] = par
        if "force_rel_err" in par:
            verify["parity_rel_err"] = max(par["force_rel_err"], par["paint_rel_err"])

    cpu = None
    if comm.rank == 0 and comm.size == 1 and not args.no_cpu:
        cpu = cpu_force_step(args, n=min(args.cpu_sample, args.nmesh), steps=2)

    if comm.rank == 0:
        line = {
            "metric": metric_name(args), "value": round(ms_step, 3), "unit": "ms", "n_gpus": comm.size,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3),
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if args.dtype == "f8" else "f32", "data": "synthetic",
            "config": {"workload": "%s PM force step, %d^3 %s particles on a %d^3 mesh (BASELINE configs[2] shape)%s"
                                   % (WINDOW_NAMES.get(args.window, args.window), M, INPUT_LABEL[args.particles], M,
                                      "" if comm.size > 1 else ", single GPU"),
                       "nmesh": M, "nparticles": ntot, "window": args.window, "paint_mode": args.paint_mode,
                       "particles": INPUT_LABEL[args.particles],
                       "decomposition": ("pencil np=%s" % pm.np) if len(pm.np) == 2 else "slab np=[%d]" % comm.size,
                       "l2": "inputs (%.1f GB positions + %.1f GB mesh per rank) are larger than L2"
                             % (n * 24 / 1e9, int(numpy.prod(pm._layout['i_shape'])) * es / 1e9)},
            "paint_readout_gparticles_per_s": main["paint_readout_gparticles_per_s"],
            "particles_per_s_force_step": round(ntot / (ms_step * 1e-3), 1),
            "cufft_library_ms_per_step": round(fft_step, 3),
            "fft_transpose": None if comm.size == 1 else {
                "ms_per_step": round(xp_ms / args.steps, 3), "nvlink_gb_per_step": round(xp_bytes / args.steps / 1e9, 3),
                "nvlink_gbs_achieved": round(xp_bytes / max(xp_ms, 1e-9) / 1e6, 1), "nvlink_gbs_peak_measured": 770.0,
                "note": "rank 0: bytes the fused transpose kernels stored into peer memory / their CUDA-event time"},
            "fused_transfer_ifft": fused_block(pm, fu_ms, fu_n, args, peak),
            "exchange": None if comm.size == 1 else {
                "mode": pm.exchange_tuner.choice, "measured_ms": getattr(pm.exchange_tuner, "measured", None),
                "note": "split: the particles a rank keeps are painted / read where they lie, only records that change rank "
                        "travel (Layout.exchange_remote); full: every record through take + alltoallv (Layout.exchange); "
                        "chosen from one timed warm-up evaluation of each, slowest rank decides (domain.ExchangeTuner)"},
            "gpu_launches": int(launches),
            "clocks": clk, "roofline": roofline, "inputs": inputs, "e2e": e2e, "cpu_baseline": cpu,
            "verify": verify,
        }
        if args.breakdown:
            line["stage_ms_per_step"] = stage_ms
        print(json.dumps(line))
        sys.stdout.flush()


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_force_step(args, n, steps, warmup=0, cores=None):
    """The reference's CPU implementation of the same force step on the host cores (oracle/cpu_arm.py:
    the reference's own compiled C paint / readout from oracle/_ref -- the oracle port when it is absent --,
    its routing arithmetic, numpy transfer, scipy.fft standing in for PFFT; one slab of the mesh per worker
    process as the reference's MPI ranks would hold it, communication free).  A full n^3 / n^3 step is
    MEASURED on all host cores; the value reported for the bench workload scales it by particle count."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_arm
    r = cpu_arm.force_step(n=n, window=args.window, cores=cores, steps=steps, warmup=warmup)
    factor = (args.nmesh / float(n)) ** 3
    return {"value": round(r["seconds_per_step"] * 1e3 * factor, 1), "unit": "ms", "cores": r["cores"], "kind": r["kind"],
            "sample": "a full %s force step of %d^3 lattice_sine particles on a %d^3 mesh measured on %d host cores "
                      "(one slab per worker process, %s as the PFFT stand-in, no MPI cost): %.2f s per step, mean of %d; "
                      "scaled x%.0f by particle count to %d^3"
                      % (args.window.upper(), n, n, r["cores"], r["fft"], r["seconds_per_step"], r["steps"], factor, args.nmesh),
            "sample_seconds": round(r["seconds_per_step"], 4), "sample_nmesh": n, "scale_factor": factor}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = min(args.cpu_sample_reference, args.nmesh)
    # each timed step is one full force step of the sample problem; the count is bounded so that the arm
    # ends within a few minutes whatever --steps / --warmup the driver passes
    steps = max(1, min(args.steps, 3))
    r = cpu_force_step(args, n=n, steps=steps, warmup=min(args.warmup, 1))
    M = args.nmesh
    line = {
        "impl": "reference", "metric": metric_name(args), "value": r["value"], "unit": "ms", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "steps_timed": steps, "ms_per_step": r["value"], "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s PM force step, %d^3 particles on a %d^3 mesh (BASELINE configs[2] shape)"
                               % (WINDOW_NAMES.get(args.window, args.window), M, M),
                   "nmesh": M, "nparticles": M ** 3, "window": args.window,
                   "measured": "%d^3 sample, scaled x%.0f" % (n, r["scale_factor"])},
        "cpu_baseline": r,
        "e2e": {"value": r["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
