"""Worker of tests/test_gpu_multirank.py: run under torch.distributed.run with one process per GPU.

Checks the NCCL legs (particle alltoallv, ghost reduction, slab FFT transposes) against the oracle's
all-ranks simulation: every rank regenerates all ranks' inputs from seeds, computes what it should
receive / hold with the oracle, and compares with what the library delivered.
"""
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import oracle  # noqa: E402
from pmesh_b200 import comm as C, transfer as T  # noqa: E402
from pmesh_b200.device import DeviceArray  # noqa: E402
from pmesh_b200.pm import ParticleMesh  # noqa: E402


def rel(a, b):
    a, b = numpy.asarray(a), numpy.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:          # a rank may own no mesh planes (e.g. 20 planes over 8 ranks)
        return 0.0
    return abs(a - b).max() / max(abs(b).max(), 1e-300)


def main():
    comm = C.world()
    P, r = comm.size, comm.rank
    assert P > 1
    for n, res, dtype in ((16, "cic", "f8"), (24, "tsc", "f8"), (16, "cic", "f4"), (20, "pcs", "f8")):
        L = 100.0
        pm = ParticleMesh(BoxSize=L, Nmesh=[n, n, n], dtype=dtype, resampler=res, comm=comm)
        tol = 1e-6 if dtype == "f8" else 2e-4
        start, shape = pm._layout["i_start"], pm._layout["i_shape"]
        edges = pm.domain.edges
        # every rank's particles, reproducible everywhere
        allpos = [numpy.random.default_rng(100 + q).uniform(-0.2 * L, 1.2 * L, (3000 + 100 * q, 3)) for q in range(P)]
        allmass = [numpy.random.default_rng(200 + q).uniform(0.5, 2.0, len(allpos[q])) for q in range(P)]
        sm = 0.5 * pm.resampler.support
        lays = [oracle.decompose(allpos[q], edges, P, smoothing=sm, assign=pm.domain.DomainAssign, scale=n / L)
                for q in range(P)]

        # 1. routing + exchange over NCCL, bit exact
        layout = pm.decompose(allpos[r])
        assert numpy.array_equal(layout.sendcounts, lays[r][0]) and numpy.array_equal(layout.indices, lays[r][1])
        want_pos = oracle.exchange_all(allpos, lays)[r]
        want_mass = oracle.exchange_all(allmass, lays)[r]
        lpos, lmass = layout.exchange(allpos[r], allmass[r])
        assert numpy.array_equal(lpos, want_pos) and numpy.array_equal(lmass, want_mass)
        dl = layout.exchange(DeviceArray.from_host(allpos[r]))
        assert numpy.array_equal(dl.to_host(), want_pos)

        # 2. decomposed deterministic paint == oracle paint of the received particles into the slab
        rho = pm.paint(allpos[r], mass=allmass[r], layout=layout, mode="deterministic")
        slab = numpy.zeros(tuple(shape), dtype)
        oracle.paint(slab, want_pos, res, mass=want_mass, scale=n / L, translate=-start, period=[n] * 3)
        assert numpy.array_equal(rho.value, slab), "decomposed paint differs"
        full = numpy.concatenate(comm.allgather(slab), axis=0)
        serial = numpy.zeros((n, n, n), dtype)
        oracle.paint(serial, numpy.concatenate(allpos), res, mass=numpy.concatenate(allmass), scale=n / L, period=[n] * 3)
        assert rel(full, serial) < (1e-12 if dtype == "f8" else 1e-5)          # not bit-equal by design

        # 3. distributed r2c / c2r against numpy on the gathered field
        ck_full = oracle.r2c(full.astype("f8"))
        rhok = rho.r2c()
        s1, m1 = rhok.start[1], rhok.shape[1]
        assert rhok.shape == (n, m1, n // 2 + 1)
        assert rel(rhok.value, ck_full[:, s1:s1 + m1, :]) < tol
        back = rhok.c2r()
        assert rel(back.value, slab) < tol
        assert rel(rhok.value, ck_full[:, s1:s1 + m1, :]) < tol          # input preserved
        rip = pm.create("real", value=slab)
        cip = rip.r2c(out=Ellipsis)
        assert rel(cip.value, ck_full[:, s1:s1 + m1, :]) < tol
        assert rel(cip.c2r(out=Ellipsis).value, slab) < tol

        # 4. transfer + c2r + readout with ghost reduction == serial oracle pipeline
        for d in range(3):
            fr_full = oracle.c2r(oracle.transfer(ck_full, [n] * 3, [L] * 3, "gravity_fd4", d), [n] * 3)
            f = rhok.apply(T.GravityFD4(d)).c2r()
            assert rel(f.value, fr_full[start[0]:start[0] + shape[0]]) < tol
            got = f.readout(allpos[r], layout=layout)
            want = oracle.readout(fr_full.astype(dtype), allpos[r], res, scale=n / L, period=[n] * 3)
            assert rel(got, want) < tol, rel(got, want)
            gd = f.readout(DeviceArray.from_host(allpos[r]), layout=layout)
            assert numpy.array_equal(gd.to_host(), got)

        # 4b. white noise does not depend on the partition: my block of the transposed complex field
        #     equals the same block of the oracle's whole field
        cdt = "complex128" if dtype == "f8" else "complex64"
        wn_full = oracle.whitenoise(numpy.zeros((n, n, n // 2 + 1), dtype=cdt), 0, (n, n, n), 120577 + n, False)
        wn = pm.generate_whitenoise(120577 + n)
        s1, m1 = wn.start[1], wn.shape[1]
        assert wn.shape == (n, m1, n // 2 + 1)
        if wn.size:
            assert abs(wn.value - wn_full[:, s1:s1 + m1, :]).max() < (1e-13 if dtype == "f8" else 5e-7)

        # 5. gather modes on real ghosts (tests/test_domain.py:229-266 semantics)
        # vector weights: 'sum' of a 3-column array on the device == bincount over the returned ghosts
        # what I hold after an exchange of per-ghost values is arbitrary data of length recvlength:
        mine = numpy.random.default_rng(400 + r).uniform(-1, 1, (int(layout.recvlength), 3))
        held = comm.allgather(mine)
        # the ghosts of MY particles that rank q holds sit at its recv segment for source r
        back = []
        for q in range(P):
            rc_q = numpy.array([lays[src][0][q] for src in range(P)])
            off = int(rc_q[:r].sum())
            back.append(held[q][off:off + int(rc_q[r])])
        want3 = oracle.bincount_sum(lays[r][1], numpy.concatenate(back), len(allpos[r]))
        got3 = layout.gather(mine, mode="sum")
        assert numpy.array_equal(got3, want3), "device ghost sum of vector weights differs from bincount"
        assert numpy.array_equal(layout.gather(DeviceArray.from_host(mine), mode="sum").to_host(), want3)
        ones = numpy.ones(layout.recvlength)
        nghost = layout.gather(ones, mode="sum")
        assert nghost.min() >= 1 and abs(nghost.sum() - lays[r][0].sum()) < 1e-9
        assert numpy.array_equal(layout.gather(lpos, mode="any"), allpos[r])
        assert numpy.allclose(layout.gather(lpos, mode="mean"), allpos[r])
        # 6. the driver's force (pmesh_b200.nbody.force <- examples/nbody.py:196-218) on my particles ==
        #    the serial numpy driver on everybody's particles
        if dtype == "f8":
            from pmesh_b200 import nbody
            Fmine = nbody.force(pm, allpos[r], factor=0.45)
            Fall = oracle.nbody_force(numpy.concatenate(allpos), 0.0, n, L, res, 0.45)
            off = sum(len(allpos[q]) for q in range(r))
            got = numpy.stack([f.to_host() for f in Fmine], axis=1)
            assert rel(got, Fall[off:off + len(allpos[r])]) < 1e-6

        # 6b. I/O order: ravel / unravel of the transposed complex field and the Fourier resample go through the
        #     distributed index exchange (the reference's mpsort, pm.py:389-448, 479-547)
        if dtype == "f8":
            import functools
            from pmesh_b200.resample import mode_table
            chunk = numpy.empty(int(rhok.size), dtype=rhok.dtype)
            rhok.ravel(out=chunk)
            whole = numpy.concatenate(comm.allgather(chunk))
            assert numpy.array_equal(whole, ck_full.astype(rhok.dtype).ravel()) or rel(whole, ck_full.ravel()) < 1e-12
            other = pm.create("complex")
            other.unravel(chunk)
            assert numpy.array_equal(other.value, rhok.value)
            n2 = n // 2
            pm2 = ParticleMesh(BoxSize=L, Nmesh=[n2, n2, n2], dtype=dtype, resampler=res, comm=comm)
            small = pm2.create("complex")
            rhok.resample(small)
            tabs = [mode_table(n, n2) for _ in range(3)]
            tabs[2] = tabs[2][:n2 // 2 + 1]
            ok = [(t >= 0) & (t < m) for t, m in zip(tabs, ck_full.shape)]
            want_s = ck_full[numpy.ix_(*[numpy.where(o, t, 0) for o, t in zip(ok, tabs)])]
            want_s = numpy.where(ok[0][:, None, None] & ok[1][None, :, None] & ok[2][None, None, :], want_s, 0)
            ii = numpy.meshgrid(*[numpy.arange(m) for m in want_s.shape], indexing="ij", sparse=True)
            selfconj = functools.reduce(numpy.bitwise_and, [(n2 - i_) % n2 == i_ for i_ in ii])
            want_s.imag[numpy.broadcast_to(selfconj, want_s.shape)] = 0
            for mm in (n2, n):
                nyq = functools.reduce(numpy.bitwise_or, [i_ == mm // 2 for i_ in ii])
                want_s[numpy.broadcast_to(nyq, want_s.shape)] = 0
            sl2 = tuple(slice(int(a), int(a + b)) for a, b in zip(small.start, small.shape))
            assert rel(small.value, want_s[sl2]) < 1e-12, "distributed Fourier resample"
            del pm2, small, other

        comm.Barrier()
        if r == 0:
            print("multirank ok: n=%d %s %s on %d ranks" % (n, res, dtype, P))
    # 6c. the gather fused with the ghost sum (pm.readout_fields(..., gather=layout): the kernel writes the own
    #     results straight into the gathered columns, only ghosts travel) == Layout.gather of separate readouts;
    #     needs >= 2^18 local particles to take the fused kernel
    from pmesh_b200.pm import apply_gradients, readout_fields
    n, L = 48, 100.0
    pm = ParticleMesh(BoxSize=L, Nmesh=[n, n, n], dtype="f8", resampler="cic", comm=comm)
    mine = numpy.random.default_rng(900 + r).uniform(0, L, (300000 + 1000 * r, 3))
    dmine = DeviceArray.from_host(mine)
    layout = pm.decompose(dmine, smoothing=1.0 * pm.resampler.support)
    lpos = layout.exchange(dmine)
    assert lpos.shape[0] >= (1 << 18)
    rhok = pm.paint(lpos).r2c()
    from pmesh_b200.pm import c2r_fields
    fields = c2r_fields(apply_gradients(rhok, [T.GravityFD4(d) for d in range(3)]), outs=[Ellipsis] * 3)
    # a second batch right behind the first (landing buffers alternate across calls), out of place, 2 fields
    again = c2r_fields(apply_gradients(rhok, [T.GravityFD4(d) for d in range(3)])[:2])
    for d in range(2):
        assert numpy.array_equal(again[d].value, fields[d].value), "overlapped c2r is not reproducible"
    del again
    fused = readout_fields(fields, lpos, gather=layout)
    for d in range(3):
        sep = layout.gather(fields[d].readout(lpos))
        assert rel(fused[d].to_host(), sep.to_host()) < 1e-14, "fused ghost sum differs"
        one = rhok.apply(T.GravityFD4(d)).c2r()
        assert numpy.array_equal(one.value, fields[d].value)          # three transfers in one pass == one at a time
    two = readout_fields(fields[:2], lpos, gather=layout)
    assert numpy.array_equal(two[1].to_host(), fused[1].to_host())
    comm.Barrier()
    if r == 0:
        print("fused gather ok on %d ranks" % P)
    # 6c'. the rank's own particles painted / read where they lie, only the records that change rank travel
    #      (Layout.exchange_remote) == the full exchange
    lrem = layout.exchange_remote(dmine)
    me_cnt = int(layout.recvcounts[r])
    assert lrem.shape[0] == lpos.shape[0] - me_cnt
    full = lpos.to_host()
    off = int(layout.recvoffsets[r])
    assert numpy.array_equal(lrem.to_host(), numpy.concatenate([full[:off], full[off + me_cnt:]])), "exchange_remote records"
    rho_a = pm.paint(lpos)
    rho_b = pm.paint(dmine)
    if lrem.shape[0]:
        pm.paint(lrem, out=rho_b, hold=True)
    sc = comm.allreduce(float(abs(rho_a.value).max()), op=C.MAX)
    assert float(abs(rho_a.value - rho_b.value).max()) <= 1e-12 * sc, "paint in place + remote ghosts"
    split = readout_fields(fields, dmine, remote=(layout, lrem))
    for d in range(3):
        assert rel(split[d].to_host(), fused[d].to_host()) < 1e-13, "readout in place + ghost sum"
    comm.Barrier()
    if r == 0:
        print("split exchange ok on %d ranks" % P)
    # 6d. transfers folded into the axis-0 pass of the backward transforms (pm.gradient_fields, pmb_ifft.cuh) on slabs:
    #     == transfer pass + cuFFT lines, for a fused length (64) and an unfused one (48, above), twice in a row
    from pmesh_b200.pm import gradient_fields
    for dt, tol, shape in (("f8", 1e-12, [64, 40, 24]), ("f4", 2e-5, [64, 40, 24]), ("f8", 1e-12, [128, 16, 12]),
                           ("f8", 1e-12, [256, 16, 12]), ("f8", 1e-12, [512, 16, 12]), ("f8", 1e-12, [1024, 16, 12]),
                           ("f8", 1e-12, [2048, 16, 12]), ("f8", 1e-12, [4096, 16, 12]), ("f4", 2e-5, [1024, 16, 12])):
        pm64 = ParticleMesh(BoxSize=[L, 0.9 * L, 1.1 * L], Nmesh=shape, dtype=dt, comm=comm)
        rk = pm64.generate_whitenoise(seed=77, type="complex")
        rk.scale(0.5)
        for make in (T.GravityFD4, T.GradientK):
            tf3 = [make(d) for d in range(3)]
            ref = c2r_fields(apply_gradients(rk, tf3), outs=[Ellipsis] * 3)
            for rep in range(2):
                got = gradient_fields(rk, tf3)
                for d in range(3):
                    sc = comm.allreduce(float(abs(ref[d].value).max()) if ref[d].value.size else 0.0, op=C.MAX)
                    err = float(abs(got[d].value - ref[d].value).max()) if ref[d].value.size else 0.0
                    assert err <= tol * sc, ("fused transfer + first pass", dt, d, err, sc)
        del pm64, rk, ref, got
    got = gradient_fields(rhok, [T.GravityFD4(d) for d in range(3)])      # n = 48: the unfused calls
    for d in range(3):
        assert numpy.array_equal(got[d].value, fields[d].value)
    comm.Barrier()
    if r == 0:
        print("fused transfer + first pass ok on %d ranks" % P)
    del pm, layout, lpos, rhok, fields, fused, two

    # 7. pencil (2-D) process meshes, the reference's default for 3-D fields (pm.py:1319-1327): real space
    #    split along axes (0, 1), complex space along (1, 2); everything against the serial oracle
    meshes = {2: [[1, 2]], 4: [[2, 2], [1, 4]], 8: [[2, 4], [4, 2]]}.get(P, [])
    for np_ in meshes:
        for n3, res, dtype in (((16, 20, 24), "cic", "f8"), ((12, 12, 12), "tsc", "f4")):
            L = 100.0
            tol = 1e-6 if dtype == "f8" else 2e-4
            pm = ParticleMesh(BoxSize=L, Nmesh=list(n3), dtype=dtype, resampler=res, comm=comm, np=np_)
            start, shape = pm._layout["i_start"], pm._layout["i_shape"]
            c0, c1 = r // np_[1], r % np_[1]
            for d, (parts, c) in enumerate(((np_[0], c0), (np_[1], c1))):
                blk = -(-n3[d] // parts)
                assert start[d] == min(c * blk, n3[d]) and shape[d] == min((c + 1) * blk, n3[d]) - start[d]
            sl = tuple(slice(int(a), int(a + b)) for a, b in zip(start, shape))
            full = numpy.random.default_rng(31).uniform(-1, 1, n3).astype(dtype)
            ck_full = oracle.r2c(full.astype("f8"))
            rho = pm.create("real", value=full[sl])
            rhok = rho.r2c()
            osl = tuple(slice(int(a), int(a + b)) for a, b in zip(rhok.start, rhok.shape))
            assert rhok.shape[0] == n3[0] and rhok.start[0] == 0
            assert rel(rhok.value, ck_full[osl]) < tol, ("pencil r2c", np_, rel(rhok.value, ck_full[osl]))
            assert rel(rhok.c2r().value, full[sl]) < tol
            assert rel(rhok.value, ck_full[osl]) < tol                     # input preserved
            cip = pm.create("real", value=full[sl]).r2c(out=Ellipsis)
            assert rel(cip.value, ck_full[osl]) < tol
            assert rel(cip.c2r(out=Ellipsis).value, full[sl]) < tol
            # cnorm / cdot over the independent modes: device reduction + allreduce == numpy on the whole field
            w = numpy.full(n3[2] // 2 + 1, 2.0)
            w[0] = 1.0
            if n3[2] % 2 == 0:
                w[-1] = 1.0
            want_norm = float((abs(ck_full) ** 2 * w).sum())
            assert abs(rhok.cnorm() - want_norm) < 1e-5 * want_norm
            # transfer + c2r on the pencil layout
            Lb = [L] * 3
            for d in range(3):
                fr_full = oracle.c2r(oracle.transfer(ck_full, list(n3), Lb, "gravity_fd4", d), list(n3))
                f = rhok.apply(T.GravityFD4(d)).c2r()
                assert rel(f.value, fr_full[sl]) < tol, ("pencil transfer", np_, d)
            # white noise is independent of the partition
            cdt = "complex128" if dtype == "f8" else "complex64"
            wn_full = oracle.whitenoise(numpy.zeros((n3[0], n3[1], n3[2] // 2 + 1), dtype=cdt), 0, n3, 4242, False)
            wn = pm.generate_whitenoise(4242)
            wsl = tuple(slice(int(a), int(a + b)) for a, b in zip(wn.start, wn.shape))
            if wn.size:
                assert abs(wn.value - wn_full[wsl]).max() < (1e-13 if dtype == "f8" else 5e-7)
            # routing on the 2-D domain grid, decomposed paint, readout with the ghost sum
            allpos = [numpy.random.default_rng(500 + q).uniform(-0.2 * L, 1.2 * L, (2000 + 50 * q, 3)) for q in range(P)]
            scale = numpy.array(n3) / L
            lays = [oracle.decompose(allpos[q], pm.domain.edges, P, smoothing=0.5 * pm.resampler.support,
                                     assign=pm.domain.DomainAssign, scale=scale) for q in range(P)]
            layout = pm.decompose(allpos[r])
            assert numpy.array_equal(layout.sendcounts, lays[r][0]) and numpy.array_equal(layout.indices, lays[r][1])
            want_pos = oracle.exchange_all(allpos, lays)[r]
            rho = pm.paint(allpos[r], layout=layout, mode="deterministic")
            blockv = numpy.zeros(tuple(shape), dtype)
            oracle.paint(blockv, want_pos, res, scale=scale, translate=-start, period=list(n3))
            assert numpy.array_equal(rho.value, blockv), "decomposed paint on pencils differs"
            serial = numpy.zeros(n3, dtype)
            oracle.paint(serial, numpy.concatenate(allpos), res, scale=scale, period=list(n3))
            assert rel(blockv, serial[sl]) < (1e-12 if dtype == "f8" else 1e-5)
            got = rho.readout(allpos[r], layout=layout)
            want = oracle.readout(serial, allpos[r], res, scale=scale, period=list(n3))
            assert rel(got, want) < tol
            # I/O order on pencils: neither the real nor the complex field is C-order local
            for fld, whole_want in ((pm.create("real", value=full[sl]), full), (rhok, ck_full.astype(rhok.dtype))):
                chunk = numpy.empty(int(fld.size), dtype=fld.dtype)
                fld.ravel(out=chunk)
                whole = numpy.concatenate(comm.allgather(chunk))
                assert rel(whole, whole_want.ravel()) < tol
                back_f = pm.create(type(fld))
                back_f.unravel(chunk)
                assert numpy.array_equal(back_f.value, fld.value)
            comm.Barrier()
            if r == 0:
                print("pencil ok: np=%s n=%s %s %s" % (np_, n3, res, dtype))
            del pm, rho, rhok, cip, f, wn, layout

    # a second mesh of a configuration used before takes the pooled peer-memory landing buffers again
    pm = None
    for rep in range(2):
        pm = ParticleMesh(BoxSize=100.0, Nmesh=[16, 16, 16], dtype="f8", resampler="cic", comm=comm)
        shape = pm._layout["i_shape"]
        x = numpy.random.default_rng(7 + r).uniform(size=tuple(shape))
        f = pm.create("real", value=x)
        assert rel(f.r2c().c2r().value, x) < 1e-12
        del f, pm
    comm.Barrier()
    print("rank %d done" % r)


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        msg = "WORKER-ERROR rank %s\n%s" % (os.environ.get("RANK"), traceback.format_exc())
        sys.stderr.write(msg)
        sys.stderr.flush()
        try:        # keep the first failure readable even if the launcher truncates the output
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "multirank_error_rank%s.txt" % os.environ.get("RANK")), "w") as f:
                f.write(msg)
        except OSError:
            pass
        os._exit(1)
