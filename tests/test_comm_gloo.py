"""Host-side N>1 logic on CPU: two processes over torch.distributed (gloo).

Covers the communicator facade (allreduce / allgather / bcast / Alltoall / Barrier), the count
exchange and offset bookkeeping of Layout (domain.py:92-123), and -- with the ORACLE standing in for
the device kernels, in this test only -- that routing + exchange + slab-local deterministic paint
summed over ranks reproduces the serial paint (the reference's distributed-correctness idiom,
tests/test_pm.py:228-264).
"""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy
    sys.path.insert(0, %(root)r)
    sys.path.insert(0, os.path.join(%(root)r, "oracle"))
    from pmesh_b200 import comm as C
    from pmesh_b200.domain import Layout
    import oracle

    comm = C.world()
    assert isinstance(comm, C.TorchComm) and comm.size == 2
    r = comm.rank
    assert comm.allreduce(r + 1) == 3
    assert comm.allreduce(numpy.array([r, 10 * r])).tolist() == [1, 10]
    assert comm.allreduce(float(r), op=C.MAX) == 1.0
    assert comm.allgather("rank%%d" %% r) == ["rank0", "rank1"]
    assert comm.bcast(numpy.dtype("f4") if r == 0 else numpy.dtype("f8")) == numpy.dtype("f4")
    send = numpy.array([10 * r + 0, 10 * r + 1], dtype="int32")
    recv = numpy.empty_like(send)
    comm.Alltoall(send, recv)
    assert recv.tolist() == [r, 10 + r]
    comm.Barrier()

    # routing of a 2-slab decomposition through the oracle, counts through the real Alltoall
    n = 8
    rng = numpy.random.default_rng(5)
    allpos = rng.uniform(0, n, (2, 400, 3))            # both ranks know both particle sets
    mypos = allpos[r]
    edges = [numpy.array([0., 4., 8.]), numpy.array([0., 8.]), numpy.array([0., 8.])]
    counts, indices = oracle.decompose(mypos, edges, 2, smoothing=1.0)
    layout = Layout(comm, len(mypos), counts, indices)
    other = oracle.decompose(allpos[1 - r], edges, 2, smoothing=1.0)
    assert layout.recvcounts.tolist() == [(counts if q == r else other[0])[r] for q in range(2)]
    assert layout.sendoffsets.tolist() == [0, int(counts[0])]
    assert layout.recvlength == layout.recvcounts.sum()
    cost = layout.get_exchange_cost()
    assert cost[r] == counts[1 - r]

    # simulated exchange (oracle) + slab-local paint; the sum over ranks equals the serial paint
    lay = [oracle.decompose(allpos[q], edges, 2, smoothing=1.0) for q in range(2)]
    recv_pos = oracle.exchange_all([allpos[0], allpos[1]], lay)[r]
    assert len(recv_pos) == layout.recvlength
    slab = numpy.zeros((4, n, n))
    oracle.paint(slab, recv_pos, "cic", translate=[-4.0 * r, 0, 0], period=[n, n, n])
    slabs = comm.allgather(slab)
    full = numpy.concatenate(slabs, axis=0)
    serial = numpy.zeros((n, n, n))
    oracle.paint(serial, numpy.concatenate([allpos[0], allpos[1]]), "cic", period=[n, n, n])
    assert abs(full - serial).max() < 1e-13, abs(full - serial).max()
    assert abs(full.sum() - 800) < 1e-9
    # I/O-order exchange behind ravel / unravel / Fourier resample on P > 1 (resample.dist_put / dist_take,
    # the reference's mpsort.sort / permute / take): a global array of 37 items cut into chunks of 20 + 17
    from pmesh_b200 import resample as R
    whole = numpy.random.default_rng(9).uniform(size=37)
    lens = [20, 17]
    first = sum(lens[:r])
    mine_g = numpy.arange(37)[r::2]                      # this rank holds the items with g mod 2 == r, any order
    mine_g = mine_g[numpy.random.default_rng(r).permutation(len(mine_g))]
    chunk = R.dist_put(comm, whole[mine_g], mine_g, lens[r])
    assert numpy.array_equal(chunk, whole[first:first + lens[r]])
    want = numpy.random.default_rng(20 + r).integers(0, 37, size=25)
    assert numpy.array_equal(R.dist_take(comm, chunk, want), whole[want])
    assert len(R.dist_take(comm, chunk, numpy.zeros(0, dtype="i8"))) == 0
    sys.stdout.write("rank%%d-ok\\n" %% r); sys.stdout.flush()
""")


def test_two_process_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = "1"
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-4000:]
    assert p.stdout.count("-ok") == 2, p.stdout[-2000:]
