"""GPU routing parity: GridND.decompose / Layout.exchange / Layout.gather on one rank and in
simulated multi-rank geometry, bit-exact against the oracle (which is pinned to the reference's
Cython gridnd_fill and numpy digitize path, tests/test_oracle.py).

Multi-rank geometry on ONE GPU: GridND accepts any communicator object; ``FakeComm(rank, size)``
gives the routing kernel the rank count of a P-rank job, and the returned counts / indices are
compared with the oracle's for the same (edges, DomainAssign, P).  The NCCL leg itself is
exercised by tests/test_gpu_multirank.py (needs >= 2 GPUs) and by bench.py --gpus N.
"""
import numpy
import pytest
from numpy.testing import assert_array_equal

pytestmark = pytest.mark.gpu


class FakeComm(object):
    """size-P communicator seen from one rank; collectives that decompose() needs are local"""
    def __init__(self, rank, size):
        self.rank, self.size = rank, size

    def Barrier(self):
        pass

    def Alltoall(self, send, recv):
        recv[...] = send          # not a real transpose: only sendcounts/indices are checked

    def allgather(self, x):
        return [x] * self.size

    def bcast(self, x, root=0):
        return x

    def allreduce(self, x, op=None):
        return x

    def ensure_device_comm(self, ctx):
        pass


@pytest.fixture(scope="module")
def D():
    from pmesh_b200 import domain
    return domain


CASES = [
    # (edges, P, smoothing, periodic)
    ([numpy.linspace(0, 4, 5)], 4, 1, True),
    ([numpy.linspace(0, 4, 3), numpy.linspace(0, 4, 3)], 4, 0.5, True),
    ([numpy.linspace(0, 64, 3), numpy.linspace(0, 64, 3), numpy.linspace(0, 64, 2)], 4, 1.0, True),
    ([numpy.linspace(0, 64, 9), numpy.array([0, 64.]), numpy.array([0, 64.])], 8, 1.5, True),
    ([numpy.linspace(0, 64, 3), numpy.linspace(0, 64, 5), numpy.array([0, 64.])], 8, [1.0, 2.0, 0.0], True),
    ([numpy.linspace(0, 10, 4), numpy.linspace(0, 10, 3)], 6, 0.7, False),
    ([numpy.array([0, 0, 2, 4., 4.]), numpy.array([0, 2., 4.])], 8, 0.3, True),       # degenerate domains
    ([numpy.linspace(0, 8, 3), numpy.linspace(0, 8, 3), numpy.linspace(0, 8, 3)], 3, 5.0, True),  # fewer ranks than domains, huge smoothing
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("posdtype", ["f8", "f4"])
def test_decompose_matches_oracle(D, oracle, case, posdtype):
    edges, P, smoothing, periodic = CASES[case]
    nd = len(edges)
    rng = numpy.random.default_rng(case)
    box = numpy.array([e[-1] for e in edges])
    pos = rng.uniform(-0.5, 1.5, (4000, nd)) * box        # includes out-of-box particles
    pos[:50] = 0.0
    pos[50:60] = -1e-17                                   # Q5: x % box == box
    pos[60:70] = box
    pos = pos.astype(posdtype)
    scale = numpy.array([1.0, 0.5, 2.0])[:nd] if case % 2 else None
    epos = pos if scale is None else pos                  # same input; the scale is applied by both sides
    sedges = edges
    for rank in (0, P - 1):
        g = D.GridND(sedges, comm=FakeComm(rank, P), periodic=periodic)
        tr = None if scale is None else D.ScaleTransform(scale)
        layout = g.decompose(pos, smoothing=smoothing, transform=tr)
        counts, indices = oracle.decompose(epos, sedges, P, smoothing=smoothing, periodic=periodic,
                                           assign=g.DomainAssign, scale=scale)
        assert layout.sendcounts.dtype == numpy.dtype("int32")
        assert_array_equal(layout.sendcounts, counts)
        assert layout.indices.dtype == numpy.dtype("int32")
        assert_array_equal(layout.indices, indices)


@pytest.mark.parametrize("case", range(len(CASES)))
def test_decompose_ordered_particles_home_cell_path(D, oracle, case):
    """particle arrays that FOLLOW the decomposition (sorted along the axes: whole warps inside one domain cell) take
    the routing kernel's home-cell shortcut; boundary layers, edge values, -0.0, out-of-box values and NaNs in between
    send single warp steps down the general path.  Counts and indices bit-exact against the oracle."""
    edges, P, smoothing, periodic = CASES[case]
    nd = len(edges)
    rng = numpy.random.default_rng(100 + case)
    box = numpy.array([e[-1] for e in edges])
    n = 60000
    pos = rng.uniform(0.0, 1.0, (n, nd)) * box
    # lexicographic order by domain cell: long runs of particles in one cell
    cell = [numpy.digitize(pos[:, d], edges[d]) for d in range(nd)]
    order = numpy.lexsort(tuple(cell[d] for d in reversed(range(nd))))
    pos = pos[order]
    # adversaries inside the runs
    k = rng.choice(n, 400, replace=False)
    pos[k[:50]] = 0.0
    pos[k[50:100]] = -0.0
    pos[k[100:150]] = box
    pos[k[150:200]] = -1e-17
    pos[k[200:260]] = rng.uniform(-0.5, 1.5, (60, nd)) * box
    for d in range(nd):
        e = numpy.asarray(edges[d], dtype="f8")
        pos[k[260 + 40 * d:300 + 40 * d], d] = rng.choice(e, 40) + rng.choice([0.0, 1e-12, -1e-12], 40)
    pos[k[380:390]] = numpy.nan
    scale = numpy.array([1.0, 0.5, 2.0])[:nd] if case % 2 else None
    for rank in (0, P - 1):
        g = D.GridND(edges, comm=FakeComm(rank, P), periodic=periodic)
        tr = None if scale is None else D.ScaleTransform(scale)
        layout = g.decompose(pos, smoothing=smoothing, transform=tr)
        counts, indices = oracle.decompose(pos, edges, P, smoothing=smoothing, periodic=periodic,
                                           assign=g.DomainAssign, scale=scale)
        assert_array_equal(layout.sendcounts, counts)
        assert_array_equal(layout.indices, indices)


def test_decompose_empty_and_device_positions(D, oracle):
    from pmesh_b200.device import DeviceArray
    g = D.GridND([numpy.linspace(0, 4, 3)] * 2, comm=FakeComm(0, 4))
    layout = g.decompose(numpy.zeros((0, 2)), smoothing=1)
    assert layout.sendcounts.sum() == 0 and len(layout.indices) == 0
    rng = numpy.random.default_rng(1)
    pos = rng.uniform(0, 4, (100000, 3))                  # extra column is ignored
    layout = g.decompose(DeviceArray.from_host(pos), smoothing=0.25)
    counts, indices = oracle.decompose(pos, g.edges, 4, smoothing=0.25)
    assert_array_equal(layout.sendcounts, counts)
    assert_array_equal(layout.indices, indices)


def test_callable_transform_compat(D, oracle):
    g = D.GridND([numpy.linspace(0, 8, 5)], comm=FakeComm(0, 4))
    pos = numpy.random.default_rng(2).uniform(0, 1, (1000, 1))
    layout = g.decompose(pos, smoothing=0.1, transform=lambda x: x * 8)
    counts, indices = oracle.decompose(pos * 8, g.edges, 4, smoothing=0.1)
    assert_array_equal(layout.sendcounts, counts)
    assert_array_equal(layout.indices, indices)


def test_single_rank_exchange_gather(D, oracle):
    """P = 1 with a periodic domain: everything is local; exchange == take, gather('sum') == identity"""
    from pmesh_b200.device import DeviceArray
    g = D.GridND([numpy.array([0, 8.])] * 3)
    assert g.comm.size == 1
    rng = numpy.random.default_rng(4)
    pos = rng.uniform(0, 8, (5000, 3))
    layout = g.decompose(pos, smoothing=1.0)
    assert layout.sendcounts[0] == 5000
    assert_array_equal(layout.indices, numpy.arange(5000))
    lpos = layout.exchange(pos)
    assert_array_equal(lpos, pos)
    mass = rng.uniform(size=5000).astype("f4")
    lp, lm = layout.exchange(pos, mass)
    assert lm.dtype == numpy.dtype("f4")
    assert_array_equal(lm, mass)
    ids = numpy.arange(5000, dtype="i8")
    assert_array_equal(layout.exchange(ids), ids)
    rec = numpy.zeros(5000, dtype=[("a", "f8"), ("b", "i4", 3)])
    rec["a"] = mass
    assert_array_equal(layout.exchange(rec), rec)          # structured records (tests/test_domain.py:94-118)
    back = layout.gather(lm, mode="sum")
    assert back.dtype == numpy.dtype("f4")
    assert_array_equal(back, mass)
    back3 = layout.gather(lpos, mode="sum")
    assert_array_equal(back3, pos)
    for mode in ("all", "any", "mean", "local", numpy.add):
        assert_array_equal(layout.gather(lm, mode=mode), mass)
    d = layout.exchange(DeviceArray.from_host(pos))
    assert_array_equal(d.to_host(), pos)
    with pytest.raises(ValueError):
        layout.exchange(pos[:10])


def test_identity_layout_fast_path(D, oracle):
    """a single periodic domain: the routing never looks at the positions (indices = arange, built on
    demand), exchange hands device arrays on without a copy and gather('sum') is out = 0.0 + data;
    counts / indices still equal the oracle's, also when the one domain belongs to a P-rank job"""
    from pmesh_b200.device import DeviceArray
    rng = numpy.random.default_rng(11)
    pos = rng.uniform(-8, 16, (3000, 3))                   # out-of-box positions wrap into the one domain
    for P in (1, 3):
        g = D.GridND([numpy.array([0, 8.])] * 3, comm=FakeComm(0, P))
        layout = g.decompose(pos, smoothing=1.0)
        assert layout.identity
        counts, indices = oracle.decompose(pos, g.edges, P, smoothing=1.0)
        assert_array_equal(layout.sendcounts, counts)
        assert_array_equal(layout.indices, indices)
        assert layout.indices.dtype == numpy.dtype("int32")
        assert_array_equal(layout.indices_device.to_host(), indices)
    g = D.GridND([numpy.array([0, 8.])] * 3)
    layout = g.decompose(pos, smoothing=1.0)
    dpos = DeviceArray.from_host(pos)
    lpos = layout.exchange(dpos)
    assert lpos.ptr == dpos.ptr                            # no copy for device-resident columns
    assert_array_equal(lpos.to_host(), pos)
    hp = layout.exchange(pos)
    assert hp is not pos and not numpy.shares_memory(hp, pos)
    assert_array_equal(hp, pos)
    for dt, odt in (("f8", "f8"), ("f4", "f4"), ("f8", "f4"), ("f4", "f8")):
        v = rng.uniform(-1, 1, 3000).astype(dt)
        v[:4] = [-0.0, 0.0, numpy.inf, -1e-300]
        out = DeviceArray.empty((3000,), odt)
        layout.gather(DeviceArray.from_host(v), mode="sum", out=out)
        want = numpy.bincount(numpy.arange(3000), weights=v, minlength=3000).astype(odt)
        got = out.to_host()
        assert_array_equal(got, want)
        assert_array_equal(numpy.signbit(got), numpy.signbit(want))   # 0.0 + -0.0 == +0.0, as bincount
    v3 = rng.uniform(-1, 1, (3000, 3))
    assert_array_equal(layout.gather(v3, mode="sum"), v3)
    # a non-periodic single domain is NOT the identity (particles outside are dropped)
    gn = D.GridND([numpy.array([0, 8.])] * 3, periodic=False)
    ln = gn.decompose(pos, smoothing=1.0)
    counts, indices = oracle.decompose(pos, gn.edges, 1, smoothing=1.0, periodic=False)
    assert_array_equal(ln.sendcounts, counts)
    assert_array_equal(ln.indices, indices)


def test_gather_sum_matches_bincount(D, oracle):
    """the ghost reduction kernel against numpy.bincount on a fabricated multi-rank layout"""
    from pmesh_b200.device import DeviceArray
    rng = numpy.random.default_rng(8)
    P, n = 4, 3000
    edges = [numpy.linspace(0, 8, 3), numpy.linspace(0, 8, 3)]
    pos = rng.uniform(0, 8, (n, 2))
    counts, indices = oracle.decompose(pos, edges, P, smoothing=1.0)
    lay = D.Layout(FakeComm(0, 1), n, numpy.array([len(indices)], dtype="int32"), indices)
    # present the layout to the kernel with its true per-rank segments
    lay.comm = FakeComm(0, P)
    lay.sendcounts = counts
    lay.sendoffsets = numpy.concatenate([[0], numpy.cumsum(counts)[:-1]]).astype("int32")
    lay.recvcounts = counts
    lay.recvoffsets = lay.sendoffsets
    lay.recvlength = counts.sum()
    # pretend the reverse alltoallv happened: what comes back is everything but the block of this rank
    # (rank 0: the first segment), packed -- the layout Layout.gather expects of its `back` buffer

    def fake_alltoallv(ctx, send, sc, so, recv, rc, ro, itemsize, skip_self=False):
        own = int(counts[0]) * int(itemsize)
        nbytes = int(counts.sum()) * int(itemsize) - own
        return DeviceArray((max(nbytes, 1),), "u1", ptr=send.ptr + own, base=send, ctx=ctx)
    lay._alltoallv = fake_alltoallv
    for dt in ("f8", "f4"):
        for trailing in ((), (3,)):
            vals = rng.uniform(-1, 1, (len(indices),) + trailing).astype(dt)
            want = oracle.bincount_sum(indices, vals, n)
            got = lay.gather(vals, mode="sum")
            assert got.dtype == numpy.dtype(dt)
            assert_array_equal(got, want)
            gd = lay.gather(DeviceArray.from_host(vals), mode="sum")
            assert_array_equal(gd.to_host(), want)
