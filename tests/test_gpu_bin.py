"""Tile-sorted copy of particle arrays without spatial order (pmb_bin.cuh): paint / readout / readout_multi on
the sorted copy + return to the caller's order == the oracle (the reference makes no assumption about particle
order, _window.pyx:157-165); the cached copy is reused only for unchanged content."""
import ctypes
import os

import numpy
import pytest
from numpy.testing import assert_allclose, assert_array_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def W():
    from pmesh_b200 import window
    return window


def _stats(ctx):
    from pmesh_b200 import _lib
    b, n = ctypes.c_int64(0), ctypes.c_int64(0)
    _lib.check(ctx.lib.pmb_bin_stats(ctx.handle, ctypes.byref(b), ctypes.byref(n)))
    return b.value, n.value


def _random_positions(N, seed, n=None):
    rng = numpy.random.default_rng(seed)
    return rng.uniform(0, N, (n or N ** 3, 3))


@pytest.mark.parametrize("name", ["cic", "tsc", "pcs"])
def test_sorted_copy_against_oracle(W, oracle, name):
    """random order, >= 2^18 particles: paint (with a per-particle mass column), readout (f8 / f4 results),
    gradient windows; full periodic canvas and a translated slab; f8, f4 and strided positions"""
    from pmesh_b200 import _lib
    from pmesh_b200.device import DeviceArray
    N = 72
    rng = numpy.random.default_rng(5)
    pos = _random_positions(N, 5)
    mass = rng.uniform(0.5, 2.0, len(pos))
    dmass = DeviceArray.from_host(mass)
    for ptype in ("f8", "f4"):
        p = pos.astype(ptype)
        dpos = DeviceArray.from_host(p)
        ctx = dpos.ctx
        _lib.check(ctx.lib.pmb_bin_release(ctx.handle))
        b0 = _stats(ctx)[0]
        for shape, translate in (((N, N, N), [0.0, 0.0, 0.0]), ((20, N, N), [-30.0, 0.0, 0.0])):
            tr = W.Affine(3, scale=1.0, translate=translate, period=N)
            for dtype, tol in (("f8", 1e-6), ("f4", 1e-4)):
                want = numpy.zeros(shape, dtype)
                oracle.paint(want, p, name, mass=mass, translate=translate, period=[N] * 3)
                mesh = DeviceArray.zeros(shape, dtype)
                W.windows[name].paint(mesh, dpos, mass=dmass, transform=tr, mode="atomic")
                assert_allclose(mesh.to_host(), want, rtol=tol, atol=tol * abs(want).max())
                field = rng.uniform(-1, 1, shape).astype(dtype)
                dfield = DeviceArray.from_host(field)
                r = W.windows[name].readout(dfield, dpos, transform=tr)
                assert_array_equal(r.to_host(), oracle.readout(field, p, name, translate=translate, period=[N] * 3))
                r = W.windows[name].readout(dfield, dpos, diffdir=1, transform=tr)
                assert_array_equal(r.to_host(), oracle.readout(field, p, name, diffdir=1, translate=translate, period=[N] * 3))
        builds = _stats(ctx)[0] - b0
        # one reorder per geometry (2), every other call re-validated the cached copy
        assert 1 <= builds <= 4, builds


def test_cached_copy_follows_the_content(W, oracle):
    """the sorted copy is keyed on the content: positions rewritten IN PLACE (same address, same count) are
    painted / read correctly, and unchanged positions do not trigger a new reorder"""
    from pmesh_b200 import _lib
    from pmesh_b200.device import DeviceArray
    N = 72
    tr = W.Affine(3, scale=1.0, translate=[0.0] * 3, period=N)
    pos = _random_positions(N, 7)
    dpos = DeviceArray.from_host(pos)
    ctx = dpos.ctx
    _lib.check(ctx.lib.pmb_bin_release(ctx.handle))
    field = numpy.random.default_rng(8).uniform(-1, 1, (N, N, N))
    dfield = DeviceArray.from_host(field)
    b0 = _stats(ctx)[0]
    r = W.CIC.readout(dfield, dpos, transform=tr)
    assert_array_equal(r.to_host(), oracle.readout(field, pos, "cic", period=[N] * 3))
    assert _stats(ctx)[0] == b0 + 1 and _stats(ctx)[1] >= pos.nbytes
    r = W.CIC.readout(dfield, dpos, transform=tr)
    assert_array_equal(r.to_host(), oracle.readout(field, pos, "cic", period=[N] * 3))
    assert _stats(ctx)[0] == b0 + 1, "unchanged content must reuse the sorted copy"
    # one coordinate of one particle changes
    pos2 = pos.copy()
    pos2[12345, 1] = (pos2[12345, 1] + 17.25) % N
    dpos.set(pos2)
    r = W.CIC.readout(dfield, dpos, transform=tr)
    assert_array_equal(r.to_host(), oracle.readout(field, pos2, "cic", period=[N] * 3))
    assert _stats(ctx)[0] == b0 + 2, "changed content must be sorted again"
    # two particles swapped: same multiset of records, different order
    pos3 = pos2.copy()
    pos3[[10, 200000]] = pos3[[200000, 10]]
    dpos.set(pos3)
    r = W.CIC.readout(dfield, dpos, transform=tr)
    assert_array_equal(r.to_host(), oracle.readout(field, pos3, "cic", period=[N] * 3))
    assert _stats(ctx)[0] == b0 + 3
    want = numpy.zeros((N, N, N))
    oracle.paint(want, pos3, "cic", period=[N] * 3)
    mesh = DeviceArray.zeros((N, N, N), "f8")
    W.CIC.paint(mesh, dpos, transform=tr, mode="atomic")
    assert_allclose(mesh.to_host(), want, rtol=1e-6, atol=1e-6)
    assert _stats(ctx)[0] == b0 + 3
    _lib.check(ctx.lib.pmb_bin_release(ctx.handle))
    assert _stats(ctx)[1] == 0


def test_three_fields_in_one_sweep_on_the_sorted_copy(W, oracle):
    """readout_multi on random order: (npart, nf) staging rows on the sorted copy, back through dest[]; odd count"""
    from pmesh_b200.device import DeviceArray
    N = 72
    rng = numpy.random.default_rng(9)
    pos = _random_positions(N, 9)[:N ** 3 - 37]
    dpos = DeviceArray.from_host(pos)
    for shape, translate in (((N, N, N), [0.0, 0.0, 0.0]), ((20, N, N), [-30.0, 0.0, 0.0])):
        tr = W.Affine(3, scale=1.0, translate=translate, period=N)
        fields = [rng.uniform(-1, 1, shape) for _ in range(3)]
        dfields = [DeviceArray.from_host(f) for f in fields]
        for nf in (1, 2, 3):
            outs = W.windows["cic"].readout_multi(dfields[:nf], dpos, transform=tr)
            for f, o in zip(fields, outs):
                assert_array_equal(o.to_host(), oracle.readout(f, pos, "cic", translate=translate, period=[N] * 3))


@pytest.mark.parametrize("switch", ["PMB_BIN=0", "PMB_BIN=2", "PMB_BIN_PAINT=0", "PMB_BIN_PAINT=1", "PMB_BIN_READOUT=1", "PMB_BIN_TILES=4096", "PMB_BIN_TILE_READOUT=0", "PMB_BIN_TZ=4", "PMB_BIN_PAINT=3", "PMB_BIN_PAINT=2", "PMB_BIN_TZ=7"])
def test_switches(W, oracle, switch):
    """PMB_BIN=0: permutation walk; PMB_BIN=2: lattice-ordered arrays are reordered too; PMB_BIN_PAINT / _READOUT:
    which kernels run on the sorted copy; PMB_BIN_TILES: coarser tiles.  Same results."""
    from pmesh_b200 import _lib
    from pmesh_b200.device import DeviceArray
    N = 72
    k, v = switch.split("=")
    old = os.environ.get(k)
    os.environ[k] = v
    try:
        rng = numpy.random.default_rng(11)
        q = numpy.indices((N, N, N)).reshape(3, -1).T + 0.5
        for pos in (_random_positions(N, 11), (q + rng.uniform(-0.4, 0.4, q.shape)) % N):
            dpos = DeviceArray.from_host(pos)
            _lib.check(dpos.ctx.lib.pmb_bin_release(dpos.ctx.handle))
            tr = W.Affine(3, scale=1.0, translate=[0.0] * 3, period=N)
            want = numpy.zeros((N, N, N))
            oracle.paint(want, pos, "cic", period=[N] * 3)
            mesh = DeviceArray.zeros((N, N, N), "f8")
            W.CIC.paint(mesh, dpos, transform=tr, mode="atomic")
            assert_allclose(mesh.to_host(), want, rtol=1e-6, atol=1e-6)
            r = W.CIC.readout(DeviceArray.from_host(want), dpos, transform=tr)
            assert_array_equal(r.to_host(), oracle.readout(want, pos, "cic", period=[N] * 3))
    finally:
        if old is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = old
