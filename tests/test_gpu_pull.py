"""Deterministic paint by sorting the PARTICLES by cell and letting every cell sum its own contributions in
particle order (pmb_pull.cuh): bit-identical to the oracle's sequential loop (the reference's order of additions,
_window_imp.c / _window_tuned_*.h) and to the pairs path it replaces."""
import os

import numpy
import pytest
from numpy.testing import assert_array_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def W():
    from pmesh_b200 import window
    return window


class _env(object):
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        for k, v in self.kw.items():
            os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _clustered(N, n, seed):
    """uniform background + two tight blobs (hundreds of particles per cell: long runs to merge), shuffled"""
    rng = numpy.random.default_rng(seed)
    a = rng.uniform(0, N, (n // 2, 3))
    b = (numpy.array([0.3, N - 0.4, N / 2.0]) + rng.normal(0, 0.7, (n // 4, 3))) % N      # straddles the period
    c = (numpy.array([N / 3.0, 5.2, 1.1]) + rng.normal(0, 1.5, (n - n // 2 - n // 4, 3))) % N
    pos = numpy.concatenate([a, b, c])
    return pos[rng.permutation(len(pos))]


@pytest.mark.parametrize("name", ["nnb", "cic", "tsc", "pcs"])
def test_pull_paint_is_the_sequential_sum(W, oracle, name):
    from pmesh_b200.device import DeviceArray
    N = 40
    rng = numpy.random.default_rng(3)
    pos = _clustered(N, 70000, 3)
    mass = rng.uniform(0.5, 2.0, len(pos))
    dmass = DeviceArray.from_host(mass)
    for ptype in ("f8", "f4"):
        p = pos.astype(ptype)
        dpos = DeviceArray.from_host(p)
        cases = [((N, N, N), dict(scale=1.0, translate=[0.0] * 3, period=N)),
                 ((12, N, N), dict(scale=1.0, translate=[-14.0, 0.0, 0.0], period=N)),          # a slab of the periodic mesh
                 ((9, N, 11), dict(scale=1.0, translate=[3.0, 0.0, -33.0], period=N)),          # slabs along two axes, one across the period
                 ((30, 44, 25), dict(scale=[0.9, 1.1, 0.5], translate=[1.0, -2.0, 3.5], period=0))]   # non-periodic
        for shape, kw in cases:
            tr = W.Affine(3, **kw)
            okw = dict(scale=kw["scale"], translate=kw["translate"], period=[kw["period"]] * 3)
            for dtype in ("f8", "f4"):
                for diffdir in (None, 0, 2):
                    if diffdir is not None and (dtype == "f4" or ptype == "f4"):
                        continue
                    start = rng.uniform(-1, 1, shape).astype(dtype)      # the canvas is accumulated into
                    want = start.copy()
                    oracle.paint(want, p, name, mass=mass, diffdir=diffdir, **okw)
                    mesh = DeviceArray.from_host(start)
                    W.windows[name].paint(mesh, dpos, mass=dmass, diffdir=diffdir, transform=tr, mode="deterministic")
                    assert_array_equal(mesh.to_host(), want)
        # scalar mass, and the pairs path gives the same bits
        tr = W.Affine(3, scale=1.0, translate=[0.0] * 3, period=N)
        want = numpy.zeros((N, N, N))
        oracle.paint(want, p, name, period=[N] * 3)
        mesh = DeviceArray.zeros((N, N, N), "f8")
        W.windows[name].paint(mesh, dpos, transform=tr, mode="deterministic")
        assert_array_equal(mesh.to_host(), want)
        with _env(PMB_PULL=0):
            mesh = DeviceArray.zeros((N, N, N), "f8")
            W.windows[name].paint(mesh, dpos, transform=tr, mode="deterministic")
            assert_array_equal(mesh.to_host(), want)


def test_pull_small_and_degenerate_canvases(W, oracle):
    """with the size threshold off every 3-D tuned paint goes through the particle sort: periods equal to the
    support (full wrap), periods below it (pairs path), empty regions, particles that reach no cell"""
    from pmesh_b200.device import DeviceArray
    rng = numpy.random.default_rng(4)
    with _env(PMB_PULL_MIN=0):
        for name, S in (("cic", 2), ("tsc", 3), ("pcs", 4)):
            for per in (S - 1, S, S + 1, 7):
                if per < 1:
                    continue
                pos = rng.uniform(-2 * per, 3 * per, (500, 3))
                mass = rng.uniform(0.5, 2.0, len(pos))
                want = numpy.zeros((per, per, per))
                oracle.paint(want, pos, name, mass=mass, period=[per] * 3)
                mesh = DeviceArray.zeros((per, per, per), "f8")
                W.windows[name].paint(mesh, DeviceArray.from_host(pos), mass=DeviceArray.from_host(mass),
                                      transform=W.Affine(3, scale=1.0, translate=[0.0] * 3, period=per), mode="deterministic")
                assert_array_equal(mesh.to_host(), want)
            # non-periodic canvas far smaller than the cloud: most particles reach no cell
            pos = rng.uniform(-20, 30, (3000, 3))
            want = numpy.zeros((6, 5, 7))
            oracle.paint(want, pos, name, period=[0] * 3)
            mesh = DeviceArray.zeros((6, 5, 7), "f8")
            W.windows[name].paint(mesh, DeviceArray.from_host(pos), transform=W.Affine(3, scale=1.0, translate=[0.0] * 3, period=0),
                                  mode="deterministic")
            assert_array_equal(mesh.to_host(), want)
