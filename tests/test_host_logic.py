"""Host-side logic that needs neither a GPU nor the oracle: index tables of the Fourier resample,
the identity layout's lazy indices, offsets bookkeeping of Layout (reference domain.py:92-123)."""
import numpy
from numpy.testing import assert_array_equal


def test_reindex_matches_the_reference_examples():
    from pmesh_b200.pm import reindex
    # the docstring examples of pm.py:1128-1144
    assert_array_equal(reindex(8, 4), [0, 1, 2, 7])
    assert_array_equal(reindex(4, 8), [0, 1, 2, -1, -1, -1, -1, 3])
    # same mesh: identity; frequencies are preserved wherever the index is valid
    assert_array_equal(reindex(6, 6), numpy.arange(6))
    for ns, nd in ((8, 4), (4, 8), (16, 6), (6, 16)):
        r = reindex(ns, nd)
        fs = numpy.fft.fftfreq(ns) * ns
        fd = numpy.fft.fftfreq(nd) * nd
        ok = r >= 0
        # apart from the Nyquist of the smaller mesh (sign ambiguous), indices map equal frequencies
        nyq = min(ns, nd) // 2
        sel = ok & (abs(fd) != nyq)
        assert_array_equal(fs[r[sel]], fd[sel])


def test_identity_layout_is_lazy_and_consistent():
    from pmesh_b200.comm import SelfComm
    from pmesh_b200.domain import Layout
    lay = Layout(SelfComm(), 1000, numpy.array([1000], dtype='int32'), None, identity=True)
    assert lay.identity and lay._indices_host is None and lay._indices_dev is None
    assert_array_equal(lay.sendoffsets, [0])
    assert_array_equal(lay.recvcounts, [1000])
    assert lay.recvlength == 1000 and lay.sendlength == 1000
    assert lay._indices_ptr() is None
    ind = lay.indices
    assert ind.dtype == numpy.dtype('int32')
    assert_array_equal(ind, numpy.arange(1000))
    assert_array_equal(lay.get_exchange_cost(), [0])


def test_layout_offsets():
    from pmesh_b200.domain import Layout

    class Comm3(object):
        rank, size = 1, 3

        def Alltoall(self, send, recv):
            recv[...] = [5, 7, 11]

        def allgather(self, x):
            return [x] * 3
    lay = Layout(Comm3(), 10, numpy.array([2, 3, 4], dtype='int32'), numpy.arange(9, dtype='int32'))
    assert_array_equal(lay.sendoffsets, [0, 2, 5])
    assert_array_equal(lay.recvcounts, [5, 7, 11])
    assert_array_equal(lay.recvoffsets, [0, 5, 12])
    assert lay.recvlength == 23 and not lay.identity
    assert_array_equal(lay.get_exchange_cost(), [6, 6, 6])


def test_exchange_tuner_measures_then_sticks():
    """domain.ExchangeTuner: evaluations run full (cold), split, full under the timer; the slowest rank's times decide
    for every rank; a single rank, or PMB_EXCHANGE, needs no measurement"""
    import os
    from pmesh_b200.domain import ExchangeTuner

    class Ctx(object):
        def __init__(self, times):
            self.times = list(times)
            self.started = 0

        def timer_start(self, slot):
            self.started += 1

        def timer_stop(self, slot):
            return self.times.pop(0)

    class Comm(object):
        def __init__(self, size, others):
            self.size, self.others = size, others

        def allgather(self, x):
            return [x] + [o for o in self.others]

    os.environ.pop("PMB_EXCHANGE", None)
    # split faster on this rank and everywhere
    t = ExchangeTuner(Comm(2, [1.0]), Ctx([90.0, 50.0, 60.0]))
    modes = []
    for _ in range(5):
        modes.append(t.begin())
        t.end()
    assert modes == ['full', 'split', 'full', 'split', 'split'] and t.choice == 'split'
    assert t.measured == {'full': 60.0, 'split': 50.0} and t.ctx.started == 3
    # another rank is slow in the split path: it decides
    t = ExchangeTuner(Comm(2, [70.0]), Ctx([90.0, 50.0, 60.0]))
    for _ in range(3):
        t.begin()
        t.end()
    assert t.choice == 'full' and t.begin() == 'full'
    # one rank: nothing to exchange; the environment fixes the choice
    assert ExchangeTuner(Comm(1, []), Ctx([])).begin() == 'full'
    os.environ["PMB_EXCHANGE"] = "split"
    try:
        assert ExchangeTuner(Comm(4, [0.0] * 3), Ctx([])).begin() == 'split'
    finally:
        del os.environ["PMB_EXCHANGE"]


def test_fd4_gradient_is_a_stencil_along_the_line():
    """the identity pmb_ifft.cuh uses for direction 0 of the finite-difference gradient (examples/nbody.py:162-170):
    multiplying the modes by i * kfinite(k) = i (8 sin w - sin 2w) / (6 C), w = k C, equals the periodic 5-point stencil
    (8 (phi[x+1] - phi[x-1]) - (phi[x+2] - phi[x-2])) / (12 C) on the inverse transform -- also with the wavenumber
    convention of pm.py:1213-1219 (negative frequencies from N/2 on, Nyquist included)"""
    import numpy
    rng = numpy.random.default_rng(5)
    for n, L in ((64, 100.0), (128, 7.0)):
        C = L / n
        i0 = numpy.arange(n)
        k = numpy.where(i0 >= n // 2, i0 - n, i0) * (2 * numpy.pi / L)
        w = k * C
        kfinite = 1.0 / C * 1 / 6.0 * (8 * numpy.sin(w) - numpy.sin(2 * w))
        modes = rng.normal(size=n) + 1j * rng.normal(size=n)
        direct = numpy.fft.ifft(1j * kfinite * modes) * n
        phi = numpy.fft.ifft(modes) * n
        sten = (8 * (numpy.roll(phi, -1) - numpy.roll(phi, 1)) - (numpy.roll(phi, -2) - numpy.roll(phi, 2))) / (12 * C)
        assert abs(sten - direct).max() <= 1e-12 * abs(direct).max()
