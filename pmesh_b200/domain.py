"""Domain decomposition -- the pmesh.domain API (reference pmesh/domain.py) with GPU routing.

``GridND.decompose`` computes, on the device, the rank set every particle's
smoothing patch touches (bit-identical to the reference's numpy ``digitize``
chunks + Cython ``gridnd_fill``, pmesh/domain.py:561-652, pmesh/_domain.pyx:9-122)
and returns a ``Layout`` whose ``exchange`` / ``gather`` move particle records
between GPUs with a gather-pack kernel + NCCL alltoallv (pmesh/domain.py:138-318).

Particle arrays may be numpy (copied in and out) or ``DeviceArray`` (resident).
"""
import ctypes
import heapq

import numpy

from . import _lib
from . import comm as _comm
from .device import DeviceArray, is_device


class ScaleTransform(object):
    """x -> scale * x ; the transform ParticleMesh.decompose hands to GridND (pm.py:1786-1790).
    Being an object (not a closure) lets the GPU routing kernel apply it itself."""
    def __init__(self, scale):
        self.scale = numpy.asarray(scale, dtype='f8')

    def __call__(self, x):
        return self.scale * x


def bincountv(x, weights, minlength=None, dtype=None, out=None):
    """ bincount with vector weights (reference domain.py:26-48) """
    weights = numpy.array(weights)
    if minlength is None:
        minlength = 0 if len(x) == 0 else x.max() + 1
    if dtype is None:
        dtype = weights.dtype
    dtype = numpy.dtype(dtype)
    shape = [minlength] + list(weights.shape[1:])
    if out is None:
        out = numpy.empty(shape, dtype=dtype)
    for index in numpy.ndindex(*shape[1:]):
        ind = tuple([Ellipsis] + list(index))
        out[ind] = numpy.bincount(x, weights[ind], minlength=minlength)
    return out


def promote(data, comm):
    """every rank adopts the dtype of rank 0 and checks trailing shapes (reference domain.py:50-57)"""
    data = numpy.asarray(data)
    dtype_root = comm.bcast(data.dtype)
    data = data.astype(dtype_root)
    shape_root = comm.bcast(data.shape)
    if shape_root[1:] != data.shape[1:]:
        raise ValueError('the shape of the data does not match across ranks.')
    return data


class ExchangeTuner(object):
    """
    Chooses, by measurement, between the two ways a force evaluation can serve the particles a rank keeps:
    ``'split'`` (paint / read them where they lie, only the records that change rank travel: Layout.exchange_remote)
    and ``'full'`` (every record through take + alltoallv: Layout.exchange, the reference's path).  Which one is faster
    depends on the machine-wide picture (2 / 4 GPUs, 1024^3: split by 12 - 15 %; 8 GPUs: full), so the first
    evaluations are timed with CUDA events -- one cold, then one of each -- the slowest rank decides for everybody
    and every later evaluation takes the winner.  ``PMB_EXCHANGE=split|full`` fixes the choice.
    """
    ORDER = ('full', 'split', 'full')      # call 0 pays the one-time costs and is not compared

    def __init__(self, comm, ctx, slot=7):
        import os
        self.comm, self.ctx, self.slot = comm, ctx, slot
        self.calls = 0
        self.ms = {}
        env = os.environ.get("PMB_EXCHANGE", "")
        self.choice = env if env in ('split', 'full') else None
        if comm.size == 1:
            self.choice = 'full'
        self._timing = False

    def begin(self):
        """-> the mode of this evaluation"""
        if self.choice is not None:
            return self.choice
        self.ctx.timer_start(self.slot)
        self._timing = True
        return self.ORDER[self.calls]

    def end(self):
        if not self._timing:
            return
        self._timing = False
        ms = self.ctx.timer_stop(self.slot)
        self.ms[self.ORDER[self.calls]] = ms          # the second 'full' overwrites the cold one
        self.calls += 1
        if self.calls == len(self.ORDER):
            worst = dict((k, max(self.comm.allgather(float(v)))) for k, v in sorted(self.ms.items()))
            self.choice = 'split' if worst['split'] < worst['full'] else 'full'
            self.measured = worst


class Layout(object):
    """
    The communication layout of a domain decomposition (reference domain.py:82-318).

    Do not create a Layout directly; use :py:meth:`GridND.decompose`.
    Useful methods are :py:meth:`exchange` and :py:meth:`gather`.
    """
    def __init__(self, comm, sendlength, sendcounts, indices, recvcounts=None, identity=False):
        self.comm = comm
        # identity: every particle is sent exactly once, in order (indices = arange(sendlength));
        # found by the routing for a single periodic domain.  `indices` is then built only on demand
        # and exchange / gather skip the index indirection.
        self.identity = bool(identity)
        sendcounts = numpy.asarray(sendcounts)
        assert self.comm.size == sendcounts.shape[0]

        self.sendcounts = numpy.array(sendcounts, order='C')
        self.recvcounts = numpy.empty_like(self.sendcounts, order='C')
        self.sendoffsets = numpy.zeros_like(self.sendcounts, order='C')
        self.recvoffsets = numpy.zeros_like(self.recvcounts, order='C')

        if recvcounts is None:
            # the reference brackets this Alltoall with Barriers (domain.py:112-114); the exchange of
            # counts is itself synchronising, so they are dropped here
            self.comm.Alltoall(self.sendcounts, self.recvcounts)
        else:
            self.recvcounts = numpy.asarray(recvcounts)
        self.sendoffsets[1:] = self.sendcounts.cumsum()[:-1]
        self.recvoffsets[1:] = self.recvcounts.cumsum()[:-1]

        self.sendlength = sendlength
        self.recvlength = self.recvcounts.sum()

        if indices is None:
            assert self.identity
            self._indices_dev = None
            self._indices_host = None
        elif is_device(indices):
            self._indices_dev = indices
            self._indices_host = None
        else:
            self._indices_host = numpy.ascontiguousarray(indices, dtype='int32')
            self._indices_dev = None

    # indices live on the device; a host copy is made when somebody looks at them
    @property
    def indices(self):
        if self._indices_host is None:
            if self._indices_dev is None:
                self._indices_host = numpy.arange(self.sendlength, dtype='int32')
            else:
                self._indices_host = self._indices_dev.to_host()
        return self._indices_host

    @property
    def indices_device(self):
        if self._indices_dev is None:
            self._indices_dev = DeviceArray.from_host(self.indices)
        return self._indices_dev

    def _indices_ptr(self):
        """device pointer of indices, or NULL for the identity layout (the library then skips the indirection)"""
        return None if self.identity else self.indices_device.ptr

    def get_exchange_cost(self):
        """ exchange cost of every rank: items sent to ranks other than itself (domain.py:125-136) """
        mask = numpy.arange(self.comm.size) != self.comm.rank
        sendcount = numpy.sum(self.sendcounts[mask])
        return numpy.array(self.comm.allgather(sendcount))

    # ------------------------------------------------------------------ exchange
    def exchange(self, *args, pack=True):
        """
        Delivers data to the intersecting domains (reference domain.py:138-171).

        Every positional argument is an array (numpy or DeviceArray) of the length and ordering of the
        positions that built the layout; ghosts are created where a particle intersects several
        domains.  ``pack`` is accepted for compatibility (columns are sent one by one; the result
        is identical).
        """
        r = tuple(self._exchange(arg) for arg in args)
        if len(args) == 0:
            return None
        if len(args) == 1:
            return r[0]
        return r

    def _to_device_records(self, data, length, what):
        """-> (DeviceArray view as (N, itemsize) bytes, record dtype, trailing shape, was_host)"""
        if is_device(data):
            # device-resident columns: local length check only (no host collective on the hot path)
            if data.shape[0] != length:
                raise ValueError(what)
            assert data.is_contiguous
            return data, data.dtype, data.shape[1:], False
        data = promote(data, self.comm)
        if any(self.comm.allgather(len(data) != length)):
            raise ValueError(what)
        return DeviceArray.from_host(numpy.ascontiguousarray(data)), data.dtype, data.shape[1:], True

    def _alltoallv(self, ctx, send, sendcounts, sendoffsets, recv, recvcounts, recvoffsets, itemsize, skip_self=False):
        """device alltoallv of `itemsize`-byte records from `send` into `recv`; with skip_self the
        block a rank sends to itself is left to the caller"""
        self.comm.ensure_device_comm(ctx)
        sc = numpy.array(sendcounts, dtype='i8')
        so = numpy.ascontiguousarray(sendoffsets, dtype='i8')
        rc = numpy.array(recvcounts, dtype='i8')
        ro = numpy.ascontiguousarray(recvoffsets, dtype='i8')
        if skip_self:
            sc[self.comm.rank] = 0
            rc[self.comm.rank] = 0
        _lib.check(ctx.lib.pmb_alltoallv(ctx.handle, send.ptr, sc.ctypes.data, so.ctypes.data,
                                         recv.ptr, rc.ctypes.data, ro.ctypes.data, int(itemsize)))
        return recv

    def _remote_offsets(self, offsets, selfcount):
        """offsets of the per-rank segments in a buffer that leaves the segment of this rank out"""
        off = numpy.array(offsets, dtype='i8')
        off[self.comm.rank + 1:] -= int(selfcount)
        off[self.comm.rank] = 0
        return off

    def _exchange(self, data):
        ddata, dtype, trailing, was_host = self._to_device_records(
            data, self.sendlength, 'the length of data does not match that used to build the layout')
        ctx = ddata.ctx
        itemsize = int(numpy.prod(trailing, dtype='i8')) * dtype.itemsize
        nsend = int(self.sendcounts.sum())
        me, P = self.comm.rank, self.comm.size
        if self.identity and P == 1:
            # take(arange) is the data itself: device arrays are handed on without a copy (a host
            # input was copied to the device just above, so the caller's array is never aliased)
            recv = ddata
        elif P == 1:
            recv = DeviceArray.empty((nsend, itemsize), 'u1')
            # buffer = data.take(indices, axis=0)  (domain.py:188)
            _lib.check(ctx.lib.pmb_take(ctx.handle, ddata.ptr, itemsize, self._indices_ptr(), nsend, recv.ptr))
        else:
            # records that stay on this rank are gathered straight into their place in the receive
            # buffer; only what really leaves is packed into `send` and goes through NCCL
            recv = DeviceArray.empty((int(self.recvlength), itemsize), 'u1')
            s0, sn = int(self.sendoffsets[me]), int(self.sendcounts[me])
            # the send buffer holds only what leaves: [records for ranks < me | records for ranks > me]
            send = DeviceArray.empty((max(nsend - sn, 1), itemsize), 'u1')
            soff = self._remote_offsets(self.sendoffsets, sn)
            ip = self._indices_ptr()
            for a, b, dst in ((0, s0, send.ptr), (s0, s0 + sn, recv.ptr + int(self.recvoffsets[me]) * itemsize),
                              (s0 + sn, nsend, send.ptr + s0 * itemsize)):
                if b > a:
                    src_idx = None if ip is None else ip + 4 * a
                    src = ddata.ptr + (a * itemsize if ip is None else 0)
                    _lib.check(ctx.lib.pmb_take(ctx.handle, src, itemsize, src_idx, b - a, dst))
            self._alltoallv(ctx, send, self.sendcounts, soff, recv,
                            self.recvcounts, self.recvoffsets, itemsize, skip_self=True)
        out = DeviceArray((int(self.recvlength),) + tuple(trailing), dtype, ptr=recv.ptr, base=recv, ctx=ctx)
        if was_host:
            return out.to_host()
        return out

    def exchange_remote(self, data):
        """
        The records OTHER ranks deliver to this rank, in rank order -- ``exchange(data)`` without the block this
        rank sends to itself (engine extension; device arrays only).

        For kernels that clip to the local canvas (every paint / readout of a decomposed mesh does) the block a
        rank keeps need not be gathered at all: painting ``data`` where it lies deposits exactly what the own
        block would -- a particle of ``data`` that does not touch the local domain is not in that block and
        deposits nothing here -- so ``paint(exchange(pos)) == paint(pos) + paint(exchange_remote(pos))`` and
        likewise for readout followed by :py:meth:`gather_add_ghosts`.  Only the few records that really change
        rank are packed and travel (domain.py:173-206 moves every record through ``take`` + ``Alltoallv``).
        """
        ddata, dtype, trailing, was_host = self._to_device_records(
            data, self.sendlength, 'the length of data does not match that used to build the layout')
        assert not was_host, "exchange_remote serves device arrays"
        ctx = ddata.ctx
        itemsize = int(numpy.prod(trailing, dtype='i8')) * dtype.itemsize
        me, P = self.comm.rank, self.comm.size
        rn, sn = int(self.recvcounts[me]), int(self.sendcounts[me])
        nrem = int(self.recvlength) - rn
        recv = DeviceArray.empty((max(nrem, 1), itemsize), 'u1')
        if P > 1:
            nsend = int(self.sendcounts.sum())
            s0 = int(self.sendoffsets[me])
            send = DeviceArray.empty((max(nsend - sn, 1), itemsize), 'u1')
            soff = self._remote_offsets(self.sendoffsets, sn)
            roff = self._remote_offsets(self.recvoffsets, rn)
            ip = self._indices_ptr()
            for a, b, dst in ((0, s0, send.ptr), (s0 + sn, nsend, send.ptr + s0 * itemsize)):
                if b > a:
                    src_idx = None if ip is None else ip + 4 * a
                    src = ddata.ptr + (a * itemsize if ip is None else 0)
                    _lib.check(ctx.lib.pmb_take(ctx.handle, src, itemsize, src_idx, b - a, dst))
            self._alltoallv(ctx, send, self.sendcounts, soff, recv, self.recvcounts, roff, itemsize, skip_self=True)
        return DeviceArray((nrem,) + tuple(trailing), dtype, ptr=recv.ptr, base=recv, ctx=ctx)

    # ------------------------------------------------------------------ fused readout + gather
    def fused_gather_plan(self):
        """(own_begin, own_count, pointer to the indices of the own block) for kernels that write the results of
        this rank's own particles straight into the gathered columns; None when there is nothing to fuse"""
        if self.identity or self.comm.size == 1:
            return None
        me = self.comm.rank
        return (int(self.recvoffsets[me]), int(self.recvcounts[me]),
                self.indices_device.ptr + 4 * int(self.sendoffsets[me]))

    def gather_add_ghosts(self, ghost_cols, own_cols):
        """second half of the fused ghost sum: ``ghost_cols`` (compact float64 columns of the ghosts this rank
        holds for others, the own block cut out) travel back through the reverse alltoallv and are added,
        in rank order, to ``own_cols`` (sendlength rows), which already hold the own results"""
        me, P = self.comm.rank, self.comm.size
        rn, sn = int(self.recvcounts[me]), int(self.sendcounts[me])
        nback = int(self.sendcounts.sum())
        roff = self._remote_offsets(self.recvoffsets, rn)
        boff = self._remote_offsets(self.sendoffsets, sn)
        offs = numpy.zeros(P + 1, dtype='i8')
        offs[1:] = numpy.cumsum(self.sendcounts)
        for g, o in zip(ghost_cols, own_cols):
            ctx = g.ctx
            back = DeviceArray.empty((max(nback - sn, 1), 8), 'u1')
            self._alltoallv(ctx, g, self.recvcounts, roff, back, self.sendcounts, boff, 8, skip_self=True)
            segs = (ctypes.c_void_p * P)()
            for q in range(P):
                segs[q] = back.ptr + int(boff[q]) * 8
            _lib.check(ctx.lib.pmb_gather_add_segments(ctx.handle, segs, 8, 1, self.indices_device.ptr,
                                                       offs.ctypes.data, P, me, int(self.sendlength), o.ptr))
        return own_cols

    # ------------------------------------------------------------------ gather
    def gather(self, data, mode='sum', out=None):
        """
        Pull the data from other ranks back to its original hosting rank (reference domain.py:208-318).

        mode : 'sum' (ghosts reduced with +, accumulated in float64 in the order of ``indices``),
               'all', 'local', 'any', 'mean', or a numpy ufunc.
        """
        if mode == 'local':
            # drop all ghosts: no communication (domain.py:245-262)
            host = numpy.asarray(data.to_host() if is_device(data) else data)
            host = promote(host, self.comm)
            if any(self.comm.allgather(len(host) != self.recvlength)):
                raise ValueError('the length of data does not match result of a domain.exchange')
            self.comm.Barrier()
            dtype = numpy.dtype((host.dtype, host.shape[1:]))
            if out is None:
                out = numpy.empty(self.sendlength, dtype=dtype)
            start2 = self.sendoffsets[self.comm.rank]
            ind = self.indices[start2:start2 + self.sendcounts[self.comm.rank]]
            start1 = self.recvoffsets[self.comm.rank]
            out[ind] = host[start1:start1 + self.recvcounts[self.comm.rank]]
            return out

        ddata, dtype, trailing, was_host = self._to_device_records(
            data, self.recvlength, 'the length of data does not match result of a domain.exchange')
        ctx = ddata.ctx
        ncomp = int(numpy.prod(trailing, dtype='i8'))
        itemsize = ncomp * dtype.itemsize
        nback = int(self.sendcounts.sum())
        # reverse Alltoallv: what I received goes back to its origin (domain.py:274-281); the ghosts
        # this rank holds of its own particles stay where they are (`segments` below)
        me, P = self.comm.rank, self.comm.size
        if P == 1:
            back = ddata
        else:
            # only the ghosts other ranks hold come back; the block of this rank is read where it lies
            sn = int(self.sendcounts[me])
            back = DeviceArray.empty((max(nback - sn, 1), max(itemsize, 1)), 'u1')
            boff = self._remote_offsets(self.sendoffsets, sn)
            back = self._alltoallv(ctx, ddata, self.recvcounts, self.recvoffsets, back,
                            self.sendcounts, boff, itemsize, skip_self=True)
        full_dtype = numpy.dtype((dtype, tuple(trailing)))

        if self.sendlength == 0:
            if out is None:
                out = numpy.empty(self.sendlength, dtype=full_dtype)
            return out

        if mode == 'sum' and self.identity and P == 1 and out is None and not was_host:
            # nothing to reduce and nowhere to move: the device column is the result (bincount would
            # only turn -0.0 into +0.0)
            return ddata

        if mode == 'sum' and dtype in (numpy.dtype('f4'), numpy.dtype('f8')):
            # bincountv(indices, recvbuffer, minlength=sendlength) on the device
            if is_device(out):
                dout = out
            else:
                odt = dtype if out is None else (out.dtype if out.dtype in (numpy.dtype('f4'), numpy.dtype('f8')) else numpy.dtype('f8'))
                dout = DeviceArray.empty((int(self.sendlength),) + tuple(trailing), odt)
            offs = numpy.zeros(self.comm.size + 1, dtype='i8')
            offs[1:] = numpy.cumsum(self.sendcounts)
            segs = (ctypes.c_void_p * P)()
            if P > 1:
                for q in range(P):
                    segs[q] = back.ptr + int(boff[q]) * itemsize
                segs[me] = ddata.ptr + int(self.recvoffsets[me]) * itemsize
            else:
                segs[0] = back.ptr
            _lib.check(ctx.lib.pmb_gather_sum_segments(ctx.handle, segs, dtype.itemsize, ncomp, self._indices_ptr(),
                                                       offs.ctypes.data, P, int(self.sendlength),
                                                       dout.ptr, dout.dtype.itemsize))
            if is_device(out):
                return out
            if out is None:
                return dout.to_host() if was_host else dout
            out[...] = dout.to_host()
            return out

        # the remaining modes are host-side bookkeeping on the returned ghosts (never on the force path):
        # assemble the full returned buffer (rank order), the block of this rank from where it lies
        if P > 1:
            full = DeviceArray.empty((nback, max(itemsize, 1)), 'u1')
            s0 = int(self.sendoffsets[me])
            if s0 > 0:
                ctx.d2d(full.ptr, back.ptr, s0 * itemsize)
            if sn > 0:
                ctx.d2d(full.ptr + s0 * itemsize, ddata.ptr + int(self.recvoffsets[me]) * itemsize, sn * itemsize)
            if nback - s0 - sn > 0:
                ctx.d2d(full.ptr + (s0 + sn) * itemsize, back.ptr + s0 * itemsize, (nback - s0 - sn) * itemsize)
            back = full
        recvbuffer = DeviceArray((nback,) + tuple(trailing), dtype, ptr=back.ptr, base=back, ctx=ctx).to_host()
        indices = self.indices
        if mode == 'all':
            if out is None:
                return recvbuffer
            out[...] = recvbuffer
            return out
        if mode == 'sum':
            return bincountv(indices, recvbuffer, minlength=self.sendlength, out=out)
        if isinstance(mode, numpy.ufunc):
            arg = indices.argsort()
            recvbuffer = recvbuffer[arg]
            N = numpy.bincount(indices, minlength=self.sendlength)
            offset = numpy.zeros(self.sendlength, 'intp')
            offset[1:] = numpy.cumsum(N)[:-1]
            return mode.reduceat(recvbuffer, offset, out=out)
        if mode == 'mean':
            N = numpy.bincount(indices, minlength=self.sendlength)
            s = [self.sendlength] + [1] * (len(recvbuffer.shape) - 1)
            N = N.reshape(s)
            out = bincountv(indices, recvbuffer, minlength=self.sendlength, out=out)
            out[...] /= N
            return out
        if mode == 'any':
            if out is None:
                out = numpy.zeros(self.sendlength, dtype=full_dtype)
            out[indices] = recvbuffer
            return out
        raise NotImplementedError


class GridND(object):
    """
    GridND is domain decomposition on a uniform grid of N dimensions (reference domain.py:320-652).

    The total number of domains is prod([ len(dir) - 1 for dir in edges]).

    Attributes
    ----------
    edges   : list  (Ndim)
        edges[i] is the edges on direction i, including 0 and BoxSize.
    comm   : communicator (default: the world communicator)
    periodic : boolean
        if periodic, edges[i][-1] is the period.
    """
    @classmethod
    def uniform(cls, BoxSize, comm=None, periodic=True):
        if comm is None:
            comm = _comm.world()
        ndim = len(BoxSize)
        # compute a optimal shape where each domain is as cubical as possible
        r = (1.0 * comm.size / numpy.prod(BoxSize) * min(BoxSize)) ** (1.0 / ndim)
        shape = [r * (BoxSize[i] / min(BoxSize)) for i in range(ndim)]
        shape = numpy.array(shape)
        imax = shape.argmax()
        shape = numpy.int32(shape)
        shape[shape < 1] = 1
        shape[imax] = 1
        shape[imax] = comm.size // numpy.prod(shape)
        assert numpy.prod(shape) <= comm.size
        edges = []
        for i in range(ndim):
            edges.append(numpy.linspace(0, BoxSize[i], shape[i] + 1, endpoint=True))
        return cls(edges, comm, periodic)

    def __init__(self, edges, comm=None, periodic=True, DomainAssign=None):
        """ DomainAssign records each domain is assigned to which rank """
        if comm is None:
            comm = _comm.world()
        self.shape = numpy.array([len(g) - 1 for g in edges], dtype='int32')
        self.ndim = len(self.shape)
        self.edges = [numpy.asarray(g) for g in edges]
        self.periodic = periodic
        self.comm = comm
        self.size = numpy.prod(self.shape)

        if DomainAssign is None:
            if comm.size >= self.size:
                DomainAssign = numpy.array(range(self.size), dtype='int32')
            else:
                DomainAssign = numpy.empty(self.size, dtype='int32')
                for i in range(comm.size):
                    start = i * self.size // comm.size
                    end = (i + 1) * self.size // comm.size
                    DomainAssign[start:end] = i
        self.DomainAssign = DomainAssign

        dd = numpy.zeros(self.shape, dtype='int16')
        for i, edge in enumerate(edges):
            edge = numpy.array(edge)
            dd1 = edge[1:] == edge[:-1]
            dd1 = dd1.reshape([-1 if ii == i else 1 for ii in range(self.ndim)])
            dd[...] |= dd1
        self.DomainDegenerate = dd.ravel()

        self._update_primary_regions()

    # ------------------------------------------------------------------ host-side helpers
    @staticmethod
    def _digitize(data, bins, right=False):
        if len(data) == 0:
            return numpy.empty((0), dtype='intp')
        return numpy.digitize(data, bins, right)

    def load(self, pos, transform=None, gamma=2):
        """ load of each domain, assuming a power law N^gamma of the particle count (domain.py:409-466) """
        pos = numpy.asarray(pos.to_host() if is_device(pos) else pos)
        assert pos.shape[1] >= self.ndim
        if transform is None:
            transform = lambda x: x
        Npoint = len(pos)
        if Npoint != 0:
            sil = numpy.empty((self.ndim, Npoint), dtype='i2', order='C')
            chunk = transform(pos)
            for j in range(self.ndim):
                if self.periodic:
                    tmp = numpy.remainder(chunk[:, j], self.edges[j][-1])
                else:
                    tmp = chunk[:, j]
                sil[j, :] = self._digitize(tmp, self.edges[j]) - 1
            mode = 'raise' if self.periodic else 'clip'
            particle_domain = numpy.ravel_multi_index(sil, self.shape, mode=mode)
            tmp = numpy.bincount(particle_domain, minlength=self.size)
        else:
            tmp = numpy.zeros(self.size)
        domainload = self.comm.allreduce(tmp)
        return domainload ** gamma

    def loadbalance(self, domainload):
        """ greedy balancing of the ranks given the load of each domain; updates DomainAssign
            (domain.py:468-500) """
        if self.size <= self.comm.size:
            return
        domains = sorted([(domainload[i], i) for i in range(self.size)], reverse=True)
        processes = [(0, i) for i in range(self.comm.size)]
        heapq.heapify(processes)
        for dload, dindex in domains:
            pload, rank = heapq.heappop(processes)
            pload += dload
            self.DomainAssign[dindex] = rank
            heapq.heappush(processes, (pload, rank))
        self._update_primary_regions()

    def _update_primary_regions(self):
        my_domains = numpy.where(self.DomainAssign == self.comm.rank)[0]
        N = len(my_domains)
        if N == 0:
            primary_region = None
        else:
            primary_region = {}
            primary_region['start'] = numpy.empty((N, self.ndim))
            primary_region['end'] = numpy.empty((N, self.ndim))
            for i in range(N):
                domain_index = numpy.unravel_index(my_domains[i], self.shape, order='C')
                primary_region['start'][i] = numpy.array([g[r] for g, r in zip(self.edges, domain_index)])
                primary_region['end'][i] = numpy.array([g[r + 1] for g, r in zip(self.edges, domain_index)])
        self.primary_region = primary_region

    def isprimary(self, pos, transform=None):
        """ True where the position falls into the primary region of the current rank (domain.py:519-559) """
        pos = numpy.asarray(pos.to_host() if is_device(pos) else pos)
        if self.primary_region is None:
            return numpy.zeros(len(pos), dtype='?')
        if transform is None:
            transform = lambda x: x
        r = numpy.zeros(len(pos), dtype='?')
        x0 = self.primary_region['start']
        x1 = self.primary_region['end']
        BoxSize = numpy.array([self.edges[j][-1] for j in range(self.ndim)])
        chunk = transform(pos)[..., :self.ndim]
        if self.periodic:
            chunk = numpy.remainder(chunk, BoxSize)
        for j in range(len(x0)):
            r[:] += ((chunk >= x0[j]) & (chunk < x1[j])).all(axis=-1)
        return r

    # ------------------------------------------------------------------ the hot path
    def decompose(self, pos, smoothing=0, transform=None):
        """
        Decompose particles into domains: returns the :py:class:`Layout` that routes every particle to
        all ranks whose domain its smoothing patch intersects (reference domain.py:561-652).

        pos       : (N, >=ndim) positions, numpy or DeviceArray (float32/float64)
        smoothing : float or per-dimension array, in the coordinate system of the edges
        transform : None, a ``ScaleTransform`` (applied on the GPU), or any callable
                    ``transform(pos) -> domain_pos`` (evaluated on the host, then routed on the GPU)
        """
        assert len(pos) < 1024 * 1024 * 1024 * 2
        _smoothing = smoothing
        smoothing = numpy.empty(self.ndim, dtype='f8')
        smoothing[:] = _smoothing
        if self.ndim > 3:
            raise _lib.PmbError("GPU routing supports domain grids of up to 3 dimensions")

        scale = numpy.ones(self.ndim, dtype='f8')
        if transform is None:
            pass
        elif isinstance(transform, ScaleTransform):
            scale[:] = numpy.broadcast_to(transform.scale, (len(scale),)) if transform.scale.ndim == 0 \
                else transform.scale[:self.ndim]
        else:
            host = numpy.asarray(pos.to_host() if is_device(pos) else pos)
            pos = numpy.asarray(transform(host))

        if is_device(pos):
            dpos = pos
        else:
            hp = numpy.asarray(pos)
            if hp.dtype not in (numpy.dtype('f4'), numpy.dtype('f8')):
                hp = hp.astype('f8')
            if hp.ndim != 2:
                hp = hp.reshape(len(hp), -1)
            dpos = DeviceArray.from_host(hp)
        assert dpos.shape[1] >= self.ndim
        Npoint = dpos.shape[0]
        ctx = dpos.ctx

        a = _lib.DecomposeArgs()
        a.pos = dpos.ptr
        a.pos_elsize = dpos.dtype.itemsize
        a.npart = Npoint
        a.pos_stride0, a.pos_stride1 = dpos.strides
        a.ndim = self.ndim
        edges = numpy.ascontiguousarray(numpy.concatenate([numpy.asarray(e, dtype='f8') for e in self.edges]))
        for d in range(self.ndim):
            a.scale[d] = scale[d]
            a.smoothing[d] = smoothing[d]
            a.nedges[d] = len(self.edges[d])
        a.edges_h = edges.ctypes.data
        a.periodic = int(bool(self.periodic))
        assign = numpy.ascontiguousarray(self.DomainAssign, dtype='int32')
        degenerate = numpy.ascontiguousarray(self.DomainDegenerate, dtype='int16')
        a.domain_assign_h = assign.ctypes.data
        a.domain_degenerate_h = degenerate.ctypes.data
        a.nranks = self.comm.size

        counts = numpy.zeros(self.comm.size, dtype='int32')
        ntotal = ctypes.c_int64(0)
        _lib.check(ctx.lib.pmb_decompose_count(ctx.handle, ctypes.byref(a), counts.ctypes.data, ctypes.byref(ntotal)))
        ident = ctypes.c_int(0)
        _lib.check(ctx.lib.pmb_decompose_identity(ctx.handle, ctypes.byref(ident)))
        if ident.value:
            return Layout(comm=self.comm, sendlength=Npoint, sendcounts=counts, indices=None, identity=True)
        indices = DeviceArray.empty((int(ntotal.value),), 'int32')
        _lib.check(ctx.lib.pmb_decompose_fill(ctx.handle, ctypes.byref(a), indices.ptr))

        return Layout(comm=self.comm, sendlength=Npoint, sendcounts=counts, indices=indices)
