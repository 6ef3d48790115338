"""Communicator facade: the subset of mpi4py's ``Comm`` that pmesh uses.

The reference talks to MPI through mpi4py (call sites: SURVEY section 2.2).  Here
there is one process per GPU, launched by ``torchrun`` (or any launcher that sets
RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT):

* small host-side collectives (counts, scalars, shapes: ``allreduce``,
  ``allgather``, ``bcast``, ``Alltoall``, ``Barrier``) go through
  ``torch.distributed`` with the ``gloo`` backend -- plumbing only;
* particle payloads and FFT transposes never touch the host: they go through
  the library's own NCCL communicator (``pmb_alltoallv`` over NVLink), which is
  bootstrapped here by broadcasting the ``ncclUniqueId`` over gloo.
"""
import ctypes
import os

import numpy

SUM, MAX, MIN = "sum", "max", "min"


class _Op(object):
    SUM = SUM
    MAX = MAX
    MIN = MIN


def _reduce(values, op):
    out = values[0]
    for v in values[1:]:
        if op == SUM:
            out = out + v
        elif op == MAX:
            out = numpy.maximum(out, v)
        elif op == MIN:
            out = numpy.minimum(out, v)
        else:
            raise ValueError("unknown reduction %r" % (op,))
    return out


class SelfComm(object):
    """Single-process communicator (size 1)."""
    rank = 0
    size = 1

    def Barrier(self):
        pass

    def allreduce(self, x, op=SUM):
        return x

    def allgather(self, x):
        return [x]

    def bcast(self, x, root=0):
        return x

    def Alltoall(self, send, recv):
        recv[...] = send

    def Allreduce_inplace(self, array, op=SUM):
        return array

    # device payload path
    def ensure_device_comm(self, ctx):
        return

    def __repr__(self):
        return "SelfComm()"


class TorchComm(object):
    """World communicator over torch.distributed (gloo) + the library's NCCL communicator."""

    def __init__(self):
        import torch.distributed as dist
        self._dist = dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group(backend="gloo")
        self._group = dist.new_group(backend="gloo") if dist.get_backend() != "gloo" else None
        self.rank = dist.get_rank()
        self.size = dist.get_world_size()
        self._nccl_ready = False

    def Barrier(self):
        self._dist.barrier(group=self._group)

    def allgather(self, x):
        out = [None] * self.size
        self._dist.all_gather_object(out, x, group=self._group)
        return out

    def allreduce(self, x, op=SUM):
        # gathered then reduced in rank order on every rank: deterministic and identical everywhere
        return _reduce(self.allgather(x), op)

    def bcast(self, x, root=0):
        box = [x]
        self._dist.broadcast_object_list(box, src=root, group=self._group)
        return box[0]

    def Alltoall(self, send, recv):
        """one item per rank (the sendcounts -> recvcounts transpose of Layout).  Whether the backend offers
        all_to_all_single is found out ONCE and agreed between the ranks (a fallback that only some ranks take would
        leave mismatched collectives behind); from then on every rank takes the same path."""
        import torch
        send = numpy.ascontiguousarray(send)
        if getattr(self, "_a2a_single", None) is None:
            ok = True
            try:
                tin = torch.zeros(self.size, dtype=torch.int64)
                tout = torch.empty_like(tin)
                self._dist.all_to_all_single(tout, tin, group=self._group)
            except (RuntimeError, NotImplementedError):
                ok = False
            self._a2a_single = all(self.allgather(ok))
        if self._a2a_single:
            tin = torch.from_numpy(send.astype("int64") if send.dtype.kind in "iu" else send.astype("float64"))
            tout = torch.empty_like(tin)
            self._dist.all_to_all_single(tout, tin, group=self._group)
            recv[...] = tout.numpy().astype(recv.dtype)
        else:
            rows = self.allgather(send.copy())
            for r in range(self.size):
                recv[r] = rows[r][self.rank]

    def Allreduce_inplace(self, array, op=SUM):
        array[...] = self.allreduce(numpy.array(array), op)
        return array

    def ensure_device_comm(self, ctx):
        """create the NCCL communicator of ``ctx`` (collective; first device payload triggers it)"""
        if self._nccl_ready:
            return
        from . import _lib
        uid = ctypes.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(ctx.lib.pmb_comm_unique_id(uid))
        raw = self.bcast(bytes(uid.raw), root=0)
        uid = ctypes.create_string_buffer(raw, 128)
        _lib.check(ctx.lib.pmb_comm_init_rank(ctx.handle, uid, self.rank, self.size))
        self._nccl_ready = True

    def __repr__(self):
        return "TorchComm(rank=%d, size=%d)" % (self.rank, self.size)


_world = None


def world():
    """COMM_WORLD: TorchComm when launched with WORLD_SIZE > 1, SelfComm otherwise."""
    global _world
    if _world is None:
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            _world = TorchComm()
        else:
            _world = SelfComm()
    return _world


class MPI(object):
    """Tiny stand-in namespace so code written as ``MPI.COMM_WORLD`` / ``MPI.SUM`` keeps working."""
    SUM = SUM
    MAX = MAX
    MIN = MIN

    class _World(object):
        def __get__(self, obj, objtype=None):
            return world()
    COMM_WORLD = _World()
