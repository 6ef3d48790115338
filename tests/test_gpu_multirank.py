"""NCCL legs on >= 2 GPUs of one node (skipped on single-GPU boxes): launches tests/multirank_worker.py
with one process per GPU through torch.distributed.run."""
import ctypes
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    from pmesh_b200 import _lib
    n = ctypes.c_int(0)
    if _lib.load().pmb_device_count(ctypes.byref(n)) != 0:
        return 0
    return n.value


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_multirank(nproc):
    if _ngpu() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multirank_worker.py")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-6000:]
    assert p.stdout.count("multirank ok") == 4
