import ctypes
import os
import subprocess
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _have_gpu():
    try:
        from pmesh_b200 import _lib
        n = ctypes.c_int(0)
        return _lib.load().pmb_device_count(ctypes.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle as _oracle
    _oracle.build()
    return _oracle


@pytest.fixture(scope="session")
def ref():
    """the compiled reference (oracle/_ref): (_window.ResampleWindow subclass factory, _domain) or skip"""
    import build_ref
    if not build_ref.have_ref():
        try:
            build_ref.build()
        except Exception:
            pass
    mods = build_ref.load()
    if mods is None:
        pytest.skip("oracle/_ref not available (no /root/reference and no prebuilt files)")
    w, d = mods

    class RW(w.ResampleWindow):
        pass
    return RW, d


@pytest.fixture(scope="session")
def harness():
    """host build of the device stencil code (tests/harness/host_harness.cpp)"""
    src = os.path.join(ROOT, "tests", "harness", "host_harness.cpp")
    so = os.path.join(ROOT, "tests", "harness", "libhost_harness.so")
    deps = [src] + [os.path.join(ROOT, "pmesh_b200", "csrc", f) for f in ("pmb_window.h", "pmb_stencil.cuh", "pmb_internal.h", "pmb_wnrng.h", "pmb_route.h", "pmb_ifft.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(f) for f in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared",
                               "-I/usr/local/cuda/include", "-o", so, src])
    lib = ctypes.CDLL(so)
    z = numpy.load(os.path.join(ROOT, "pmesh_b200", "data", "window_tables.npz"))
    from pmesh_b200.window import KINDS
    for name in z.files:
        if name.endswith("_meta"):
            continue
        vals = numpy.ascontiguousarray(z[name], dtype="f8")
        step, support, hs = [float(v) for v in z[name + "_meta"]]
        lib.hh_set_table(ctypes.c_int(KINDS[name]), ctypes.c_void_p(vals.ctypes.data), ctypes.c_int(len(vals)),
                         ctypes.c_double(step), ctypes.c_double(hs))
    return lib
