"""The C-ABI shared library loads without a GPU and exports every symbol include/pmesh_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pmesh_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pmb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from pmesh_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) > 40
    for name in names:
        assert hasattr(lib, name), "libpmesh_b200.so does not export %s" % name
    # and the ctypes binding knows all of them
    assert sorted(_lib.SYMBOLS) == names


def test_struct_layouts_match_the_header():
    """sizeof of the ctypes mirrors == sizeof of the C structs, as compiled by gcc from the header"""
    import subprocess
    import tempfile
    from pmesh_b200 import _lib
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "sz.c")
        with open(src, "w") as f:
            f.write('#include <stdio.h>\n#include "pmesh_b200.h"\nint main(){printf("%zu %zu\\n", '
                    'sizeof(pmb_resample_args), sizeof(pmb_decompose_args));return 0;}\n')
        exe = os.path.join(tmp, "sz")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        a, b = [int(x) for x in subprocess.check_output([exe]).split()]
    assert ctypes.sizeof(_lib.ResampleArgs) == a
    assert ctypes.sizeof(_lib.DecomposeArgs) == b


def test_window_queries_run_on_the_host():
    from pmesh_b200 import window
    assert window.CIC.support == 2 and window.TSC.support == 3 and window.PCS.support == 4
    assert window.LANCZOS2.support == 4 and window.LANCZOS3.support == 6 and window.ACG3.support == 3
    assert window.DB12.support == 10 and window.DB20.support == 13 and window.SYM20.support == 12
    assert window.ResampleWindow("linear", 4).support == 4
    assert window.CUBIC.resize(8).support == 8 and window.CUBIC.resize(8).nativesupport == 4
    assert window.windows["cic"] is window.CIC and window.methods is window.windows
    assert window.FindResampler("tsc") is window.TSC
    with pytest.raises(TypeError):
        window.FindResampler(3)
    assert len(window.windows) == 48


def test_no_silent_fallback_without_gpu():
    """without a CUDA device the product must raise, not compute on the CPU"""
    from pmesh_b200 import _lib
    n = ctypes.c_int(0)
    rc = _lib.load().pmb_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    import numpy
    from pmesh_b200 import window
    with pytest.raises(_lib.PmbError):
        window.CIC.paint(numpy.zeros((4, 4)), [[1.0, 1.0]])


def test_header_is_plain_c_and_cites_the_reference():
    """include/pmesh_b200.h compiles as C99 (no C++, no CUDA, no torch types in the boundary) and every
    entry point group cites the reference interface it replaces (file:line)"""
    import subprocess
    hdr = os.path.join(ROOT, "include", "pmesh_b200.h")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    txt = open(hdr).read()
    code = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    assert "torch" not in code.lower() and "tensor" not in code.lower()
    for cite in ("pmesh/_window.pyx:", "pmesh/_window_imp.h:", "pmesh/domain.py:", "pmesh/_domain.pyx:", "pmesh/pm.py:",
                 "examples/nbody.py:", "pmesh/_whitenoise.pyx:"):
        assert cite in txt, cite


def test_every_exported_entry_point_is_declared():
    """the other direction: no pmb_* function is exported that the header does not declare"""
    import subprocess
    so = os.path.join(ROOT, "pmesh_b200", "csrc", "libpmesh_b200.so")
    out = subprocess.check_output(["nm", "-D", "--defined-only", so], text=True)
    exported = sorted(set(l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("pmb_")))
    internal = {"pmb_set_error", "pmb_cuda_fail", "pmb_scratch", "pmb_resolve_window", "pmb_stream_barrier", "pmb_allgather_host"}
    undeclared = [n for n in exported if n not in declared_symbols() and n not in internal]
    assert not undeclared, undeclared
