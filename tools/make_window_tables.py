"""Generate ``pmesh_b200/data/window_tables.npz`` -- the lookup tables of the
table-driven resampling windows (lanczos2-6, acg2-6, db6/12/20, sym6/12/20).

Why a data file: the reference evaluates these windows by linear interpolation
in tables printed to 8 decimals with a 7-significant-digit ``step`` literal
(reference ``pmesh/_window_lanczos.h:2058-2084``, ``_window_acg.h``,
``_window_wavelets.h:460-475``).  Parity requires the *same rounded numbers*,
not a re-derivation of sinc(): so the tables are data, the same way golden
vectors are.

* lanczosN / acgN are RE-GENERATED here from their defining formulas
  (restating reference ``makelanczos.py:3-8`` and ``makeacg.py:4-24``) and
  rounded through the same ``%.8f`` / ``%e`` text formats.  When
  ``/root/reference`` is present the result is compared value-by-value with
  the numbers in the reference headers and the script aborts on any mismatch.
* db*/sym* need PyWavelets' cascade output (``makewavelets.py:4-22``); pywt is
  not installed in this image, so those six tables are parsed out of the
  reference header as numeric data (values only; no code is taken).

Run:  python tools/make_window_tables.py   (needs /root/reference for the
wavelet tables and for the cross-check; the committed .npz is the artefact).
"""
import os
import re
import sys

import numpy

REF = os.environ.get("PMESH_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pmesh_b200", "data",
                   "window_tables.npz")

NTAB = 8192


def _round8(a):
    """numbers exactly as the C compiler parses the '%.8f' literals"""
    return numpy.array([float("%.8f" % v) for v in a], dtype="f8")


def _trapz(y, x):
    return float(numpy.sum((y[1:] + y[:-1]) * numpy.diff(x)) * 0.5) if not hasattr(numpy, "trapezoid") \
        else float(numpy.trapezoid(y, x))


def lanczos_table(n):
    x = numpy.linspace(0, n, NTAB, endpoint=False)
    phi = numpy.sinc(x) * numpy.sinc(x / n)
    phi = phi / (2 * _trapz(phi, x))
    step = float("%e" % numpy.diff(x).mean())
    return _round8(phi), step, 2 * n


def acg_table(n):
    a = (n - 1) / 2.0
    x = numpy.linspace(0, n * 0.5, NTAB, endpoint=True)
    y = x + a

    def g(t):
        return numpy.exp(-0.25 * (t - a) ** 2)

    phi = g(y) - g(-0.5) * (g(y + n) + g(y - n)) / (g(-0.5 + n) + g(-0.5 - n))
    phi = phi / (2 * _trapz(phi, x))
    step = float("%e" % numpy.diff(x).mean())
    return _round8(phi), step, n


_TABLE_RE = re.compile(
    r"static double _(\w+?)_v?table\[\] = \{(.*?)\};\s*"
    r"static double _\w+_nativesupport = ([0-9.eE+-]+);(.*?)static double _\w+_diff", re.S)


def parse_header(path):
    """name -> (values, step, nativesupport, hsupport-or-0) from a generated reference header"""
    txt = open(path).read()
    out = {}
    for m in _TABLE_RE.finditer(txt):
        name, body, support, kernel = m.group(1), m.group(2), float(m.group(3)), m.group(4)
        vals = numpy.array([float(t) for t in body.replace("\n", " ").split(",") if t.strip()], dtype="f8")
        step = float(re.search(r"double f = x / ([0-9.eE+-]+);", kernel).group(1))
        hs = re.search(r"x \+= ([0-9.eE+-]+);", kernel)
        out[name] = (vals, step, support, float(hs.group(1)) if hs else 0.0)
    return out


def main():
    tables = {}
    for n in range(2, 7):
        tables["lanczos%d" % n] = lanczos_table(n) + (0.0,)
        tables["acg%d" % n] = acg_table(n) + (0.0,)

    have_ref = os.path.isdir(os.path.join(REF, "pmesh"))
    if not have_ref:
        print("no reference tree: cannot obtain the wavelet tables; keeping the committed file")
        return 1
    ref = {}
    for h in ("_window_lanczos.h", "_window_acg.h", "_window_wavelets.h"):
        ref.update(parse_header(os.path.join(REF, "pmesh", h)))
    bad = 0
    for name, (vals, step, support, hs) in sorted(tables.items()):
        rv, rstep, rsup, rhs = ref[name]
        same = len(rv) == len(vals) and numpy.array_equal(rv, vals) and rstep == step and rsup == support
        print("%-9s regenerated: %d entries step=%.6e support=%g  %s" %
              (name, len(vals), step, support, "== reference header" if same else "MISMATCH"))
        bad += not same
    if bad:
        print("regenerated tables differ from the reference headers")
        return 2
    for name in ("db6", "db12", "db20", "sym6", "sym12", "sym20"):
        tables[name] = ref[name]
        print("%-9s from header data: %d entries step=%.6e support=%g hsupport=%g" %
              (name, len(ref[name][0]), ref[name][1], ref[name][2], ref[name][3]))

    save = {}
    for name, (vals, step, support, hs) in tables.items():
        save[name] = vals
        save[name + "_meta"] = numpy.array([step, support, hs], dtype="f8")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    numpy.savez_compressed(OUT, **save)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
