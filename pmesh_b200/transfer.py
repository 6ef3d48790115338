"""Recognised k-space transfer functions.

The reference's ``Field.apply(func)`` (pmesh/pm.py:617-648) evaluates an arbitrary
Python callable slab by slab on the host.  A GPU engine cannot run arbitrary
Python, so the transfer functions of the PM force step are provided as objects:
they are *callables with the reference's ``func(k, v)`` signature* (so they also
work with any code that evaluates them on numpy slabs) and carry the kernel id
that ``ComplexField.apply`` dispatches to ``pmb_transfer`` on the device.

Formulas are those of the reference's canonical caller ``examples/nbody.py:154-181``
(and ``pmesh/transfer.py:75-112,208-240`` of the deprecated API).
"""
import numpy

from . import _lib


class Transfer(object):
    """base: kind id + direction + parameters of pmb_transfer; ``apply_kind`` is the kind of
    coordinates the equivalent python callable expects (pm.py:1047-1070)."""
    kind = None
    apply_kind = "wavenumber"
    direction = 0

    def params(self):
        return (0.0, 0.0, 0.0, 0.0)


class Scale(Transfer):
    kind = _lib.TF_SCALE

    def __init__(self, factor):
        self.factor = float(factor)

    def params(self):
        return (self.factor, 0.0, 0.0, 0.0)

    def __call__(self, k, v):
        return self.factor * v


class GravityFD4(Transfer):
    """i * kfinite_d / k^2 with the 4th-order finite-difference gradient -- force_transfer(d),
    examples/nbody.py:162-170."""
    kind = _lib.TF_GRAVITY_FD4

    def __init__(self, direction):
        self.direction = int(direction)

    def __call__(self, k, v):
        k2 = sum(ki ** 2 for ki in k)
        k2[k2 == 0] = 1.0
        C = (v.BoxSize / v.Nmesh)[self.direction]
        w = k[self.direction] * C
        kfinite = 1.0 / C * 1 / 6.0 * (8 * numpy.sin(w) - numpy.sin(2 * w))
        return 1j * kfinite / k2 * v


class GradientK(Transfer):
    """i * k_d / k^2 -- dx1_transfer(d), examples/nbody.py:154-160."""
    kind = _lib.TF_GRADIENT_K

    def __init__(self, direction):
        self.direction = int(direction)

    def __call__(self, k, v):
        k2 = sum(ki ** 2 for ki in k)
        k2[k2 == 0] = 1.0
        return 1j * k[self.direction] / k2 * v


class InverseLaplace(Transfer):
    """-1 / k^2 -- pot_transfer, examples/nbody.py:172-175."""
    kind = _lib.TF_INV_LAPLACE

    def __call__(self, k, v):
        k2 = sum(ki ** 2 for ki in k)
        k2[k2 == 0] = 1.0
        return -1. / k2 * v


class GaussianLowpass(Transfer):
    """exp(-k^2 r^2 / 2) -- lowpass_transfer(r), examples/nbody.py:177-181."""
    kind = _lib.TF_GAUSS_LOWPASS

    def __init__(self, r):
        self.r = float(r)

    def params(self):
        return (self.r, 0.0, 0.0, 0.0)

    def __call__(self, k, v):
        k2 = sum(ki ** 2 for ki in k)
        return numpy.exp(-0.5 * k2 * self.r ** 2) * v


class PowerLaw(Transfer):
    """|k|^p (0 at k = 0): shapes white noise into a power-law spectrum P(k) ~ k^(2p)."""
    kind = _lib.TF_POWERLAW

    def __init__(self, p):
        self.p = float(p)

    def params(self):
        return (self.p, 0.0, 0.0, 0.0)

    def __call__(self, k, v):
        k2 = sum(ki ** 2 for ki in k)
        with numpy.errstate(divide='ignore'):
            f = numpy.where(k2 == 0, 0.0, k2 ** (0.5 * self.p))
        return f * v


class GradientIK(Transfer):
    """i * k_d (plain spectral derivative)."""
    kind = _lib.TF_IK

    def __init__(self, direction):
        self.direction = int(direction)

    def __call__(self, k, v):
        return 1j * k[self.direction] * v


class Compensate(Transfer):
    """1 / prod_d fwindow(w_d): deconvolve the resampling window (window.py:65-80), kind='circular'."""
    kind = _lib.TF_COMPENSATE
    apply_kind = "circular"

    def __init__(self, resampler):
        from .window import FindResampler
        self.resampler = FindResampler(resampler)

    def params(self):
        return (float(self.resampler._kind), float(self.resampler.support), 0.0, 0.0)

    def __call__(self, w, v):
        tf = 1.0
        for wi in w:
            tf = tf * self.resampler.get_fwindow(wi)
        return v / tf


def find_transfer(func):
    """Transfer object behind ``func`` or None (plain python callable -> host compatibility path)."""
    if isinstance(func, Transfer):
        return func
    return getattr(func, "transfer", None)
