"""The field resampling / I/O-order operations, restating the reference's own tests:
pmesh/tests/test_pm.py:392-538 (sort/ravel, Fourier resample, upsample/downsample, cmean) and
:756-812 (ctranspose, preview).  They are compositions of the hot-path operators (decompose /
exchange / paint / readout / r2c / c2r), so these tests also exercise those on 2-D meshes."""
import numpy
import pytest
from numpy.testing import assert_array_equal, assert_allclose, assert_almost_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def PM():
    from pmesh_b200 import pm
    return pm


def test_sort_ravel_unravel(PM):
    pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[8, 6], dtype='f8')
    real = PM.RealField(pm)
    truth = numpy.arange(8 * 6)
    real[...] = truth.reshape(8, 6)[real.slices]
    unsorted = real.copy()
    with pytest.warns(DeprecationWarning):
        real.sort(out=Ellipsis)
    assert_array_equal(real.value.ravel(), truth)
    real.unravel(real)
    assert_array_equal(real, unsorted)
    complex = PM.ComplexField(pm)
    truth = numpy.arange(8 * 4)
    complex[...] = truth.reshape(8, 4)[complex.slices]
    complex.ravel(out=Ellipsis)
    assert_array_equal(complex.value.ravel(), truth)
    r2 = pm.unravel('real', numpy.arange(48.0))
    assert_array_equal(r2.value, numpy.arange(48.0).reshape(8, 6))


def _truth_fields(PM, zero_high):
    pm1 = PM.ParticleMesh(BoxSize=8.0, Nmesh=[8, 8], dtype='f8')
    pm2 = PM.ParticleMesh(BoxSize=8.0, Nmesh=[4, 4], dtype='f8')
    numpy.random.seed(3333)
    truth = numpy.fft.rfftn(numpy.random.normal(size=(8, 8)))
    complex1 = PM.ComplexField(pm1)
    for ind in numpy.ndindex(*complex1.cshape):
        complex1.csetitem(ind, truth[ind])
        if zero_high:
            if any(i == 4 for i in ind):
                complex1.csetitem(ind, 0)
            if any(i >= 2 and i < 7 for i in ind):
                complex1.csetitem(ind, 0)
    complex2 = PM.ComplexField(pm2)
    for ind in numpy.ndindex(*complex2.cshape):
        newind = tuple([i if i <= 2 else 8 - (4 - i) for i in ind])
        if any(i == 2 for i in ind):
            complex2.csetitem(ind, 0)
        else:
            complex2.csetitem(ind, truth[newind])
    return pm1, pm2, complex1, complex2


def test_fdownsample(PM):
    pm1, pm2, complex1, complex2 = _truth_fields(PM, False)
    assert_almost_equal(complex1[...], complex1.c2r().r2c()[...])
    tmpr = PM.RealField(pm2)
    tmp = PM.ComplexField(pm2)
    complex1.resample(tmp)
    assert_almost_equal(complex2[...], tmp[...], decimal=5)
    complex1.c2r().resample(tmp)
    assert_almost_equal(complex2[...], tmp[...], decimal=5)
    complex1.resample(tmpr)
    assert_almost_equal(tmpr.r2c()[...], tmp[...])
    complex1.c2r().resample(tmpr)
    assert_almost_equal(tmpr.r2c()[...], tmp[...])


def test_fupsample(PM):
    pm1, pm2, complex1, complex2 = _truth_fields(PM, True)
    assert_almost_equal(complex1[...], complex1.c2r().r2c()[...])
    tmpr = PM.RealField(pm1)
    tmp = PM.ComplexField(pm1)
    complex2.resample(tmp)
    assert_almost_equal(complex1[...], tmp[...], decimal=5)
    complex2.c2r().resample(tmp)
    assert_almost_equal(complex1[...], tmp[...], decimal=5)
    complex2.resample(tmpr)
    assert_almost_equal(tmpr.r2c()[...], tmp[...])
    complex2.c2r().resample(tmpr)
    assert_almost_equal(tmpr.r2c()[...], tmp[...])


def test_real_resample_conserves_mass(PM):
    pmh = PM.ParticleMesh(BoxSize=8.0, Nmesh=[8, 8], dtype='f8')
    pml = PM.ParticleMesh(BoxSize=8.0, Nmesh=[4, 4], dtype='f8')
    reall = pml.create(type='real')
    reall.apply(lambda i, v: (i[0] % 2) * (i[1] % 2), kind='index', out=Ellipsis)
    for resampler in ['nearest', 'cic', 'tsc', 'cubic']:
        realh = pmh.upsample(reall, resampler=resampler, keep_mean=False)
        reall2 = pml.downsample(realh, resampler=resampler)
        assert_almost_equal(reall.csum(), realh.csum())
        assert_almost_equal(reall.csum(), reall2.csum())
    # three dimensions, different factors per axis
    pmh = PM.ParticleMesh(BoxSize=[8.0, 4.0, 6.0], Nmesh=[16, 8, 12], dtype='f8')
    pml = PM.ParticleMesh(BoxSize=[8.0, 4.0, 6.0], Nmesh=[8, 4, 6], dtype='f8')
    reall = pml.create(type='real', value=numpy.random.default_rng(1).uniform(size=(8, 4, 6)))
    realh = pmh.upsample(reall, resampler='cic')
    assert_almost_equal(reall.csum(), realh.csum())
    assert_almost_equal(pml.downsample(realh, resampler='cic').csum(), reall.csum())


def test_cmean_preserved_by_resample(PM):
    pm1 = PM.ParticleMesh(BoxSize=8.0, Nmesh=[8, 8], dtype='f8')
    pm2 = PM.ParticleMesh(BoxSize=8.0, Nmesh=[4, 4], dtype='f8')
    complex1 = PM.ComplexField(pm1)
    real2 = PM.RealField(pm2)
    real1 = PM.RealField(pm1)
    for i, kk, slab in zip(complex1.slabs.i, complex1.slabs.x, complex1.slabs):
        slab[...] = sum([k ** 2 for k in kk]) ** 0.5
    complex1.c2r(real1)
    real1.resample(real2)
    assert_almost_equal(real1.cmean(), real2.cmean())


def test_ctranspose(PM):
    pm = PM.ParticleMesh(BoxSize=[8.0, 16.0, 32.0], Nmesh=[4, 6, 8], dtype='f8')
    comp1 = pm.generate_whitenoise(1234, type='real')
    comp1t = comp1.ctranspose([0, 1, 2])
    assert_array_equal(comp1t.Nmesh, comp1.Nmesh)
    assert_array_equal(comp1t.BoxSize, comp1.BoxSize)
    assert_array_equal(comp1t.cnorm(), comp1.cnorm())
    comp1t = comp1.ctranspose([1, 2, 0])
    assert_array_equal(comp1t.Nmesh, comp1.Nmesh[[1, 2, 0]])
    assert_array_equal(comp1t.BoxSize, comp1.BoxSize[[1, 2, 0]])
    assert_array_equal(comp1t.value, comp1.value.transpose(1, 2, 0))
    comp1ttt = comp1t.ctranspose([1, 2, 0]).ctranspose([1, 2, 0])
    assert_allclose(comp1ttt, comp1)


def test_preview(PM):
    pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[4, 4, 4], dtype='f8')
    comp1 = pm.generate_whitenoise(1234, type='real')
    preview = comp1.preview(axes=(0, 1, 2))
    preview = comp1.preview(Nmesh=4, axes=(0, 1, 2))
    for ind1 in numpy.ndindex(*(list(comp1.cshape))):
        assert_allclose(preview[ind1], comp1.cgetitem(ind1))
    assert_allclose(comp1.preview(Nmesh=4, axes=(0, 1)), preview.sum(axis=2))
    assert_allclose(comp1.preview(Nmesh=4, axes=(1, 2)), preview.sum(axis=0))
    assert_allclose(comp1.preview(Nmesh=4, axes=(0, 2)), preview.sum(axis=1))
    assert_allclose(comp1.preview(Nmesh=4, axes=(2, 0)), preview.sum(axis=1).T)
    assert_allclose(comp1.preview(Nmesh=4, axes=(0,)), preview.sum(axis=(1, 2)))
    comp1.value[...] += 2.0                       # a non-zero mean to look at
    p8 = comp1.preview(Nmesh=8, axes=(0,))
    assert p8.shape == (8,)
    assert_allclose(p8.mean(), comp1.cmean() * 64, rtol=1e-10)      # keep_mean upsampling, summed over 8 x 8
    p2 = comp1.preview(Nmesh=2, axes=(0, 1, 2))
    assert p2.shape == (2, 2, 2)
    assert_allclose(p2.mean(), comp1.cmean(), rtol=1e-10)
