"""The reference's own API tests, restated for one rank (pmesh/tests/test_pm.py, test_gradient.py):
what a user of pmesh relies on besides the numbers -- shapes, coordinate conventions, operators,
in-place transforms, Hermitian indexing, iteration over slabs, apply() with Python callables,
collective reductions, white-noise invariances.  Each test names the reference test it restates.
"""
import numpy
import pytest
from numpy.testing import assert_array_equal, assert_allclose, assert_almost_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def PM():
    from pmesh_b200 import pm
    return pm


def test_asarray_and_shapes(PM):
    """test_asarray, test_shape_real, test_shape_complex (test_pm.py:11-42)"""
    for dt in ('f8', 'f4'):
        pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[8, 8], dtype=dt)
        real = PM.RealField(pm)
        a = numpy.asarray(real)
        assert a is real.value
    pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[8, 8], dtype='f8')
    real = PM.RealField(pm)
    assert tuple(real.cshape) == (8, 8) and real.csize == 64
    comp = PM.ComplexField(pm)
    assert tuple(comp.cshape) == (8, 5) and comp.csize == 40


def test_negnyquist_and_indices(PM):
    """test_negnyquist (:44-53), test_indices (:266-275): the Nyquist wavenumber is negative"""
    pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[8, 8], dtype='f8')
    c = pm.create(type='complex')
    assert (c.x[-1][0][-1] < 0).all()
    assert (c.x[-1][0][:-1] >= 0).all()
    pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[4, 4], dtype='f8')
    comp = pm.create(type='complex')
    real = pm.create(type='real')
    assert_almost_equal(comp.x[0], [[0], [0.785], [-1.571], [-0.785]], decimal=3)
    assert_almost_equal(comp.x[1], [[0, 0.785, -1.571]], decimal=3)
    assert_almost_equal(real.x[0], [[0], [2], [-4], [-2]], decimal=3)
    assert_almost_equal(real.x[1], [[0, 2, -4, -2]], decimal=3)
    assert comp.compressed == True and real.compressed == False          # test_field_compressed (:288-300)


def test_operators(PM):
    """test_operators (:78-111): ufuncs keep the field type, comparisons and reductions do not"""
    pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[4, 4], dtype='f4')
    real = pm.create(type='real', value=0)
    complex = pm.create(type='complex', value=0)
    real = real + 1
    real = 1 + real
    real = real + real.value
    real = real * real.value
    real = real * real
    assert isinstance(real, PM.RealField)
    assert_array_equal(real.value, 256.0)
    complex = 1 + complex
    assert isinstance(complex, PM.ComplexField)
    complex = complex + 1
    assert isinstance(complex, PM.ComplexField)
    complex = complex + complex.value
    assert isinstance(complex, PM.ComplexField)
    complex = numpy.conj(complex) * complex
    assert isinstance(complex, PM.ComplexField)
    assert (real == real).dtype == numpy.dtype('?')
    assert not isinstance(real == real, PM.RealField)
    assert not isinstance(complex == complex, PM.ComplexField)
    assert not isinstance(numpy.sum(real), PM.RealField)


def test_create_typenames_and_fft(PM):
    """test_create_typenames (:113-125), test_fft (:127-142)"""
    pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[4, 4], dtype='f4')
    real = pm.create(type=PM.RealField, value=0)
    real = pm.create(type=PM._typestr_to_type('real'), value=0)
    real.cast(type=PM._typestr_to_type('real'))
    real.cast(type=PM.RealField)
    real = pm.create(type='real', value=0)
    real[...] = 2
    real[::2, ::2] = -2
    real3 = real.copy()
    complex = real.r2c()
    assert_almost_equal(numpy.asarray(real), numpy.asarray(real3), decimal=7)
    real2 = complex.c2r()
    assert_almost_equal(numpy.asarray(real), numpy.asarray(real2), decimal=7)


def test_inplace_fft(PM):
    """test_inplace_fft (:165-192)"""
    pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[8, 8], dtype='f8')
    Npar = 100
    pos = 1.0 * (numpy.arange(Npar * len(pm.Nmesh))).reshape(-1, len(pm.Nmesh)) * (7, 7)
    pos %= (pm.Nmesh + 1)
    layout = pm.decompose(pos)
    npos = layout.exchange(pos)
    real = pm.paint(npos)
    complex = real.r2c()
    complex2 = real.r2c(out=Ellipsis)
    assert real._base in complex2._base
    assert_almost_equal(numpy.asarray(complex), numpy.asarray(complex2), decimal=7)
    real = complex2.c2r()
    real2 = complex2.c2r(out=Ellipsis)
    assert real2._base in complex2._base
    assert_almost_equal(numpy.asarray(real), numpy.asarray(real2), decimal=7)


def test_decompose_paints_like_serial(PM):
    """test_decompose (:228-264): decompose + exchange + paint == a serial paint, cic / tsc / db12"""
    from pmesh_b200 import window
    pm = PM.ParticleMesh(BoxSize=4.0, Nmesh=[4, 4, 4], dtype='f8')
    pos = pm.generate_uniform_particle_grid(shift=0.5)
    for resampler in ['cic', 'tsc', 'db12']:
        truth = numpy.zeros(pm.Nmesh, dtype='f8')
        affine = window.Affine(ndim=3, period=4)
        window.FindResampler(resampler).paint(truth, pos, transform=affine)
        layout = pm.decompose(pos, smoothing=resampler)
        npos = layout.exchange(pos)
        real = pm.paint(npos, resampler=resampler)
        assert_almost_equal(real.value, truth)


def test_slab_iteration_and_apply(PM):
    """test_real_iter (:310-324), test_real_apply (:326-339), test_complex_apply (:341-354),
    test_untransposed_complex_apply (:356-369), test_complex_iter (:542-551)"""
    pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[8, 8], dtype='f8')
    real = PM.RealField(pm)
    for i, x, slab in zip(real.slabs.i, real.slabs.x, real.slabs):
        assert_almost_equal(slab.shape, sum(x[d] ** 2 for d in range(len(pm.Nmesh))).shape)
        for a, b in zip(slab.x, x):
            assert_array_equal(a, b)
        for a, b in zip(slab.i, i):
            assert_array_equal(a, b)

    def rfilter(x, v):
        assert_allclose(x.normp(), sum(xi ** 2 for xi in x))
        return x[0] * 10 + x[1]
    real.apply(rfilter, out=Ellipsis)
    for i, x, slab in zip(real.slabs.i, real.slabs.x, real.slabs):
        assert_array_equal(slab, x[0] * 10 + x[1])

    complex = PM.ComplexField(pm)

    def cfilter(k, v):
        assert_allclose(k.normp(), sum(ki ** 2 for ki in k))
        return k[0] + k[1] * 1j
    complex.apply(cfilter, out=Ellipsis)
    for i, x, slab in zip(complex.slabs.i, complex.slabs.x, complex.slabs):
        assert_array_equal(slab, x[0] + x[1] * 1j)
    for x, slab in zip(complex.slabs.x, complex.slabs):
        assert_array_equal(slab.shape, sum(x[d] ** 2 for d in range(len(pm.Nmesh))).shape)
        for a, b in zip(slab.x, x):
            assert_almost_equal(a, b)

    pm3 = PM.ParticleMesh(BoxSize=8.0, Nmesh=[8, 8, 8], dtype='f8')
    ucomplex = PM.UntransposedComplexField(pm3)

    def ufilter(k, v):
        assert_allclose(k.normp(), sum(ki ** 2 for ki in k))
        return k[0] + k[1] * 1j + k[2]
    ucomplex = ucomplex.apply(ufilter)
    for i, x, slab in zip(ucomplex.slabs.i, ucomplex.slabs.x, ucomplex.slabs):
        assert_array_equal(slab, x[0] + x[1] * 1j + x[2])


def test_ctol_and_cgetitem(PM):
    """test_ctol (:553-559), test_cgetitem (:561-630): Hermitian-aware collective indexing"""
    pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[4, 4], dtype='f8')
    complex = PM.ComplexField(pm)
    value, local = complex._ctol((3, 3))
    assert local is None
    for i in numpy.ndindex((4, 4)):
        f = PM.RealField(pm)
        f[...] = 0
        v2 = f.csetitem(i, 100.)
        v1 = f.cgetitem(i)
        assert v2 == 100.
        assert_array_equal(v1, v2)
    for i in numpy.ndindex((4, 3)):
        complex = PM.ComplexField(pm)
        complex[...] = 0
        v2 = complex.csetitem(i, 100. + 10j)
        complex.c2r(out=Ellipsis).r2c(out=Ellipsis)
        v1 = complex.cgetitem(i)
        total = complex.value.sum()
        if i in ((0, 0), (0, 2), (2, 0), (2, 2)):
            assert v2 == 100.
            assert_allclose(total, 100., atol=1e-12)
        elif i in ((1, 0), (3, 0), (3, 2), (1, 2)):
            assert v2 == 100 + 10j
            assert_allclose(total, 200., atol=1e-12)
        else:
            assert v2 == 100. + 10j
            assert_allclose(total, 100. + 10j, atol=1e-12)
        assert_allclose(v1, v2, atol=1e-12)
    for i in numpy.ndindex((4, 3, 2)):
        complex = PM.ComplexField(pm)
        complex[...] = 0
        v2 = complex.csetitem(i, 100.)
        complex.c2r(out=Ellipsis).r2c(out=Ellipsis)
        v1 = complex.cgetitem(i)
        if i in ((0, 0, 1), (0, 2, 1), (2, 0, 1), (2, 2, 1)):
            assert v2 == 0.
        else:
            assert v2 == 100.
        assert_allclose(v1, v2, atol=1e-12)


def test_whitenoise_invariances(PM):
    """test_whitenoise (:632-648): the large scales do not depend on the resolution;
    test_whitenoise_mean (:650-658); test_whitenoise_untransposed (:144-163)"""
    pm0 = PM.ParticleMesh(BoxSize=8.0, Nmesh=[8, 8, 8], dtype='f8')
    pm1 = PM.ParticleMesh(BoxSize=8.0, Nmesh=[16, 16, 16], dtype='f8')
    pm2 = PM.ParticleMesh(BoxSize=8.0, Nmesh=[32, 32, 32], dtype='f8')
    complex1_down = PM.ComplexField(pm0)
    complex2_down = PM.ComplexField(pm0)
    complex1 = pm1.generate_whitenoise(seed=8, unitary=True)
    complex2 = pm2.generate_whitenoise(seed=8, unitary=True)
    complex1.resample(complex1_down)
    complex2.resample(complex2_down)
    assert_array_equal(complex1_down.value, complex2_down.value)
    complex1 = pm0.generate_whitenoise(seed=8, unitary=True, mean=1.0)
    assert_allclose(complex1.c2r().cmean(), 1.0)
    pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[4, 4, 4], dtype='f4')
    f1 = pm.generate_whitenoise(seed=3333, type='untransposedcomplex')
    f2 = pm.generate_whitenoise(seed=3333, type='transposedcomplex')
    assert_array_equal(numpy.array(f1.ravel()), numpy.array(f2.ravel()))
    assert_array_equal(numpy.array(f1.c2r().ravel()), numpy.array(f2.c2r().ravel()))


def test_readout_dtypes(PM):
    """test_readout (:660-677): every position / output dtype combination"""
    pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[8, 8], dtype='f8')
    real = PM.RealField(pm)
    real.value[...] = 1.0
    for pdt, odt in (('f8', 'f8'), ('f4', 'f8'), ('f4', 'f4')):
        pos = numpy.ones((1, 2), dtype=pdt)
        out = numpy.empty((1), dtype=odt)
        real.readout(pos, out=out)
        assert_allclose(out, 1.0)


def test_cdot_cnorm(PM):
    """test_cdot_cnorm (:679-689), test_cnorm_log (:692-700), test_cdot (:702-718), test_cdot_types (:739-754)"""
    pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[4, 4, 4], dtype='f8')
    comp1 = pm.generate_whitenoise(1234, type='complex')
    norm1 = comp1.cdot(comp1)
    norm2 = comp1.cnorm()
    norm3 = (abs(numpy.fft.fftn(numpy.fft.irfftn(comp1.value, s=(4, 4, 4), axes=(0, 1, 2)))) ** 2).sum()
    assert_allclose(norm2, norm3)
    assert_allclose(norm2, norm1)
    compm = pm.generate_whitenoise(1234, type='complex', mean=1.0)
    n2 = compm.cnorm(norm=lambda x: numpy.log(x.real ** 2 + x.imag ** 2))
    n3 = (numpy.log(abs(numpy.fft.fftn(numpy.fft.irfftn(compm.value, s=(4, 4, 4), axes=(0, 1, 2)))) ** 2)).sum()
    assert_allclose(n2, n3)
    comp2 = pm.generate_whitenoise(1239, type='complex')
    n12 = comp1.cdot(comp2)
    n21 = comp2.cdot(comp1)
    norm_r = comp1.c2r().cdot(comp2.c2r()) / pm.Nmesh.prod()
    assert_allclose(n21.real, norm_r)
    assert_allclose(n12.real, norm_r)
    assert_allclose(n12.imag, -n21.imag, atol=1e-12)
    compu = pm.generate_whitenoise(1239, type='untransposedcomplex')
    with pytest.raises(TypeError):
        comp1.cdot(compu)
    comp1.cdot(compu.value)
    compu.cdot(comp1.value)


def test_grid_and_coords(PM):
    """test_grid x2 (:826-845), test_grid_shifted (:847-867), test_coords (:869-889)"""
    pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[4, 4, 4], dtype='f8')
    grid = pm.generate_uniform_particle_grid(shift=0.5)
    assert grid.shape[0] == pm.Nmesh.prod()
    real = pm.paint(grid)
    assert_array_equal(real, 1.0)
    grid, id = pm.generate_uniform_particle_grid(shift=0.5, return_id=True)
    assert len(id) == len(grid)
    assert len(numpy.unique(id)) == len(id) and numpy.max(id) == len(id) - 1 and numpy.min(id) == 0
    grid = grid + 4.0
    layout = pm.decompose(grid)
    layout.exchange(grid)
    real = pm.paint(grid, layout=layout)
    assert_allclose(real, 1.0)
    grid = grid - 6.1
    layout = pm.decompose(grid)
    real = pm.paint(grid, layout=layout)
    assert_allclose(real, 1.0)
    for dt, names in (('f8', ('real', 'complex')), ('f4', ('transposedcomplex',))):
        pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[4, 4, 4], dtype=dt)
        for name in names:
            grid_x = pm.create_coords(name)
            grid_i = pm.create_coords(name, return_indices=True)
            assert len(grid_x) == 3 and len(grid_i) == 3
            assert grid_x[0].dtype == pm.dtype


def test_c2r_vjp(PM):
    """test_c2r_vjp (test_gradient.py:70-101): analytic gradient == finite differences"""
    pm = PM.ParticleMesh(BoxSize=8.0, Nmesh=[4, 4], dtype='f8')
    # the reference draws the field with its 2-D white-noise helper (a numpy RandomState); any real
    # field with a non-zero mean serves here
    real = pm.create(type='real', value=1.0 + numpy.random.RandomState(1234).normal(size=(4, 4)))
    comp = real.r2c()

    def objective(comp):
        r = comp.c2r()
        return (r.value ** 2).sum()

    def perturb(comp, mode, value):
        comp = comp.copy()
        old = comp.cgetitem(mode)
        new = comp.csetitem(mode, value + old)
        return new - old, comp

    grad_real = PM.RealField(pm)
    grad_real[...] = real[...] * 2
    grad_comp = grad_real.c2r_vjp(grad_real)
    grad_comp.decompress_vjp(grad_comp)
    ng, ag = [], []
    dx = 1e-7
    for ind1 in numpy.ndindex(*(list(grad_comp.cshape) + [2])):
        dx1, c1 = perturb(comp, ind1, dx)
        ng.append((objective(c1) - objective(comp)) / dx)
        ag.append(grad_comp.cgetitem(ind1) * dx1 / dx)
    assert_allclose(ng, ag, rtol=1e-5, atol=1e-5)
