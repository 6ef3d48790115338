// host_harness.cpp -- TEST-ONLY host build of the device stencil code.
//
// The window / stencil arithmetic of the CUDA kernels lives in headers that compile for host and
// device (pmesh_b200/csrc/pmb_window.h, pmb_stencil.cuh).  This file walks particles serially with
// exactly those routines so that `pytest -m "not gpu"` can check, in a container without a GPU,
// that the arithmetic the kernels execute is bit-identical to the oracle (g++ -ffp-contract=off
// mirrors nvcc -fmad=false).  It is not part of libpmesh_b200.so and is never used by the product.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../pmesh_b200/csrc/pmb_stencil.cuh"

static pmb_table g_tables[PMB_NKINDS];
static std::vector<double> g_store[PMB_NKINDS];

extern "C" int hh_set_table(int kind, const double *values, int n, double step, double hsupport)
{
    if (kind < 0 || kind >= PMB_NKINDS) return -1;
    g_store[kind].assign(values, values + n);
    g_tables[kind].d_values = g_store[kind].data();
    g_tables[kind].n = n;
    g_tables[kind].step = step;
    g_tables[kind].hsupport = hsupport;
    return 0;
}

static int resolve(const pmb_resample_args *a, PmbWindow *w)
{
    if (pmb_window_resolve(a->kind, a->support, a->ndim, a->order, w) != 0) return -1;
    if (w->family == PMB_FAM_SYMTABLE || w->family == PMB_FAM_WAVELET) {
        if (!g_tables[a->kind].d_values) return -2;
        w->table = g_tables[a->kind].d_values;
        w->tablesize = g_tables[a->kind].n;
        w->step = g_tables[a->kind].step;
        w->hsupport = g_tables[a->kind].hsupport;
    }
    return 0;
}

static void geom(const pmb_resample_args *a, PmbGeom *g)
{
    memset(g, 0, sizeof(*g));
    g->ndim = a->ndim;
    for (int d = 0; d < a->ndim; d++) {
        g->order[d] = a->order[d]; g->scale[d] = a->scale[d]; g->translate[d] = a->translate[d];
        g->period[d] = a->period[d]; g->size[d] = a->size[d]; g->strides[d] = a->strides[d];
    }
}

static void parts(const pmb_resample_args *a, PmbParticles *p)
{
    p->pos = a->pos; p->pos_elsize = a->pos_elsize; p->ps0 = a->pos_stride0; p->ps1 = a->pos_stride1;
    p->mass = a->mass; p->mass_elsize = a->mass_elsize; p->ms = a->mass_stride; p->mass_scalar = a->mass_scalar;
    p->hsml = a->hsml; p->hsml_elsize = a->hsml_elsize; p->hs = a->hsml_stride; p->hsml_scalar = a->hsml_scalar;
}

static int fixed_family(const PmbWindow &w, const pmb_resample_args *a)
{
    if (!w.tuned || a->hsml) return 0;
    PmbWinInfo info;
    pmb_window_info(w.nativesupport, w.support * a->hsml_scalar, &info);
    return info.support == w.tuned ? w.tuned : 0;
}

// visit the stencil of particle i the way the kernels do: f(ordinal, off, value_for_paint, weightprod)
template <int NDIM, int FAM, class F>
static void visit_fixed(const PmbGeom &g, const PmbParticles &p, int64_t i, int pcsfix, F &&f)
{
    double x[NDIM];
    pmb_load_pos<NDIM>(p, i, x);
    const double m = pmb_load_mass(p, i);
    PmbAxes<NDIM, FAM> A;
    auto cb = [&](int ord, int64_t off, double v0, double v1, double v2) {
        f(ord, off, pmb_paint_value(true, m, v0, v1, v2), (v0 * v1) * v2);
    };
    if (pmb_geom_needs_check(g)) {          // the same CHECK dispatch as the kernels
        pmb_axes_tuned<NDIM, FAM, true>(g, g.order, x, pcsfix, A);
        pmb_for_points_fixed<NDIM, FAM, true>(A, cb);
    } else {
        pmb_axes_tuned<NDIM, FAM, false>(g, g.order, x, pcsfix, A);
        pmb_for_points_fixed<NDIM, FAM, false>(A, cb);
    }
}

template <int NDIM, class F>
static void visit_dyn(const PmbGeom &g, const PmbWindow &w, const PmbParticles &p, int64_t i, int pcsfix, F &&f)
{
    double x[NDIM];
    pmb_load_pos<NDIM>(p, i, x);
    const double m = pmb_load_mass(p, i);
    const double h = pmb_load_hsml(p, i);
    PmbWinInfo info;
    pmb_window_info(w.nativesupport, w.support * h, &info);
    if (info.support <= PMB_MAX_SUPPORT) {
        PmbAxes<NDIM, PMB_MAX_SUPPORT> A;
        pmb_axes_dyn<NDIM>(g, w, info, g.order, x, pcsfix, A);
        const bool tuned = A.tuned;
        pmb_for_points_dyn<NDIM, PMB_MAX_SUPPORT>(A, [&](int ord, int64_t off, double v0, double v1, double v2) {
            f(ord, off, pmb_paint_value(tuned, m, v0, v1, v2), (v0 * v1) * v2);
        });
    } else {
        pmb_for_points_wide<NDIM>(g, w, info, g.order, x, [&](int ord, int64_t off, double v0, double v1, double v2) {
            f(ord, off, pmb_paint_value(false, m, v0, v1, v2), (v0 * v1) * v2);
        });
    }
}

template <int NDIM, class F>
static void visit(const PmbGeom &g, const PmbWindow &w, int fam, const PmbParticles &p, int64_t i, int pcsfix, F &&f)
{
    switch (fam) {
    case 0: visit_dyn<NDIM>(g, w, p, i, pcsfix, f); break;
    case 1: visit_fixed<NDIM, 1>(g, p, i, pcsfix, f); break;
    case 2: visit_fixed<NDIM, 2>(g, p, i, pcsfix, f); break;
    case 3: visit_fixed<NDIM, 3>(g, p, i, pcsfix, f); break;
    default: visit_fixed<NDIM, 4>(g, p, i, pcsfix, f); break;
    }
}

template <int NDIM>
static int paint_nd(const pmb_resample_args *a, const PmbWindow &w)
{
    PmbGeom g;
    geom(a, &g);
    PmbParticles p;
    parts(a, &p);
    const int fam = fixed_family(w, a);
    char *mesh = (char *) a->mesh;
    if (a->mode == PMB_MODE_ATOMIC) {
        for (int64_t i = 0; i < a->npart; i++)
            visit<NDIM>(g, w, fam, p, i, a->pcs_gradient_scale_fix, [&](int, int64_t off, double f, double) {
                if (off == PMB_OFF_INVALID) return;
                if (a->mesh_elsize == 8) *(double *) (mesh + off) += f;
                else *(float *) (mesh + off) = *(float *) (mesh + off) + (float) f;   // red.global.add.f32 semantics
            });
        return 0;
    }
    // deterministic: (cell, value) pairs in particle/point order, stable sort by cell, sequential sums
    PmbGeom gd = g;
    int64_t acc = 1;
    for (int d = NDIM - 1; d >= 0; d--) { gd.strides[d] = acc; acc *= gd.size[d]; }
    std::vector<std::pair<int64_t, double>> pairs;
    for (int64_t i = 0; i < a->npart; i++)
        visit<NDIM>(gd, w, fam, p, i, a->pcs_gradient_scale_fix, [&](int, int64_t off, double f, double) {
            if (off != PMB_OFF_INVALID) pairs.emplace_back(off, f);
        });
    std::stable_sort(pairs.begin(), pairs.end(),
                     [](const std::pair<int64_t, double> &l, const std::pair<int64_t, double> &r) { return l.first < r.first; });
    for (size_t s = 0; s < pairs.size();) {
        int64_t lin = pairs[s].first, rem = lin;
        int64_t off = 0;
        for (int d = NDIM - 1; d >= 0; d--) { off += (rem % g.size[d]) * g.strides[d]; rem /= g.size[d]; }
        size_t e = s;
        if (a->mesh_elsize == 8) {
            double c = *(double *) (mesh + off);
            for (; e < pairs.size() && pairs[e].first == lin; e++) c = (double) ((double) c + pairs[e].second);
            *(double *) (mesh + off) = c;
        } else {
            float c = *(float *) (mesh + off);
            for (; e < pairs.size() && pairs[e].first == lin; e++) c = (float) ((double) c + pairs[e].second);
            *(float *) (mesh + off) = c;
        }
        s = e;
    }
    return 0;
}

template <int NDIM>
static int readout_nd(const pmb_resample_args *a, const PmbWindow &w)
{
    PmbGeom g;
    geom(a, &g);
    PmbParticles p;
    parts(a, &p);
    const int fam = fixed_family(w, a);
    const char *mesh = (const char *) a->mesh;
    for (int64_t i = 0; i < a->npart; i++) {
        double value = 0;
        visit<NDIM>(g, w, fam, p, i, a->pcs_gradient_scale_fix, [&](int, int64_t off, double, double wp) {
            if (off == PMB_OFF_INVALID) return;
            const double c = a->mesh_elsize == 8 ? *(const double *) (mesh + off) : (double) *(const float *) (mesh + off);
            value += c * wp;
        });
        pmb_st_real(a->out, i * a->out_stride, a->out_elsize, value);
    }
    return 0;
}

extern "C" int hh_paint(const pmb_resample_args *a)
{
    PmbWindow w;
    int rc = resolve(a, &w);
    if (rc) return rc;
    for (int d = 0; d < a->ndim; d++) if (a->size[d] == 0) return 0;
    switch (a->ndim) {
    case 1: return paint_nd<1>(a, w);
    case 2: return paint_nd<2>(a, w);
    case 3: return paint_nd<3>(a, w);
    }
    return -3;
}

extern "C" int hh_readout(const pmb_resample_args *a)
{
    PmbWindow w;
    int rc = resolve(a, &w);
    if (rc) return rc;
    switch (a->ndim) {
    case 1: return readout_nd<1>(a, w);
    case 2: return readout_nd<2>(a, w);
    case 3: return readout_nd<3>(a, w);
    }
    return -3;
}

// ---- white noise: the column routines of pmb_whitenoise.cu, walked serially on the host ------------
#include "../../pmesh_b200/csrc/pmb_wnrng.h"

extern "C" void hh_wn_stream(unsigned int seed, double *out, int n)
{
    WnRng g;
    wn_seed(g, seed);
    for (int i = 0; i < n; i++) out[i] = wn_uniform(g);
}

extern "C" int hh_whitenoise(void *canvas, int elsize, const int64_t *nmesh, const int64_t *start, const int64_t *size,
                             const int64_t *strides, unsigned int seed, int unitary)
{
    const int64_t ncol = size[0] * size[1];
    if (ncol == 0) return 0;
    std::vector<unsigned int> tab((size_t) ncol * 2, 0u);
    WnTables T;
    T.N0 = nmesh[0]; T.N1 = nmesh[1]; T.s0 = start[0]; T.s1 = start[1]; T.m0 = size[0]; T.m1 = size[1];
    T.t00 = tab.data(); T.t11 = tab.data() + ncol;
    wn_build_tables(T, seed);
    WnArgs a;
    for (int d = 0; d < 3; d++) { a.N[d] = nmesh[d]; a.start[d] = start[d]; a.size[d] = size[d]; a.strides[d] = strides[d]; }
    a.t00 = T.t00; a.t11 = T.t11; a.unitary = unitary; a.fast_axis = 1;
    for (int64_t li = 0; li < size[0]; li++)
        for (int64_t lj = 0; lj < size[1]; lj++) {
            WnColumn c;
            wn_column_init(c, a, li, lj);
            for (int64_t k = 0; k <= nmesh[2] / 2; k++) {
                double re, im;
                bool in = elsize == 16 ? wn_column_mode<double>(c, a, k, re, im) : wn_column_mode<float>(c, a, k, re, im);
                if (!in) continue;
                char *p = (char *) canvas + li * strides[0] + lj * strides[1] + (k - start[2]) * strides[2];
                if (elsize == 16) { ((double *) p)[0] = re; ((double *) p)[1] = im; }
                else { ((float *) p)[0] = (float) re; ((float *) p)[1] = (float) im; }
            }
        }
    return 0;
}

// ---- routing: the rank-mask arithmetic of pmb_domain.cu, walked serially on the host ----------------
#include "../../pmesh_b200/csrc/pmb_route.h"

// counts[nranks] and indices (grouped by rank ascending, particle ascending; capacity `cap`) of
// GridND.decompose for host arrays; returns the number of indices written or -1
extern "C" int64_t hh_decompose(const void *pos, int pos_elsize, int64_t npart, int64_t ps0, int64_t ps1, int ndim,
                                const double *scale, const double *smoothing, const double *edges, const int *nedges,
                                int periodic, const int32_t *assign, const int16_t *degenerate, int nranks,
                                int32_t *counts, int32_t *indices, int64_t cap)
{
    if (ndim < 1 || ndim > 3 || nranks < 1 || nranks > ROUTE_MAXRANKS) return -1;
    RouteGeom g;
    memset(&g, 0, sizeof(g));
    g.ndim = ndim; g.periodic = periodic; g.nranks = nranks;
    int nd = 1, o = 0;
    const double *e[3] = {NULL, NULL, NULL};
    for (int d = 0; d < ndim; d++) {
        g.shape[d] = nedges[d] - 1; g.nedges[d] = nedges[d];
        g.scale[d] = scale[d]; g.smoothing[d] = smoothing[d];
        e[d] = edges + o; g.edges[d] = e[d]; o += nedges[d];
        {
            const double span = e[d][nedges[d] - 1] - e[d][0];
            g.inv_width[d] = span > 0 ? (double) (nedges[d] - 1) / span : 0.0;
        }
        nd *= g.shape[d];
    }
    g.ndomains = nd;
    int st = 1;
    for (int d = ndim - 1; d >= 0; d--) { g.dstride[d] = st; st *= g.shape[d]; }
    g.assign = assign; g.degenerate = degenerate;
    std::vector<uint64_t> masks((size_t) npart);
    for (int64_t i = 0; i < npart; i++) {
        if (ndim == 1) masks[i] = pmb_route_mask<1>(g, e, pos, pos_elsize, ps0, ps1, i);
        else if (ndim == 2) masks[i] = pmb_route_mask<2>(g, e, pos, pos_elsize, ps0, ps1, i);
        else masks[i] = pmb_route_mask<3>(g, e, pos, pos_elsize, ps0, ps1, i);
    }
    int64_t n = 0;
    for (int r = 0; r < nranks; r++) {
        counts[r] = 0;
        for (int64_t i = 0; i < npart; i++)
            if ((masks[i] >> r) & 1) {
                if (n >= cap) return -1;
                indices[n++] = (int32_t) i;
                counts[r]++;
            }
    }
    return n;
}

// ---- the line transform of pmb_ifft.cuh, "thread" by "thread" and phase by phase on the host --------------
#include "../../pmesh_b200/csrc/pmb_ifft.cuh"

template <typename C, int N, bool CONTIG>
static int hh_ifft_run(const C *in, C *out, const C *twtab)
{
    constexpr int B = pmb_ifft_bundle<C, N, CONTIG>::B;
    typedef pmb_ifft_line<C, N, CONTIG, B> F;
    std::vector<C> sm((size_t) F::SMEM_ELEMS);
    std::vector<C> regs((size_t) B * F::TPL * 16);
    auto tw2 = [&](int r, int k) -> C { return twtab[k * r * F::R3]; };
    auto tw3 = [&](int r, int j) -> C { return twtab[j * r]; };
    auto V = [&](int b, int t) -> C * { return regs.data() + ((size_t) b * F::TPL + t) * 16; };
    for (int b = 0; b < B; b++)
        for (int t = 0; t < F::TPL; t++) {
            C *v = V(b, t);
            for (int r = 0; r < 16; r++) v[r] = in[(size_t) b * N + t + r * F::TPL];
            F::p1(sm.data(), b, t, v);
        }
    for (int b = 0; b < B; b++) for (int t = 0; t < F::TPL; t++) F::p2_load(sm.data(), b, t, V(b, t));
    for (int b = 0; b < B; b++)
        for (int t = 0; t < F::TPL; t++) {
            C *v = V(b, t);
            F::p2_compute(t, v, tw2);
            if (F::R3 == 1) {
                for (int m = 0; m < 16 / F::R2; m++)
                    for (int q = 0; q < F::R2; q++) out[(size_t) b * N + F::p2_out(t, m, q)] = v[m * F::R2 + q];
            } else {
                F::p2_store(sm.data(), b, t, v);
            }
        }
    if (F::R3 > 1) {
        for (int b = 0; b < B; b++) for (int t = 0; t < F::TPL; t++) F::p3_load(sm.data(), b, t, V(b, t));
        for (int b = 0; b < B; b++)
            for (int t = 0; t < F::TPL; t++) {
                C *v = V(b, t);
                F::p3_compute(t, v, tw3);
                for (int m = 0; m < 16 / F::R3; m++)
                    for (int q = 0; q < F::R3; q++) out[(size_t) b * N + F::p3_out(t, m, q)] = v[m * F::R3 + q];
            }
    }
    return B;
}

// inverse transform of a bundle of lines of n points ([line][point], complex, elsize 8 or 16 bytes per complex);
// returns the number of lines in a bundle (the caller supplies at least that many), -1: unsupported n
extern "C" int hh_ifft_lines(int n, int elsize, int contig, const void *in, void *out, const void *tw)
{
#define HH_IFFT_CASE(NN)                                                                                              \
    case NN:                                                                                                          \
        if (elsize == 16) return contig ? hh_ifft_run<double2, NN, true>((const double2 *) in, (double2 *) out, (const double2 *) tw) \
                                        : hh_ifft_run<double2, NN, false>((const double2 *) in, (double2 *) out, (const double2 *) tw); \
        return contig ? hh_ifft_run<float2, NN, true>((const float2 *) in, (float2 *) out, (const float2 *) tw)      \
                      : hh_ifft_run<float2, NN, false>((const float2 *) in, (float2 *) out, (const float2 *) tw);
    switch (n) {
        HH_IFFT_CASE(64) HH_IFFT_CASE(128) HH_IFFT_CASE(256) HH_IFFT_CASE(512) HH_IFFT_CASE(1024) HH_IFFT_CASE(2048) HH_IFFT_CASE(4096)
    }
#undef HH_IFFT_CASE
    return -1;
}
