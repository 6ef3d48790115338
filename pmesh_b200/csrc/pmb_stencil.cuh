// pmb_stencil.cuh -- per-particle stencil construction and traversal used by the paint,
// readout and deterministic-expand kernels.
//
// A stencil is, per axis, S unwrapped->wrapped mesh indices (as byte offsets, or OFF_INVALID when
// the point lies outside the local canvas) and S weights.  Points are visited in C order of
// (a, b, c) -- the order of the reference's tuned ACCESS3 sequences (_window_tuned_cic.h:43-50)
// and of its generic odometer (_window_generics.h:33-72) -- so sums round identically.
#pragma once
#include "pmb_internal.h"

#define PMB_OFF_INVALID INT64_MIN

struct PmbParticles {
    const void *pos; int pos_elsize; int64_t ps0, ps1;
    const void *mass; int mass_elsize; int64_t ms; double mass_scalar;
    const void *hsml; int hsml_elsize; int64_t hs; double hsml_scalar;
};

template <int NDIM, int SMAX>
struct PmbAxes {
    int S;
    bool tuned;                 // product convention of the tuned routines: ((V0*w)*V1)*V2
    double V[NDIM][SMAX];
    int64_t off[NDIM][SMAX];
};

template <int NDIM>
PMB_HD void pmb_load_pos(const PmbParticles &p, int64_t i, double *x, int stream = 1)
{
#pragma unroll
    for (int d = 0; d < NDIM; d++)
        x[d] = stream ? pmb_ld_real_stream(p.pos, i * p.ps0 + d * p.ps1, p.pos_elsize)
                      : pmb_ld_real(p.pos, i * p.ps0 + d * p.ps1, p.pos_elsize);
}
PMB_HD double pmb_load_mass(const PmbParticles &p, int64_t i)
{
    return p.mass ? pmb_ld_real_stream(p.mass, i * p.ms, p.mass_elsize) : p.mass_scalar;
}
PMB_HD double pmb_load_hsml(const PmbParticles &p, int64_t i)
{
    return p.hsml ? pmb_ld_real_stream(p.hsml, i * p.hs, p.hsml_elsize) : p.hsml_scalar;
}

template <int FAM>
PMB_HD void pmb_axis_tuned(double X, int order, double scale, int pcsfix, int *I, double *V)
{
    if (FAM == 1) pmb_axis_nnb(X, order, scale, I, V);
    else if (FAM == 2) pmb_axis_cic(X, order, scale, I, V);
    else if (FAM == 3) pmb_axis_tsc(X, order, scale, I, V);
    else pmb_axis_pcs(X, order, scale, I, V, pcsfix);
}

// true when some stencil point can fall outside the local canvas: a non-periodic axis, or a canvas
// that is a slab of the periodic mesh.  Otherwise (single-rank periodic mesh) the kernels skip
// every bounds test.
static inline bool pmb_geom_needs_check(const PmbGeom &g)
{
    for (int d = 0; d < g.ndim; d++)
        if (g.period[d] <= 0 || g.size[d] != g.period[d]) return true;
    return false;
}

// tuned stencil with compile-time support FAM (1 nnb, 2 cic, 3 tsc, 4 pcs).  The FAM mesh indices
// of an axis are consecutive: wrap the first, then step with a compare (no division).
template <int NDIM, int FAM, bool CHECK>
PMB_HD void pmb_axes_tuned(const PmbGeom &g, const int *order, const double *x, int pcsfix,
                           PmbAxes<NDIM, FAM> &A)
{
    A.S = FAM;
    A.tuned = true;
#pragma unroll
    for (int d = 0; d < NDIM; d++) {
        double X = pmb_gridpos(x[d], g.scale[d], g.translate[d]);
        int I[FAM];
        pmb_axis_tuned<FAM>(X, order[d], g.scale[d], pcsfix, I, A.V[d]);
        const int per = (int) g.period[d];
        const int sz = (int) g.size[d];
        int t = I[0];
        if (per > 0) t = pmb_wrap32(t, per);
#pragma unroll
        for (int s = 0; s < FAM; s++) {
            const bool ok = !CHECK || (t >= 0 && t < sz);
            A.off[d][s] = ok ? (int64_t) t * g.strides[d] : PMB_OFF_INVALID;
            t += 1;
            if (per > 0 && t == per) t = 0;
        }
    }
}

// run-time support (<= PMB_MAX_SUPPORT): tuned formulas when the per-particle support equals the
// native one (getfastmethod, _window_tuned_cic.h:135-157), the generic _fill_k otherwise.
template <int NDIM>
PMB_HD void pmb_axes_dyn(const PmbGeom &g, const PmbWindow &w, const PmbWinInfo &info,
                                             const int *order, const double *x, int pcsfix,
                                             PmbAxes<NDIM, PMB_MAX_SUPPORT> &A)
{
    A.S = info.support;
    A.tuned = (w.tuned != 0 && info.support == w.tuned);
    for (int d = 0; d < NDIM; d++) {
        double X = pmb_gridpos(x[d], g.scale[d], g.translate[d]);
        int I0;
        if (A.tuned) {
            int I[4];
            double V[4];
            switch (w.tuned) {
            case 1: pmb_axis_nnb(X, order[d], g.scale[d], I, V); break;
            case 2: pmb_axis_cic(X, order[d], g.scale[d], I, V); break;
            case 3: pmb_axis_tsc(X, order[d], g.scale[d], I, V); break;
            default: pmb_axis_pcs(X, order[d], g.scale[d], I, V, pcsfix); break;
            }
            I0 = I[0];
#pragma unroll 1
            for (int s = 0; s < A.S; s++) A.V[d][s] = V[s];
        } else {
            double dx;
            I0 = pmb_axis_generic_origin(X, info, &dx);
#pragma unroll 1
            for (int s = 0; s < A.S; s++)
                A.V[d][s] = pmb_axis_generic_weight(w, info, dx, s, order[d], g.scale[d]);
        }
#pragma unroll 1
        for (int s = 0; s < A.S; s++) {
            int64_t t = pmb_wrap_clip((int64_t) I0 + s, g.period[d], g.size[d]);
            A.off[d][s] = t < 0 ? PMB_OFF_INVALID : t * g.strides[d];
        }
    }
}

// visit all S^NDIM points in C order: f(ordinal, offset_or_INVALID, v0, v1, v2)
// FIXED: compile-time support SMAX, fully unrolled (tuned kernels).
template <int NDIM, int SMAX, bool CHECK, class F>
PMB_HD void pmb_for_points_fixed(const PmbAxes<NDIM, SMAX> &A, F &&f)
{
    int ord = 0;
#pragma unroll
    for (int a = 0; a < SMAX; a++) {
        const int64_t o0 = A.off[0][a];
        if (NDIM == 1) {
            f(ord++, o0, A.V[0][a], 1.0, 1.0);
        } else {
#pragma unroll
            for (int b = 0; b < SMAX; b++) {
                const int64_t o1 = A.off[NDIM > 1 ? 1 : 0][b];
                const bool bad01 = CHECK && (o0 == PMB_OFF_INVALID || o1 == PMB_OFF_INVALID);
                if (NDIM == 2) {
                    f(ord++, bad01 ? PMB_OFF_INVALID : o0 + o1, A.V[0][a], A.V[NDIM > 1 ? 1 : 0][b], 1.0);
                } else {
#pragma unroll
                    for (int c = 0; c < SMAX; c++) {
                        const int64_t o2 = A.off[NDIM > 2 ? 2 : 0][c];
                        const int64_t o = (CHECK && (bad01 || o2 == PMB_OFF_INVALID)) ? PMB_OFF_INVALID : o0 + o1 + o2;
                        f(ord++, o, A.V[0][a], A.V[NDIM > 1 ? 1 : 0][b], A.V[NDIM > 2 ? 2 : 0][c]);
                    }
                }
            }
        }
    }
}

// run-time support A.S: plain loops, no unrolling
template <int NDIM, int SMAX, class F>
PMB_HD void pmb_for_points_dyn(const PmbAxes<NDIM, SMAX> &A, F &&f)
{
    const int S = A.S;
    int ord = 0;
#pragma unroll 1
    for (int a = 0; a < S; a++) {
        const int64_t o0 = A.off[0][a];
        if (NDIM == 1) {
            f(ord++, o0, A.V[0][a], 1.0, 1.0);
        } else {
#pragma unroll 1
            for (int b = 0; b < S; b++) {
                const int64_t o1 = A.off[NDIM > 1 ? 1 : 0][b];
                const bool bad01 = (o0 == PMB_OFF_INVALID || o1 == PMB_OFF_INVALID);
                if (NDIM == 2) {
                    f(ord++, bad01 ? PMB_OFF_INVALID : o0 + o1, A.V[0][a], A.V[NDIM > 1 ? 1 : 0][b], 1.0);
                } else {
#pragma unroll 1
                    for (int c = 0; c < S; c++) {
                        const int64_t o2 = A.off[NDIM > 2 ? 2 : 0][c];
                        const int64_t o = (bad01 || o2 == PMB_OFF_INVALID) ? PMB_OFF_INVALID : o0 + o1 + o2;
                        f(ord++, o, A.V[0][a], A.V[NDIM > 1 ? 1 : 0][b], A.V[NDIM > 2 ? 2 : 0][c]);
                    }
                }
            }
        }
    }
}

// stencils wider than PMB_MAX_SUPPORT (e.g. LANCZOS2.resize(400), tests/test_window.py:215-219):
// nothing is cached, each point re-evaluates its per-axis weights.  Never tuned.
template <int NDIM, class F>
PMB_HD void pmb_for_points_wide(const PmbGeom &g, const PmbWindow &w, const PmbWinInfo &info,
                                    const int *order, const double *x, F &&f)
{
    int I0[NDIM];
    double dx[NDIM];
    for (int d = 0; d < NDIM; d++) {
        double X = pmb_gridpos(x[d], g.scale[d], g.translate[d]);
        I0[d] = pmb_axis_generic_origin(X, info, &dx[d]);
    }
    const int S = info.support;
    int64_t npts = 1;
    for (int d = 0; d < NDIM; d++) npts *= S;
    for (int64_t q = 0; q < npts; q++) {
        int r[3] = {0, 0, 0};
        int64_t t = q;
        for (int d = NDIM - 1; d >= 0; d--) { r[d] = (int) (t % S); t /= S; }
        double v[3] = {1.0, 1.0, 1.0};
        int64_t o = 0;
        bool bad = false;
        for (int d = 0; d < NDIM; d++) {
            v[d] = pmb_axis_generic_weight(w, info, dx[d], r[d], order[d], g.scale[d]);
            int64_t tt = pmb_wrap_clip((int64_t) I0[d] + r[d], g.period[d], g.size[d]);
            if (tt < 0) bad = true; else o += tt * g.strides[d];
        }
        f((int) q, bad ? PMB_OFF_INVALID : o, v[0], v[1], v[2]);
    }
}

// value deposited by paint for one point (ref: tuned `V0[0] *= weight; Va*Vb*Vc`, _window_tuned_cic.h:41-50;
// generic `weight * kernel`, _window_generics.h:61)
PMB_HD double pmb_paint_value(bool tuned, double m, double v0, double v1, double v2)
{
    return tuned ? ((v0 * m) * v1) * v2 : m * ((v0 * v1) * v2);
}
