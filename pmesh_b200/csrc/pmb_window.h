// pmb_window.h -- resampling-window arithmetic shared by every paint / readout
// kernel of the library (device) and by the host-side test harness.
//
// Everything here is written so that, compiled with FMA contraction OFF
// (nvcc -fmad=false / gcc -ffp-contract=off), each floating point expression
// rounds exactly like the reference pmesh C code does on x86-64:
//   window descriptor ........ pmesh/_window_imp.c:24-47   (pmesh_window_info_init)
//   analytic kernels ......... pmesh/_window_imp.c:108-244
//   table kernels ............ pmesh/_window_lanczos.h:2058-2084, _window_acg.h (same form),
//                              pmesh/_window_wavelets.h:460-486
//   generic per-axis weights . pmesh/_window_imp.c:50-83    (_fill_k)
//   tuned NNB/CIC/TSC/PCS .... pmesh/_window_tuned_{nnb,cic,tsc,pcs}.h (SETUP_KERNEL_*)
// The file is a restatement: one per-axis weight routine per family, no macros,
// no function pointers; the kernels decide how to traverse the stencil.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PMB_HD __host__ __device__ __forceinline__
#else
#define PMB_HD inline
#endif

// window kinds; numbering follows the public enum of the reference C header
// (pmesh/_window_imp.h:4-28) so integer kinds mean the same thing.
enum {
    PMB_NEAREST = 0, PMB_LINEAR = 1, PMB_CUBIC = 2, PMB_QUADRATIC = 3,
    PMB_LANCZOS2 = 4, PMB_LANCZOS3 = 5, PMB_LANCZOS4 = 6, PMB_LANCZOS5 = 7, PMB_LANCZOS6 = 8,
    PMB_ACG2 = 9, PMB_ACG3 = 10, PMB_ACG4 = 11, PMB_ACG5 = 12, PMB_ACG6 = 13,
    PMB_DB6 = 14, PMB_DB12 = 15, PMB_DB20 = 16, PMB_SYM6 = 17, PMB_SYM12 = 18, PMB_SYM20 = 19,
    PMB_TUNED_NNB = 20, PMB_TUNED_CIC = 21, PMB_TUNED_TSC = 22, PMB_TUNED_PCS = 23,
    PMB_NKINDS = 24
};

// kernel families (how kernel(x) / diff(x) are evaluated)
enum { PMB_FAM_NEAREST = 0, PMB_FAM_LINEAR = 1, PMB_FAM_QUADRATIC = 2, PMB_FAM_CUBIC = 3,
       PMB_FAM_SYMTABLE = 4, PMB_FAM_WAVELET = 5 };

#define PMB_MAXDIM 3
#define PMB_MAX_SUPPORT 32   // widest stencil kept in per-thread arrays; wider ones are evaluated on the fly

typedef struct PmbWindow {
    int kind;
    int family;
    int tuned;          // 0, or the native support (1,2,3,4) of the tuned fast path when eligible
    int nativesupport;
    int support;        // integer support after init (painter->support), multiplied by hsml per particle
    int tablesize;
    double step;
    double hsupport;
    const double *table;
} PmbWindow;

typedef struct PmbGeom {
    int ndim;
    int order[PMB_MAXDIM];
    double scale[PMB_MAXDIM];
    double translate[PMB_MAXDIM];
    int64_t period[PMB_MAXDIM];
    int64_t size[PMB_MAXDIM];
    int64_t strides[PMB_MAXDIM];   // bytes
} PmbGeom;

typedef struct PmbWinInfo {
    int support;
    int left;
    double shift;
    double vfactor;
} PmbWinInfo;

// ref: pmesh/_window_imp.c:24-47
PMB_HD void pmb_window_info(int nativesupport, double support, PmbWinInfo *info)
{
    if (support <= 0) {
        info->support = nativesupport;
        support = nativesupport;
    } else {
        info->support = (int) support;
        info->support += (support != (double) info->support);
    }
    info->left = (info->support - 1) / 2;
    info->shift = support / 2.0 - info->support / 2;
    info->vfactor = nativesupport / (1. * support);
}

// family / native support / tuned family (0 none, else 1 nnb, 2 cic, 3 tsc, 4 pcs) of a kind.
// ref: the switch of pmesh_painter_init, pmesh/_window_imp.c:259-452.  Returns 0, or -1 for a bad kind.
PMB_HD int pmb_kind_info(int kind, int *family, int *native, int *tuned)
{
    *tuned = 0;
    if (kind == PMB_NEAREST || kind == PMB_TUNED_NNB) { *family = PMB_FAM_NEAREST; *native = 1; *tuned = kind == PMB_TUNED_NNB ? 1 : 0; return 0; }
    if (kind == PMB_LINEAR || kind == PMB_TUNED_CIC) { *family = PMB_FAM_LINEAR; *native = 2; *tuned = kind == PMB_TUNED_CIC ? 2 : 0; return 0; }
    if (kind == PMB_QUADRATIC || kind == PMB_TUNED_TSC) { *family = PMB_FAM_QUADRATIC; *native = 3; *tuned = kind == PMB_TUNED_TSC ? 3 : 0; return 0; }
    if (kind == PMB_CUBIC || kind == PMB_TUNED_PCS) { *family = PMB_FAM_CUBIC; *native = 4; *tuned = kind == PMB_TUNED_PCS ? 4 : 0; return 0; }
    if (kind >= PMB_LANCZOS2 && kind <= PMB_LANCZOS6) { *family = PMB_FAM_SYMTABLE; *native = 2 * (kind - PMB_LANCZOS2 + 2); return 0; }
    if (kind >= PMB_ACG2 && kind <= PMB_ACG6) { *family = PMB_FAM_SYMTABLE; *native = kind - PMB_ACG2 + 2; return 0; }
    if (kind == PMB_DB6 || kind == PMB_SYM6) { *family = PMB_FAM_WAVELET; *native = 7; return 0; }
    if (kind == PMB_DB12 || kind == PMB_SYM12) { *family = PMB_FAM_WAVELET; *native = 10; return 0; }
    if (kind == PMB_DB20) { *family = PMB_FAM_WAVELET; *native = 13; return 0; }
    if (kind == PMB_SYM20) { *family = PMB_FAM_WAVELET; *native = 12; return 0; }
    return -1;
}

// resolve everything of a window but its lookup table (<- pmesh_painter_init, _window_imp.c:246-456):
// integer support from the request, and whether the tuned fast path is eligible
// (tuned kind, order <= 1 on every axis, ndim <= 3).
PMB_HD int pmb_window_resolve(int kind, int support_req, int ndim, const int *order, PmbWindow *w)
{
    int family, native, tuned;
    if (pmb_kind_info(kind, &family, &native, &tuned) != 0) return -1;
    w->kind = kind;
    w->family = family;
    w->nativesupport = native;
    w->tablesize = 0;
    w->step = 0;
    w->hsupport = 0;
    w->table = 0;
    PmbWinInfo info;
    pmb_window_info(native, (double) support_req, &info);
    w->support = info.support;
    w->tuned = 0;
    if (tuned && ndim <= 3) {
        int ok = 1;
        for (int d = 0; d < ndim; d++) if (order && order[d] > 1) ok = 0;
        if (ok) w->tuned = tuned;
    }
    return 0;
}

// ---------------------------------------------------------------- kernels
PMB_HD double pmb_table_lerp(const double *t, int i, double f)
{
    return t[i] * (1 - f) + t[i + 1] * f;
}

// ref: analytic kernels pmesh/_window_imp.c:108-236; tables _window_lanczos.h:2058-2070,
// _window_wavelets.h:460-471
PMB_HD double pmb_kernel(const PmbWindow &w, double x)
{
    switch (w.family) {
    case PMB_FAM_NEAREST:
        if (x < 0.5 && x >= -0.5) return 1.0;
        return 0;
    case PMB_FAM_LINEAR:
        x = fabs(x);
        if (x < 1.0) return 1.0 - x;
        return 0;
    case PMB_FAM_QUADRATIC:
        x = fabs(x);
        if (x <= 0.5) return 0.75 - x * x;
        if (x < 1.5) { x = 1.5 - x; return (x * x) * 0.5; }
        return 0;
    case PMB_FAM_CUBIC: {
        x = fabs(x);
        double xx = x * x;
        if (x < 1.0) return 1.0 / 6.0 * (4 - 6 * xx + 3 * xx * x);
        if (x < 2) return 1.0 / 6.0 * (2 - x) * (2 - x) * (2 - x);
        return 0;
    }
    case PMB_FAM_SYMTABLE: {
        x = fabs(x);
        double f = x / w.step;
        int i = (int) f;
        if (i < 0) return 0;
        if (i >= w.tablesize - 1) return 0;
        f -= i;
        return pmb_table_lerp(w.table, i, f);
    }
    default: { // PMB_FAM_WAVELET
        x += w.hsupport;
        double f = x / w.step;
        if (f < 0) return 0;
        int i = (int) f;
        f -= i;
        if (i >= w.tablesize - 1) return 0;
        return pmb_table_lerp(w.table, i, f);
    }
    }
}

// ref: pmesh/_window_imp.c:128-236 (analytic), _window_lanczos.h:2071-2084, _window_wavelets.h:472-486
PMB_HD double pmb_diff(const PmbWindow &w, double x)
{
    double factor;
    switch (w.family) {
    case PMB_FAM_NEAREST:
        return 0;
    case PMB_FAM_LINEAR:
        if (x < 0) { factor = 1; x = -x; }
        else if (x > 0) factor = -1;
        else factor = 0;
        if (x < 1.0) return factor;
        return 0;
    case PMB_FAM_QUADRATIC:
        if (x < 0) { x = -x; factor = -1; } else factor = +1;
        if (x <= 0.5) return factor * (-2 * x);
        if (x < 1.5) return factor * (-(1.5 - x));
        return 0;
    case PMB_FAM_CUBIC: {
        if (x < 0) { factor = -1; x = -x; } else factor = +1;
        double xx = x * x;
        if (x < 1.0) return factor * (1.0 / 6.0) * (-12 * x + 9 * xx);
        if (x < 2.0) return factor * (-1.0 / 2.0) * (2 - x) * (2 - x);
        return 0;
    }
    case PMB_FAM_SYMTABLE: {
        if (x >= 0) factor = 1; else { factor = -1; x = -x; }
        int i = (int) (x / w.step);
        if (i < 0) return 0;
        if (i >= w.tablesize - 1) return 0;
        double f = w.table[i + 1] - w.table[i];
        return factor * f / w.step;
    }
    default: { // PMB_FAM_WAVELET
        x += w.hsupport;
        int i = (int) (x / w.step);
        if (i < 0) return 0;
        if (i >= w.tablesize - 1) return 0;
        double f0 = w.table[i];
        double f1 = w.table[i + 1];
        return (f1 - f0) / w.step;
    }
    }
}

// ------------------------------------------------- per-axis stencil weights
// All routines return the un-wrapped left-to-right mesh indices I[0..S) and the
// weights V[0..S) of one axis.  X is the particle coordinate in mesh units.

// grid coordinate: two roundings, never an FMA (ref: _window_tuned_cic.h:8, _window_imp.c:58)
PMB_HD double pmb_gridpos(double pos, double scale, double translate)
{
    return pos * scale + translate;
}

// ref: pmesh/_window_tuned_nnb.h:1-27
PMB_HD void pmb_axis_nnb(double X, int order, double scale, int *I, double *V)
{
    (void) scale;
    I[0] = (int) floor(X + 0.5);
    V[0] = (order == 0) ? 1 : 0;
}

// ref: pmesh/_window_tuned_cic.h:1-32
PMB_HD void pmb_axis_cic(double X, int order, double scale, int *I, double *V)
{
    I[0] = (int) floor(X);
    I[1] = I[0] + 1;
    if (order == 0) {
        V[1] = X - I[0];
        V[0] = 1. - V[1];
    } else {
        V[1] = scale;
        V[0] = -scale;
    }
}

// ref: pmesh/_window_tuned_tsc.h:1-37
PMB_HD void pmb_axis_tsc(double X, int order, double scale, int *I, double *V)
{
    I[1] = (int) floor(X + 0.5);
    I[0] = I[1] - 1;
    I[2] = I[1] + 1;
    if (order == 0) {
        V[1] = 0.75 - (X - I[1]) * (X - I[1]);
        V[0] = (1.5 - (X - I[0])) * (1.5 - (X - I[0])) * 0.5;
        V[2] = (1.5 + (X - I[2])) * (1.5 + (X - I[2])) * 0.5;
    } else {
        V[1] = -2 * (X - I[1]) * scale;
        V[0] = -(1.5 - (X - I[0])) * scale;
        V[2] = (1.5 + (X - I[2])) * scale;
    }
}

// ref: pmesh/_window_tuned_pcs.h:1-52.  The derivative branch has NO scale[d]
// factor in the reference; kept bug-compatible unless pcs_scale_fix.
PMB_HD void pmb_axis_pcs(double X, int order, double scale, int *I, double *V, int pcs_scale_fix)
{
    I[1] = (int) floor(X);
    I[0] = I[1] - 1;
    I[2] = I[1] + 1;
    I[3] = I[2] + 1;
    if (order == 0) {
        V[1] = 1.0 / 6.0 * (4 - 6 * (X - I[1]) * (X - I[1])
                              + 3 * (X - I[1]) * (X - I[1]) * (X - I[1]));
        V[2] = 1.0 / 6.0 * (4 - 6 * (X - I[2]) * (X - I[2])
                              - 3 * (X - I[2]) * (X - I[2]) * (X - I[2]));
        V[0] = 1.0 / 6.0 * (2 - (X - I[0])) * (2 - (X - I[0])) * (2 - (X - I[0]));
        V[3] = 1.0 / 6.0 * (2 + (X - I[3])) * (2 + (X - I[3])) * (2 + (X - I[3]));
    } else {
        V[1] = +1.0 / 6.0 * (-12 * (X - I[1]) + 9 * (X - I[1]) * (X - I[1]));
        V[2] = -1.0 / 6.0 * (+12 * (X - I[2]) + 9 * (X - I[2]) * (X - I[2]));
        V[0] = -1.0 / 2.0 * (2 - (X - I[0])) * (2 - (X - I[0]));
        V[3] = +1.0 / 2.0 * (2 + (X - I[3])) * (2 + (X - I[3]));
        if (pcs_scale_fix) {
            V[0] *= scale; V[1] *= scale; V[2] *= scale; V[3] *= scale;
        }
    }
}

// generic axis: left-most index and dx (ref: pmesh/_window_imp.c:58-60)
PMB_HD int pmb_axis_generic_origin(double X, const PmbWinInfo &info, double *dx)
{
    int ipos = (int) (floor(X + info.shift) - info.left);
    *dx = X - ipos;
    return ipos;
}

// generic weight of stencil point i of one axis (ref: pmesh/_window_imp.c:62-68)
PMB_HD double pmb_axis_generic_weight(const PmbWindow &w, const PmbWinInfo &info, double dx, int i,
                                      int order, double scale)
{
    double x = (dx - i) * info.vfactor;
    if (order == 0)
        return pmb_kernel(w, x) * info.vfactor;
    return pmb_diff(w, x) * scale * info.vfactor * info.vfactor;
}

// periodic wrap then local bounds test (ref: _window_generics.h:42-55, _window_tuned_cic.h:24-31)
// returns the wrapped index, or -1 when the point falls outside the local canvas.  The common
// case (at most one period away) costs a compare and an add; no integer division.
PMB_HD int64_t pmb_wrap_clip(int64_t t, int64_t period, int64_t size)
{
    if (period > 0) {
        if (t >= period) {
            t -= period;
            if (t >= period) t %= period;
        } else if (t < 0) {
            t += period;
            if (t < 0) { t %= period; if (t < 0) t += period; }
        }
    }
    if (t < 0 || t >= size) return -1;
    return t;
}

// 32-bit variant used by the tuned kernels (canvas extents are checked to be < 2^31)
PMB_HD int pmb_wrap32(int t, int period)
{
    if (t >= period) {
        t -= period;
        if (t >= period) t %= period;
    } else if (t < 0) {
        t += period;
        if (t < 0) { t %= period; if (t < 0) t += period; }
    }
    return t;
}
