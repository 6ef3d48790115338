/* pm_oracle.c -- CPU ORACLE (test infrastructure, NOT part of the product).
 *
 * A plain-C restatement of the reference pmesh particle-mesh kernels, written to
 * be read next to the reference and to round exactly like it (compile with
 * -O2 -ffp-contract=off; the reference is built for generic x86-64, no FMA).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this file's shared object.  Parity status: PINNED -- tests/test_oracle.py checks
 * it bit-for-bit against the compiled reference (oracle/_ref) and against the
 * reference's own known-answer tests (pmesh/tests/test_window.py, test_domain.py).
 *
 * What follows what:
 *   ora_window_init ........ pmesh/_window_imp.c:24-47    pmesh_window_info_init
 *   ora_kernel / ora_diff .. pmesh/_window_imp.c:108-236  analytic kernels
 *                            pmesh/_window_lanczos.h:2058-2084, _window_acg.h (same shape),
 *                            pmesh/_window_wavelets.h:460-486  table kernels
 *   ora_axis_generic ....... pmesh/_window_imp.c:50-83    _fill_k
 *   ora_axis_tuned ......... pmesh/_window_tuned_nnb.h:1-27, _window_tuned_cic.h:1-32,
 *                            _window_tuned_tsc.h:1-37, _window_tuned_pcs.h:1-52
 *   ora_paint / ora_readout  pmesh/_window.pyx:128-205 (particle loop),
 *                            pmesh/_window_generics.h:4-142 (generic walk, bounds, +=),
 *                            pmesh/_window_tuned_cic.h:34-72 (tuned product order)
 *   ora_gridnd_fill ........ pmesh/_domain.pyx:9-122
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#define ORA_MAXDIM 8

enum { K_NEAREST, K_LINEAR, K_CUBIC, K_QUADRATIC,
       K_LANCZOS2, K_LANCZOS3, K_LANCZOS4, K_LANCZOS5, K_LANCZOS6,
       K_ACG2, K_ACG3, K_ACG4, K_ACG5, K_ACG6,
       K_DB6, K_DB12, K_DB20, K_SYM6, K_SYM12, K_SYM20,
       K_TUNED_NNB, K_TUNED_CIC, K_TUNED_TSC, K_TUNED_PCS };

typedef struct {
    int kind;
    int support;                    /* requested (<=0: native) */
    int ndim;
    int order[ORA_MAXDIM];
    double scale[ORA_MAXDIM];
    double translate[ORA_MAXDIM];
    ptrdiff_t period[ORA_MAXDIM];
    void *canvas;
    int elsize;
    ptrdiff_t size[ORA_MAXDIM];
    ptrdiff_t strides[ORA_MAXDIM];  /* bytes */
    const double *table;            /* lookup table of the kind, or NULL */
    int tablesize;
    double step;
    double hsupport;
    int pcs_scale_fix;              /* 0 = reference behaviour */
} ora_painter;

typedef struct { int support; int left; double shift; double vfactor; } ora_window;

static int native_support(int kind)
{
    switch (kind) {
    case K_NEAREST: case K_TUNED_NNB: return 1;
    case K_LINEAR: case K_TUNED_CIC: return 2;
    case K_QUADRATIC: case K_TUNED_TSC: return 3;
    case K_CUBIC: case K_TUNED_PCS: return 4;
    case K_LANCZOS2: return 4; case K_LANCZOS3: return 6; case K_LANCZOS4: return 8;
    case K_LANCZOS5: return 10; case K_LANCZOS6: return 12;
    case K_ACG2: return 2; case K_ACG3: return 3; case K_ACG4: return 4; case K_ACG5: return 5; case K_ACG6: return 6;
    case K_DB6: case K_SYM6: return 7;
    case K_DB12: case K_SYM12: return 10;
    case K_DB20: return 13;
    case K_SYM20: return 12;
    }
    return -1;
}

static void ora_window_init(ora_window *w, int native, double support)
{
    if (support <= 0) {
        w->support = native;
        support = native;
    } else {
        w->support = support;
        w->support += (support != (double) w->support);
    }
    w->left = (w->support - 1) / 2;
    w->shift = support / 2.0 - w->support / 2;
    w->vfactor = native / (1. * support);
}

/* shape of kernel(x): 0 nearest, 1 linear, 2 quadratic, 3 cubic, 4 symmetric table, 5 wavelet table */
static int kernel_shape(int kind)
{
    switch (kind) {
    case K_NEAREST: case K_TUNED_NNB: return 0;
    case K_LINEAR: case K_TUNED_CIC: return 1;
    case K_QUADRATIC: case K_TUNED_TSC: return 2;
    case K_CUBIC: case K_TUNED_PCS: return 3;
    case K_DB6: case K_DB12: case K_DB20: case K_SYM6: case K_SYM12: case K_SYM20: return 5;
    default: return 4;
    }
}

static double ora_kernel(const ora_painter *p, double x)
{
    const double *t = p->table;
    double f, xx;
    int i;
    switch (kernel_shape(p->kind)) {
    case 0:
        return (x < 0.5 && x >= -0.5) ? 1.0 : 0.0;
    case 1:
        x = fabs(x);
        return x < 1.0 ? 1.0 - x : 0.0;
    case 2:
        x = fabs(x);
        if (x <= 0.5) return 0.75 - x * x;
        if (x < 1.5) { x = 1.5 - x; return (x * x) * 0.5; }
        return 0;
    case 3:
        x = fabs(x);
        xx = x * x;
        if (x < 1.0) return 1.0 / 6.0 * (4 - 6 * xx + 3 * xx * x);
        if (x < 2) return 1.0 / 6.0 * (2 - x) * (2 - x) * (2 - x);
        return 0;
    case 4:
        x = fabs(x);
        f = x / p->step;
        i = f;
        if (i < 0) return 0;
        if (i >= p->tablesize - 1) return 0;
        f -= i;
        return t[i] * (1 - f) + t[i + 1] * f;
    default:
        x += p->hsupport;
        f = x / p->step;
        if (f < 0) return 0;
        i = f;
        f -= i;
        if (i >= p->tablesize - 1) return 0;
        return t[i] * (1 - f) + t[i + 1] * f;
    }
}

static double ora_diff(const ora_painter *p, double x)
{
    const double *t = p->table;
    double sgn, xx, f;
    int i;
    switch (kernel_shape(p->kind)) {
    case 0:
        return 0;
    case 1:
        if (x < 0) { sgn = 1; x = -x; } else if (x > 0) sgn = -1; else sgn = 0;
        return x < 1.0 ? sgn : 0.0;
    case 2:
        if (x < 0) { x = -x; sgn = -1; } else sgn = +1;
        if (x <= 0.5) return sgn * (-2 * x);
        if (x < 1.5) return sgn * (-(1.5 - x));
        return 0;
    case 3:
        if (x < 0) { sgn = -1; x = -x; } else sgn = +1;
        xx = x * x;
        if (x < 1.0) return sgn * (1.0 / 6.0) * (-12 * x + 9 * xx);
        if (x < 2.0) return sgn * (-1.0 / 2.0) * (2 - x) * (2 - x);
        return 0;
    case 4:
        if (x >= 0) sgn = 1; else { sgn = -1; x = -x; }
        i = x / p->step;
        if (i < 0) return 0;
        if (i >= p->tablesize - 1) return 0;
        f = t[i + 1] - t[i];
        return sgn * f / p->step;
    default:
        x += p->hsupport;
        i = x / p->step;
        if (i < 0) return 0;
        if (i >= p->tablesize - 1) return 0;
        return (t[i + 1] - t[i]) / p->step;
    }
}

/* which tuned routine applies (0 = none): needs a tuned kind, order <= 1 on the first three axes,
 * ndim <= 3 (pmesh/_window_imp.c:381-452) and, per particle, support == native. */
static int tuned_family(const ora_painter *p)
{
    int fam = p->kind == K_TUNED_NNB ? 1 : p->kind == K_TUNED_CIC ? 2 : p->kind == K_TUNED_TSC ? 3
              : p->kind == K_TUNED_PCS ? 4 : 0;
    int d;
    if (!fam || p->ndim > 3) return 0;
    for (d = 0; d < p->ndim; d++) if (p->order[d] > 1) return 0;
    return fam;
}

/* per-axis tuned stencil: first mesh index and `fam` weights */
static int ora_axis_tuned(const ora_painter *p, int fam, int d, double x, double *V)
{
    const double X = x * p->scale[d] + p->translate[d];
    const double sc = p->scale[d];
    const int diff = p->order[d] != 0;
    int I0, I1, I2, I3;
    switch (fam) {
    case 1:
        I0 = (int) floor(X + 0.5);
        V[0] = diff ? 0 : 1;
        return I0;
    case 2:
        I0 = (int) floor(X);
        if (!diff) { V[1] = X - I0; V[0] = 1. - V[1]; }
        else { V[1] = sc; V[0] = -sc; }
        return I0;
    case 3:
        I1 = (int) floor(X + 0.5); I0 = I1 - 1; I2 = I1 + 1;
        if (!diff) {
            V[1] = 0.75 - (X - I1) * (X - I1);
            V[0] = (1.5 - (X - I0)) * (1.5 - (X - I0)) * 0.5;
            V[2] = (1.5 + (X - I2)) * (1.5 + (X - I2)) * 0.5;
        } else {
            V[1] = -2 * (X - I1) * sc;
            V[0] = -(1.5 - (X - I0)) * sc;
            V[2] = (1.5 + (X - I2)) * sc;
        }
        return I0;
    default:
        I1 = (int) floor(X); I0 = I1 - 1; I2 = I1 + 1; I3 = I2 + 1;
        if (!diff) {
            V[1] = 1.0 / 6.0 * (4 - 6 * (X - I1) * (X - I1) + 3 * (X - I1) * (X - I1) * (X - I1));
            V[2] = 1.0 / 6.0 * (4 - 6 * (X - I2) * (X - I2) - 3 * (X - I2) * (X - I2) * (X - I2));
            V[0] = 1.0 / 6.0 * (2 - (X - I0)) * (2 - (X - I0)) * (2 - (X - I0));
            V[3] = 1.0 / 6.0 * (2 + (X - I3)) * (2 + (X - I3)) * (2 + (X - I3));
        } else {
            /* no scale[d] factor here in the reference */
            V[1] = +1.0 / 6.0 * (-12 * (X - I1) + 9 * (X - I1) * (X - I1));
            V[2] = -1.0 / 6.0 * (+12 * (X - I2) + 9 * (X - I2) * (X - I2));
            V[0] = -1.0 / 2.0 * (2 - (X - I0)) * (2 - (X - I0));
            V[3] = +1.0 / 2.0 * (2 + (X - I3)) * (2 + (X - I3));
            if (p->pcs_scale_fix) { V[0] *= sc; V[1] *= sc; V[2] *= sc; V[3] *= sc; }
        }
        return I0;
    }
}

static int ora_axis_generic(const ora_painter *p, const ora_window *w, int d, double x, double *k)
{
    const double g = x * p->scale[d] + p->translate[d];
    const int ipos = floor(g + w->shift) - w->left;
    const double dx = g - ipos;
    int i;
    for (i = 0; i < w->support; i++) {
        const double u = (dx - i) * w->vfactor;
        if (p->order[d] == 0) k[i] = ora_kernel(p, u) * w->vfactor;
        else k[i] = ora_diff(p, u) * p->scale[d] * w->vfactor * w->vfactor;
    }
    return ipos;
}

/* one particle: mode 0 paint (adds weight*W into the canvas), mode 1 readout (returns sum W*canvas) */
static double ora_particle(const ora_painter *p, const double *x, double weight, double hsml, int mode)
{
    ora_window w;
    const int nd = p->ndim;
    int first[ORA_MAXDIM], rel[ORA_MAXDIM], d;
    double value = 0;
    ora_window_init(&w, native_support(p->kind), (p->support <= 0 ? native_support(p->kind) : p->support) * hsml);
    const int fam = tuned_family(p);
    const int tuned = fam && w.support == fam;
    const int S = w.support;
    double *k = (double *) malloc(sizeof(double) * nd * S);
    for (d = 0; d < nd; d++) {
        first[d] = tuned ? ora_axis_tuned(p, fam, d, x[d], k + d * S)
                         : ora_axis_generic(p, &w, d, x[d], k + d * S);
        rel[d] = 0;
    }
    if (tuned && mode == 0) for (d = 0; d < S; d++) k[d] *= weight;   /* V*[0] *= weight */
    while (rel[0] != S) {
        double prod = tuned ? k[rel[0]] : 1.0 * k[rel[0]];
        ptrdiff_t off = 0;
        int inside = 1;
        for (d = 0; d < nd; d++) {
            ptrdiff_t t = (ptrdiff_t) first[d] + rel[d];
            if (d > 0) prod *= k[d * S + rel[d]];
            if (p->period[d] > 0) {
                while (t >= p->period[d]) t -= p->period[d];
                while (t < 0) t += p->period[d];
            }
            if (t < 0 || t >= p->size[d]) { inside = 0; break; }
            off += t * p->strides[d];
        }
        if (inside) {
            char *cell = (char *) p->canvas + off;
            if (mode == 0) {
                const double f = tuned ? prod : weight * prod;
                if (p->elsize == 8) *(double *) cell += f; else *(float *) cell += f;
            } else {
                const double c = p->elsize == 8 ? *(double *) cell : (double) *(float *) cell;
                value += tuned ? c * prod : prod * c;
            }
        }
        rel[nd - 1]++;
        for (d = nd - 1; d > 0; d--)
            if (rel[d] == S) { rel[d - 1]++; rel[d] = 0; }
    }
    free(k);
    return value;
}

/* pos: (n, ndim) C-contiguous doubles (the reference promotes f4 positions to double per element,
 * _window.pyx:159); mass/hsml may be NULL (1.0). */
void ora_paint(const ora_painter *p, const double *pos, const double *mass, const double *hsml, ptrdiff_t n)
{
    ptrdiff_t i;
    for (i = 0; i < n; i++)
        ora_particle(p, pos + i * p->ndim, mass ? mass[i] : 1.0, hsml ? hsml[i] : 1.0, 0);
}

void ora_readout(const ora_painter *p, const double *pos, const double *hsml, double *out, ptrdiff_t n)
{
    ptrdiff_t i;
    for (i = 0; i < n; i++)
        out[i] = ora_particle(p, pos + i * p->ndim, 1.0, hsml ? hsml[i] : 1.0, 1);
}

int ora_window_support(int kind, int support_req, int *support, int *native)
{
    ora_window w;
    int ns = native_support(kind);
    if (ns < 0) return -1;
    ora_window_init(&w, ns, (double) support_req);
    *support = w.support;
    *native = ns;
    return 0;
}

/* mode 0: counts[rank] += 1 per (particle, distinct target rank); mode 1: indices grouped by rank
 * (offsets from counts), particle ids ascending inside a rank.  sil/sir are (ndim, npoint) int16. */
int ora_gridnd_fill(int mode, int32_t *counts, int nranks, const int32_t *dims, int ndim,
                    const int16_t *sil, const int16_t *sir, ptrdiff_t npoint, int periodic,
                    const int16_t *degenerate, const int32_t *assign, int32_t *indices)
{
    int32_t *offset = NULL;
    ptrdiff_t i;
    int strides[ORA_MAXDIM], j;
    ptrdiff_t cap = 16;
    ptrdiff_t *list = (ptrdiff_t *) malloc(sizeof(ptrdiff_t) * cap);
    if (mode == 1) {
        offset = (int32_t *) malloc(sizeof(int32_t) * (nranks + 1));
        offset[0] = 0;
        for (j = 1; j <= nranks; j++) offset[j] = offset[j - 1] + counts[j - 1];
    }
    strides[ndim - 1] = 1;
    for (j = ndim - 2; j >= 0; j--) strides[j] = strides[j + 1] * dims[j + 1];
    for (i = 0; i < npoint; i++) {
        int p[ORA_MAXDIM];
        long patch = 1, q;
        ptrdiff_t nk = 0, m, last;
        for (j = 0; j < ndim; j++) {
            patch *= sir[j * npoint + i] - sil[j * npoint + i];
            p[j] = sil[j * npoint + i];
        }
        if (patch > cap) { cap = patch; list = (ptrdiff_t *) realloc(list, sizeof(ptrdiff_t) * cap); }
        for (q = 0; q < patch; q++) {
            ptrdiff_t target = 0;
            for (j = 0; j < ndim; j++) {
                int t = p[j];
                if (periodic) {
                    while (t >= dims[j]) t -= dims[j];
                    while (t < 0) t += dims[j];
                }
                target += t * strides[j];
            }
            target = assign[target];
            if (!degenerate[target]) {       /* indexed by rank, as in the reference */
                m = nk++;
                while (m > 0 && list[m - 1] > target) { list[m] = list[m - 1]; m--; }
                list[m] = target;
            }
            p[ndim - 1]++;
            for (j = ndim - 1; j > 0; j--) {
                if (p[j] == sir[j * npoint + i]) { p[j] = sil[j * npoint + i]; p[j - 1]++; }
                else break;
            }
        }
        last = -1;
        for (m = 0; m < nk; m++) {
            if (list[m] == last) continue;
            last = list[m];
            if (mode == 0) counts[last]++;
            else indices[offset[last]++] = (int32_t) i;
        }
    }
    free(list);
    free(offset);
    return 0;
}
