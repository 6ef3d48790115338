// pmb_route.h -- the per-particle routing arithmetic of GridND.decompose, compiled for host and device.
//
// pmb_domain.cu evaluates it in its count kernel; tests/harness/host_harness.cpp walks particles
// serially with the same routines, so the CPU-only test-suite checks the integer results (rank sets,
// counts, indices) bit for bit against the oracle and the reference's golden vectors.
//
// Reference: pmesh/domain.py:561-652 (numpy floored modulo, digitize, sil / sir), pmesh/_domain.pyx:62-100
// (patch walk, DomainAssign, the degenerate-by-rank quirk).
#pragma once
#include <math.h>
#include <stdint.h>

#include "pmb_internal.h"

#define ROUTE_BLOCK 256
#define ROUTE_MAXRANKS 64
#define ROUTE_MAXEDGES 1024

struct RouteGeom {
    int ndim;
    int periodic;
    int nranks;
    int ndomains;
    int shape[3];
    int dstride[3];
    double scale[3];
    double smoothing[3];
    double inv_width[3];         // (nedges - 1) / (edges[-1] - edges[0]): the guess of pmb_digitize_near
    const double *edges[3];      // device
    int nedges[3];
    const int32_t *assign;       // device [ndomains]
    const int16_t *degenerate;   // device [ndomains]
    int home;                    // 1: the count kernel may take its home-cell shortcut (PMB_ROUTE_HOME=0 switches it off)
    int all_trivial;             // every axis is a single periodic domain: the mask is a constant
    uint64_t const_mask;
};

// numpy floored modulo for doubles (npy_divmod): result has the sign of b, may round up to b itself
// (-1e-17 % 64.0 == 64.0); ref use: pmesh/domain.py:616-619
PMB_HD double pmb_pymod(double a, double b)
{
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0) != (m < 0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

// numpy.digitize(x, bins, right=False) for increasing bins: number of bins <= x (len(bins) for NaN)
PMB_HD int pmb_digitize(double x, const double *bins, int n)
{
    if (x != x) return n;
    int lo = 0, hi = n;            // first index with bins[idx] > x
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (bins[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// the same count found from a guess: edges of a GridND are (nearly) equally spaced, so
// (int)(x * inv_width) + 1 is the answer or next to it; the two loops move it until
// bins[g - 1] <= x < bins[g] holds, which is exactly what the binary search returns for increasing bins.
PMB_HD int pmb_digitize_near(double x, const double *bins, int n, double inv_width)
{
    if (x != x) return n;
    const double t = (x - bins[0]) * inv_width;
    int g;
    if (!(t >= 0.0)) g = 0;                 // also NaN (degenerate grids: 0 * inf)
    else if (t >= (double) n) g = n;
    else g = (int) t + 1;
    if (g > n) g = n;
    while (g < n && bins[g] <= x) g++;
    while (g > 0 && bins[g - 1] > x) g--;
    return g;
}

// python non-negative integer modulo
PMB_HD int pmb_imod(int a, int n)
{
    int m = a % n;
    return m < 0 ? m + n : m;
}

// c = x % box with numpy semantics; for 0 <= x < box fmod(x, box) == x exactly, so the (slow,
// iterative) fmod is only taken by out-of-box coordinates.  -0.0 maps to +0.0 like npy_divmod.
PMB_HD double pmb_pymod_fast(double a, double b)
{
    if (a >= 0.0 && a < b) return a + 0.0;
    return pmb_pymod(a, b);
}

// per-particle rank mask; ref: pmesh/domain.py:609-630 (sil/sir) + pmesh/_domain.pyx:62-100 (patch walk)
// NDIM is a compile-time constant so that sil / sir / the patch odometer live in registers; `edges`
// point into shared memory (the binary searches of digitize never leave the SM).
// does the routing look at coordinate d at all?  (one periodic domain on an axis: no)
PMB_HD bool pmb_route_axis_used(const RouteGeom &g, int d)
{
    return !(g.periodic && g.shape[d] == 1);
}

// xin[d]: the raw coordinates of the particle (only the axes pmb_route_axis_used says are read)
template <int NDIM>
PMB_HD uint64_t pmb_route_mask_x(const RouteGeom &g, const double *const *edges, const double *xin)
{
    int sil[NDIM], sir[NDIM];
    bool single = true;          // every axis so far has a patch of exactly one domain
#pragma unroll
    for (int d = 0; d < NDIM; d++) {
        if (!pmb_route_axis_used(g, d)) {
            // one periodic domain on this axis: sil = p - 1, sir = p, and the single patch cell wraps
            // to domain 0 whatever the coordinate is (also for NaN: digitize gives len(edges))
            sil[d] = 0; sir[d] = 1;
            continue;
        }
        const double x = g.scale[d] * xin[d];
        const double sm = g.smoothing[d];
        const double *e = edges[d];
        const int ne = g.nedges[d];
        int l, r;
        const double inv_w = g.inv_width[d];
        if (g.periodic) {
            const double box = e[ne - 1];
            const double c = pmb_pymod_fast(x, box);
            const int p = pmb_digitize_near(c, e, ne, inv_w);
            // particles well inside their domain (the common case): c - sm and c + sm stay in [0, box)
            // and between the same two edges, so l = r = p without two more searches
            const double cl = c - sm, cr = c + sm;
            if (sm >= 0.0 && p >= 1 && p < ne && cl >= 0.0 && cr < box && e[p - 1] <= cl && cr < e[p]) {
                // l = r = p: sil = p - (0 % shape) - 1, sir = p + (0 % shape), without the two integer divisions
                l = p - 1; r = p;
            } else {
                l = pmb_digitize_near(pmb_pymod_fast(cl, box), e, ne, inv_w);
                r = pmb_digitize_near(pmb_pymod_fast(cr, box), e, ne, inv_w);
                l = p - pmb_imod(p - l, g.shape[d]) - 1;
                r = p + pmb_imod(r - p, g.shape[d]);
                single = false;
            }
        } else {
            l = pmb_digitize_near(x - sm, e, ne, inv_w);
            r = pmb_digitize_near(x + sm, e, ne, inv_w);
            l = l - 1;
            l = l < 0 ? 0 : (l > g.shape[d] ? g.shape[d] : l);
            r = r < 0 ? 0 : (r > g.shape[d] ? g.shape[d] : r);
            single = false;
        }
        sil[d] = (int) (int16_t) l;     // the reference stores sil/sir as int16 (domain.py:603-604)
        sir[d] = (int) (int16_t) r;
    }
    if (single) {
        // the common case, particles well inside their domain on every axis: ONE patch cell, sil[d] = p - 1 in
        // [0, shape) (1 <= p < nedges), no periodic wrap to apply -- the walk below reduces to one lookup
        int target = 0;
#pragma unroll
        for (int d = 0; d < NDIM; d++) target += sil[d] * g.dstride[d];
        const int rank = g.assign[target];
        const int deg = (rank >= 0 && rank < g.ndomains) ? g.degenerate[rank] : 0;
        return (!deg && rank >= 0 && rank < ROUTE_MAXRANKS) ? (uint64_t) 1 << rank : 0;
    }
    long long patch = 1;
    int p[NDIM];
#pragma unroll
    for (int d = 0; d < NDIM; d++) { patch *= (sir[d] - sil[d]); p[d] = sil[d]; }
    uint64_t mask = 0;
    for (long long q = 0; q < patch; q++) {
        int target = 0;
#pragma unroll
        for (int d = 0; d < NDIM; d++) {
            int t = p[d];
            if (g.periodic) {      // sil >= -shape - 1 and sir <= 2 shape: a few conditional steps, no division
                const int n = g.shape[d];
                if (t >= n) { t -= n; if (t >= n) t = pmb_imod(t, n); }
                else if (t < 0) { t += n; if (t < 0) t = pmb_imod(t, n); }
            }
            target += t * g.dstride[d];
        }
        if (target >= 0 && target < g.ndomains) {
            const int rank = g.assign[target];
            // quirk kept: the degenerate flag is looked up by RANK, not by domain (_domain.pyx:81-83)
            const int deg = (rank >= 0 && rank < g.ndomains) ? g.degenerate[rank] : 0;
            if (!deg && rank >= 0 && rank < ROUTE_MAXRANKS) mask |= (uint64_t) 1 << rank;
        }
        p[NDIM - 1] += 1;
#pragma unroll
        for (int d = NDIM - 1; d > 0; d--) {
            if (p[d] == sir[d]) { p[d] = sil[d]; p[d - 1] += 1; }
        }
    }
    return mask;
}

// the same from the particle array (host harness, and callers that do not prefetch)
template <int NDIM>
PMB_HD uint64_t pmb_route_mask(const RouteGeom &g, const double *const *edges, const void *pos,
                                                   int elsize, int64_t ps0, int64_t ps1, int64_t i)
{
    double x[NDIM];
#pragma unroll
    for (int d = 0; d < NDIM; d++)
        x[d] = pmb_route_axis_used(g, d) ? pmb_ld_real_stream(pos, i * ps0 + d * ps1, elsize) : 0.0;
    return pmb_route_mask_x<NDIM>(g, edges, x);
}
