// pmb_domain.cu -- GPU particle routing: rank/cell assignment, stable compaction, take, ghost sum.
//
// Replaces the host path  GridND.decompose (numpy digitize chunks, pmesh/domain.py:606-639)
//                       + gridnd_fill       (two-pass count/fill,   pmesh/_domain.pyx:9-122)
//                       + ndarray.take      (pmesh/domain.py:188)
//                       + bincountv         (pmesh/domain.py:26-48, gather mode 'sum').
// All integer results are bit-identical to the reference: `indices` lists particle ids grouped by
// target rank ascending and, within a rank, by particle id ascending (the reference's fill order).
//
// Design: the per-particle target set is a 64-bit rank mask (<= 64 ranks), which makes the
// reference's insertion sort + de-duplication (_domain.pyx:86-113) implicit.  A block owns a
// contiguous range of particles; per-(block, rank) histograms + one tiny scan give every block its
// write cursor, and ballots inside the block keep the particle order stable.
#include "pmb_internal.h"

#include "pmb_route.h"

// WARP w owns the contiguous particles [w*per_unit, (w+1)*per_unit): no block-level synchronisation
// anywhere.  Lane r (and r + 32) of a warp keeps the count / cursor of rank r in a register.
// MaskT: the narrowest unsigned type that holds one bit per rank (1 byte per particle up to 8 ranks).
#define ROUTE_UNROLL 8          // granularity of a unit: 32 * ROUTE_UNROLL particles
template <int NDIM, typename MaskT, int UNROLL>
__global__ void __launch_bounds__(ROUTE_BLOCK)
pmb_k_route_count(RouteGeom g, const void *pos, int elsize, int64_t ps0, int64_t ps1, int64_t npart,
                  int64_t per_unit, MaskT *masks, int32_t *unithist)
{
    extern __shared__ double s_edges[];
    const double *edges[NDIM];
    {
        int o = 0;
#pragma unroll
        for (int d = 0; d < NDIM; d++) {
            for (int k = threadIdx.x; k < g.nedges[d]; k += ROUTE_BLOCK) s_edges[o + k] = g.edges[d][k];
            edges[d] = s_edges + o;
            o += g.nedges[d];
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t unit = (int64_t) blockIdx.x * (ROUTE_BLOCK / 32) + (threadIdx.x >> 5);
    const int64_t begin = unit * per_unit;
    const int64_t end = min(begin + per_unit, npart);
    int c0 = 0, c1 = 0;
    // HOME CELL of the warp: the domain cell of its first particle.  A particle whose smoothing interval lies inside that
    // cell on every routed axis takes pmb_route_mask_x's interior path (l = r = p, one patch cell) with exactly this
    // cell, so its mask is the cell's -- known without a search, a table lookup or a divergent branch.  When ALL 32
    // particles of a step pass the test (nearly every step of a particle array that follows the decomposition) the
    // general arithmetic is skipped; otherwise the whole warp takes it as before.  Same predicates, same results.
    bool home_ok = false;
    __shared__ double s_home[ROUTE_BLOCK / 32][3 * NDIM];       // per warp: the cell's edges and the box, per axis
    double *hlo = s_home[threadIdx.x >> 5], *hhi = hlo + NDIM, *hbox = hlo + 2 * NDIM;
    uint64_t home_mask = 0;
    int home_rank = -1;
    if (g.home && g.periodic && begin < end) {
        home_ok = true;
        int target = 0;
#pragma unroll
        for (int d = 0; d < NDIM; d++) {
            if (!pmb_route_axis_used(g, d)) continue;
            const double x = g.scale[d] * pmb_ld_real_stream(pos, begin * ps0 + d * ps1, elsize);
            const double *e = edges[d];
            const int ne = g.nedges[d];
            const double box = e[ne - 1];
            const int p = pmb_digitize_near(pmb_pymod_fast(x, box), e, ne, g.inv_width[d]);
            if (!(g.smoothing[d] >= 0.0) || p < 1 || p >= ne) { home_ok = false; continue; }
            if (lane == 0) { hlo[d] = e[p - 1]; hhi[d] = e[p]; hbox[d] = box; }
            target += (p - 1) * g.dstride[d];
        }
        __syncwarp();
        if (home_ok) {
            const int rank = g.assign[target];
            const int deg = (rank >= 0 && rank < g.ndomains) ? g.degenerate[rank] : 0;
            if (!deg && rank >= 0 && rank < ROUTE_MAXRANKS) { home_mask = (uint64_t) 1 << rank; home_rank = rank; }
        }
    }
    for (int64_t base = begin; base < end; base += 32 * UNROLL) {
        // all coordinates of the UNROLL particles of this lane are requested before the first one is used:
        // the routing arithmetic is branchy (fmod fall-backs, edge searches) and the compiler does not hoist
        // loads across it
        double xs[UNROLL][NDIM];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int64_t i = base + u * 32 + lane;
#pragma unroll
            for (int d = 0; d < NDIM; d++)
                xs[u][d] = (i < end && pmb_route_axis_used(g, d)) ? pmb_ld_real_stream(pos, i * ps0 + d * ps1, elsize) : 0.0;
        }
        uint64_t mask[UNROLL];
        bool fast[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int64_t i = base + u * 32 + lane;
            bool in = home_ok;
#pragma unroll
            for (int d = 0; d < NDIM; d++) {
                if (!pmb_route_axis_used(g, d)) continue;
                // the conditions of the interior path of pmb_route_mask_x, with the home cell's edges for e[p - 1], e[p]
                const double x = g.scale[d] * xs[u][d];
                const double c = x + 0.0, cl = c - g.smoothing[d], cr = c + g.smoothing[d];
                in = in && x >= 0.0 && x < hbox[d] && cl >= 0.0 && cr < hbox[d] && hlo[d] <= cl && cr < hhi[d];
            }
            fast[u] = __all_sync(0xffffffffu, in || i >= end);
            if (fast[u]) mask[u] = i < end ? home_mask : 0;
            else mask[u] = i < end ? pmb_route_mask_x<NDIM>(g, edges, xs[u]) : 0;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int64_t i = base + u * 32 + lane;
            if (i < end) __stcs(masks + i, (MaskT) mask[u]);
            if (fast[u]) {
                // one rank (or none) for the whole step
                const int cnt = __popc(__ballot_sync(0xffffffffu, i < end));
                if (home_rank >= 0) {
                    if (lane == home_rank) c0 += cnt;
                    if (lane + 32 == home_rank) c1 += cnt;
                }
                continue;
            }
            // ranks that any lane of the warp targets (usually one or two)
            unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned) mask[u]);
            unsigned hi = g.nranks > 32 ? __reduce_or_sync(0xffffffffu, (unsigned) (mask[u] >> 32)) : 0u;
            while (lo) {
                const int r = __ffs(lo) - 1;
                lo &= lo - 1;
                const unsigned bal = __ballot_sync(0xffffffffu, (mask[u] >> r) & 1);
                if (lane == r) c0 += __popc(bal);
            }
            while (hi) {
                const int r = __ffs(hi) - 1;
                hi &= hi - 1;
                const unsigned bal = __ballot_sync(0xffffffffu, (mask[u] >> (r + 32)) & 1);
                if (lane == r) c1 += __popc(bal);
            }
        }
    }
    if (lane < g.nranks) unithist[unit * g.nranks + lane] = c0;
    if (lane + 32 < g.nranks) unithist[unit * g.nranks + lane + 32] = c1;
}

// block r scans the unit counts of rank r: thread t sums its chunk of units; the exclusive scan of
// the chunk sums goes to partial[r][t], the total to totals[r]
__global__ void __launch_bounds__(256)
pmb_k_route_scan(const int32_t *unithist, int64_t nunits, int nranks, int64_t *totals, int64_t *partial)
{
    __shared__ int64_t sums[256];
    const int r = blockIdx.x, t = threadIdx.x;
    const int64_t chunk = (nunits + 255) / 256;
    const int64_t u0 = min((int64_t) t * chunk, nunits), u1 = min(u0 + chunk, nunits);
    int64_t sum = 0;
    for (int64_t u = u0; u < u1; u++) sum += unithist[u * nranks + r];
    sums[t] = sum;
    __syncthreads();
    if (t == 0) {
        int64_t run = 0;
        for (int k = 0; k < 256; k++) { const int64_t v = sums[k]; sums[k] = run; run += v; }
        totals[r] = run;
    }
    __syncthreads();
    partial[(int64_t) r * 256 + t] = sums[t];
}

// counts -> write cursors, rank-major output order (rank ascending, then particle ascending)
__global__ void __launch_bounds__(256)
pmb_k_route_cursors(int32_t *unithist, int64_t nunits, int nranks, const int64_t *totals, const int64_t *partial)
{
    const int r = blockIdx.x, t = threadIdx.x;
    int64_t base = 0;
    for (int q = 0; q < r; q++) base += totals[q];
    base += partial[(int64_t) r * 256 + t];
    const int64_t chunk = (nunits + 255) / 256;
    const int64_t u0 = min((int64_t) t * chunk, nunits), u1 = min(u0 + chunk, nunits);
    for (int64_t u = u0; u < u1; u++) {
        const int32_t c = unithist[u * nranks + r];
        unithist[u * nranks + r] = (int32_t) base;
        base += c;
    }
}

template <typename MaskT>
__global__ void __launch_bounds__(ROUTE_BLOCK)
pmb_k_route_fill(int nranks, int64_t npart, int64_t per_unit, const MaskT *masks,
                 const int32_t *unitcursor, int32_t *indices)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t unit = (int64_t) blockIdx.x * (ROUTE_BLOCK / 32) + (threadIdx.x >> 5);
    const int64_t begin = unit * per_unit;
    const int64_t end = min(begin + per_unit, npart);
    if (begin >= end) return;
    int c0 = lane < nranks ? unitcursor[unit * nranks + lane] : 0;
    int c1 = lane + 32 < nranks ? unitcursor[unit * nranks + lane + 32] : 0;
    for (int64_t base = begin; base < end; base += 32 * ROUTE_UNROLL) {
        uint64_t mask[ROUTE_UNROLL];
#pragma unroll
        for (int u = 0; u < ROUTE_UNROLL; u++) {
            const int64_t i = base + u * 32 + lane;
            mask[u] = i < end ? (uint64_t) __ldcs(masks + i) : 0;
        }
#pragma unroll
        for (int u = 0; u < ROUTE_UNROLL; u++) {
            const int64_t i = base + u * 32 + lane;
            unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned) mask[u]);
            unsigned hi = nranks > 32 ? __reduce_or_sync(0xffffffffu, (unsigned) (mask[u] >> 32)) : 0u;
            while (lo) {
                const int r = __ffs(lo) - 1;
                lo &= lo - 1;
                const bool mine = (mask[u] >> r) & 1;
                const unsigned bal = __ballot_sync(0xffffffffu, mine);
                const int cur = __shfl_sync(0xffffffffu, c0, r);
                if (mine) indices[cur + __popc(bal & lt)] = (int32_t) i;
                if (lane == r) c0 += __popc(bal);
            }
            while (hi) {
                const int r = __ffs(hi) - 1;
                hi &= hi - 1;
                const bool mine = (mask[u] >> (r + 32)) & 1;
                const unsigned bal = __ballot_sync(0xffffffffu, mine);
                const int cur = __shfl_sync(0xffffffffu, c1, r);
                if (mine) indices[cur + __popc(bal & lt)] = (int32_t) i;
                if (lane == r) c1 += __popc(bal);
            }
        }
    }
}

static int route_setup(pmb_ctx *ctx, const pmb_decompose_args *a, RouteGeom *g, void **dev_tables)
{
    PMB_REQUIRE(ctx && a, "null argument");
    PMB_REQUIRE(a->ndim >= 1 && a->ndim <= 3, "domain grid must be 1..3 dimensional");
    PMB_REQUIRE(a->nranks >= 1 && a->nranks <= ROUTE_MAXRANKS, "routing supports 1..%d ranks", ROUTE_MAXRANKS);
    PMB_REQUIRE(a->pos_elsize == 4 || a->pos_elsize == 8, "pos must be float32 or float64");
    PMB_REQUIRE(a->npart >= 0 && a->npart < ((int64_t) 1 << 31), "npart must be < 2^31 (domain.py:590)");
    PMB_REQUIRE(a->edges_h && a->domain_assign_h && a->domain_degenerate_h, "null domain tables");
    memset(g, 0, sizeof(*g));
    g->ndim = a->ndim;
    g->periodic = a->periodic;
    g->nranks = a->nranks;
    int ndomains = 1, tot_edges = 0;
    for (int d = 0; d < a->ndim; d++) {
        PMB_REQUIRE(a->nedges[d] >= 2 && a->nedges[d] <= ROUTE_MAXEDGES, "bad number of edges on axis %d", d);
        g->shape[d] = a->nedges[d] - 1;
        g->nedges[d] = a->nedges[d];
        g->scale[d] = a->scale[d];
        g->smoothing[d] = a->smoothing[d];
        {
            // guess of pmb_digitize_near: domains per unit length (0 on a degenerate grid: the fix-up loops decide)
            const double *ed = a->edges_h + tot_edges;
            const double span = ed[a->nedges[d] - 1] - ed[0];
            g->inv_width[d] = span > 0 ? (double) (a->nedges[d] - 1) / span : 0.0;
        }
        ndomains *= g->shape[d];
        tot_edges += a->nedges[d];
    }
    g->ndomains = ndomains;
    {
        static int home = -1;
        if (home < 0) { const char *e = getenv("PMB_ROUTE_HOME"); home = e ? atoi(e) : 1; }
        g->home = home;
    }
    int st = 1;
    for (int d = a->ndim - 1; d >= 0; d--) { g->dstride[d] = st; st *= g->shape[d]; }
    // one small device allocation holds edges | assign | degenerate
    size_t b_edges = sizeof(double) * tot_edges;
    size_t b_assign = (sizeof(int32_t) * ndomains + 7) & ~(size_t) 7;
    size_t b_deg = (sizeof(int16_t) * ndomains + 7) & ~(size_t) 7;
    char *dev;
    PMB_CUDA(cudaMalloc(&dev, b_edges + b_assign + b_deg));
    cudaError_t e = cudaMemcpyAsync(dev, a->edges_h, b_edges, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dev + b_edges, a->domain_assign_h, sizeof(int32_t) * ndomains, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dev + b_edges + b_assign, a->domain_degenerate_h, sizeof(int16_t) * ndomains, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFree(dev); return pmb_cuda_fail(e, "route tables", __FILE__, __LINE__); }
    int o = 0;
    for (int d = 0; d < a->ndim; d++) { g->edges[d] = (const double *) dev + o; o += a->nedges[d]; }
    // one periodic domain on every axis (a single rank): every particle goes to the rank of domain 0
    g->all_trivial = a->periodic ? 1 : 0;
    for (int d = 0; d < a->ndim; d++) if (g->shape[d] != 1) g->all_trivial = 0;
    if (g->all_trivial) {
        const int rank = a->domain_assign_h[0];
        const int deg = (rank >= 0 && rank < ndomains) ? a->domain_degenerate_h[rank] : 0;
        g->const_mask = (!deg && rank >= 0 && rank < ROUTE_MAXRANKS) ? ((uint64_t) 1 << rank) : 0;
    }
    g->assign = (const int32_t *) (dev + b_edges);
    g->degenerate = (const int16_t *) (dev + b_edges + b_assign);
    *dev_tables = dev;
    return PMB_OK;
}

static int ensure(void **buf, size_t *have, size_t need, cudaStream_t s)
{
    if (need <= *have) return PMB_OK;
    if (*buf) { PMB_CUDA(cudaStreamSynchronize(s)); PMB_CUDA(cudaFree(*buf)); *buf = NULL; *have = 0; }
    size_t want = need + (need >> 3) + 256;
    PMB_CUDA(cudaMalloc(buf, want));
    *have = want;
    return PMB_OK;
}

extern "C" int pmb_decompose_count(pmb_ctx *ctx, const pmb_decompose_args *a, int32_t *counts_h, int64_t *ntotal)
{
    PMB_REQUIRE(counts_h && ntotal, "null output");
    RouteGeom g;
    void *tables = NULL;
    PMB_CHECK(route_setup(ctx, a, &g, &tables));
    for (int r = 0; r < a->nranks; r++) counts_h[r] = 0;
    *ntotal = 0;
    ctx->route_npart = a->npart;
    ctx->route_nblocks = 0;
    ctx->route_identity = 0;
    if (a->npart == 0) { cudaFree(tables); return PMB_OK; }
    if (g.all_trivial) {
        // one periodic domain on every axis: every particle goes, once and in order, to the rank of
        // domain 0 (or nowhere if that domain is flagged degenerate): counts are known without
        // looking at a single position and indices = arange(npart)
        cudaFree(tables);
        if (g.const_mask) {
            int rank = 0;
            while (!((g.const_mask >> rank) & 1)) rank++;
            if (rank < a->nranks) {
                counts_h[rank] = (int32_t) a->npart;
                *ntotal = a->npart;
                ctx->route_identity = 1;
            }
        }
        return PMB_OK;
    }

    // units = warps; every unit owns a contiguous, 128-aligned range of particles
    const int wpb = ROUTE_BLOCK / 32;
    const int64_t gran = 32 * ROUTE_UNROLL;
    int64_t nblocks = (a->npart + ROUTE_BLOCK * ROUTE_UNROLL - 1) / (ROUTE_BLOCK * ROUTE_UNROLL);
    const int64_t cap = (int64_t) ctx->sm_count * 8;
    if (nblocks > cap) nblocks = cap;
    int64_t nunits = nblocks * wpb;
    int64_t per_unit = (a->npart + nunits - 1) / nunits;
    per_unit = (per_unit + gran - 1) / gran * gran;
    nunits = (a->npart + per_unit - 1) / per_unit;
    nblocks = (nunits + wpb - 1) / wpb;
    const int64_t nunits_alloc = nblocks * wpb;          // trailing units of the last block are empty
    const size_t b_hist = (sizeof(int32_t) * nunits_alloc * a->nranks + 7) & ~(size_t) 7;
    int rc = ensure(&ctx->route_masks, &ctx->route_masks_bytes, sizeof(uint64_t) * a->npart, ctx->stream);
    if (rc == PMB_OK)
        rc = ensure(&ctx->route_blockhist, &ctx->route_blockhist_bytes,
                    b_hist + sizeof(int64_t) * ROUTE_MAXRANKS * 257, ctx->stream);
    if (rc != PMB_OK) { cudaFree(tables); return rc; }
    int32_t *hist = (int32_t *) ctx->route_blockhist;
    int64_t *totals = (int64_t *) ((char *) ctx->route_blockhist + b_hist);
    int64_t *partial = totals + ROUTE_MAXRANKS;
    int tot_edges = 0;
    for (int d = 0; d < a->ndim; d++) tot_edges += a->nedges[d];
    const size_t smem = sizeof(double) * tot_edges;
    ctx->route_maskbytes = a->nranks <= 8 ? 1 : (a->nranks <= 16 ? 2 : 8);
    // coordinates requested ahead per lane (PMB_ROUTE_UNROLL = 2, 4, 8; measured in profiles/r2_route_unroll.json)
    static int route_unroll = -1;
    // 268 M particles, 2 slabs (B200, ms): 2 -> 4.12, 4 -> 5.01, 8 -> 8.56: registers (48 / 64 / 93), i.e. resident
    // warps, matter more than loads in flight per lane
    if (route_unroll < 0) { const char *e = getenv("PMB_ROUTE_UNROLL"); route_unroll = e ? atoi(e) : 2; }
#define ROUTE_COUNT(ND, MT)                                                                          \
    do {                                                                                             \
        if (route_unroll >= 8)                                                                       \
            pmb_k_route_count<ND, MT, 8><<<(int) nblocks, ROUTE_BLOCK, smem, ctx->stream>>>(          \
                g, a->pos, a->pos_elsize, a->pos_stride0, a->pos_stride1, a->npart, per_unit, (MT *) ctx->route_masks, hist); \
        else if (route_unroll >= 4)                                                                  \
            pmb_k_route_count<ND, MT, 4><<<(int) nblocks, ROUTE_BLOCK, smem, ctx->stream>>>(          \
                g, a->pos, a->pos_elsize, a->pos_stride0, a->pos_stride1, a->npart, per_unit, (MT *) ctx->route_masks, hist); \
        else                                                                                         \
            pmb_k_route_count<ND, MT, 2><<<(int) nblocks, ROUTE_BLOCK, smem, ctx->stream>>>(          \
                g, a->pos, a->pos_elsize, a->pos_stride0, a->pos_stride1, a->npart, per_unit, (MT *) ctx->route_masks, hist); \
    } while (0)
#define ROUTE_COUNT_ND(MT)                                                                           \
    do {                                                                                             \
        if (a->ndim == 1) ROUTE_COUNT(1, MT); else if (a->ndim == 2) ROUTE_COUNT(2, MT); else ROUTE_COUNT(3, MT); \
    } while (0)
    if (ctx->route_maskbytes == 1) ROUTE_COUNT_ND(uint8_t);
    else if (ctx->route_maskbytes == 2) ROUTE_COUNT_ND(uint16_t);
    else ROUTE_COUNT_ND(unsigned long long);
#undef ROUTE_COUNT_ND
#undef ROUTE_COUNT
    ctx->launches++;
    pmb_k_route_scan<<<a->nranks, 256, 0, ctx->stream>>>(hist, nunits_alloc, a->nranks, totals, partial);
    ctx->launches++;
    pmb_k_route_cursors<<<a->nranks, 256, 0, ctx->stream>>>(hist, nunits_alloc, a->nranks, totals, partial);
    ctx->launches++;
    int64_t totals_h[ROUTE_MAXRANKS];
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(totals_h, totals, sizeof(int64_t) * a->nranks, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(tables);
    if (e != cudaSuccess) return pmb_cuda_fail(e, "route count", __FILE__, __LINE__);
    int64_t tot = 0;
    for (int r = 0; r < a->nranks; r++) {
        PMB_REQUIRE(totals_h[r] < ((int64_t) 1 << 31), "more than 2^31 particles routed to rank %d", r);
        counts_h[r] = (int32_t) totals_h[r];
        tot += totals_h[r];
    }
    PMB_REQUIRE(tot < ((int64_t) 1 << 31), "routed particle count overflows int32 offsets (domain.py:590)");
    *ntotal = tot;
    ctx->route_nblocks = (int) nblocks;
    ctx->route_per_block = per_unit;
    return PMB_OK;
}

__global__ void __launch_bounds__(256) pmb_k_iota(int32_t *out, int64_t n)
{
    int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; t < n; t += stride) __stcs(out + t, (int32_t) t);
}

extern "C" int pmb_decompose_identity(pmb_ctx *ctx, int *flag)
{
    PMB_REQUIRE(ctx && flag, "null argument");
    *flag = ctx->route_identity;
    return PMB_OK;
}

extern "C" int pmb_decompose_fill(pmb_ctx *ctx, const pmb_decompose_args *a, int32_t *indices)
{
    PMB_REQUIRE(ctx && a, "null argument");
    PMB_REQUIRE(ctx->route_npart == a->npart, "pmb_decompose_fill must follow pmb_decompose_count with the same particles");
    if (ctx->route_identity) {
        PMB_REQUIRE(indices, "null indices");
        pmb_k_iota<<<pmb_grid(ctx, a->npart, 256, 8), 256, 0, ctx->stream>>>(indices, a->npart);
        PMB_LAUNCH_CHECK(ctx);
        return PMB_OK;
    }
    if (a->npart == 0 || ctx->route_nblocks == 0) return PMB_OK;
    PMB_REQUIRE(indices, "null indices");
    const int64_t per_block = ctx->route_per_block;
    if (ctx->route_maskbytes == 1)
        pmb_k_route_fill<uint8_t><<<ctx->route_nblocks, ROUTE_BLOCK, 0, ctx->stream>>>(
            a->nranks, a->npart, per_block, (const uint8_t *) ctx->route_masks, (const int32_t *) ctx->route_blockhist, indices);
    else if (ctx->route_maskbytes == 2)
        pmb_k_route_fill<uint16_t><<<ctx->route_nblocks, ROUTE_BLOCK, 0, ctx->stream>>>(
            a->nranks, a->npart, per_block, (const uint16_t *) ctx->route_masks, (const int32_t *) ctx->route_blockhist, indices);
    else
        pmb_k_route_fill<unsigned long long><<<ctx->route_nblocks, ROUTE_BLOCK, 0, ctx->stream>>>(
            a->nranks, a->npart, per_block, (const unsigned long long *) ctx->route_masks, (const int32_t *) ctx->route_blockhist, indices);
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}

// ------------------------------------------------------------------ take (gather-pack)
// WORDS > 0: one thread per record of WORDS machine words (no integer division on the hot path);
// WORDS == 0: generic, one thread per word.
template <typename W, int WORDS>
__global__ void __launch_bounds__(256)
pmb_k_take(const W *__restrict__ data, int64_t words, const int32_t *__restrict__ indices, int64_t n, W *__restrict__ out)
{
    int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    if (WORDS > 0) {
        for (; t < n; t += stride) {
            const W *src = data + (int64_t) __ldcs(indices + t) * WORDS;
            W v[WORDS > 0 ? WORDS : 1];
#pragma unroll
            for (int w = 0; w < WORDS; w++) v[w] = __ldcs(src + w);
#pragma unroll
            for (int w = 0; w < WORDS; w++) __stcs(out + t * WORDS + w, v[w]);
        }
    } else {
        const int64_t total = n * words;
        for (; t < total; t += stride) {
            const int64_t j = t / words, w = t - j * words;
            out[t] = data[(int64_t) indices[j] * words + w];
        }
    }
}

template <typename W>
static void take_launch(pmb_ctx *ctx, const void *data, int64_t words, const int32_t *indices, int64_t n, void *out)
{
    const W *d = (const W *) data;
    W *o = (W *) out;
    const int grid = pmb_grid(ctx, n, 256, 8);
    switch (words) {
    case 1: pmb_k_take<W, 1><<<grid, 256, 0, ctx->stream>>>(d, words, indices, n, o); break;
    case 2: pmb_k_take<W, 2><<<grid, 256, 0, ctx->stream>>>(d, words, indices, n, o); break;
    case 3: pmb_k_take<W, 3><<<grid, 256, 0, ctx->stream>>>(d, words, indices, n, o); break;
    case 4: pmb_k_take<W, 4><<<grid, 256, 0, ctx->stream>>>(d, words, indices, n, o); break;
    case 6: pmb_k_take<W, 6><<<grid, 256, 0, ctx->stream>>>(d, words, indices, n, o); break;
    default:
        pmb_k_take<W, 0><<<pmb_grid(ctx, n * words, 256, 8), 256, 0, ctx->stream>>>(d, words, indices, n, o);
    }
}

extern "C" int pmb_take(pmb_ctx *ctx, const void *data, int64_t itemsize, const int32_t *indices, int64_t n, void *out)
{
    PMB_REQUIRE(ctx && itemsize > 0 && n >= 0, "bad take arguments");
    if (n == 0) return PMB_OK;
    PMB_REQUIRE(data && out, "null argument");
    if (!indices) {   // identity layout: take(arange(n)) is a copy
        PMB_CUDA(cudaMemcpyAsync(out, data, (size_t) n * itemsize, cudaMemcpyDeviceToDevice, ctx->stream));
        return PMB_OK;
    }
    const uintptr_t al = (uintptr_t) data | (uintptr_t) out | (uintptr_t) itemsize;
    if ((al & 7) == 0) take_launch<unsigned long long>(ctx, data, itemsize / 8, indices, n, out);
    else if ((al & 3) == 0) take_launch<unsigned int>(ctx, data, itemsize / 4, indices, n, out);
    else take_launch<unsigned char>(ctx, data, itemsize, indices, n, out);
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}

// ------------------------------------------------------------------ ghost reduction (gather 'sum')
// numpy.bincount(indices, weights) adds weights[j] into out[indices[j]] for ascending j, in float64,
// starting from 0.0.  `indices` is grouped by rank and a particle appears at most once per rank
// segment, so one streaming pass per segment (in rank order) performs exactly those additions in
// exactly that order without atomics: acc[indices[j]] = acc[indices[j]] + data[j].
template <bool ASSIGN>
__global__ void __launch_bounds__(256)
pmb_k_gather_pass(const void *__restrict__ seg, int data_elsize, int ncomp, const int32_t *__restrict__ indices,
                  int64_t begin, int64_t end, double *__restrict__ acc)
{
    // seg points at the first record of the segment [begin, end) of `indices`
    int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int64_t total = (end - begin) * ncomp;
    for (; t < total; t += stride) {
        const int64_t j = ncomp == 1 ? t : t / ncomp;
        const int c = ncomp == 1 ? 0 : (int) (t - j * ncomp);
        const int64_t o = (int64_t) __ldcs(indices + begin + j) * ncomp + c;
        const double v = data_elsize == 8 ? __ldcs((const double *) seg + t) : (double) __ldcs((const float *) seg + t);
        acc[o] = ASSIGN ? 0.0 + v : acc[o] + v;
    }
}

__global__ void pmb_k_f64_to_f32(const double *__restrict__ in, float *__restrict__ out, int64_t n)
{
    int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; t < n; t += stride) out[t] = (float) in[t];
}

// identity layout (indices = arange): bincount degenerates to out[j] = 0.0 + data[j]
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256) pmb_k_gather_identity(const Tin *__restrict__ in, Tout *__restrict__ out, int64_t n)
{
    int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; t < n; t += stride) __stcs(out + t, (Tout) (0.0 + (double) __ldcs(in + t)));
}

extern "C" int pmb_gather_sum(pmb_ctx *ctx, const void *data, int data_elsize, int ncomp, const int32_t *indices,
                              const int64_t *offsets_h, int nranks, int64_t nout, void *out, int out_elsize)
{
    PMB_REQUIRE(ctx && offsets_h && ncomp >= 1, "bad gather arguments");
    PMB_REQUIRE(nranks >= 1 && nranks <= ROUTE_MAXRANKS, "gather supports 1..%d ranks", ROUTE_MAXRANKS);
    PMB_REQUIRE(data_elsize == 4 || data_elsize == 8, "float32/float64 only");
    PMB_REQUIRE(offsets_h[nranks] == 0 || data, "null data");
    const void *segs[ROUTE_MAXRANKS];
    for (int r = 0; r < nranks; r++) segs[r] = (const char *) data + offsets_h[r] * ncomp * data_elsize;
    return pmb_gather_sum_segments(ctx, segs, data_elsize, ncomp, indices, offsets_h, nranks, nout, out, out_elsize);
}

extern "C" int pmb_gather_sum_segments(pmb_ctx *ctx, const void *const *segments_h, int data_elsize, int ncomp,
                                       const int32_t *indices, const int64_t *offsets_h, int nranks, int64_t nout,
                                       void *out, int out_elsize)
{
    PMB_REQUIRE(ctx && offsets_h && segments_h && ncomp >= 1 && nout >= 0, "bad gather arguments");
    PMB_REQUIRE(nranks >= 1 && nranks <= ROUTE_MAXRANKS, "gather supports 1..%d ranks", ROUTE_MAXRANKS);
    PMB_REQUIRE((data_elsize == 4 || data_elsize == 8) && (out_elsize == 4 || out_elsize == 8), "float32/float64 only");
    if (nout == 0) return PMB_OK;
    PMB_REQUIRE(out, "null out");
    for (int r = 0; r < nranks; r++)
        PMB_REQUIRE(offsets_h[r + 1] <= offsets_h[r] || segments_h[r], "null data segment %d", r);
    const int64_t n = nout * ncomp;
    if (!indices && offsets_h[nranks] != 0) {
        // identity layout: one segment that lists every row once, in order
        int only = -1;
        for (int r = 0; r < nranks; r++)
            if (offsets_h[r + 1] > offsets_h[r]) { PMB_REQUIRE(only < 0, "identity gather needs a single segment"); only = r; }
        PMB_REQUIRE(offsets_h[nranks] == nout, "identity gather needs exactly one record per output row");
        const void *data = segments_h[only];
        const int grid = pmb_grid(ctx, n, 256, 8);
        if (data_elsize == 8 && out_elsize == 8)
            pmb_k_gather_identity<<<grid, 256, 0, ctx->stream>>>((const double *) data, (double *) out, n);
        else if (data_elsize == 8)
            pmb_k_gather_identity<<<grid, 256, 0, ctx->stream>>>((const double *) data, (float *) out, n);
        else if (out_elsize == 8)
            pmb_k_gather_identity<<<grid, 256, 0, ctx->stream>>>((const float *) data, (double *) out, n);
        else
            pmb_k_gather_identity<<<grid, 256, 0, ctx->stream>>>((const float *) data, (float *) out, n);
        PMB_LAUNCH_CHECK(ctx);
        return PMB_OK;
    }
    PMB_REQUIRE(offsets_h[nranks] == 0 || indices, "null indices");
    double *acc = (double *) out;
    if (out_elsize == 4) {
        void *tmp;
        PMB_CHECK(pmb_scratch(ctx, sizeof(double) * n, &tmp));
        acc = (double *) tmp;
    }
    // A first segment that holds nout entries lists every particle exactly once (ascending, unique,
    // < nout): it can assign `0.0 + data[j]` instead of read-modify-write, and no memset is needed.
    bool first = true;
    for (int r = 0; r < nranks; r++) {
        const int64_t b = offsets_h[r], e = offsets_h[r + 1];
        if (e <= b) continue;
        const bool assign = first && (e - b) == nout;
        if (first && !assign) PMB_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * n, ctx->stream));
        first = false;
        if (assign)
            pmb_k_gather_pass<true><<<pmb_grid(ctx, (e - b) * ncomp, 256, 8), 256, 0, ctx->stream>>>(
                segments_h[r], data_elsize, ncomp, indices, b, e, acc);
        else
            pmb_k_gather_pass<false><<<pmb_grid(ctx, (e - b) * ncomp, 256, 8), 256, 0, ctx->stream>>>(
                segments_h[r], data_elsize, ncomp, indices, b, e, acc);
        PMB_LAUNCH_CHECK(ctx);
    }
    if (first) PMB_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * n, ctx->stream));
    if (out_elsize == 4) {
        pmb_k_f64_to_f32<<<pmb_grid(ctx, n, 256, 8), 256, 0, ctx->stream>>>(acc, (float *) out, n);
        PMB_LAUNCH_CHECK(ctx);
    }
    return PMB_OK;
}

// out[indices[j]] += data[j] for the segments of every rank but `skip_rank`, in rank order: the ghosts that
// came back through the reverse alltoallv are added to columns that already hold the rank's own results
// (pmb_readout_multi_gather).  out is float64, (nout, ncomp) rows.
extern "C" int pmb_gather_add_segments(pmb_ctx *ctx, const void *const *segments_h, int data_elsize, int ncomp,
                                       const int32_t *indices, const int64_t *offsets_h, int nranks, int skip_rank,
                                       int64_t nout, void *out)
{
    PMB_REQUIRE(ctx && offsets_h && segments_h && ncomp >= 1 && nout >= 0, "bad gather arguments");
    PMB_REQUIRE(nranks >= 1 && nranks <= ROUTE_MAXRANKS, "gather supports 1..%d ranks", ROUTE_MAXRANKS);
    PMB_REQUIRE(data_elsize == 4 || data_elsize == 8, "float32/float64 only");
    if (nout == 0) return PMB_OK;
    PMB_REQUIRE(out && indices, "null out / indices");
    for (int r = 0; r < nranks; r++) {
        if (r == skip_rank) continue;
        const int64_t b = offsets_h[r], e = offsets_h[r + 1];
        if (e <= b) continue;
        PMB_REQUIRE(segments_h[r], "null data segment %d", r);
        pmb_k_gather_pass<false><<<pmb_grid(ctx, (e - b) * ncomp, 256, 8), 256, 0, ctx->stream>>>(
            segments_h[r], data_elsize, ncomp, indices, b, e, (double *) out);
        PMB_LAUNCH_CHECK(ctx);
    }
    return PMB_OK;
}
