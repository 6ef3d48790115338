// pmb_pull.cuh -- the deterministic paint at size: particles sorted by the cell of their FIRST stencil point,
// every mesh cell then sums its own contributions, in ascending particle number.
//
// The reference paints with one sequential loop over the particles (_window_imp.c, _window_tuned_*.h): cell c
// receives `*(FLOAT *) p += f` once per (particle, stencil point) that lands on it, in particle order, a
// particle's own points in C order.  Round 1 reproduced that by expanding every particle into support^3
// (cell, value) pairs and radix-sorting the PAIRS (8 / 27 / 64 x the data; CIC 2.6 G particles/s).  Here only the
// particles are sorted:
//
//   keys     key = linear number of the cell of the particle's first stencil point (per axis: wrapped into the
//            period when the canvas covers the whole period, else relative to the canvas with room for points
//            that enter it from below), value = particle number                       1 pass over the positions
//   sort     stable LSD radix sort of the (key, number) pairs on the bits of the key space (cub, library):
//            inside a cell the particle numbers ascend
//   runs     (first, last + 1) of every non-empty cell, from the boundaries of the sorted keys
//   records  (x, y, z, number) of the sorted particles as 32-byte records (+ the sorted mass column)
//   pull     one thread per mesh cell: the support^3 cells whose particles can reach it are the neighbours at
//            offsets (ka, kb, kc) < support below it; their runs are merged by particle number (ties -- the same
//            particle reaching the cell twice through a period shorter than its stencil cannot happen here, such
//            canvases take the pairs path -- would break towards the lower offset, the reference's point order)
//            and summed as acc = (T) ((double) acc + f), f = ((V0 * m) * V1) * V2 evaluated like every other
//            tuned kernel.  No atomics, every mesh cell is read and written once, coalesced.
//
// Bit-identical to the reference and to the pairs path for every tuned window (nnb, cic, tsc, pcs, and their
// gradient windows) on 3-D canvases; everything else (run-time supports, per-particle hsml, 1-D / 2-D, periods
// shorter than the stencil, key spaces >= 2^31) keeps the pairs path.
#pragma once
#include <cub/cub.cuh>

#include "pmb_sched.cuh"

struct PmbPullGeom {
    int full[3];           // 1: the canvas covers the whole period of the axis (keys wrap), 0: keys relative to the canvas
    int E[3];              // key extent per axis: period (full) or size + S - 1
};

// first stencil index of a coordinate -> key coordinate in [0, E), or -1 when no stencil point can land on the canvas
template <int FAM>
__device__ __forceinline__ int pmb_pull_base(int I0, int full, int per, int sz)
{
    if (full) return pmb_wrap32(I0, per);
    int cb = I0;
    if (per > 0) {
        cb = pmb_wrap32(I0, per);
        if (cb > per - FAM) cb -= per;         // the stencil runs across the period: it enters the canvas from below
    }
    if (cb <= -FAM || cb >= sz) return -1;
    return cb + FAM - 1;
}

template <typename KeyT, int FAM>
__global__ void __launch_bounds__(256)
pmb_k_pull_keys(PmbGeom g, PmbPullGeom pg, PmbParticles p, int64_t npart, int pcsfix, KeyT *keys, uint32_t *ids)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; i < npart; i += stride) {
        double x[3];
        pmb_load_pos<3>(p, i, x);
        int64_t key = 0;
        bool ok = true;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double X = pmb_gridpos(x[d], g.scale[d], g.translate[d]);
            int I[FAM];
            double V[FAM];
            pmb_axis_tuned<FAM>(X, g.order[d], g.scale[d], pcsfix, I, V);
            const int b = pmb_pull_base<FAM>(I[0], pg.full[d], (int) g.period[d], (int) g.size[d]);
            ok = ok && b >= 0;
            key = key * pg.E[d] + b;
        }
        keys[i] = ok ? (KeyT) key : (KeyT) ~(KeyT) 0;
        ids[i] = (uint32_t) i;
    }
}

// runs of equal keys: se[key] = (first, last + 1); cells without particles keep (0, 0) from the memset
template <typename KeyT>
__global__ void __launch_bounds__(256)
pmb_k_pull_runs(const KeyT *__restrict__ keys, int64_t n, int64_t nkeys, uint2 *se)
{
    int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; j < n; j += stride) {
        const KeyT k = keys[j];
        if ((int64_t) k >= nkeys) continue;           // particles that reach no cell of the canvas sort last
        if (j == 0 || keys[j - 1] != k) se[k].x = (uint32_t) j;
        if (j + 1 == n || keys[j + 1] != k) se[k].y = (uint32_t) (j + 1);
    }
}

// sorted records: (x, y, z, particle number) as one 32-byte store, and the mass column in the same order
__global__ void __launch_bounds__(256)
pmb_k_pull_records(PmbParticles p, const uint32_t *__restrict__ ids, int64_t n, double *__restrict__ recs, double *__restrict__ smass)
{
    int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; j < n; j += stride) {
        const int64_t i = ids[j];
        double x[3];
        pmb_load_pos<3>(p, i, x);
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(recs + 4 * j), "d"(x[0]), "d"(x[1]), "d"(x[2]),
                     "d"(__longlong_as_double((long long) i)) : "memory");
        if (smass) smass[j] = pmb_load_mass(p, i);
    }
}

template <int FAM>
__device__ __forceinline__ double pmb_pull_pick(const double *V, int k)
{
    double v = V[0];
#pragma unroll
    for (int s = 1; s < FAM; s++) if (k == s) v = V[s];
    return v;
}

// contribution of sorted particle j to the cell at stencil offsets (ka, kb, kc) above its first point
struct PmbPullRec { double x0, x1, x2, idb; };

__device__ __forceinline__ PmbPullRec pmb_pull_load(const double *__restrict__ recs, uint32_t j)
{
    PmbPullRec r;
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.x0), "=d"(r.x1), "=d"(r.x2), "=d"(r.idb) : "l"(recs + 4 * (int64_t) j));
    return r;
}

// contribution of a sorted particle (its record already loaded) to the cell at stencil offsets (ka, kb, kc)
template <int FAM>
__device__ __forceinline__ double pmb_pull_value_rec(const PmbGeom &g, int pcsfix, const PmbPullRec &r, double m, int ka, int kb, int kc)
{
    int I[FAM];
    double V0[FAM], V1[FAM], V2[FAM];
    pmb_axis_tuned<FAM>(pmb_gridpos(r.x0, g.scale[0], g.translate[0]), g.order[0], g.scale[0], pcsfix, I, V0);
    pmb_axis_tuned<FAM>(pmb_gridpos(r.x1, g.scale[1], g.translate[1]), g.order[1], g.scale[1], pcsfix, I, V1);
    pmb_axis_tuned<FAM>(pmb_gridpos(r.x2, g.scale[2], g.translate[2]), g.order[2], g.scale[2], pcsfix, I, V2);
    return pmb_paint_value(true, m, pmb_pull_pick<FAM>(V0, ka), pmb_pull_pick<FAM>(V1, kb), pmb_pull_pick<FAM>(V2, kc));
}

template <int FAM>
__device__ __forceinline__ double pmb_pull_value(const PmbGeom &g, int pcsfix, const double *__restrict__ recs,
                                                 const double *__restrict__ smass, double mass_scalar, uint32_t j,
                                                 int ka, int kb, int kc)
{
    double x0, x1, x2, idb;
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(x0), "=d"(x1), "=d"(x2), "=d"(idb) : "l"(recs + 4 * (int64_t) j));
    (void) idb;
    const double m = smass ? smass[j] : mass_scalar;
    int I[FAM];
    double V0[FAM], V1[FAM], V2[FAM];
    pmb_axis_tuned<FAM>(pmb_gridpos(x0, g.scale[0], g.translate[0]), g.order[0], g.scale[0], pcsfix, I, V0);
    pmb_axis_tuned<FAM>(pmb_gridpos(x1, g.scale[1], g.translate[1]), g.order[1], g.scale[1], pcsfix, I, V1);
    pmb_axis_tuned<FAM>(pmb_gridpos(x2, g.scale[2], g.translate[2]), g.order[2], g.scale[2], pcsfix, I, V2);
    return pmb_paint_value(true, m, pmb_pull_pick<FAM>(V0, ka), pmb_pull_pick<FAM>(V1, kb), pmb_pull_pick<FAM>(V2, kc));
}

__device__ __forceinline__ uint32_t pmb_pull_id(const double *__restrict__ recs, uint32_t j)
{
    return (uint32_t) __double_as_longlong(__ldg(recs + 4 * (int64_t) j + 3));
}

// the run of neighbour q = (ka * FAM + kb) * FAM + kc of cell (c0, c1, c2)
template <int FAM>
__device__ __forceinline__ uint2 pmb_pull_run(const PmbPullGeom &pg, const uint2 *__restrict__ se, int c0, int c1, int c2, int q)
{
    const int ka = q / (FAM * FAM), kb = (q / FAM) % FAM, kc = q % FAM;
    int b0 = pg.full[0] ? c0 - ka : c0 - ka + FAM - 1;
    if (pg.full[0] && b0 < 0) b0 += pg.E[0];
    int b1 = pg.full[1] ? c1 - kb : c1 - kb + FAM - 1;
    if (pg.full[1] && b1 < 0) b1 += pg.E[1];
    int b2 = pg.full[2] ? c2 - kc : c2 - kc + FAM - 1;
    if (pg.full[2] && b2 < 0) b2 += pg.E[2];
    return __ldg(se + (((int64_t) b0 * pg.E[1] + b1) * pg.E[2] + b2));
}

// Weight records (PRE): per sorted particle RS doubles -- particle number, V0[k] * m, V1[k], V2[k] (k < FAM), padding to
// whole 32-byte sectors -- written once by pmb_k_pull_records_w.  The pull kernel visits every particle FAM^3 times;
// the position records make it evaluate 3 * FAM weights at every visit (CIC 8 x, PCS 64 x redundant arithmetic).
// MEASURED (512^3, B200): slower -- CIC pull 12.4 ms against 8.5, PCS 160 against 115: the wider records (64 - 128 bytes,
// each read FAM^3 times) double the L1 / L2 traffic the kernel is really bound by.  Kept behind PMB_PULL_WEIGHTS=1.
template <int FAM>
struct PmbPullCap {
    static constexpr bool SMEM = FAM <= 2;
    static constexpr int CAP = FAM == 1 ? 8 : (FAM == 2 ? 16 : (FAM == 3 ? 56 : 112));
};

template <int FAM>
struct PmbPullW { static constexpr int RS = FAM == 1 ? 4 : (FAM == 2 ? 8 : (FAM == 3 ? 12 : 16)); };

template <int FAM>
__global__ void __launch_bounds__(256)
pmb_k_pull_records_w(PmbGeom g, int pcsfix, PmbParticles p, const uint32_t *__restrict__ ids, int64_t n, double *__restrict__ recs)
{
    constexpr int RS = PmbPullW<FAM>::RS;
    int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; j < n; j += stride) {
        const int64_t i = ids[j];
        double x[3];
        pmb_load_pos<3>(p, i, x);
        const double m = pmb_load_mass(p, i);
        int I[FAM];
        double V0[FAM], V1[FAM], V2[FAM];
        pmb_axis_tuned<FAM>(pmb_gridpos(x[0], g.scale[0], g.translate[0]), g.order[0], g.scale[0], pcsfix, I, V0);
        pmb_axis_tuned<FAM>(pmb_gridpos(x[1], g.scale[1], g.translate[1]), g.order[1], g.scale[1], pcsfix, I, V1);
        pmb_axis_tuned<FAM>(pmb_gridpos(x[2], g.scale[2], g.translate[2]), g.order[2], g.scale[2], pcsfix, I, V2);
        double w[RS];
#pragma unroll
        for (int k = 0; k < RS; k++) w[k] = 0.0;
        w[0] = __longlong_as_double((long long) i);
#pragma unroll
        for (int k = 0; k < FAM; k++) { w[1 + k] = V0[k] * m; w[1 + FAM + k] = V1[k]; w[1 + 2 * FAM + k] = V2[k]; }
#pragma unroll
        for (int k = 0; k < RS; k += 4)
            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(recs + RS * j + k), "d"(w[k]), "d"(w[k + 1]), "d"(w[k + 2]), "d"(w[k + 3]) : "memory");
    }
}

// (particle number, value) of the contribution of sorted particle j at stencil offsets (ka, kb, kc)
template <int FAM, bool PRE>
__device__ __forceinline__ void pmb_pull_contrib(const PmbGeom &g, int pcsfix, const double *__restrict__ recs,
                                                 const double *__restrict__ smass, double mass_scalar, uint32_t j,
                                                 int ka, int kb, int kc, uint32_t &id, double &val)
{
    if (PRE) {
        const double *r = recs + (int64_t) PmbPullW<FAM>::RS * j;
        id = (uint32_t) __double_as_longlong(__ldg(r));
        val = (__ldg(r + 1 + ka) * __ldg(r + 1 + FAM + kb)) * __ldg(r + 1 + 2 * FAM + kc);      // ((V0 * m) * V1) * V2
    } else {
        const PmbPullRec rr = pmb_pull_load(recs, j);
        id = (uint32_t) __double_as_longlong(rr.idb);
        val = pmb_pull_value_rec<FAM>(g, pcsfix, rr, smass ? smass[j] : mass_scalar, ka, kb, kc);
    }
}

// One thread per mesh cell.  Common case (at most CAP contributions): gather (particle number, value) of every
// contribution -- neighbours from the highest offset down, which for particles kept in lattice order is already
// nearly ascending particle number -- insertion-sort by number, add in order.  Cells with more contributions
// (clustered particles) merge their sorted runs head by head.  Both paths add in ascending particle number.
template <typename MeshT, int FAM, bool PRE>
__global__ void __launch_bounds__(128, (FAM <= 2 ? 8 : 5))
pmb_k_pull(PmbGeom g, PmbPullGeom pg, int pcsfix, const uint2 *__restrict__ se, const double *__restrict__ recs,
           const double *__restrict__ smass, double mass_scalar, char *mesh)
{
    constexpr int NR = FAM * FAM * FAM;
    // (particle number, value) of a cell's contributions live in SHARED memory, entry k of thread t at [k][t]: as
    // per-thread local arrays they thrashed L1 (1024 threads x 384 bytes per SM; ncu: 28 % of the local loads missed,
    // 5 L2 requests per cell) and spilled to DRAM for the wider windows
    // (measured at 512^3: CIC 8.1 -> 7.5 ms).  The wider windows keep local arrays: their lists (56 / 112 entries) in
    // shared memory leave 3 / 1 resident CTAs per SM and the kernel twice / three times slower (TSC 84 ms against 46).
    constexpr int CAP = PmbPullCap<FAM>::CAP;
    constexpr bool SM = PmbPullCap<FAM>::SMEM;
    constexpr int ST = SM ? 128 : 1;
    extern __shared__ double pull_smem[];
    uint32_t lids[SM ? 1 : CAP];
    double lvals[SM ? 1 : CAP];
    double *vals = SM ? pull_smem + threadIdx.x : lvals;
    uint32_t *ids = SM ? (uint32_t *) (pull_smem + CAP * 128) + threadIdx.x : lids;
    constexpr int RW = PRE ? PmbPullW<FAM>::RS : 4;          // words per record; the particle number is word 0 (PRE) or 3
    constexpr int IDW = PRE ? 0 : 3;
    // a warp owns a block of 2 x 2 x 8 cells (8 along the contiguous axis: two full sectors per mesh row): the FAM^3
    // neighbourhoods of its lanes overlap, so most of their run descriptors and records are the SAME sectors and the
    // requests of one instruction coalesce (one cell per lane along z alone: 5 L2 requests per cell, each record
    // fetched 2 x from DRAM).  Blocks are numbered z-fastest: the four warps of a CTA work on neighbouring blocks.
    const uint32_t nt1 = (uint32_t) ((g.size[1] + 1) >> 1), nt2 = (uint32_t) ((g.size[2] + 7) >> 3);
    const uint32_t nblk = (uint32_t) ((g.size[0] + 1) >> 1) * nt1 * nt2;       // < 2^31 / 8
    const uint32_t wstride = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint32_t blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; blk < nblk; blk += wstride) {
        const uint32_t t01 = blk / nt2;
        const int c2 = (int) ((blk - t01 * nt2) << 3) + (lane & 7);
        const uint32_t t0 = t01 / nt1;
        const int c1 = (int) ((t01 - t0 * nt1) << 1) + ((lane >> 3) & 1);
        const int c0 = (int) (t0 << 1) + (lane >> 4);
        if (c0 >= (int) g.size[0] || c1 >= (int) g.size[1] || c2 >= (int) g.size[2]) continue;
        int cnt = 0;
        uint32_t total = 0;
        // one (ka, kb) row of neighbours at a time: its FAM run descriptors are requested together
#pragma unroll 1
        for (int row = FAM * FAM - 1; row >= 0; row--) {
            uint2 r[FAM];
#pragma unroll
            for (int kc = 0; kc < FAM; kc++) r[kc] = pmb_pull_run<FAM>(pg, se, c0, c1, c2, row * FAM + kc);
#pragma unroll
            for (int kc = FAM - 1; kc >= 0; kc--) {
                total += r[kc].y - r[kc].x;
                for (uint32_t j = r[kc].x; j < r[kc].y && cnt < CAP; j++) {
                    uint32_t id;
                    double v;
                    pmb_pull_contrib<FAM, PRE>(g, pcsfix, recs, smass, mass_scalar, j, row / FAM, row % FAM, kc, id, v);
                    ids[cnt * ST] = id;
                    vals[cnt * ST] = v;
                    cnt++;
                }
            }
        }
        if (!total) continue;
        MeshT *cell = (MeshT *) (mesh + c0 * g.strides[0] + c1 * g.strides[1] + c2 * g.strides[2]);
        MeshT acc = *cell;
        if (total <= (uint32_t) CAP) {
            for (int k = 1; k < cnt; k++) {
                const uint32_t id = ids[k * ST];
                const double v = vals[k * ST];
                int m = k - 1;
                while (m >= 0 && ids[m * ST] > id) { ids[(m + 1) * ST] = ids[m * ST]; vals[(m + 1) * ST] = vals[m * ST]; m--; }
                ids[(m + 1) * ST] = id; vals[(m + 1) * ST] = v;
            }
            for (int k = 0; k < cnt; k++) acc = (MeshT) ((double) acc + vals[k * ST]);
        } else {
            // merge the runs by particle number; on equal numbers the lower offset goes first
            uint32_t cur[NR], end[NR], head[NR];
#pragma unroll 1
            for (int q = 0; q < NR; q++) {
                const uint2 r = pmb_pull_run<FAM>(pg, se, c0, c1, c2, q);
                cur[q] = r.x; end[q] = r.y;
                head[q] = r.x < r.y ? (uint32_t) __double_as_longlong(__ldg(recs + (int64_t) RW * r.x + IDW)) : 0xFFFFFFFFu;
            }
            for (;;) {
                int best = -1;
                uint32_t bid = 0xFFFFFFFFu;
#pragma unroll 1
                for (int q = 0; q < NR; q++) {
                    const bool take = cur[q] < end[q] && (best < 0 || head[q] < bid);
                    if (take) { best = q; bid = head[q]; }
                }
                if (best < 0) break;
                const uint32_t j = cur[best];
                uint32_t id;
                double v;
                pmb_pull_contrib<FAM, PRE>(g, pcsfix, recs, smass, mass_scalar, j, best / (FAM * FAM), (best / FAM) % FAM, best % FAM, id, v);
                acc = (MeshT) ((double) acc + v);
                cur[best] = j + 1;
                head[best] = j + 1 < end[best] ? (uint32_t) __double_as_longlong(__ldg(recs + (int64_t) RW * (j + 1) + IDW)) : 0xFFFFFFFFu;
            }
        }
        *cell = acc;
    }
}

// ---- host side ------------------------------------------------------------------------------------------
// *done = false: this canvas / input takes the pairs path
template <typename MeshT, int FAM>
static int pmb_pull_paint_fam(pmb_ctx *ctx, const pmb_resample_args *a, const PmbGeom &g, const PmbParticles &p, bool *done)
{
    *done = false;
    PmbPullGeom pg;
    int64_t nkeys = 1;
    for (int d = 0; d < 3; d++) {
        const int64_t per = g.period[d], sz = g.size[d];
        if (per > 0 && sz == per) {
            if (per < FAM) return PMB_OK;                 // a stencil longer than the period lands on a cell twice
            pg.full[d] = 1; pg.E[d] = (int) per;
        } else {
            if (per > 0 && sz + FAM - 1 > per) return PMB_OK;
            pg.full[d] = 0; pg.E[d] = (int) (sz + FAM - 1);
        }
        nkeys *= pg.E[d];
        if (nkeys >= ((int64_t) 1 << 31) - 1) return PMB_OK;
    }
    const int64_t n = a->npart;
    if (n >= ((int64_t) 1 << 31)) return PMB_OK;
    int bits = 1;
    while (((int64_t) 1 << bits) <= nkeys) bits++;        // the all-ones key of dropped particles stays above every real key
    typedef uint32_t KeyT;
    size_t temp = 0;
    PMB_CUDA(cub::DeviceRadixSort::SortPairs(NULL, temp, (KeyT *) NULL, (KeyT *) NULL, (uint32_t *) NULL, (uint32_t *) NULL,
                                             (int) n, 0, bits, ctx->stream));
    auto up = [](size_t b) { return (b + 255) & ~(size_t) 255; };
    const size_t b_keys = up(sizeof(KeyT) * n), b_ids = up(sizeof(uint32_t) * n), b_se = up(sizeof(uint2) * (size_t) nkeys);
    const size_t b_temp = up(temp);
    // weight records (3 * FAM + 1 doubles per particle, padded) when they fit beside everything else, else position
    // records (32 bytes) whose weights are evaluated at every visit
    size_t free_b = 0, total_b = 0;
    PMB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t fixed = 2 * b_keys + 2 * b_ids + b_temp + b_se + 256;
    const size_t b_wrecs = up(sizeof(double) * PmbPullW<FAM>::RS * (size_t) n);
    const bool pre = pmb_env_flag("PMB_PULL_WEIGHTS", 0) && fixed + b_wrecs + ((size_t) 2 << 30) < free_b + ctx->scratch_bytes;
    const size_t b_recs = pre ? b_wrecs : up(32 * (size_t) n), b_mass = (a->mass && !pre) ? up(sizeof(double) * n) : 0;
    // keys in | keys out | ids in | ids out | cub temp | runs | records | mass   (records overlay nothing: simple and safe)
    const size_t total = fixed + b_recs + b_mass;
    void *ws = NULL;
    if (pmb_scratch(ctx, total, &ws) != PMB_OK) { cudaGetLastError(); return PMB_OK; }
    char *b = (char *) ws;
    KeyT *k0 = (KeyT *) b, *k1 = (KeyT *) (b + b_keys);
    uint32_t *i0 = (uint32_t *) (b + 2 * b_keys), *i1 = (uint32_t *) (b + 2 * b_keys + b_ids);
    void *tmp = b + 2 * b_keys + 2 * b_ids;
    uint2 *se = (uint2 *) (b + 2 * b_keys + 2 * b_ids + b_temp);
    double *recs = (double *) ((char *) se + b_se);
    double *smass = (a->mass && !pre) ? (double *) ((char *) recs + b_recs) : NULL;
    pmb_k_pull_keys<KeyT, FAM><<<pmb_grid(ctx, n, 256, 8), 256, 0, ctx->stream>>>(g, pg, p, n, a->pcs_gradient_scale_fix, k0, i0);
    PMB_LAUNCH_CHECK(ctx);
    PMB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, temp, k0, k1, i0, i1, (int) n, 0, bits, ctx->stream));
    ctx->launches += 4;
    PMB_CUDA(cudaMemsetAsync(se, 0, sizeof(uint2) * (size_t) nkeys, ctx->stream));
    pmb_k_pull_runs<KeyT><<<pmb_grid(ctx, n, 256, 8), 256, 0, ctx->stream>>>(k1, n, nkeys, se);
    PMB_LAUNCH_CHECK(ctx);
    const int64_t ncell = g.size[0] * g.size[1] * g.size[2];
    const size_t smem = PmbPullCap<FAM>::SMEM ? (size_t) PmbPullCap<FAM>::CAP * 128 * (sizeof(double) + sizeof(uint32_t)) : 0;
    {
        static bool attr_done = false;
        if (!attr_done) {
            PMB_CUDA(cudaFuncSetAttribute(pmb_k_pull<MeshT, FAM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            PMB_CUDA(cudaFuncSetAttribute(pmb_k_pull<MeshT, FAM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            attr_done = true;
        }
    }
    if (pre) {
        pmb_k_pull_records_w<FAM><<<pmb_grid(ctx, n, 256, 8), 256, 0, ctx->stream>>>(g, a->pcs_gradient_scale_fix, p, i1, n, recs);
        PMB_LAUNCH_CHECK(ctx);
        pmb_k_pull<MeshT, FAM, true><<<pmb_grid(ctx, ncell, 128, 16), 128, smem, ctx->stream>>>(g, pg, a->pcs_gradient_scale_fix, se, recs, NULL,
                                                                                               a->mass_scalar, (char *) a->mesh);
    } else {
        pmb_k_pull_records<<<pmb_grid(ctx, n, 256, 8), 256, 0, ctx->stream>>>(p, i1, n, recs, smass);
        PMB_LAUNCH_CHECK(ctx);
        pmb_k_pull<MeshT, FAM, false><<<pmb_grid(ctx, ncell, 128, 16), 128, smem, ctx->stream>>>(g, pg, a->pcs_gradient_scale_fix, se, recs, smass,
                                                                                                a->mass_scalar, (char *) a->mesh);
    }
    PMB_LAUNCH_CHECK(ctx);
    if (total > ((size_t) 40 << 30)) {
        // a large workspace goes back to the device instead of staying in the context's scratch
        PMB_CUDA(cudaStreamSynchronize(ctx->stream));
        PMB_CUDA(cudaFree(ctx->scratch));
        ctx->scratch = NULL;
        ctx->scratch_bytes = 0;
    }
    *done = true;
    return PMB_OK;
}

template <typename MeshT>
static int pmb_pull_paint(pmb_ctx *ctx, const pmb_resample_args *a, const PmbGeom &g, const PmbParticles &p, int fam, bool *done)
{
    *done = false;
    if (a->ndim != 3 || fam < 1 || fam > 4 || !pmb_env_flag("PMB_PULL", 1)) return PMB_OK;
    if (a->npart < pmb_env_flag("PMB_PULL_MIN", 1 << 15)) return PMB_OK;
    switch (fam) {
    case 1: return pmb_pull_paint_fam<MeshT, 1>(ctx, a, g, p, done);
    case 2: return pmb_pull_paint_fam<MeshT, 2>(ctx, a, g, p, done);
    case 3: return pmb_pull_paint_fam<MeshT, 3>(ctx, a, g, p, done);
    default: return pmb_pull_paint_fam<MeshT, 4>(ctx, a, g, p, done);
    }
}
