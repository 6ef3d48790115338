// pmb_sched.cuh -- chunk-scheduled paint / readout kernels for the tuned windows on 3-D meshes.
//
// Why: one thread per particle in a grid-stride loop is only as cache-friendly as the particle
// order AND as the lock-step of the CTAs.  PM simulations keep particles in (displaced) lattice
// order, so a 256-particle chunk is a compact z-segment, but chunks that touch the SAME mesh rows
// (lattice planes x and x+1) are a whole plane of particles apart in memory, and a CTA that lags one
// grid-stride iteration lags by 300k particles.  ncu at 1024^3 (profiles/): the atomic paint
// re-fetched every mesh sector ~4x and wrote it back ~3.8x (93 GB of DRAM traffic for 34 GB of
// algorithmic bytes); the gather re-read the mesh 2.5x.
//
// What:
// (1) CHUNKS + DYNAMIC TICKETS.  Work is handed out in 256-particle chunks through a global ticket
//     counter, so the chunks in flight are always one compact window of the traversal order however
//     unevenly CTAs progress; the next ticket is drawn at the top of an iteration and resolved at
//     its end, off the critical path.  For the READOUT this brings DRAM traffic down to the
//     algorithmic 43 GB and is the fastest variant (12.4 ms at 1024^3).
// (2) SPATIAL SCHEDULE.  Each chunk gets a key (y-block, x, y, z) from its middle particle; keys are
//     radix sorted (cub, library, ~4M keys) and chunks are walked in key order.  Any permutation is
//     correct, so a stale schedule costs speed, never correctness (it is rebuilt every 8 uses).
// (3) WARP-AGGREGATED ATOMICS.  Lanes whose stencils are shifted by c cells along the contiguous
//     axis pass their overlapping contributions to the owning lane by shuffle: one red.global.add
//     per cell of the warp instead of one per (particle, point): 8 -> ~4.5 per CIC particle.  The
//     paint is bound by red issue (~0.8 clk per active lane per SM, measured) so this is the lever.
// Measured trade-off for the PAINT (1024^3 CIC f8, ms / DRAM GB): tickets give the ideal 43.6 GB
// but 23.0 ms -- neighbouring chunks run concurrently and their reds serialise on the same L2
// lines; a static chunk stride over the spatial schedule lets CTAs drift apart just enough:
// 16.1 ms / 63 GB.  The paint therefore uses (2)+(3) with a static stride, the readout (1).
#pragma once
#include <cub/cub.cuh>

#include "pmb_stencil.cuh"

#define PMB_CHUNK 256

static int pmb_env_flag(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

// ---- schedule ---------------------------------------------------------------------------------
__global__ void pmb_k_chunk_keys(PmbGeom g, PmbParticles p, int64_t npart, int64_t nchunks,
                                 uint64_t *keys, uint32_t *ids, int keymode)
{
    int64_t c = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    int64_t i = c * PMB_CHUNK + PMB_CHUNK / 2;       // a particle in the middle of the chunk
    if (i >= npart) i = npart - 1;
    double x[3];
    pmb_load_pos<3>(p, i, x);
    uint64_t cell[3];
    for (int d = 0; d < 3; d++) {
        double X = pmb_gridpos(x[d], g.scale[d], g.translate[d]);
        int64_t t = (int64_t) floor(X);
        int64_t per = g.period[d] > 0 ? g.period[d] : (g.size[d] > 0 ? g.size[d] : 1);
        t %= per;
        if (t < 0) t += per;
        cell[d] = (uint64_t) t;
    }
    // (y-block of 32 rows, x, y within the block, z / 64): 16 bits per field is plenty
    if (keymode == 0)     // (y-block, x, y, z/64): spatial neighbours close in the order
        keys[c] = ((cell[1] >> 5) << 48) | (cell[0] << 26) | ((cell[1] & 31) << 21) | (cell[2] >> 6);
    else                  // (y-block, x, z/256, y): consecutive chunks are y-neighbours at the same z range
        keys[c] = ((cell[1] >> 5) << 48) | (cell[0] << 26) | ((cell[2] >> 8) << 5) | (cell[1] & 31);
    ids[c] = (uint32_t) c;
}

// returns the device array of chunk ids in traversal order (NULL: natural order)
static int pmb_sched_ticket(pmb_ctx *ctx, unsigned long long **ticket)
{
    // a tiny dedicated allocation (first 256 bytes of sched_buf are reserved for it)
    if (!ctx->sched_buf) {
        PMB_CUDA(cudaMalloc(&ctx->sched_buf, 4096));
        ctx->sched_bytes = 4096;
        ctx->sched_sig = 0;
    }
    *ticket = (unsigned long long *) ctx->sched_buf;
    PMB_CUDA(cudaMemsetAsync(*ticket, 0, sizeof(unsigned long long), ctx->stream));
    return PMB_OK;
}

static int pmb_sched_prepare(pmb_ctx *ctx, const PmbGeom &g, const PmbParticles &p, int64_t npart,
                             const uint32_t **order, int64_t *nchunks_out, int keymode = 0)
{
    const int64_t nchunks = (npart + PMB_CHUNK - 1) / PMB_CHUNK;
    *nchunks_out = nchunks;
    *order = NULL;
    if (!pmb_env_flag("PMB_SCHED", 1) || nchunks < 4096 || nchunks >= ((int64_t) 1 << 31)) return PMB_OK;
    uint64_t sig = (uint64_t) (uintptr_t) p.pos * 0x9E3779B97F4A7C15ull ^ (uint64_t) npart * 0xD6E8FEB86659FD93ull
                   ^ (uint64_t) p.ps0 ^ ((uint64_t) g.size[0] << 40) ^ ((uint64_t) g.size[1] << 20) ^ (uint64_t) g.size[2];
    uint64_t tr;
    memcpy(&tr, &g.translate[0], sizeof(tr));
    sig ^= tr * 0x94D049BB133111EBull;
    sig += (uint64_t) keymode;
    const size_t b_keys = (sizeof(uint64_t) * nchunks + 255) & ~(size_t) 255;
    const size_t b_ids = (sizeof(uint32_t) * nchunks + 255) & ~(size_t) 255;
    size_t temp = 0;
    PMB_CUDA(cub::DeviceRadixSort::SortPairs(NULL, temp, (uint64_t *) NULL, (uint64_t *) NULL, (uint32_t *) NULL,
                                             (uint32_t *) NULL, (int) nchunks, 0, 64, ctx->stream));
    const size_t need = 256 + 2 * b_keys + 2 * b_ids + temp + 256;
    if (need > ctx->sched_bytes) {
        if (ctx->sched_buf) { PMB_CUDA(cudaStreamSynchronize(ctx->stream)); PMB_CUDA(cudaFree(ctx->sched_buf)); ctx->sched_buf = NULL; }
        PMB_CUDA(cudaMalloc(&ctx->sched_buf, need + need / 8));
        ctx->sched_bytes = need + need / 8;
        ctx->sched_sig = 0;
    }
    char *b = (char *) ctx->sched_buf + 256;     // the first 256 bytes hold the ticket counter
    uint32_t *sorted_ids = (uint32_t *) (b + 2 * b_keys + b_ids);
    // a schedule stays useful while the particles move slowly: rebuild every few uses
    if (ctx->sched_sig == sig && ctx->sched_nchunks == nchunks && ctx->sched_uses < 8) {
        ctx->sched_uses++;
        *order = sorted_ids;
        return PMB_OK;
    }
    uint64_t *keys = (uint64_t *) b, *keys2 = (uint64_t *) (b + b_keys);
    uint32_t *ids = (uint32_t *) (b + 2 * b_keys);
    pmb_k_chunk_keys<<<(int) ((nchunks + 255) / 256), 256, 0, ctx->stream>>>(g, p, npart, nchunks, keys, ids, keymode);
    PMB_LAUNCH_CHECK(ctx);
    PMB_CUDA(cub::DeviceRadixSort::SortPairs(b + 2 * b_keys + 2 * b_ids, temp, keys, keys2, ids, sorted_ids,
                                             (int) nchunks, 0, 64, ctx->stream));
    ctx->launches += 8;
    ctx->sched_sig = sig;
    ctx->sched_nchunks = nchunks;
    ctx->sched_uses = 1;
    *order = sorted_ids;
    return PMB_OK;
}

// ---- dynamic chunk tickets ----------------------------------------------------------------------
__device__ __forceinline__ long long pmb_resolve_chunk(unsigned long long tk, const uint32_t *order, int64_t nchunks)
{
    if ((int64_t) tk >= nchunks) return -1;
    return order ? (long long) order[tk] : (long long) tk;
}
__device__ __forceinline__ long long pmb_next_chunk(unsigned long long *ticket, const uint32_t *order, int64_t nchunks)
{
    return pmb_resolve_chunk(atomicAdd(ticket, 1ull), order, nchunks);
}

// ---- paint -------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t pmb_shfl_up_i64(int64_t v, int delta)
{
    int lo = __shfl_up_sync(0xffffffffu, (int) (v & 0xffffffff), delta);
    int hi = __shfl_up_sync(0xffffffffu, (int) (v >> 32), delta);
    return ((int64_t) hi << 32) | (uint32_t) lo;
}
__device__ __forceinline__ double pmb_shfl_up_f64(double v, int delta)
{
    return __shfl_up_sync(0xffffffffu, v, delta);
}

template <typename MeshT>
__device__ __forceinline__ void pmb_red(char *mesh, int64_t off, double f, uint64_t policy)
{
    if (sizeof(MeshT) == 8)
        asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(mesh + off), "d"(f), "l"(policy) : "memory");
    else
        asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(mesh + off), "f"((float) f), "l"(policy) : "memory");
}

// One thread per particle of a chunk; MERGE aggregates along the last mesh axis inside the warp.
template <typename MeshT, int FAM, bool CHECK, bool MERGE>
__global__ void __launch_bounds__(PMB_CHUNK)
pmb_k_paint_sched(PmbGeom g, PmbParticles p, char *mesh, int64_t npart, int pcsfix,
                  const uint32_t *__restrict__ order, int64_t nchunks, int dbg, unsigned long long *ticket)
{
    uint64_t policy;
    if ((dbg & 3) == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    else if ((dbg & 3) == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(policy));
    const int pos_stream = (dbg >> 2) & 1;
    const int lane = threadIdx.x & 31;
    // dynamic tickets: chunks are handed out strictly in schedule order, so the set of chunks in
    // flight is always one compact window of the schedule however unevenly the CTAs progress (with
    // a static stride, a CTA that lags by one iteration lags by a whole window).  The ticket of the
    // NEXT iteration is drawn at the top of the current one and resolved at its end, off the
    // critical path.
    __shared__ long long s_chunk[2];
    const bool dynamic = (dbg >> 3) & 1;
    if (threadIdx.x == 0)
        s_chunk[0] = dynamic ? pmb_next_chunk(ticket, order, nchunks) : pmb_resolve_chunk(blockIdx.x, order, nchunks);
    __syncthreads();
    for (int it = 0;; it++) {
        const int64_t chunk = s_chunk[it & 1];
        if (chunk < 0) break;
        unsigned long long tk = 0;
        if (threadIdx.x == 0) tk = dynamic ? atomicAdd(ticket, 1ull) : (unsigned long long) blockIdx.x + (unsigned long long) (it + 1) * gridDim.x;
        const int64_t i = chunk * PMB_CHUNK + threadIdx.x;
        const bool active = i < npart;
        double x[3] = {0, 0, 0};
        double m = 0;
        if (active) {
            pmb_load_pos<3>(p, i, x, pos_stream);
            m = pmb_load_mass(p, i);
        }
        PmbAxes<3, FAM> A;
        pmb_axes_tuned<3, FAM, CHECK>(g, g.order, x, pcsfix, A);
        if (!MERGE) {
            if (active)
                pmb_for_points_fixed<3, FAM, CHECK>(A, [&](int, int64_t off, double v0, double v1, double v2) {
                    if (!CHECK || off != PMB_OFF_INVALID) pmb_red<MeshT>(mesh, off, pmb_paint_value(true, m, v0, v1, v2), policy);
                });
        } else {
        // cell (0,0,c) of my stencil as a byte offset; invalid -> a value no other lane can equal
        int64_t base[FAM];
#pragma unroll
        for (int c = 0; c < FAM; c++) {
            const bool ok = active && (!CHECK || (A.off[0][0] != PMB_OFF_INVALID && A.off[1][0] != PMB_OFF_INVALID &&
                                                  A.off[2][c] != PMB_OFF_INVALID));
            base[c] = ok ? A.off[0][0] + A.off[1][0] + A.off[2][c] : PMB_OFF_INVALID;
        }
        // accept[c]: the lane c below me owns, as its point c, the cell that is my point 0
        // taken[c]:  my point c is delivered by the lane c above me
        bool accept[FAM], taken[FAM];
        accept[0] = taken[0] = false;
#pragma unroll
        for (int c = 1; c < FAM; c++) {
            const int64_t theirs = pmb_shfl_up_i64(base[c], c);
            accept[c] = lane >= c && base[0] != PMB_OFF_INVALID && theirs == base[0];
            taken[c] = __shfl_down_sync(0xffffffffu, (int) accept[c], c) != 0 && lane + c < 32;
        }
#pragma unroll
        for (int a = 0; a < FAM; a++) {
#pragma unroll
            for (int b = 0; b < FAM; b++) {
                const int64_t o0 = A.off[0][a], o1 = A.off[1][b];
                const bool rowok = active && (!CHECK || (o0 != PMB_OFF_INVALID && o1 != PMB_OFF_INVALID));
                const double w01 = A.V[0][a] * m;       // ((V0 * mass) * V1) * V2, as the reference's tuned path
                double v[FAM];
#pragma unroll
                for (int c = 0; c < FAM; c++) v[c] = (w01 * A.V[1][b]) * A.V[2][c];
                double acc = v[0];
#pragma unroll
                for (int c = 1; c < FAM; c++) {
                    const double r = pmb_shfl_up_f64(v[c], c);
                    if (accept[c]) acc += r;
                }
                if (rowok) {
                    if (!CHECK || A.off[2][0] != PMB_OFF_INVALID) pmb_red<MeshT>(mesh, o0 + o1 + A.off[2][0], acc, policy);
#pragma unroll
                    for (int c = 1; c < FAM; c++)
                        if (!taken[c] && (!CHECK || A.off[2][c] != PMB_OFF_INVALID))
                            pmb_red<MeshT>(mesh, o0 + o1 + A.off[2][c], v[c], policy);
                }
            }
        }
        }
        if (threadIdx.x == 0) s_chunk[(it + 1) & 1] = pmb_resolve_chunk(tk, order, nchunks);
        __syncthreads();
    }
}

// ---- CIC paint with a register carry along y ------------------------------------------------------
// With the (y-block, x, z/256, y) schedule a CTA that walks a run of consecutive chunks sees, in
// thread t, the particles (x', y', k), (x', y'+1, k), (x', y'+2, k), ... of the lattice: the b = 1
// row of one particle's stencil is the b = 0 row of the next.  That row is CARRIED in registers to
// the next iteration and merged there, so only the b = 0 rows are ever written: with the warp
// aggregation along z this is ~2.2 red.global.add per particle instead of 8.  A carry whose target
// row does not match (row ends, displaced neighbours, end of a run) is flushed with plain reds.
// CTAs own whole runs of UNIT chunks (static stride over runs), so concurrently running CTAs work
// on different mesh rows and their reds do not serialise on the same L2 lines.
template <typename MeshT, bool CHECK, bool PREFETCH>
__global__ void __launch_bounds__(PMB_CHUNK, 4)
pmb_k_paint_cic_carry(PmbGeom g, PmbParticles p, char *mesh, int64_t npart,
                      const uint32_t *__restrict__ order, int64_t nchunks, int unit)
{
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    const int lane = threadIdx.x & 31;
    const int64_t nunits = (nchunks + unit - 1) / unit;
    for (int64_t u = blockIdx.x; u < nunits; u += gridDim.x) {
        // carry: values of the (a, b = 1, c) points of the previous particle and where they go
        double cv[2][2] = {{0, 0}, {0, 0}};
        int64_t coff[2][2] = {{PMB_OFF_INVALID, PMB_OFF_INVALID}, {PMB_OFF_INVALID, PMB_OFF_INVALID}};
        int64_t ckey = PMB_OFF_INVALID;
        const int64_t cend = min((u + 1) * (int64_t) unit, nchunks);
        int64_t chunk = order ? (int64_t) order[u * unit] : u * unit;
        // software pipeline: the positions of the next chunk are loaded while this one is deposited
        double xn[3] = {0, 0, 0};
        double mn = 0;
        if (PREFETCH && chunk * PMB_CHUNK + threadIdx.x < npart) {
            pmb_load_pos<3>(p, chunk * PMB_CHUNK + threadIdx.x, xn);
            mn = pmb_load_mass(p, chunk * PMB_CHUNK + threadIdx.x);
        }
        for (int64_t cb = u * unit; cb < cend; cb++) {
            const int64_t i = chunk * PMB_CHUNK + threadIdx.x;
            const bool active = i < npart;
            if (!PREFETCH && active) {
                pmb_load_pos<3>(p, i, xn);
                mn = pmb_load_mass(p, i);
            }
            const double x[3] = {xn[0], xn[1], xn[2]};
            const double m = mn;
            if (cb + 1 < cend) {
                chunk = order ? (int64_t) order[cb + 1] : cb + 1;
                const int64_t in = chunk * PMB_CHUNK + threadIdx.x;
                if (PREFETCH && in < npart) {
                    pmb_load_pos<3>(p, in, xn);
                    mn = pmb_load_mass(p, in);
                }
            }
            PmbAxes<3, 2> A;
            pmb_axes_tuned<3, 2, CHECK>(g, g.order, x, 0, A);
            // this particle's values: ((V0 * m) * V1) * V2
            double v[2][2][2];
            int64_t off[2][2][2];
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++)
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        v[a][b][c] = ((A.V[0][a] * m) * A.V[1][b]) * A.V[2][c];
                        const bool ok = active && (!CHECK || (A.off[0][a] != PMB_OFF_INVALID && A.off[1][b] != PMB_OFF_INVALID &&
                                                              A.off[2][c] != PMB_OFF_INVALID));
                        off[a][b][c] = ok ? A.off[0][a] + A.off[1][b] + A.off[2][c] : PMB_OFF_INVALID;
                    }
            // merge the carried row into my b = 0 row, or flush it
            if (ckey != PMB_OFF_INVALID && ckey == off[0][0][0] && coff[0][1] == off[0][0][1] &&
                coff[1][0] == off[1][0][0] && coff[1][1] == off[1][0][1]) {
#pragma unroll
                for (int a = 0; a < 2; a++)
#pragma unroll
                    for (int c = 0; c < 2; c++) v[a][0][c] += cv[a][c];
            } else {
#pragma unroll
                for (int a = 0; a < 2; a++)
#pragma unroll
                    for (int c = 0; c < 2; c++)
                        if (coff[a][c] != PMB_OFF_INVALID) pmb_red<MeshT>(mesh, coff[a][c], cv[a][c], policy);
            }
            // new carry = my b = 1 row
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int c = 0; c < 2; c++) { cv[a][c] = v[a][1][c]; coff[a][c] = off[a][1][c]; }
            ckey = off[0][1][0];
            // b = 0 row: warp aggregation along z, then one red per owned cell
            const int64_t theirs = pmb_shfl_up_i64(off[0][0][1], 1);
            const bool accept = lane >= 1 && off[0][0][0] != PMB_OFF_INVALID && theirs == off[0][0][0];
            const bool taken = __shfl_down_sync(0xffffffffu, (int) accept, 1) != 0 && lane < 31;
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const double r = pmb_shfl_up_f64(v[a][0][1], 1);
                double acc = v[a][0][0];
                if (accept) acc += r;
                if (off[a][0][0] != PMB_OFF_INVALID) pmb_red<MeshT>(mesh, off[a][0][0], acc, policy);
                if (!taken && off[a][0][1] != PMB_OFF_INVALID) pmb_red<MeshT>(mesh, off[a][0][1], v[a][0][1], policy);
            }
        }
        // end of the run: whatever is still carried goes out
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int c = 0; c < 2; c++)
                if (coff[a][c] != PMB_OFF_INVALID) pmb_red<MeshT>(mesh, coff[a][c], cv[a][c], policy);
    }
}

// ---- the same kernel with 32-bit element indices ---------------------------------------------------
// ncu (profiles/) showed the 64-bit version issue-bound: 428 warp instructions per 32 particles, 29 %
// IMAD + 14 % ISETP from 64-bit byte-offset arithmetic and compares.  When the canvas has fewer than
// 2^31 elements (1024^3 padded has 1.08e9) every cell is a 32-bit element index: one IMAD per axis
// point, 32-bit adds / compares / shuffles, and a single IMAD.WIDE per red for the address.
struct PmbGeom32 {
    double scale[3], translate[3];
    int period[3], size[3], estride[3];
};

// positions of particle i; POS8: contiguous (N, 3) float64 rows (the common case) -- no stride
// arithmetic, no element-size branch
template <bool POS8>
__device__ __forceinline__ void pmb_load_pos3(const PmbParticles &p, int64_t i, double &x0, double &x1, double &x2)
{
    if (POS8) {
        const double *pp = (const double *) p.pos + 3 * i;
        x0 = __ldcs(pp); x1 = __ldcs(pp + 1); x2 = __ldcs(pp + 2);
    } else {
        double t[3];
        pmb_load_pos<3>(p, i, t);
        x0 = t[0]; x1 = t[1]; x2 = t[2];
    }
}
static inline bool pmb_pos_is_f8_rows(const PmbParticles &p)
{
    return p.pos_elsize == 8 && p.ps1 == 8 && p.ps0 == 24 && ((uintptr_t) p.pos & 7) == 0;
}
// the 32-byte records (x, y, z, particle number) of the tile-sorted copy (pmb_bin.cuh)
static inline bool pmb_pos_is_f8_rec4(const PmbParticles &p)
{
    return p.pos_elsize == 8 && p.ps1 == 8 && p.ps0 == 32 && ((uintptr_t) p.pos & 31) == 0;
}

template <bool CHECK>
__device__ __forceinline__ void pmb_cic_axis32(double xin, double scale, double translate, int per, int sz, int es,
                                               double &V0, double &V1, int &e0, int &e1)
{
    const double X = pmb_gridpos(xin, scale, translate);
    const int I0 = (int) floor(X);
    V1 = X - I0;
    V0 = 1. - V1;
    int t0 = I0;
    if (per > 0) t0 = pmb_wrap32(t0, per);
    int t1 = t0 + 1;
    if (per > 0 && t1 == per) t1 = 0;
    e0 = (!CHECK || (unsigned) t0 < (unsigned) sz) ? t0 * es : -1;
    e1 = (!CHECK || (unsigned) t1 < (unsigned) sz) ? t1 * es : -1;
}

template <typename MeshT, bool CHECK, bool POS8, int MINB = 4>
__global__ void __launch_bounds__(PMB_CHUNK, MINB)
pmb_k_paint_cic_carry32(PmbGeom32 g, PmbParticles p, MeshT *mesh, int64_t npart,
                        const uint32_t *__restrict__ order, int64_t nchunks, int unit)
{
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    const int lane = threadIdx.x & 31;
    const int64_t nunits = (nchunks + unit - 1) / unit;
    for (int64_t u = blockIdx.x; u < nunits; u += gridDim.x) {
        double cv00 = 0, cv01 = 0, cv10 = 0, cv11 = 0;     // carried (a, b = 1, c) values
        int co00 = -1, co01 = -1, co10 = -1, co11 = -1;     // and their element indices (-1: none)
        const int64_t cend = min((u + 1) * (int64_t) unit, nchunks);
        int64_t chunk = order ? (int64_t) order[u * unit] : u * unit;
        double xn0 = 0, xn1 = 0, xn2 = 0, mn = 0;
        {
            const int64_t i0 = chunk * PMB_CHUNK + threadIdx.x;
            if (i0 < npart) {
                pmb_load_pos3<POS8>(p, i0, xn0, xn1, xn2);
                mn = pmb_load_mass(p, i0);
            }
        }
        for (int64_t cb = u * unit; cb < cend; cb++) {
            const bool active = chunk * PMB_CHUNK + threadIdx.x < npart;
            const double x0 = xn0, x1 = xn1, x2 = xn2, m = mn;
            if (cb + 1 < cend) {
                chunk = order ? (int64_t) order[cb + 1] : cb + 1;
                const int64_t in = chunk * PMB_CHUNK + threadIdx.x;
                if (in < npart) {
                    pmb_load_pos3<POS8>(p, in, xn0, xn1, xn2);
                    mn = pmb_load_mass(p, in);
                }
            }
            double Vx0, Vx1, Vy0, Vy1, Vz0, Vz1;
            int ex0, ex1, ey0, ey1, ez0, ez1;
            pmb_cic_axis32<CHECK>(x0, g.scale[0], g.translate[0], g.period[0], g.size[0], g.estride[0], Vx0, Vx1, ex0, ex1);
            pmb_cic_axis32<CHECK>(x1, g.scale[1], g.translate[1], g.period[1], g.size[1], g.estride[1], Vy0, Vy1, ey0, ey1);
            pmb_cic_axis32<CHECK>(x2, g.scale[2], g.translate[2], g.period[2], g.size[2], g.estride[2], Vz0, Vz1, ez0, ez1);
            // element index of point (a, b, c); -1 when outside the canvas or the lane is idle
            auto idx = [&](int ea, int eb, int ec) -> int {
                if (CHECK) return (active && ea >= 0 && eb >= 0 && ec >= 0) ? ea + eb + ec : -1;
                return active ? ea + eb + ec : -1;
            };
            const int o000 = idx(ex0, ey0, ez0), o001 = idx(ex0, ey0, ez1), o100 = idx(ex1, ey0, ez0), o101 = idx(ex1, ey0, ez1);
            const int o010 = idx(ex0, ey1, ez0), o011 = idx(ex0, ey1, ez1), o110 = idx(ex1, ey1, ez0), o111 = idx(ex1, ey1, ez1);
            // ((V0 * m) * V1) * V2, the tuned routine's order
            const double wx0 = Vx0 * m, wx1 = Vx1 * m;
            const double w00 = wx0 * Vy0, w01 = wx0 * Vy1, w10 = wx1 * Vy0, w11 = wx1 * Vy1;
            double v000 = w00 * Vz0, v001 = w00 * Vz1, v100 = w10 * Vz0, v101 = w10 * Vz1;
            const double v010 = w01 * Vz0, v011 = w01 * Vz1, v110 = w11 * Vz0, v111 = w11 * Vz1;
            // carried row: merge into my b = 0 row when it is the same mesh row, else flush it
            const bool same = co00 >= 0 && co00 == o000 && co01 == o001 && co10 == o100 && (!CHECK || co11 == o101);
            if (same) {
                v000 += cv00; v001 += cv01; v100 += cv10; v101 += cv11;
            } else {
                if (co00 >= 0) pmb_red<MeshT>((char *) mesh, (int64_t) co00 * sizeof(MeshT), cv00, policy);
                if (co01 >= 0) pmb_red<MeshT>((char *) mesh, (int64_t) co01 * sizeof(MeshT), cv01, policy);
                if (co10 >= 0) pmb_red<MeshT>((char *) mesh, (int64_t) co10 * sizeof(MeshT), cv10, policy);
                if (co11 >= 0) pmb_red<MeshT>((char *) mesh, (int64_t) co11 * sizeof(MeshT), cv11, policy);
            }
            cv00 = v010; cv01 = v011; cv10 = v110; cv11 = v111;
            co00 = o010; co01 = o011; co10 = o110; co11 = o111;
            // b = 0 row: aggregate along z inside the warp, one red per owned cell
            const int theirs = __shfl_up_sync(0xffffffffu, o001, 1);
            const bool accept = lane >= 1 && o000 >= 0 && theirs == o000;
            const bool taken = __shfl_down_sync(0xffffffffu, (int) accept, 1) != 0 && lane < 31;
            const double r0 = __shfl_up_sync(0xffffffffu, v001, 1);
            const double r1 = __shfl_up_sync(0xffffffffu, v101, 1);
            if (accept) { v000 += r0; v100 += r1; }
            if (o000 >= 0) pmb_red<MeshT>((char *) mesh, (int64_t) o000 * sizeof(MeshT), v000, policy);
            if (o100 >= 0) pmb_red<MeshT>((char *) mesh, (int64_t) o100 * sizeof(MeshT), v100, policy);
            if (!taken) {
                if (o001 >= 0) pmb_red<MeshT>((char *) mesh, (int64_t) o001 * sizeof(MeshT), v001, policy);
                if (o101 >= 0) pmb_red<MeshT>((char *) mesh, (int64_t) o101 * sizeof(MeshT), v101, policy);
            }
        }
        if (co00 >= 0) pmb_red<MeshT>((char *) mesh, (int64_t) co00 * sizeof(MeshT), cv00, policy);
        if (co01 >= 0) pmb_red<MeshT>((char *) mesh, (int64_t) co01 * sizeof(MeshT), cv01, policy);
        if (co10 >= 0) pmb_red<MeshT>((char *) mesh, (int64_t) co10 * sizeof(MeshT), cv10, policy);
        if (co11 >= 0) pmb_red<MeshT>((char *) mesh, (int64_t) co11 * sizeof(MeshT), cv11, policy);
    }
}

// ---- y-carry paint for every tuned window (FAM = 2, 3, 4) and gradient windows -----------------------
// The CIC trick above, generalised: a FAM-wide stencil shares FAM - 1 of its y rows with the stencil of
// the next particle along y, so the partial sums of rows b = 1 .. FAM-1 (FAM x FAM values each) are
// carried in registers to the next chunk and only row b = 0 -- complete at that point -- is written,
// after the warp has merged the overlapping z cells of neighbouring lanes.  Global reds per particle
// drop from FAM^2 rows to FAM rows (TSC: 9 -> 3, PCS: 16 -> 4), and the wide windows are bound by
// exactly that number (profiles/README.md: ~0.4 T fp64 reds/s whatever the window).  A carry that
// does not match the next particle (row ends, large displacements, end of a run) is flushed with
// plain reds: any particle order gives the right answer, lattice-like order gives the speed.
template <int FAM, bool CHECK>
__device__ __forceinline__ void pmb_axis32(double xin, int order, double scale, double translate, int pcsfix,
                                           int per, int sz, int es, double *V, int *e)
{
    const double X = pmb_gridpos(xin, scale, translate);
    int I[FAM];
    pmb_axis_tuned<FAM>(X, order, scale, pcsfix, I, V);
    int t = I[0];
    if (per > 0) t = pmb_wrap32(t, per);
#pragma unroll
    for (int s = 0; s < FAM; s++) {
        e[s] = (!CHECK || (unsigned) t < (unsigned) sz) ? t * es : -1;
        t += 1;
        if (per > 0 && t == per) t = 0;
    }
}

struct PmbGeom32o {
    PmbGeom32 g;
    int order[3];
    int pcsfix;
};

template <typename MeshT, int FAM, bool CHECK, bool POS8>
__global__ void __launch_bounds__(PMB_CHUNK, (FAM == 2 ? 4 : (FAM == 3 ? 2 : 1)))
pmb_k_paint_carry32(PmbGeom32o go, PmbParticles p, MeshT *mesh, int64_t npart,
                    const uint32_t *__restrict__ order, int64_t nchunks, int unit)
{
    const PmbGeom32 &g = go.g;
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    const int lane = threadIdx.x & 31;
    const int64_t nunits = (nchunks + unit - 1) / unit;
    for (int64_t u = blockIdx.x; u < nunits; u += gridDim.x) {
        // carry: partial sums of the y rows b = 1 .. FAM-1 of the previous particle of this thread,
        // stored at row index b - 1, and where they go (per-axis element indices; cey < 0: no carry)
        double cv[FAM - 1][FAM][FAM];
        int cex[FAM], cez[FAM], cey[FAM - 1];
#pragma unroll
        for (int b = 0; b < FAM - 1; b++) {
            cey[b] = -2;
#pragma unroll
            for (int a = 0; a < FAM; a++)
#pragma unroll
                for (int c = 0; c < FAM; c++) cv[b][a][c] = 0;
        }
#pragma unroll
        for (int a = 0; a < FAM; a++) { cex[a] = -2; cez[a] = -2; }
        bool have = false;
        const int64_t cend = min((u + 1) * (int64_t) unit, nchunks);
        int64_t chunk = order ? (int64_t) order[u * unit] : u * unit;
        double xn0 = 0, xn1 = 0, xn2 = 0, mn = 0;
        {
            const int64_t i0 = chunk * PMB_CHUNK + threadIdx.x;
            if (i0 < npart) {
                pmb_load_pos3<POS8>(p, i0, xn0, xn1, xn2);
                mn = pmb_load_mass(p, i0);
            }
        }
        for (int64_t cb = u * unit; cb < cend; cb++) {
            const bool active = chunk * PMB_CHUNK + threadIdx.x < npart;
            const double x0 = xn0, x1 = xn1, x2 = xn2, m = mn;
            if (cb + 1 < cend) {
                chunk = order ? (int64_t) order[cb + 1] : cb + 1;
                const int64_t in = chunk * PMB_CHUNK + threadIdx.x;
                if (in < npart) {
                    pmb_load_pos3<POS8>(p, in, xn0, xn1, xn2);
                    mn = pmb_load_mass(p, in);
                }
            }
            double Vx[FAM], Vy[FAM], Vz[FAM];
            int ex[FAM], ey[FAM], ez[FAM];
            pmb_axis32<FAM, CHECK>(x0, go.order[0], g.scale[0], g.translate[0], go.pcsfix, g.period[0], g.size[0], g.estride[0], Vx, ex);
            pmb_axis32<FAM, CHECK>(x1, go.order[1], g.scale[1], g.translate[1], go.pcsfix, g.period[1], g.size[1], g.estride[1], Vy, ey);
            pmb_axis32<FAM, CHECK>(x2, go.order[2], g.scale[2], g.translate[2], go.pcsfix, g.period[2], g.size[2], g.estride[2], Vz, ez);
            // does the carry continue into this particle?  same x and z cells, y shifted by one
            bool same = have && active;
#pragma unroll
            for (int a = 0; a < FAM; a++) same = same && cex[a] == ex[a] && cez[a] == ez[a];
#pragma unroll
            for (int b = 0; b < FAM - 1; b++) same = same && cey[b] == ey[b];
            if (have && !same) {
#pragma unroll
                for (int b = 0; b < FAM - 1; b++)
#pragma unroll
                    for (int a = 0; a < FAM; a++)
#pragma unroll
                        for (int c = 0; c < FAM; c++)
                            if (!CHECK || (cex[a] >= 0 && cey[b] >= 0 && cez[c] >= 0))
                                pmb_red<MeshT>((char *) mesh, (int64_t) (cex[a] + cey[b] + cez[c]) * sizeof(MeshT), cv[b][a][c], policy);
            }
            // row b = 0 of this particle (+ the carried partial sums): complete, goes out below
            double r0[FAM][FAM];
            double wx[FAM];
#pragma unroll
            for (int a = 0; a < FAM; a++) wx[a] = Vx[a] * m;          // ((V0 * m) * V1) * V2, the tuned routines' order
#pragma unroll
            for (int a = 0; a < FAM; a++)
#pragma unroll
                for (int c = 0; c < FAM; c++) {
                    const double v = (wx[a] * Vy[0]) * Vz[c];
                    r0[a][c] = same ? v + cv[0][a][c] : v;
                }
            // rows b = 1 .. FAM-1 become the new carry (ascending b: slot b - 1 was consumed already)
#pragma unroll
            for (int b = 1; b < FAM; b++)
#pragma unroll
                for (int a = 0; a < FAM; a++)
#pragma unroll
                    for (int c = 0; c < FAM; c++) {
                        const double v = (wx[a] * Vy[b]) * Vz[c];
                        cv[b - 1][a][c] = (same && b < FAM - 1) ? v + cv[b][a][c] : v;
                    }
            have = active;
#pragma unroll
            for (int a = 0; a < FAM; a++) { cex[a] = ex[a]; cez[a] = ez[a]; }
#pragma unroll
            for (int b = 0; b < FAM - 1; b++) cey[b] = ey[b + 1];
            // row 0: merge along z inside the warp.  accept[c]: the lane c below me holds, as its z point
            // c, the cell that is my z point 0 (same x / y cells); taken[c]: my z point c is delivered
            // by the lane c above me.
            const bool row_ok = active && (!CHECK || ey[0] >= 0);
            int base[FAM];
#pragma unroll
            for (int c = 0; c < FAM; c++)
                base[c] = (row_ok && (!CHECK || (ex[0] >= 0 && ez[c] >= 0))) ? ex[0] + ey[0] + ez[c] : -1;
            bool accept[FAM], taken[FAM];
            accept[0] = taken[0] = false;
#pragma unroll
            for (int c = 1; c < FAM; c++) {
                const int theirs = __shfl_up_sync(0xffffffffu, base[c], c);
                accept[c] = lane >= c && base[0] >= 0 && theirs == base[0];
                taken[c] = __shfl_down_sync(0xffffffffu, (int) accept[c], c) != 0 && lane + c < 32;
            }
#pragma unroll
            for (int a = 0; a < FAM; a++) {
                double acc = r0[a][0];
#pragma unroll
                for (int c = 1; c < FAM; c++) {
                    const double r = __shfl_up_sync(0xffffffffu, r0[a][c], c);
                    if (accept[c]) acc += r;
                }
                if (row_ok && (!CHECK || ex[a] >= 0)) {
                    if (!CHECK || ez[0] >= 0)
                        pmb_red<MeshT>((char *) mesh, (int64_t) (ex[a] + ey[0] + ez[0]) * sizeof(MeshT), acc, policy);
#pragma unroll
                    for (int c = 1; c < FAM; c++)
                        if (!taken[c] && (!CHECK || ez[c] >= 0))
                            pmb_red<MeshT>((char *) mesh, (int64_t) (ex[a] + ey[0] + ez[c]) * sizeof(MeshT), r0[a][c], policy);
                }
            }
        }
        if (have) {
#pragma unroll
            for (int b = 0; b < FAM - 1; b++)
#pragma unroll
                for (int a = 0; a < FAM; a++)
#pragma unroll
                    for (int c = 0; c < FAM; c++)
                        if (!CHECK || (cex[a] >= 0 && cey[b] >= 0 && cez[c] >= 0))
                            pmb_red<MeshT>((char *) mesh, (int64_t) (cex[a] + cey[b] + cez[c]) * sizeof(MeshT), cv[b][a][c], policy);
        }
    }
}

// ---- readout -----------------------------------------------------------------------------------
template <typename MeshT, bool VOL>
__device__ __forceinline__ double pmb_mesh_load(const char *mesh, int64_t off, uint64_t policy)
{
    if (sizeof(MeshT) == 8) {
        double v;
        if (VOL) asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(mesh + off), "l"(policy));
        else asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(mesh + off), "l"(policy));
        return v;
    } else {
        float v;
        if (VOL) asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(mesh + off), "l"(policy));
        else asm("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(mesh + off), "l"(policy));
        return (double) v;
    }
}

// Chunks come from the dynamic ticket counter; the positions of the NEXT chunk are loaded while the
// current one is gathered (software pipeline: the DRAM latency of the particle stream is hidden
// behind the mesh gathers of the previous chunk).
template <typename MeshT, int FAM, bool CHECK, int VARIANT>
__global__ void __launch_bounds__(PMB_CHUNK, (VARIANT == 2 ? 5 : 1))
pmb_k_readout_sched(PmbGeom g, PmbParticles p, const char *mesh, int64_t npart, int pcsfix,
                    void *out, int out_elsize, int64_t out_stride,
                    const uint32_t *__restrict__ order, int64_t nchunks, unsigned long long *ticket)
{
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    __shared__ long long s_chunk[3];
    if (threadIdx.x == 0) {
        s_chunk[0] = pmb_next_chunk(ticket, order, nchunks);
        s_chunk[1] = pmb_next_chunk(ticket, order, nchunks);
    }
    __syncthreads();
    int64_t cur = s_chunk[0], nxt = s_chunk[1];
    double x[3] = {0, 0, 0};
    if (VARIANT >= 2 && cur >= 0 && cur * PMB_CHUNK + threadIdx.x < npart) pmb_load_pos<3>(p, cur * PMB_CHUNK + threadIdx.x, x);
    for (int it = 0; cur >= 0; it++) {
        unsigned long long tk = 0;
        if (threadIdx.x == 0) tk = atomicAdd(ticket, 1ull);
        // prefetch the next chunk's positions
        double xn[3] = {0, 0, 0};
        const int64_t in = nxt * PMB_CHUNK + threadIdx.x;
        if (VARIANT >= 2 && nxt >= 0 && in < npart) pmb_load_pos<3>(p, in, xn);
        const int64_t i = cur * PMB_CHUNK + threadIdx.x;
        if (VARIANT < 2 && i < npart) pmb_load_pos<3>(p, i, x);
        if (i < npart) {
            PmbAxes<3, FAM> A;
            pmb_axes_tuned<3, FAM, CHECK>(g, g.order, x, pcsfix, A);
            double value = 0;
            pmb_for_points_fixed<3, FAM, CHECK>(A, [&](int, int64_t off, double v0, double v1, double v2) {
                if (!CHECK || off != PMB_OFF_INVALID) value += pmb_mesh_load<MeshT, VARIANT == 0>(mesh, off, policy) * ((v0 * v1) * v2);
            });
            pmb_st_real_stream(out, i * out_stride, out_elsize, value);
        }
        if (threadIdx.x == 0) s_chunk[(it + 2) % 3] = pmb_resolve_chunk(tk, order, nchunks);
        __syncthreads();
        cur = nxt;
        nxt = s_chunk[(it + 2) % 3];
        x[0] = xn[0]; x[1] = xn[1]; x[2] = xn[2];
    }
}

// Rejected gather variants (measured on B200, kept in the git history, profiles/README.md):
//  * three chunks in flight per CTA (positions of k+2, gathers of k+1, sums of k): 13.3 ms vs 11.4 at
//    1024^3 -- the extra registers cost two resident CTAs and warps in flight matter more;
//  * y-carry of the mesh rows in registers + z sharing by shuffle (the mirror image of the scatter's
//    carry): CIC 1.96 vs 1.30 ms, TSC 4.38 vs 3.07, PCS 9.2 vs 5.5 at 512^3 -- the repeated loads of
//    the plain gather hit L1 and are cheaper than the register traffic that avoids them.
// ---- CIC gather with 32-bit element indices ---------------------------------------------------------
// Same pipeline as pmb_k_readout_sched (dynamic tickets, positions of the next chunk prefetched) with
// the lean index arithmetic of pmb_k_paint_cic_carry32.  Sums in the reference's point order:
// bit-identical results.
template <typename MeshT, bool CHECK, bool POS8, int MINB = 5>
__global__ void __launch_bounds__(PMB_CHUNK, MINB)
pmb_k_readout_cic32(PmbGeom32 g, PmbParticles p, const MeshT *__restrict__ mesh, int64_t npart,
                    void *out, int out_elsize, int64_t out_stride,
                    const uint32_t *__restrict__ order, int64_t nchunks, unsigned long long *ticket)
{
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    __shared__ long long s_chunk[3];
    if (threadIdx.x == 0) {
        s_chunk[0] = pmb_next_chunk(ticket, order, nchunks);
        s_chunk[1] = pmb_next_chunk(ticket, order, nchunks);
    }
    __syncthreads();
    int64_t cur = s_chunk[0], nxt = s_chunk[1];
    double x0 = 0, x1 = 0, x2 = 0;
    if (cur >= 0 && cur * PMB_CHUNK + threadIdx.x < npart) pmb_load_pos3<POS8>(p, cur * PMB_CHUNK + threadIdx.x, x0, x1, x2);
    for (int it = 0; cur >= 0; it++) {
        unsigned long long tk = 0;
        if (threadIdx.x == 0) tk = atomicAdd(ticket, 1ull);
        double xn0 = 0, xn1 = 0, xn2 = 0;
        const int64_t in = nxt * PMB_CHUNK + threadIdx.x;
        if (nxt >= 0 && in < npart) pmb_load_pos3<POS8>(p, in, xn0, xn1, xn2);
        const int64_t i = cur * PMB_CHUNK + threadIdx.x;
        if (i < npart) {
            double Vx[2], Vy[2], Vz[2];
            int ex[2], ey[2], ez[2];
            pmb_cic_axis32<CHECK>(x0, g.scale[0], g.translate[0], g.period[0], g.size[0], g.estride[0], Vx[0], Vx[1], ex[0], ex[1]);
            pmb_cic_axis32<CHECK>(x1, g.scale[1], g.translate[1], g.period[1], g.size[1], g.estride[1], Vy[0], Vy[1], ey[0], ey[1]);
            pmb_cic_axis32<CHECK>(x2, g.scale[2], g.translate[2], g.period[2], g.size[2], g.estride[2], Vz[0], Vz[1], ez[0], ez[1]);
            double mv[2][2][2];
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++)
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const bool ok = !CHECK || (ex[a] >= 0 && ey[b] >= 0 && ez[c] >= 0);
                        mv[a][b][c] = ok ? pmb_mesh_load<MeshT, false>((const char *) mesh,
                                                                       (int64_t) (ex[a] + ey[b] + ez[c]) * sizeof(MeshT), policy)
                                         : 0.0;
                    }
            double value = 0;
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++)
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const bool ok = !CHECK || (ex[a] >= 0 && ey[b] >= 0 && ez[c] >= 0);
                        if (ok) value += mv[a][b][c] * ((Vx[a] * Vy[b]) * Vz[c]);
                    }
            pmb_st_real_stream(out, i * out_stride, out_elsize, value);
        }
        if (threadIdx.x == 0) s_chunk[(it + 2) % 3] = pmb_resolve_chunk(tk, order, nchunks);
        __syncthreads();
        cur = nxt;
        nxt = s_chunk[(it + 2) % 3];
        x0 = xn0; x1 = xn1; x2 = xn2;
    }
}

