// pmb_tile.cuh -- CIC gather and scatter on the TILE-SORTED copy of a particle array (pmb_bin.cuh) with the mesh
// tile in shared memory.
//
// Inside a tile the sorted particles are in arrival order: no two lanes of a warp share a mesh sector, so the
// ordinary kernels pay one L2 request per (particle, stencil point, field) -- measured at 1024^3 uniform random:
// three-field ring gather 69 ms (17 G L2 requests), plain scatter 43.5 ms with the L2 at 82 % of its throughput.
// Here a CTA takes one tile at a time (tickets):
//
//   gather   the tile's mesh cells + a one-cell upper halo of every field are loaded into shared memory (coalesced
//            rows along the contiguous axis), every particle of the tile then reads its 8 * NF values from shared
//            memory and its result goes to the staging rows in sorted order (coalesced).  Weights, order of additions
//            and the treatment of cells outside the canvas are those of pmb_k_readout_cic32_ring: bit-identical.
//   scatter  see pmb_k_paint_cic_tile below.
//
// A particle whose stencil does not start inside the tile it was filed under (positions outside the canvas are
// parked in the last tile of their row) takes the ordinary global-memory path inside the same kernel.
#pragma once
#include "pmb_bin.cuh"

struct PmbTileGeom {
    int s[3];              // log2 of the tile extent per axis
    int n1, n2;            // tiles along axes 1, 2
    int ntiles;
};

// mesh cell of local halo index `off` along an axis whose tile starts at cell o: wrapped into the period, -1 outside the canvas
__device__ __forceinline__ int pmb_tile_cell(int o, int off, int per, int sz)
{
    int c = o + off;
    if (per > 0 && c >= per) c -= per;
    return (unsigned) c < (unsigned) sz ? c : -1;
}

// pmb_cic_axis32 (pmb_sched.cuh) that also hands out the wrapped cell of the first stencil point (-1: outside the canvas)
template <bool CHECK>
__device__ __forceinline__ void pmb_cic_axis32_cell(double xin, double scale, double translate, int per, int sz, int es,
                                                    double &V0, double &V1, int &e0, int &e1, int &cell)
{
    const double X = pmb_gridpos(xin, scale, translate);
    const int I0 = (int) floor(X);
    V1 = X - I0;
    V0 = 1. - V1;
    int t0 = I0;
    if (per > 0) t0 = pmb_wrap32(t0, per);
    int t1 = t0 + 1;
    if (per > 0 && t1 == per) t1 = 0;
    cell = (unsigned) t0 < (unsigned) sz ? t0 : -1;
    e0 = (!CHECK || (unsigned) t0 < (unsigned) sz) ? t0 * es : -1;
    e1 = (!CHECK || (unsigned) t1 < (unsigned) sz) ? t1 * es : -1;
}

template <typename MeshT, bool CHECK, int NF>
__global__ void __launch_bounds__(256)
pmb_k_readout_cic_tile(PmbGeom32 g, PmbTileGeom t, const uint32_t *__restrict__ counts, const uint32_t *__restrict__ ends,
                       const double *__restrict__ recs, PmbFields f, unsigned long long *ticket)
{
    extern __shared__ double tile_smem[];
    __shared__ unsigned long long s_tk;
    const int E0 = (1 << t.s[0]) + 1, E1 = (1 << t.s[1]) + 1, E2 = (1 << t.s[2]) + 1;
    const int ncell = E0 * E1 * E2;
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_tk = atomicAdd(ticket, 1ull);
        __syncthreads();
        const int tile = (int) s_tk;
        if (s_tk >= (unsigned long long) t.ntiles) break;
        const uint32_t cnt = counts[tile];
        if (cnt == 0) continue;
        const uint32_t first = ends[tile] - cnt;
        const int i2 = tile % t.n2, i01 = tile / t.n2, i1 = i01 % t.n1, i0 = i01 / t.n1;
        const int o0 = i0 << t.s[0], o1 = i1 << t.s[1], o2 = i2 << t.s[2];
        // ---- the tile's cells + upper halo of every field ----
        for (int idx = threadIdx.x; idx < ncell; idx += blockDim.x) {
            const int c = idx % E2, ab = idx / E2, b = ab % E1, a = ab / E1;
            const int m0 = pmb_tile_cell(o0, a, g.period[0], g.size[0]);
            const int m1 = pmb_tile_cell(o1, b, g.period[1], g.size[1]);
            const int m2 = pmb_tile_cell(o2, c, g.period[2], g.size[2]);
            const bool ok = m0 >= 0 && m1 >= 0 && m2 >= 0;
            const int64_t off = ok ? (int64_t) (m0 * g.estride[0] + m1 * g.estride[1] + m2 * g.estride[2]) * (int64_t) sizeof(MeshT) : 0;
#pragma unroll
            for (int q = 0; q < NF; q++)
                tile_smem[q * ncell + idx] = ok ? pmb_mesh_load<MeshT, false>((const char *) f.mesh[q], off, policy) : 0.0;
        }
        __syncthreads();
        // ---- the tile's particles ----
        for (uint32_t k = threadIdx.x; k < cnt; k += blockDim.x) {
            const int64_t j = (int64_t) first + k;
            double x0, x1, x2, idb;
            asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(x0), "=d"(x1), "=d"(x2), "=d"(idb) : "l"(recs + 4 * j));
            (void) idb;
            double Vx[2], Vy[2], Vz[2];
            int ex[2], ey[2], ez[2];
            int c0, c1, c2;
            pmb_cic_axis32_cell<CHECK>(x0, g.scale[0], g.translate[0], g.period[0], g.size[0], g.estride[0], Vx[0], Vx[1], ex[0], ex[1], c0);
            pmb_cic_axis32_cell<CHECK>(x1, g.scale[1], g.translate[1], g.period[1], g.size[1], g.estride[1], Vy[0], Vy[1], ey[0], ey[1], c1);
            pmb_cic_axis32_cell<CHECK>(x2, g.scale[2], g.translate[2], g.period[2], g.size[2], g.estride[2], Vz[0], Vz[1], ez[0], ez[1], c2);
            // local index of the first stencil cell (the upper halo holds the second one)
            const int l0 = c0 >= 0 ? c0 - o0 : -1, l1 = c1 >= 0 ? c1 - o1 : -1, l2 = c2 >= 0 ? c2 - o2 : -1;
            const bool local = (unsigned) l0 < (unsigned) (E0 - 1) && (unsigned) l1 < (unsigned) (E1 - 1) && (unsigned) l2 < (unsigned) (E2 - 1);
            double w[2][2][2];
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++)
#pragma unroll
                    for (int c = 0; c < 2; c++) w[a][b][c] = (Vx[a] * Vy[b]) * Vz[c];
            double res[NF];
#pragma unroll
            for (int q = 0; q < NF; q++) {
                double value = 0;
                if (local) {
                    const double *m = tile_smem + q * ncell + (l0 * E1 + l1) * E2 + l2;
#pragma unroll
                    for (int a = 0; a < 2; a++)
#pragma unroll
                        for (int b = 0; b < 2; b++)
#pragma unroll
                            for (int c = 0; c < 2; c++) value += m[(a * E1 + b) * E2 + c] * w[a][b][c];
                } else {
#pragma unroll
                    for (int a = 0; a < 2; a++)
#pragma unroll
                        for (int b = 0; b < 2; b++)
#pragma unroll
                            for (int c = 0; c < 2; c++) {
                                const bool ok = !CHECK || (ex[a] >= 0 && ey[b] >= 0 && ez[c] >= 0);
                                if (ok) value += pmb_mesh_load<MeshT, false>((const char *) f.mesh[q], (int64_t) (ex[a] + ey[b] + ez[c]) * sizeof(MeshT), policy) * w[a][b][c];
                            }
                }
                res[q] = value;
            }
            if (NF == 3 && f.packed_rows) {
                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"((double *) f.out[0] + 4 * j), "d"(res[0]), "d"(res[1]),
                             "d"(res[NF - 1 > 1 ? 2 : 0]), "d"(0.0) : "memory");
            } else {
#pragma unroll
                for (int q = 0; q < NF; q++) pmb_store_result(f, q, j, res[q]);
            }
        }
    }
}

// ---- scatter --------------------------------------------------------------------------------------------------
// Shared-memory float64 atomics are CAS loops on sm_100a, so the deposit into the shared tile is made conflict-free
// by construction: the tile's particles are counting-sorted by cell (32-bit shared atomics), then the cells are
// processed in 8 COLOURS (parity of the local cell index per axis): two cells of one colour are at least two apart
// along every axis, their 2 x 2 x 2 stencils are disjoint, so one thread per cell adds its particles' contributions
// with plain shared-memory read-add-writes.  After the 8 colours the tile (+ upper halo) goes to the mesh with one
// red per touched cell, rows coalesced along the contiguous axis: ~1.3 reds per cell of the tile instead of 8 per
// particle.  Tiles with more particles than a batch holds are deposited batch after batch into the same shared tile.
// Measured at 1024^3 uniform random: 41 - 43 ms, the speed of the plain ticketed kernel, with 7 x fewer L2 requests:
// the ~16 barrier-separated phases of a tile, eight of them waiting on a record re-read from L2, leave it latency-bound.
// Tried: the particles' weights in shared memory as well (no re-read, but one resident CTA per SM): 85 ms; particles in
// registers with deposit passes by (colour, arrival rank in the cell) -- no sort, no re-read, ~48 passes per tile in
// which 1 / 48 of the particles is active: 94 ms.
#define PMB_TILE_BATCH 2560

template <typename MeshT, bool CHECK>
__global__ void __launch_bounds__(256)
pmb_k_paint_cic_tile(PmbGeom32 g, PmbTileGeom t, const uint32_t *__restrict__ counts, const uint32_t *__restrict__ ends,
                     const double *__restrict__ recs, const double *__restrict__ smass, double mass_scalar, MeshT *mesh,
                     unsigned long long *ticket)
{
    extern __shared__ double tile_smem[];
    __shared__ unsigned long long s_tk;
    __shared__ uint32_t s_warp[32];
    const int T0 = 1 << t.s[0], T1 = 1 << t.s[1], T2 = 1 << t.s[2];
    const int E1 = T1 + 1, E2 = T2 + 1;
    const int nhalo = (T0 + 1) * E1 * E2, ncells = T0 * T1 * T2;
    double *acc = tile_smem;                                         // [nhalo]
    uint32_t *ccount = (uint32_t *) (tile_smem + nhalo);             // [ncells] particles per cell, then fill cursors
    uint32_t *cstart = ccount + ncells;                              // [ncells + 1]
    uint16_t *order = (uint16_t *) (cstart + ncells + 1);            // [PMB_TILE_BATCH] batch slots in cell order
    uint16_t *pcell = order + PMB_TILE_BATCH;                        // [PMB_TILE_BATCH] cell of batch slot k (0xFFFF: not local)
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_tk = atomicAdd(ticket, 1ull);
        __syncthreads();
        const int tile = (int) s_tk;
        if (s_tk >= (unsigned long long) t.ntiles) break;
        const uint32_t cnt = counts[tile];
        if (cnt == 0) continue;
        const uint32_t first = ends[tile] - cnt;
        const int i2 = tile % t.n2, i01 = tile / t.n2, i1 = i01 % t.n1, i0 = i01 / t.n1;
        const int o0 = i0 << t.s[0], o1 = i1 << t.s[1], o2 = i2 << t.s[2];
        for (int idx = threadIdx.x; idx < nhalo; idx += blockDim.x) acc[idx] = 0.0;
        for (uint32_t b0 = 0; b0 < cnt; b0 += PMB_TILE_BATCH) {
            const uint32_t bn = min((uint32_t) PMB_TILE_BATCH, cnt - b0);
            for (int c = threadIdx.x; c < ncells; c += blockDim.x) ccount[c] = 0;
            __syncthreads();
            // ---- A: cell of every particle of the batch; particles that do not start in this tile go out directly ----
            for (uint32_t k = threadIdx.x; k < bn; k += blockDim.x) {
                const int64_t j = (int64_t) first + b0 + k;
                double x0, x1, x2, idb;
                asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(x0), "=d"(x1), "=d"(x2), "=d"(idb) : "l"(recs + 4 * j));
                (void) idb;
                double Vx[2], Vy[2], Vz[2];
                int ex[2], ey[2], ez[2], c0, c1, c2;
                pmb_cic_axis32_cell<CHECK>(x0, g.scale[0], g.translate[0], g.period[0], g.size[0], g.estride[0], Vx[0], Vx[1], ex[0], ex[1], c0);
                pmb_cic_axis32_cell<CHECK>(x1, g.scale[1], g.translate[1], g.period[1], g.size[1], g.estride[1], Vy[0], Vy[1], ey[0], ey[1], c1);
                pmb_cic_axis32_cell<CHECK>(x2, g.scale[2], g.translate[2], g.period[2], g.size[2], g.estride[2], Vz[0], Vz[1], ez[0], ez[1], c2);
                const int l0 = c0 >= 0 ? c0 - o0 : -1, l1 = c1 >= 0 ? c1 - o1 : -1, l2 = c2 >= 0 ? c2 - o2 : -1;
                if ((unsigned) l0 < (unsigned) T0 && (unsigned) l1 < (unsigned) T1 && (unsigned) l2 < (unsigned) T2) {
                    const int ci = (l0 * T1 + l1) * T2 + l2;
                    pcell[k] = (uint16_t) ci;
                    atomicAdd(&ccount[ci], 1u);
                } else {
                    pcell[k] = 0xFFFF;
                    const double m = smass ? smass[j] : mass_scalar;
#pragma unroll
                    for (int a = 0; a < 2; a++)
#pragma unroll
                        for (int b = 0; b < 2; b++)
#pragma unroll
                            for (int c = 0; c < 2; c++) {
                                if (CHECK && (ex[a] < 0 || ey[b] < 0 || ez[c] < 0)) continue;
                                pmb_red<MeshT>((char *) mesh, (int64_t) (ex[a] + ey[b] + ez[c]) * sizeof(MeshT), ((Vx[a] * m) * Vy[b]) * Vz[c], policy);
                            }
                }
            }
            __syncthreads();
            // ---- B: exclusive scan of the cell counts (each thread a run of cells, warp scan, scan of the warp totals) ----
            {
                const int per = (ncells + blockDim.x - 1) / blockDim.x;
                const int c_lo = threadIdx.x * per;
                uint32_t sum = 0;
                for (int c = c_lo; c < c_lo + per && c < ncells; c++) sum += ccount[c];
                uint32_t inc = sum;
                for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
                if (lane == 31) s_warp[warp] = inc;
                __syncthreads();
                if (warp == 0) {
                    uint32_t w = lane < (int) (blockDim.x >> 5) ? s_warp[lane] : 0u;
                    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += v; }
                    if (lane < (int) (blockDim.x >> 5)) s_warp[lane] = w;
                }
                __syncthreads();
                uint32_t run = inc - sum + (warp > 0 ? s_warp[warp - 1] : 0u);
                for (int c = c_lo; c < c_lo + per && c < ncells; c++) { const uint32_t n = ccount[c]; cstart[c] = run; ccount[c] = run; run += n; }
                if (threadIdx.x == blockDim.x - 1) cstart[ncells] = run;
            }
            __syncthreads();
            // ---- C: batch slots in cell order ----
            for (uint32_t k = threadIdx.x; k < bn; k += blockDim.x) {
                const uint32_t ci = pcell[k];
                if (ci != 0xFFFF) order[atomicAdd(&ccount[ci], 1u)] = (uint16_t) k;
            }
            __syncthreads();
            // ---- D: the 8 colours ----
            const int h1 = T1 >> 1, h2 = T2 >> 1, ncol = ncells >> 3;
            for (int col = 0; col < 8; col++) {
                for (int cc = threadIdx.x; cc < ncol; cc += blockDim.x) {
                    const int q2 = cc % h2, q01 = cc / h2, q1 = q01 % h1, q0 = q01 / h1;
                    const int l0 = 2 * q0 + (col >> 2), l1 = 2 * q1 + ((col >> 1) & 1), l2 = 2 * q2 + (col & 1);
                    const int ci = (l0 * T1 + l1) * T2 + l2;
                    double *m000 = acc + (l0 * E1 + l1) * E2 + l2;
                    // up to 4 records of the cell are requested before the first one is used (a cell holds ~1 particle
                    // on average, 3 - 4 at most in a warp: one round trip to L2 per colour instead of one per particle)
                    for (uint32_t s = cstart[ci], e = cstart[ci + 1]; s < e; s += 4) {
                        double rx[4], ry[4], rz[4], rm[4];
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (s + u < e) {
                                const int64_t j = (int64_t) first + b0 + order[s + u];
                                double idb;
                                asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(rx[u]), "=d"(ry[u]), "=d"(rz[u]), "=d"(idb) : "l"(recs + 4 * j));
                                (void) idb;
                                rm[u] = smass ? smass[j] : mass_scalar;
                            }
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (s + u < e) {
                                const double m = rm[u];
                                double Vx[2], Vy[2], Vz[2];
                                int ex[2], ey[2], ez[2];
                                pmb_cic_axis32<CHECK>(rx[u], g.scale[0], g.translate[0], g.period[0], g.size[0], g.estride[0], Vx[0], Vx[1], ex[0], ex[1]);
                                pmb_cic_axis32<CHECK>(ry[u], g.scale[1], g.translate[1], g.period[1], g.size[1], g.estride[1], Vy[0], Vy[1], ey[0], ey[1]);
                                pmb_cic_axis32<CHECK>(rz[u], g.scale[2], g.translate[2], g.period[2], g.size[2], g.estride[2], Vz[0], Vz[1], ez[0], ez[1]);
#pragma unroll
                                for (int a = 0; a < 2; a++)
#pragma unroll
                                    for (int b = 0; b < 2; b++)
#pragma unroll
                                        for (int c = 0; c < 2; c++) {
                                            if (CHECK && (ex[a] < 0 || ey[b] < 0 || ez[c] < 0)) continue;     // that cell is outside the canvas
                                            m000[(a * E1 + b) * E2 + c] += ((Vx[a] * m) * Vy[b]) * Vz[c];
                                        }
                            }
                    }
                }
                __syncthreads();
            }
        }
        // ---- E: the tile + upper halo goes to the mesh: one red per touched cell ----
        for (int idx = threadIdx.x; idx < nhalo; idx += blockDim.x) {
            const double v = acc[idx];
            if (v == 0.0) continue;
            const int c = idx % E2, ab = idx / E2, b = ab % E1, a = ab / E1;
            const int m0 = pmb_tile_cell(o0, a, g.period[0], g.size[0]);
            const int m1 = pmb_tile_cell(o1, b, g.period[1], g.size[1]);
            const int m2 = pmb_tile_cell(o2, c, g.period[2], g.size[2]);
            if (m0 >= 0 && m1 >= 0 && m2 >= 0)
                pmb_red<MeshT>((char *) mesh, (int64_t) (m0 * g.estride[0] + m1 * g.estride[1] + m2 * g.estride[2]) * (int64_t) sizeof(MeshT), v, policy);
        }
    }
}

