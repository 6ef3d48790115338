// pmb_ring.cuh -- CIC paint / readout kernels whose particle stream arrives through a shared-memory
// ring filled by the copy engine (cp.async.bulk + mbarrier, pmb_tma.cuh).
//
// Round-1 profiles (profiles/r1_cic32_paint_readout_1024_hotspots.txt): in both CIC kernels a third of
// all stall samples sat on the first use of the particle position although the next chunk was requested
// one iteration ahead, and the positions came in as three stride-24-byte scalar LDGs per lane.  Here a
// CTA is 8 consumer warps + 1 producer warp.  The producer (one lane) draws chunks, arms the stage's
// `full` mbarrier with the chunk's byte count and issues ONE bulk copy of the 6 KB chunk, RING stages
// ahead of the consumers; a consumer warp waits on `full`, reads its 32 positions from shared memory
// (stride 24 B: conflict-free for 64-bit accesses), arrives on the stage's `empty` mbarrier and goes on
// to the mesh work.  There is no block-wide barrier in the loop: warps drift apart freely.
//
// Arithmetic, index conventions and the order of additions are those of pmb_k_paint_cic_carry32 /
// pmb_k_readout_cic32 (pmb_sched.cuh): results are bit-identical to them (readout: to the reference).
#pragma once
#include "pmb_sched.cuh"
#include "pmb_tma.cuh"

#define PMB_RING 4
#define PMB_RING_THREADS (PMB_CHUNK + 32)

// RW: float64 words per particle record -- 3: (N, 3) position rows; 4: the 32-byte records of the tile-sorted
// copy (pmb_bin.cuh: x, y, z, particle number)
template <int RW>
struct PmbRingSmem {
    double pos[PMB_RING][PMB_CHUNK * RW];    // 4 x 6 KB (RW = 3)
    uint64_t full[PMB_RING], empty[PMB_RING];
    long long chunk[PMB_RING];               // chunk id of the stage, -1: end of work
    int first[PMB_RING];                     // paint: the chunk starts a new run (carry must be flushed)
};

// producer side: copy chunk `c` (particles [c * CHUNK, ...)) into stage s.  A chunk whose byte count is
// not a multiple of 16 (odd particle count at the very end) gets its last 8 bytes by a plain store.
template <int RW>
__device__ __forceinline__ void pmb_ring_fill(PmbRingSmem<RW> &sm, int s, const double *pos, int64_t c, int64_t npart, uint64_t policy)
{
    const int64_t first = c * PMB_CHUNK;
    const int cnt = (int) min((int64_t) PMB_CHUNK, npart - first);
    const uint32_t bytes = (uint32_t) cnt * (8u * RW);
    const uint32_t b16 = bytes & ~15u;
    if (bytes != b16) sm.pos[s][cnt * RW - 1] = __ldcs(pos + RW * first + cnt * RW - 1);
    if (b16) {
        pmb_mbar_arrive_expect_tx(&sm.full[s], b16);
        pmb_bulk_g2s(sm.pos[s], pos + RW * first, b16, &sm.full[s], policy);
    } else {
        pmb_mbar_arrive(&sm.full[s]);
    }
}

struct PmbFields {
    const void *mesh[3];
    void *out[3];
    int64_t out_stride[3];
    int out_elsize;
    // fused ghost-sum routing of the results (Layout.gather('sum') without its pass over the particles):
    // particles j in [sel0, sel1) are this rank's own (the block it sent to itself); their value goes
    // straight to row oidx[j - sel0] of the float64 column out2[q] (the original particle order); every other
    // particle is a ghost held for another rank: its value goes to the compact column out[q] at
    // j (j < sel0) or j - (sel1 - sel0).  sel1 <= sel0: plain readout, out[q][j].
    double *out2[3];
    const int32_t *oidx;
    int64_t sel0, sel1;
    // staging rows of the tile-sorted copy (pmb_bin.cuh): out[q] = row + q, rows of 4 doubles (32 bytes) for three
    // fields -- a kernel that has the three values of a particle at hand writes / reads the row as ONE 32-byte access
    int packed_rows;
};

// where the result of particle j for field q goes (see PmbFields)
__device__ __forceinline__ void pmb_store_result(const PmbFields &f, int q, int64_t j, double value)
{
    if (f.sel1 > f.sel0) {
        if (j >= f.sel0 && j < f.sel1) {
            f.out2[q][f.oidx[j - f.sel0]] = 0.0 + value;      // bincount starts from 0.0: -0.0 becomes +0.0
            return;
        }
        if (j >= f.sel1) j -= f.sel1 - f.sel0;
    }
    pmb_st_real_stream(f.out[q], j * f.out_stride[q], f.out_elsize, value);
}

// ---- readout of NF fields in one sweep ---------------------------------------------------------------
// indices and weights are computed once per particle and used for NF gathers (the three force components
// of the PM step share one pass over the positions: 24 + NF * 16 bytes per particle instead of NF * 40)
template <typename MeshT, bool CHECK, int NF, int MINB, int RW = 3>
__global__ void __launch_bounds__(PMB_RING_THREADS, MINB)
pmb_k_readout_cic32_ring(PmbGeom32 g, const double *__restrict__ pos, PmbFields f, int64_t npart,
                         int64_t nchunks, unsigned long long *ticket, const uint32_t *__restrict__ order = NULL)
{
    __shared__ __align__(128) PmbRingSmem<RW> sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < PMB_RING; s++) { pmb_mbar_init(&sm.full[s], 1); pmb_mbar_init(&sm.empty[s], PMB_CHUNK / 32); }
        pmb_mbar_init_fence();
    }
    __syncthreads();
    if (warp == PMB_CHUNK / 32) {
        // ---- producer ----
        if (lane == 0) {
            const uint64_t pol = pmb_policy_evict_first();
            for (int it = 0;; it++) {
                const int s = it % PMB_RING;
                if (it >= PMB_RING) pmb_mbar_wait(&sm.empty[s], ((it / PMB_RING) - 1) & 1);
                const unsigned long long tk = atomicAdd(ticket, 1ull);
                if ((int64_t) tk >= nchunks) {
                    sm.chunk[s] = -1;
                    pmb_mbar_arrive(&sm.full[s]);
                    break;
                }
                // `order`: the spatial schedule of the chunks (pmb_sched_prepare) instead of memory order
                const int64_t c = order ? (int64_t) order[tk] : (int64_t) tk;
                sm.chunk[s] = (long long) c;
                pmb_ring_fill(sm, s, pos, c, npart, pol);
            }
        }
        return;
    }
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    for (int it = 0;; it++) {
        const int s = it % PMB_RING;
        pmb_mbar_wait(&sm.full[s], (it / PMB_RING) & 1);
        const long long c = sm.chunk[s];
        if (c < 0) break;
        const int64_t i = c * PMB_CHUNK + threadIdx.x;
        const bool active = i < npart;
        double x0 = 0, x1 = 0, x2 = 0;
        if (active) {
            x0 = sm.pos[s][RW * threadIdx.x];
            x1 = sm.pos[s][RW * threadIdx.x + 1];
            x2 = sm.pos[s][RW * threadIdx.x + 2];
        }
        __syncwarp();
        if (lane == 0) pmb_mbar_arrive(&sm.empty[s]);
        if (!active) continue;
        double Vx[2], Vy[2], Vz[2];
        int ex[2], ey[2], ez[2];
        pmb_cic_axis32<CHECK>(x0, g.scale[0], g.translate[0], g.period[0], g.size[0], g.estride[0], Vx[0], Vx[1], ex[0], ex[1]);
        pmb_cic_axis32<CHECK>(x1, g.scale[1], g.translate[1], g.period[1], g.size[1], g.estride[1], Vy[0], Vy[1], ey[0], ey[1]);
        pmb_cic_axis32<CHECK>(x2, g.scale[2], g.translate[2], g.period[2], g.size[2], g.estride[2], Vz[0], Vz[1], ez[0], ez[1]);
        double mv[NF][2][2][2];
#pragma unroll
        for (int q = 0; q < NF; q++)
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++)
#pragma unroll
                    for (int cc = 0; cc < 2; cc++) {
                        const bool ok = !CHECK || (ex[a] >= 0 && ey[b] >= 0 && ez[cc] >= 0);
                        mv[q][a][b][cc] = ok ? pmb_mesh_load<MeshT, false>((const char *) f.mesh[q],
                                                                          (int64_t) (ex[a] + ey[b] + ez[cc]) * sizeof(MeshT), policy)
                                             : 0.0;
                    }
        // the reference's sum: value += mesh * ((Vx * Vy) * Vz), points in C order
        double w[2][2][2];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++)
#pragma unroll
                for (int cc = 0; cc < 2; cc++) w[a][b][cc] = (Vx[a] * Vy[b]) * Vz[cc];
#pragma unroll
        for (int q = 0; q < NF; q++) {
            double value = 0;
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++)
#pragma unroll
                    for (int cc = 0; cc < 2; cc++) {
                        const bool ok = !CHECK || (ex[a] >= 0 && ey[b] >= 0 && ez[cc] >= 0);
                        if (ok) value += mv[q][a][b][cc] * w[a][b][cc];
                    }
            pmb_store_result(f, q, i, value);
        }
    }
}

// ---- paint: y-carry + z aggregation (pmb_k_paint_cic_carry32) fed by the ring -------------------------
template <typename MeshT, bool CHECK, int MINB, int RW = 3>
__global__ void __launch_bounds__(PMB_RING_THREADS, MINB)
pmb_k_paint_cic_carry32_ring(PmbGeom32 g, const double *__restrict__ pos, PmbParticles p, MeshT *mesh, int64_t npart,
                             const uint32_t *__restrict__ order, int64_t nchunks, int unit)
{
    __shared__ __align__(128) PmbRingSmem<RW> sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < PMB_RING; s++) { pmb_mbar_init(&sm.full[s], 1); pmb_mbar_init(&sm.empty[s], PMB_CHUNK / 32); }
        pmb_mbar_init_fence();
    }
    __syncthreads();
    const int64_t nunits = (nchunks + unit - 1) / unit;
    if (warp == PMB_CHUNK / 32) {
        if (lane == 0) {
            const uint64_t pol = pmb_policy_evict_first();
            int it = 0;
            for (int64_t u = blockIdx.x; u < nunits; u += gridDim.x) {
                const int64_t cend = min((u + 1) * (int64_t) unit, nchunks);
                for (int64_t cb = u * unit; cb < cend; cb++, it++) {
                    const int s = it % PMB_RING;
                    if (it >= PMB_RING) pmb_mbar_wait(&sm.empty[s], ((it / PMB_RING) - 1) & 1);
                    const int64_t c = order ? (int64_t) order[cb] : cb;
                    sm.chunk[s] = (long long) c;
                    sm.first[s] = cb == u * unit;
                    pmb_ring_fill(sm, s, pos, c, npart, pol);
                }
            }
            const int s = it % PMB_RING;
            if (it >= PMB_RING) pmb_mbar_wait(&sm.empty[s], ((it / PMB_RING) - 1) & 1);
            sm.chunk[s] = -1;
            pmb_mbar_arrive(&sm.full[s]);
        }
        return;
    }
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    double cv00 = 0, cv01 = 0, cv10 = 0, cv11 = 0;     // carried (a, b = 1, c) values
    int co00 = -1, co01 = -1, co10 = -1, co11 = -1;     // and their element indices (-1: none)
    for (int it = 0;; it++) {
        const int s = it % PMB_RING;
        pmb_mbar_wait(&sm.full[s], (it / PMB_RING) & 1);
        const long long c = sm.chunk[s];
        const bool newrun = c < 0 || sm.first[s];
        double x0 = 0, x1 = 0, x2 = 0;
        const int64_t i = c * PMB_CHUNK + threadIdx.x;
        const bool active = c >= 0 && i < npart;
        if (active) {
            x0 = sm.pos[s][RW * threadIdx.x];
            x1 = sm.pos[s][RW * threadIdx.x + 1];
            x2 = sm.pos[s][RW * threadIdx.x + 2];
        }
        __syncwarp();
        if (lane == 0 && c >= 0) pmb_mbar_arrive(&sm.empty[s]);
        if (newrun) {
            // end of a run: whatever is still carried goes out
            if (co00 >= 0) pmb_red<MeshT>((char *) mesh, (int64_t) co00 * sizeof(MeshT), cv00, policy);
            if (co01 >= 0) pmb_red<MeshT>((char *) mesh, (int64_t) co01 * sizeof(MeshT), cv01, policy);
            if (co10 >= 0) pmb_red<MeshT>((char *) mesh, (int64_t) co10 * sizeof(MeshT), cv10, policy);
            if (co11 >= 0) pmb_red<MeshT>((char *) mesh, (int64_t) co11 * sizeof(MeshT), cv11, policy);
            co00 = co01 = co10 = co11 = -1;
        }
        if (c < 0) break;
        const double m = active ? pmb_load_mass(p, i) : 0.0;
        double Vx0, Vx1, Vy0, Vy1, Vz0, Vz1;
        int ex0, ex1, ey0, ey1, ez0, ez1;
        pmb_cic_axis32<CHECK>(x0, g.scale[0], g.translate[0], g.period[0], g.size[0], g.estride[0], Vx0, Vx1, ex0, ex1);
        pmb_cic_axis32<CHECK>(x1, g.scale[1], g.translate[1], g.period[1], g.size[1], g.estride[1], Vy0, Vy1, ey0, ey1);
        pmb_cic_axis32<CHECK>(x2, g.scale[2], g.translate[2], g.period[2], g.size[2], g.estride[2], Vz0, Vz1, ez0, ez1);
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
}
