// pmb_perm.cuh -- tile-binned traversal for particle sets WITHOUT spatial order in memory.
//
// The chunk schedule of pmb_sched.cuh reorders 256-particle chunks; it needs every chunk to be compact
// in space (particles kept in displaced-lattice order).  For a particle array in random order it does
// nothing: every stencil is a DRAM miss (measured, 1024^3 uniform random, B200: paint 241 ms, readout
// 156 ms against 11 ms each on lattice order -- profiles/r2_inputs_before.json).  Here the particles
// themselves are binned: a key (tile of T0 x T1 x 128 cells, C order, <= 2^16 tiles) per particle,
// stable radix sort of (key, particle id) (cub, library; 2 digit passes), and the kernels walk the
// particles THROUGH the permutation.  Concurrently running CTAs then work on a handful of neighbouring
// tiles: mesh accesses hit L2, each DRAM line of the mesh moves once.  What remains random is the 24-byte
// position fetch and the 8-byte result store per particle.
//
// Like the chunk schedule, the permutation is cached per (position array, count, geometry) and rebuilt
// every few uses: ANY permutation gives the right answer (the scatter is atomic, the gather writes
// out[perm[j]]), so a stale one costs speed, never correctness.  A probe of a few hundred chunks decides
// which of the two schedules applies.
#pragma once
#include <cub/cub.cuh>

#include "pmb_ring.cuh"

struct PmbTiling {
    int t0, t1;            // tile extent along axes 0, 1 (cells); axis 2: 128 cells
    int n0, n1, n2;        // tiles per axis
};

__device__ __forceinline__ int pmb_cell_of(double x, double scale, double translate, int64_t period, int64_t size)
{
    const double X = pmb_gridpos(x, scale, translate);
    int64_t t = (int64_t) floor(X);
    const int64_t per = period > 0 ? period : (size > 0 ? size : 1);
    t %= per;
    if (t < 0) t += per;
    if (t >= size) t = size - 1;     // cells beyond a slab canvas: park them in the last tile
    return (int) t;
}

// ---- probe: are 256-particle chunks compact in space? -----------------------------------------------
// sample chunk s: cells of its first, middle and last particle; "scattered" when they are more than 32
// cells apart (periodic distance) along axis 0 or 1
__global__ void pmb_k_probe_chunks(PmbGeom g, PmbParticles p, int64_t npart, int64_t nchunks, int nsamples, unsigned int *scattered)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsamples) return;
    const int64_t c = (int64_t) ((double) s * (double) nchunks / (double) nsamples);
    const int64_t i0 = c * PMB_CHUNK;
    int64_t i2 = i0 + PMB_CHUNK - 1;
    if (i2 >= npart) i2 = npart - 1;
    const int64_t i1 = (i0 + i2) / 2;
    const int64_t ii[3] = {i0, i1, i2};
    int cell[3][2];
    for (int k = 0; k < 3; k++) {
        double x[3];
        pmb_load_pos<3>(p, ii[k], x, 0);
        for (int d = 0; d < 2; d++) cell[k][d] = pmb_cell_of(x[d], g.scale[d], g.translate[d], g.period[d], g.size[d]);
    }
    bool far = false;
    for (int d = 0; d < 2; d++) {
        const int per = (int) (g.period[d] > 0 ? g.period[d] : 0);
        for (int k = 1; k < 3; k++) {
            int dist = abs(cell[k][d] - cell[0][d]);
            if (per > 0 && dist > per / 2) dist = per - dist;
            if (dist > 32) far = true;
        }
    }
    if (far) atomicAdd(scattered, 1u);
}

__global__ void __launch_bounds__(256)
pmb_k_tile_keys(PmbGeom g, PmbParticles p, int64_t npart, PmbTiling t, uint16_t *keys, uint32_t *ids)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; i < npart; i += stride) {
        double x[3];
        pmb_load_pos<3>(p, i, x, 0);
        const int c0 = pmb_cell_of(x[0], g.scale[0], g.translate[0], g.period[0], g.size[0]) / t.t0;
        const int c1 = pmb_cell_of(x[1], g.scale[1], g.translate[1], g.period[1], g.size[1]) / t.t1;
        const int c2 = pmb_cell_of(x[2], g.scale[2], g.translate[2], g.period[2], g.size[2]) >> 7;
        keys[i] = (uint16_t) ((c0 * t.n1 + c1) * t.n2 + c2);
        ids[i] = (uint32_t) i;
    }
}

static void pmb_tiling(const PmbGeom &g, PmbTiling *t)
{
    t->t0 = t->t1 = 8;
    for (;;) {
        t->n0 = (int) ((g.size[0] + t->t0 - 1) / t->t0);
        t->n1 = (int) ((g.size[1] + t->t1 - 1) / t->t1);
        t->n2 = (int) ((g.size[2] + 127) / 128);
        if ((int64_t) t->n0 * t->n1 * t->n2 <= 65536) break;
        if (t->t0 <= t->t1) t->t0 *= 2; else t->t1 *= 2;
    }
}

// *perm = device permutation (tile order) when the particle array is scattered in memory, NULL when its
// chunks are compact (the chunk schedule then applies).  PMB_PERM=0 disables, 2 forces the permutation.
static int pmb_perm_prepare(pmb_ctx *ctx, const PmbGeom &g, const PmbParticles &p, int64_t npart, const uint32_t **perm)
{
    *perm = NULL;
    const int mode = pmb_env_flag("PMB_PERM", 1);
    if (!mode || npart < ((int64_t) 1 << 18) || npart >= ((int64_t) 1 << 31)) return PMB_OK;
    uint64_t sig = (uint64_t) (uintptr_t) p.pos * 0x9E3779B97F4A7C15ull ^ (uint64_t) npart * 0xD6E8FEB86659FD93ull
                   ^ (uint64_t) p.ps0 ^ ((uint64_t) g.size[0] << 40) ^ ((uint64_t) g.size[1] << 20) ^ (uint64_t) g.size[2];
    uint64_t tr;
    memcpy(&tr, &g.translate[0], sizeof(tr));
    sig ^= tr * 0x94D049BB133111EBull;
    if (ctx->perm_sig == sig && ctx->perm_npart == npart && ctx->perm_uses < 8) {
        ctx->perm_uses++;
        if (ctx->perm_state == 1) *perm = (const uint32_t *) ctx->perm_ids;
        return PMB_OK;
    }
    // ---- probe ----
    const int64_t nchunks = (npart + PMB_CHUNK - 1) / PMB_CHUNK;
    const int nsamples = 512;
    unsigned int *d_count;
    PMB_CHECK(pmb_scratch(ctx, 256, (void **) &d_count));
    PMB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned int), ctx->stream));
    pmb_k_probe_chunks<<<(nsamples + 127) / 128, 128, 0, ctx->stream>>>(g, p, npart, nchunks, nsamples, d_count);
    PMB_LAUNCH_CHECK(ctx);
    unsigned int scattered = 0;
    PMB_CUDA(cudaMemcpyAsync(&scattered, d_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    PMB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->perm_sig = sig;
    ctx->perm_npart = npart;
    ctx->perm_uses = 1;
    ctx->perm_state = (mode >= 2 || scattered * 4 > (unsigned) nsamples) ? 1 : 0;
    if (ctx->perm_state == 0) return PMB_OK;
    // ---- keys + stable sort ----
    PmbTiling t;
    pmb_tiling(g, &t);
    int bits = 1;
    while (((int64_t) 1 << bits) < (int64_t) t.n0 * t.n1 * t.n2) bits++;
    const size_t b_keys = (sizeof(uint16_t) * npart + 255) & ~(size_t) 255;
    const size_t b_ids = (sizeof(uint32_t) * npart + 255) & ~(size_t) 255;
    size_t temp = 0;
    PMB_CUDA(cub::DeviceRadixSort::SortPairs(NULL, temp, (uint16_t *) NULL, (uint16_t *) NULL, (uint32_t *) NULL,
                                             (uint32_t *) NULL, npart, 0, bits, ctx->stream));
    // persistent: the sorted ids; transient (scratch): keys in / out, ids in, cub temp
    if (b_ids > ctx->perm_bytes) {
        if (ctx->perm_ids) { PMB_CUDA(cudaStreamSynchronize(ctx->stream)); PMB_CUDA(cudaFree(ctx->perm_ids)); ctx->perm_ids = NULL; ctx->perm_bytes = 0; }
        PMB_CUDA(cudaMalloc(&ctx->perm_ids, b_ids));
        ctx->perm_bytes = b_ids;
    }
    void *ws;
    PMB_CHECK(pmb_scratch(ctx, 2 * b_keys + b_ids + temp + 256, &ws));
    uint16_t *k0 = (uint16_t *) ws, *k1 = (uint16_t *) ((char *) ws + b_keys);
    uint32_t *ids = (uint32_t *) ((char *) ws + 2 * b_keys);
    pmb_k_tile_keys<<<pmb_grid(ctx, npart, 256, 8), 256, 0, ctx->stream>>>(g, p, npart, t, k0, ids);
    PMB_LAUNCH_CHECK(ctx);
    PMB_CUDA(cub::DeviceRadixSort::SortPairs((char *) ws + 2 * b_keys + b_ids, temp, k0, k1, ids, (uint32_t *) ctx->perm_ids,
                                             npart, 0, bits, ctx->stream));
    ctx->launches += 4;
    *perm = (const uint32_t *) ctx->perm_ids;
    return PMB_OK;
}

// ---- kernels that walk the particles through the permutation -----------------------------------------
// CIC, 32-bit element indices (the lean arithmetic of pmb_k_paint_cic_carry32); plain reds: the tile's
// mesh lines are L2 resident
// `ticket` non-NULL: 256-particle chunks are handed out through a ticket counter instead of the grid-stride loop
// (all CTAs stay inside one compact window of the array)
template <typename MeshT, bool CHECK>
__global__ void __launch_bounds__(256)
pmb_k_paint_cic32_perm(PmbGeom32 g, PmbParticles p, MeshT *mesh, int64_t npart, const uint32_t *__restrict__ perm,
                       unsigned long long *ticket)
{
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    auto body = [&](int64_t j) {
        const int64_t i = perm ? (int64_t) perm[j] : j;     // NULL: the array is already in tile order (pmb_bin.cuh)
        double x[3];
        pmb_load_pos<3>(p, i, x);
        const double m = pmb_load_mass(p, i);
        double Vx[2], Vy[2], Vz[2];
        int ex[2], ey[2], ez[2];
        pmb_cic_axis32<CHECK>(x[0], g.scale[0], g.translate[0], g.period[0], g.size[0], g.estride[0], Vx[0], Vx[1], ex[0], ex[1]);
        pmb_cic_axis32<CHECK>(x[1], g.scale[1], g.translate[1], g.period[1], g.size[1], g.estride[1], Vy[0], Vy[1], ey[0], ey[1]);
        pmb_cic_axis32<CHECK>(x[2], g.scale[2], g.translate[2], g.period[2], g.size[2], g.estride[2], Vz[0], Vz[1], ez[0], ez[1]);
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++)
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    if (CHECK && (ex[a] < 0 || ey[b] < 0 || ez[c] < 0)) continue;
                    pmb_red<MeshT>((char *) mesh, (int64_t) (ex[a] + ey[b] + ez[c]) * sizeof(MeshT), ((Vx[a] * m) * Vy[b]) * Vz[c], policy);
                }
    };
    if (ticket) {
        __shared__ unsigned long long s_tk;
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) s_tk = atomicAdd(ticket, 1ull);
            __syncthreads();
            const int64_t j = (int64_t) s_tk * 256 + threadIdx.x;
            if ((int64_t) s_tk * 256 >= npart) break;
            if (j < npart) body(j);
        }
        return;
    }
    int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; j < npart; j += stride) body(j);
}

template <typename MeshT, bool CHECK, int NF>
__global__ void __launch_bounds__(256)
pmb_k_readout_cic32_perm(PmbGeom32 g, PmbParticles p, PmbFields f, int64_t npart, const uint32_t *__restrict__ perm)
{
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; j < npart; j += stride) {
        const int64_t i = perm ? (int64_t) perm[j] : j;     // NULL: the array is already in tile order (pmb_bin.cuh)
        double x[3];
        pmb_load_pos<3>(p, i, x);
        double Vx[2], Vy[2], Vz[2];
        int ex[2], ey[2], ez[2];
        pmb_cic_axis32<CHECK>(x[0], g.scale[0], g.translate[0], g.period[0], g.size[0], g.estride[0], Vx[0], Vx[1], ex[0], ex[1]);
        pmb_cic_axis32<CHECK>(x[1], g.scale[1], g.translate[1], g.period[1], g.size[1], g.estride[1], Vy[0], Vy[1], ey[0], ey[1]);
        pmb_cic_axis32<CHECK>(x[2], g.scale[2], g.translate[2], g.period[2], g.size[2], g.estride[2], Vz[0], Vz[1], ez[0], ez[1]);
#pragma unroll
        for (int q = 0; q < NF; q++) {
            double value = 0;
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++)
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        if (CHECK && (ex[a] < 0 || ey[b] < 0 || ez[c] < 0)) continue;
                        value += pmb_mesh_load<MeshT, false>((const char *) f.mesh[q], (int64_t) (ex[a] + ey[b] + ez[c]) * sizeof(MeshT), policy)
                                 * ((Vx[a] * Vy[b]) * Vz[c]);
                    }
            pmb_store_result(f, q, i, value);
        }
    }
}

// every tuned window (and gradient windows): the generic fixed-support stencil walk
template <typename MeshT, int FAM, bool CHECK>
__global__ void __launch_bounds__(256)
pmb_k_paint_perm(PmbGeom g, PmbParticles p, char *mesh, int64_t npart, int pcsfix, const uint32_t *__restrict__ perm,
                 unsigned long long *ticket)
{
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    auto body = [&](int64_t j) {
        const int64_t i = perm ? (int64_t) perm[j] : j;     // NULL: the array is already in tile order (pmb_bin.cuh)
        double x[3];
        pmb_load_pos<3>(p, i, x);
        const double m = pmb_load_mass(p, i);
        PmbAxes<3, FAM> A;
        pmb_axes_tuned<3, FAM, CHECK>(g, g.order, x, pcsfix, A);
        pmb_for_points_fixed<3, FAM, CHECK>(A, [&](int, int64_t off, double v0, double v1, double v2) {
            if (!CHECK || off != PMB_OFF_INVALID) pmb_red<MeshT>(mesh, off, pmb_paint_value(true, m, v0, v1, v2), policy);
        });
    };
    if (ticket) {
        __shared__ unsigned long long s_tk;
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) s_tk = atomicAdd(ticket, 1ull);
            __syncthreads();
            const int64_t j = (int64_t) s_tk * 256 + threadIdx.x;
            if ((int64_t) s_tk * 256 >= npart) break;
            if (j < npart) body(j);
        }
        return;
    }
    int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; j < npart; j += stride) body(j);
}

template <typename MeshT, int FAM, bool CHECK>
__global__ void __launch_bounds__(256)
pmb_k_readout_perm(PmbGeom g, PmbParticles p, const char *mesh, int64_t npart, int pcsfix,
                   void *out, int out_elsize, int64_t out_stride, const uint32_t *__restrict__ perm)
{
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; j < npart; j += stride) {
        const int64_t i = perm ? (int64_t) perm[j] : j;     // NULL: the array is already in tile order (pmb_bin.cuh)
        double x[3];
        pmb_load_pos<3>(p, i, x);
        PmbAxes<3, FAM> A;
        pmb_axes_tuned<3, FAM, CHECK>(g, g.order, x, pcsfix, A);
        double value = 0;
        pmb_for_points_fixed<3, FAM, CHECK>(A, [&](int, int64_t off, double v0, double v1, double v2) {
            if (!CHECK || off != PMB_OFF_INVALID) value += pmb_mesh_load<MeshT, false>(mesh, off, policy) * ((v0 * v1) * v2);
        });
        pmb_st_real_stream(out, i * out_stride, out_elsize, value);
    }
}
