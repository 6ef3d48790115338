// pmb_resample.cu -- paint (scatter-add) and readout (gather) kernels.
//
// Replaces the per-particle Cython/C loops of the reference
// (pmesh/_window.pyx:157-165,198-205 -> pmesh/_window_imp.c:461-471 -> _window_generics.h /
// _window_tuned_*.h).  Three paint paths:
//   * atomic   : one thread per particle, red.global.add.{f64,f32} per stencil point.
//   * deterministic : expand (cell, value) pairs in particle order -> stable radix sort by cell ->
//                 sequential per-cell sum `acc = (T)((double)acc + f)`; bit-equal to the reference,
//                 which visits particles in index order (proved in SURVEY section 7).
// Readout is a gather of the same stencil, summed in the reference's point order.
#include <cub/cub.cuh>
#include <stdlib.h>

#include "pmb_sched.cuh"
#include "pmb_ring.cuh"
#include "pmb_bin.cuh"
#include "pmb_pull.cuh"
#include "pmb_tile.cuh"

// ------------------------------------------------------------------ atomic paint
// L2 residency control: mesh cells are re-touched by particles of neighbouring lattice rows /
// planes, the particle columns are touched once.  Mesh accesses carry an evict_last policy, the
// particle stream is loaded with ld.global.cs (evict-first), so the stream cannot push the mesh
// working set out of L2.  PMB_CACHE_MODE=0 disables the mesh policy (for A/B measurements).
static int pmb_cache_mode(void)
{
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("PMB_CACHE_MODE");
        mode = e ? atoi(e) : 1;
    }
    return mode;
}

__device__ __forceinline__ uint64_t pmb_policy_evict_last(void)
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

template <typename MeshT>
__device__ __forceinline__ void pmb_red_add(char *mesh, int64_t off, double f, uint64_t policy, int hint)
{
    if (hint) {
        if (sizeof(MeshT) == 8)
            asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(mesh + off), "d"(f), "l"(policy) : "memory");
        else
            asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(mesh + off), "f"((float) f), "l"(policy) : "memory");
    } else {
        atomicAdd((MeshT *) (mesh + off), (MeshT) f);
    }
}

template <typename MeshT, int NDIM, int FAM, bool CHECK>
__global__ void __launch_bounds__(256)
pmb_k_paint_tuned(PmbGeom g, PmbParticles p, char *mesh, int64_t npart, int pcsfix, int hint)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const uint64_t policy = pmb_policy_evict_last();
    for (; i < npart; i += stride) {
        double x[NDIM];
        pmb_load_pos<NDIM>(p, i, x);
        const double m = pmb_load_mass(p, i);
        PmbAxes<NDIM, FAM> A;
        pmb_axes_tuned<NDIM, FAM, CHECK>(g, g.order, x, pcsfix, A);
        pmb_for_points_fixed<NDIM, FAM, CHECK>(A, [&](int, int64_t off, double v0, double v1, double v2) {
            if (!CHECK || off != PMB_OFF_INVALID) pmb_red_add<MeshT>(mesh, off, pmb_paint_value(true, m, v0, v1, v2), policy, hint);
        });
    }
}

template <typename MeshT, int NDIM>
__global__ void __launch_bounds__(128)
pmb_k_paint_dyn(PmbGeom g, PmbWindow w, PmbParticles p, char *mesh, int64_t npart, int pcsfix)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; i < npart; i += stride) {
        double x[NDIM];
        pmb_load_pos<NDIM>(p, i, x);
        const double m = pmb_load_mass(p, i);
        const double h = pmb_load_hsml(p, i);
        PmbWinInfo info;
        pmb_window_info(w.nativesupport, w.support * h, &info);
        if (info.support <= PMB_MAX_SUPPORT) {
            PmbAxes<NDIM, PMB_MAX_SUPPORT> A;
            pmb_axes_dyn<NDIM>(g, w, info, g.order, x, pcsfix, A);
            const bool tuned = A.tuned;
            pmb_for_points_dyn<NDIM, PMB_MAX_SUPPORT>(A, [&](int, int64_t off, double v0, double v1, double v2) {
                if (off != PMB_OFF_INVALID) pmb_red_add<MeshT>(mesh, off, pmb_paint_value(tuned, m, v0, v1, v2), 0, 0);
            });
        } else {
            pmb_for_points_wide<NDIM>(g, w, info, g.order, x, [&](int, int64_t off, double v0, double v1, double v2) {
                if (off != PMB_OFF_INVALID) pmb_red_add<MeshT>(mesh, off, pmb_paint_value(false, m, v0, v1, v2), 0, 0);
            });
        }
    }
}

// ------------------------------------------------------------------ readout
template <typename MeshT>
__device__ __forceinline__ double pmb_mesh_ld(const char *mesh, int64_t off)
{
    return (double) __ldg((const MeshT *) (mesh + off));
}

template <typename MeshT>
__device__ __forceinline__ double pmb_mesh_ld_hint(const char *mesh, int64_t off, uint64_t policy, int hint)
{
    if (!hint) return (double) __ldg((const MeshT *) (mesh + off));
    if (sizeof(MeshT) == 8) {
        double v;
        asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(mesh + off), "l"(policy));
        return v;
    } else {
        float v;
        asm("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(mesh + off), "l"(policy));
        return (double) v;
    }
}

template <typename MeshT, int NDIM, int FAM, bool CHECK>
__global__ void __launch_bounds__(256)
pmb_k_readout_tuned(PmbGeom g, PmbParticles p, const char *mesh, int64_t npart, int pcsfix,
                    void *out, int out_elsize, int64_t out_stride, int hint)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const uint64_t policy = pmb_policy_evict_last();
    for (; i < npart; i += stride) {
        double x[NDIM];
        pmb_load_pos<NDIM>(p, i, x);
        PmbAxes<NDIM, FAM> A;
        pmb_axes_tuned<NDIM, FAM, CHECK>(g, g.order, x, pcsfix, A);
        double value = 0;
        pmb_for_points_fixed<NDIM, FAM, CHECK>(A, [&](int, int64_t off, double v0, double v1, double v2) {
            if (!CHECK || off != PMB_OFF_INVALID) value += pmb_mesh_ld_hint<MeshT>(mesh, off, policy, hint) * ((v0 * v1) * v2);
        });
        pmb_st_real_stream(out, i * out_stride, out_elsize, value);
    }
}

template <typename MeshT, int NDIM>
__global__ void __launch_bounds__(128)
pmb_k_readout_dyn(PmbGeom g, PmbWindow w, PmbParticles p, const char *mesh, int64_t npart, int pcsfix,
                  void *out, int out_elsize, int64_t out_stride)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; i < npart; i += stride) {
        double x[NDIM];
        pmb_load_pos<NDIM>(p, i, x);
        const double h = pmb_load_hsml(p, i);
        PmbWinInfo info;
        pmb_window_info(w.nativesupport, w.support * h, &info);
        double value = 0;
        auto acc = [&](int, int64_t off, double v0, double v1, double v2) {
            if (off != PMB_OFF_INVALID) value += ((v0 * v1) * v2) * pmb_mesh_ld<MeshT>(mesh, off);
        };
        if (info.support <= PMB_MAX_SUPPORT) {
            PmbAxes<NDIM, PMB_MAX_SUPPORT> A;
            pmb_axes_dyn<NDIM>(g, w, info, g.order, x, pcsfix, A);
            pmb_for_points_dyn<NDIM, PMB_MAX_SUPPORT>(A, acc);
        } else {
            pmb_for_points_wide<NDIM>(g, w, info, g.order, x, acc);
        }
        pmb_st_real_stream(out, i * out_stride, out_elsize, value);
    }
}

// fused value + NDIM gradients: one sweep over the neighbourhood, four accumulators.  Each
// accumulator sees the same addends in the same order as a separate readout(gradient=d) would.
template <typename MeshT, int NDIM, int FAM, bool CHECK>
__global__ void __launch_bounds__(256)
pmb_k_readout_grad_tuned(PmbGeom g, PmbParticles p, const char *mesh, int64_t npart, int pcsfix,
                         void *out, int out_elsize, int64_t out_stride,
                         void *grad, int64_t gs0, int64_t gs1)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int zero[3] = {0, 0, 0};
    const int one[3] = {1, 1, 1};
    for (; i < npart; i += stride) {
        double x[NDIM];
        pmb_load_pos<NDIM>(p, i, x);
        PmbAxes<NDIM, FAM> A, D;
        pmb_axes_tuned<NDIM, FAM, CHECK>(g, zero, x, pcsfix, A);
        pmb_axes_tuned<NDIM, FAM, CHECK>(g, one, x, pcsfix, D);
        double value = 0, gr[3] = {0, 0, 0};
        // walk the stencil by ordinal to address both weight sets
        int ord = 0;
#pragma unroll
        for (int a = 0; a < FAM; a++) {
#pragma unroll
            for (int b = 0; b < (NDIM > 1 ? FAM : 1); b++) {
#pragma unroll
                for (int c = 0; c < (NDIM > 2 ? FAM : 1); c++, ord++) {
                    int64_t o0 = A.off[0][a];
                    int64_t o1 = NDIM > 1 ? A.off[NDIM > 1 ? 1 : 0][b] : 0;
                    int64_t o2 = NDIM > 2 ? A.off[NDIM > 2 ? 2 : 0][c] : 0;
                    if (CHECK && (o0 == PMB_OFF_INVALID || o1 == PMB_OFF_INVALID || o2 == PMB_OFF_INVALID)) continue;
                    const double mval = pmb_mesh_ld<MeshT>(mesh, o0 + o1 + o2);
                    const double v0 = A.V[0][a], d0 = D.V[0][a];
                    const double v1 = NDIM > 1 ? A.V[NDIM > 1 ? 1 : 0][b] : 1.0;
                    const double d1 = NDIM > 1 ? D.V[NDIM > 1 ? 1 : 0][b] : 1.0;
                    const double v2 = NDIM > 2 ? A.V[NDIM > 2 ? 2 : 0][c] : 1.0;
                    const double d2 = NDIM > 2 ? D.V[NDIM > 2 ? 2 : 0][c] : 1.0;
                    value += mval * ((v0 * v1) * v2);
                    gr[0] += mval * ((d0 * v1) * v2);
                    if (NDIM > 1) gr[1] += mval * ((v0 * d1) * v2);
                    if (NDIM > 2) gr[2] += mval * ((v0 * v1) * d2);
                }
            }
        }
        if (out) pmb_st_real_stream(out, i * out_stride, out_elsize, value);
#pragma unroll
        for (int d = 0; d < NDIM; d++) pmb_st_real_stream(grad, i * gs0 + d * gs1, out_elsize, gr[d]);
    }
}

// the same for every window with a run-time support (lanczos, acg, wavelets, hsml-scaled windows): the mesh
// neighbourhood (216 loads for lanczos3) is swept ONCE for the value and the NDIM gradients instead of
// 1 + NDIM times; per-axis value weights and derivative weights are evaluated once per particle.  Every
// accumulator sees the addends of the separate readout(gradient = d) in the same order, with the same
// product ((v0 * v1) * v2) * mesh: the results are bit-identical to pmb_k_readout_dyn's.
template <typename MeshT, int NDIM>
__global__ void __launch_bounds__(128)
pmb_k_readout_grad_dyn(PmbGeom g, PmbWindow w, PmbParticles p, const char *mesh, int64_t npart, int pcsfix,
                       void *out, int out_elsize, int64_t out_stride, void *grad, int64_t gs0, int64_t gs1)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int zero[3] = {0, 0, 0};
    for (; i < npart; i += stride) {
        double x[NDIM];
        pmb_load_pos<NDIM>(p, i, x);
        const double h = pmb_load_hsml(p, i);
        PmbWinInfo info;
        pmb_window_info(w.nativesupport, w.support * h, &info);
        PmbAxes<NDIM, PMB_MAX_SUPPORT> A;
        pmb_axes_dyn<NDIM>(g, w, info, zero, x, pcsfix, A);
        const int S = A.S;
        // derivative weights of every axis: the weights pmb_axes_dyn gives that axis for order[d] = 1
        double D[NDIM][PMB_MAX_SUPPORT];
        {
            PmbAxes<NDIM, PMB_MAX_SUPPORT> B;
            int one[3] = {1, 1, 1};
            pmb_axes_dyn<NDIM>(g, w, info, one, x, pcsfix, B);
            for (int d = 0; d < NDIM; d++)
                for (int s = 0; s < S; s++) D[d][s] = B.V[d][s];
        }
        double value = 0, gr[3] = {0, 0, 0};
#pragma unroll 1
        for (int a = 0; a < S; a++) {
            const int64_t o0 = A.off[0][a];
#pragma unroll 1
            for (int b = 0; b < (NDIM > 1 ? S : 1); b++) {
                const int64_t o1 = NDIM > 1 ? A.off[NDIM > 1 ? 1 : 0][b] : 0;
#pragma unroll 1
                for (int c = 0; c < (NDIM > 2 ? S : 1); c++) {
                    const int64_t o2 = NDIM > 2 ? A.off[NDIM > 2 ? 2 : 0][c] : 0;
                    if (o0 == PMB_OFF_INVALID || o1 == PMB_OFF_INVALID || o2 == PMB_OFF_INVALID) continue;
                    const double mval = pmb_mesh_ld<MeshT>(mesh, o0 + o1 + o2);
                    const double v0 = A.V[0][a], d0 = D[0][a];
                    const double v1 = NDIM > 1 ? A.V[NDIM > 1 ? 1 : 0][b] : 1.0;
                    const double d1 = NDIM > 1 ? D[NDIM > 1 ? 1 : 0][b] : 1.0;
                    const double v2 = NDIM > 2 ? A.V[NDIM > 2 ? 2 : 0][c] : 1.0;
                    const double d2 = NDIM > 2 ? D[NDIM > 2 ? 2 : 0][c] : 1.0;
                    value += ((v0 * v1) * v2) * mval;
                    gr[0] += ((d0 * v1) * v2) * mval;
                    if (NDIM > 1) gr[1] += ((v0 * d1) * v2) * mval;
                    if (NDIM > 2) gr[2] += ((v0 * v1) * d2) * mval;
                }
            }
        }
        if (out) pmb_st_real_stream(out, i * out_stride, out_elsize, value);
#pragma unroll
        for (int d = 0; d < NDIM; d++) pmb_st_real_stream(grad, i * gs0 + d * gs1, out_elsize, gr[d]);
    }
}

// ------------------------------------------------------------------ deterministic paint
// per-axis dense indices variant of the stencil walk for the deterministic path: we need the
// C-order cell number (sort key), not the byte offset.  To share all arithmetic with the atomic
// path the geometry handed to these kernels has strides replaced by the dense C-order strides in
// ELEMENTS (g.strides[d] = prod(size[d+1:])), so `off` IS the linear cell index.
template <typename KeyT, int NDIM, int FAM>
__global__ void __launch_bounds__(256)
pmb_k_expand_tuned(PmbGeom g, PmbParticles p, int64_t first, int64_t count, int pcsfix,
                   KeyT *keys, double *vals)
{
    int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    int npts = 1;
#pragma unroll
    for (int d = 0; d < NDIM; d++) npts *= FAM;
    for (; j < count; j += stride) {
        const int64_t i = first + j;
        double x[NDIM];
        pmb_load_pos<NDIM>(p, i, x);
        const double m = pmb_load_mass(p, i);
        PmbAxes<NDIM, FAM> A;
        pmb_axes_tuned<NDIM, FAM, true>(g, g.order, x, pcsfix, A);
        KeyT *kk = keys + j * npts;
        double *vv = vals + j * npts;
        pmb_for_points_fixed<NDIM, FAM, true>(A, [&](int ord, int64_t off, double v0, double v1, double v2) {
            kk[ord] = off == PMB_OFF_INVALID ? (KeyT) ~(KeyT) 0 : (KeyT) off;
            vv[ord] = pmb_paint_value(true, m, v0, v1, v2);
        });
    }
}

template <typename KeyT, int NDIM>
__global__ void __launch_bounds__(128)
pmb_k_expand_dyn(PmbGeom g, PmbWindow w, PmbParticles p, int64_t first, int64_t count, int pcsfix,
                 int64_t npts_max, KeyT *keys, double *vals)
{
    int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; j < count; j += stride) {
        const int64_t i = first + j;
        double x[NDIM];
        pmb_load_pos<NDIM>(p, i, x);
        const double m = pmb_load_mass(p, i);
        const double h = pmb_load_hsml(p, i);
        PmbWinInfo info;
        pmb_window_info(w.nativesupport, w.support * h, &info);
        KeyT *kk = keys + j * npts_max;
        double *vv = vals + j * npts_max;
        // slots beyond this particle's own point count keep the sentinel written by the memset
        if (info.support <= PMB_MAX_SUPPORT) {
            PmbAxes<NDIM, PMB_MAX_SUPPORT> A;
            pmb_axes_dyn<NDIM>(g, w, info, g.order, x, pcsfix, A);
            const bool tuned = A.tuned;
            pmb_for_points_dyn<NDIM, PMB_MAX_SUPPORT>(A, [&](int ord, int64_t off, double v0, double v1, double v2) {
                kk[ord] = off == PMB_OFF_INVALID ? (KeyT) ~(KeyT) 0 : (KeyT) off;
                vv[ord] = pmb_paint_value(tuned, m, v0, v1, v2);
            });
        } else {
            pmb_for_points_wide<NDIM>(g, w, info, g.order, x, [&](int ord, int64_t off, double v0, double v1, double v2) {
                kk[ord] = off == PMB_OFF_INVALID ? (KeyT) ~(KeyT) 0 : (KeyT) off;
                vv[ord] = pmb_paint_value(false, m, v0, v1, v2);
            });
        }
    }
}

// one thread per sorted pair; the thread sitting on the first pair of a cell walks the cell's run
// and performs the reference's `*(FLOAT*)p += f` sequence (_window_generics.h:155).
template <typename KeyT, typename MeshT>
__global__ void __launch_bounds__(256)
pmb_k_segment_sum(const KeyT *keys, const double *vals, int64_t n, char *mesh,
                  int64_t sz1, int64_t sz2, int64_t st0, int64_t st1, int64_t st2)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const KeyT sentinel = (KeyT) ~(KeyT) 0;
    for (; i < n; i += stride) {
        const KeyT k = keys[i];
        if (k == sentinel) continue;
        if (i > 0 && keys[i - 1] == k) continue;
        int64_t lin = (int64_t) k;
        int64_t c = lin % sz2;
        int64_t r = lin / sz2;
        int64_t b = r % sz1;
        int64_t a = r / sz1;
        MeshT *cell = (MeshT *) (mesh + a * st0 + b * st1 + c * st2);
        MeshT acc = *cell;
        for (int64_t j = i; j < n && keys[j] == k; j++) acc = (MeshT) ((double) acc + vals[j]);
        *cell = acc;
    }
}

template <typename T>
__global__ void pmb_k_max_real(const void *a, int elsize, int64_t stride_bytes, int64_t n, double *out)
{
    __shared__ double sh[32];
    double acc = -1e300;
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) acc = fmax(acc, pmb_ld_real(a, i * stride_bytes, elsize));
    for (int o = 16; o > 0; o >>= 1) acc = fmax(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        acc = lane < (blockDim.x >> 5) ? sh[lane] : -1e300;
        for (int o = 16; o > 0; o >>= 1) acc = fmax(acc, __shfl_xor_sync(0xffffffffu, acc, o));
        if (lane == 0) out[blockIdx.x] = acc;
    }
}

// ------------------------------------------------------------------ host side
static int check_args(pmb_ctx *ctx, const pmb_resample_args *a, int need_out)
{
    PMB_REQUIRE(ctx && a, "null argument");
    PMB_REQUIRE(a->ndim >= 1 && a->ndim <= 3, "ndim %d not supported (1..3)", a->ndim);
    PMB_REQUIRE(a->mesh_elsize == 4 || a->mesh_elsize == 8, "canvas must be float32 or float64");
    PMB_REQUIRE(a->pos_elsize == 4 || a->pos_elsize == 8, "pos must be float32 or float64");
    PMB_REQUIRE(a->npart >= 0, "negative particle count");
    PMB_REQUIRE(a->npart == 0 || (a->pos && a->mesh), "null pos / mesh");
    if (a->mass) PMB_REQUIRE(a->mass_elsize == 4 || a->mass_elsize == 8, "mass must be float32 or float64");
    if (a->hsml) PMB_REQUIRE(a->hsml_elsize == 4 || a->hsml_elsize == 8, "hsml must be float32 or float64");
    if (need_out) {
        PMB_REQUIRE(a->npart == 0 || a->out, "null out");
        PMB_REQUIRE(a->out_elsize == 4 || a->out_elsize == 8, "out must be float32 or float64");
    }
    for (int d = 0; d < a->ndim; d++) {
        PMB_REQUIRE(a->size[d] < ((int64_t) 1 << 31) && a->period[d] < ((int64_t) 1 << 31), "canvas extent / period must be < 2^31");
        PMB_REQUIRE(a->size[d] >= 0, "negative canvas size");
        PMB_REQUIRE(a->order[d] >= 0, "negative order");
    }
    return PMB_OK;
}

static void fill_geom(const pmb_resample_args *a, PmbGeom *g)
{
    memset(g, 0, sizeof(*g));
    g->ndim = a->ndim;
    for (int d = 0; d < a->ndim; d++) {
        g->order[d] = a->order[d];
        g->scale[d] = a->scale[d];
        g->translate[d] = a->translate[d];
        g->period[d] = a->period[d];
        g->size[d] = a->size[d];
        g->strides[d] = a->strides[d];
    }
}

static void fill_particles(const pmb_resample_args *a, PmbParticles *p)
{
    p->pos = a->pos; p->pos_elsize = a->pos_elsize; p->ps0 = a->pos_stride0; p->ps1 = a->pos_stride1;
    p->mass = a->mass; p->mass_elsize = a->mass_elsize; p->ms = a->mass_stride; p->mass_scalar = a->mass_scalar;
    p->hsml = a->hsml; p->hsml_elsize = a->hsml_elsize; p->hs = a->hsml_stride; p->hsml_scalar = a->hsml_scalar;
}

// the compile-time tuned kernels apply when no per-particle hsml is given and the (scalar-hsml
// scaled) support equals the native support of a tuned kind
static int fixed_family(const PmbWindow &w, const pmb_resample_args *a)
{
    if (!w.tuned || a->hsml) return 0;
    PmbWinInfo info;
    pmb_window_info(w.nativesupport, w.support * a->hsml_scalar, &info);
    return info.support == w.tuned ? w.tuned : 0;
}

#define PMB_DISPATCH_FAM(FAMV, NDIMV, CALL)                 \
    switch (FAMV) {                                         \
    case 1: { constexpr int FAM = 1; CALL; } break;         \
    case 2: { constexpr int FAM = 2; CALL; } break;         \
    case 3: { constexpr int FAM = 3; CALL; } break;         \
    default: { constexpr int FAM = 4; CALL; } break;        \
    }

#define PMB_DISPATCH_CHECK(CHECKV, CALL)                    \
    if (CHECKV) { constexpr bool CHECK = true; CALL; }      \
    else { constexpr bool CHECK = false; CALL; }

#define PMB_DISPATCH_NDIM(NDIMV, CALL)                      \
    switch (NDIMV) {                                        \
    case 1: { constexpr int NDIM = 1; CALL; } break;        \
    case 2: { constexpr int NDIM = 2; CALL; } break;        \
    default: { constexpr int NDIM = 3; CALL; } break;       \
    }

// 32-bit element-index geometry of a canvas (false: strides / extent do not allow it)
template <typename MeshT>
static bool geom32(const pmb_resample_args *a, const PmbGeom &g, PmbGeom32 *g32)
{
    int64_t span = 0;
    for (int d = 0; d < 3; d++) {
        if (a->strides[d] < 0 || a->strides[d] % (int64_t) sizeof(MeshT)) return false;
        span += (a->size[d] - 1) * (a->strides[d] / (int64_t) sizeof(MeshT));
        g32->scale[d] = g.scale[d]; g32->translate[d] = g.translate[d];
        g32->period[d] = (int) g.period[d]; g32->size[d] = (int) g.size[d];
        g32->estride[d] = (int) (a->strides[d] / (int64_t) sizeof(MeshT));
    }
    return span < ((int64_t) 1 << 31) - 1;
}

// How a large 3-D particle array is traversed.  bn->pos: a tile-sorted copy of the records serves it
// (pmb_bin.cuh); *walk: the kernels that walk the array through *perm do (pmb_perm.cuh; *perm NULL: the array
// IS the sorted copy and the plain kernels serve it); neither: the chunk-scheduled kernels.
// On the sorted copy (measured at 1024^3 uniform random, B200): the scatter takes the plain kernel with chunk
// tickets (PMB_BIN_PAINT=2: 46 ms with the algorithmic DRAM traffic; the carry kernels' static chunk stride puts
// every CTA into a tile of its own: 98 ms, 190 GB of DRAM traffic; 1: grid-stride loop, 0: carry kernels; 3: one CTA
// per tile, conflict-free deposit into a shared-memory tile, pmb_tile.cuh: 41 - 43 ms with 7 x fewer L2 requests -- the
// serial phases of a tile leave it latency-bound; keeping the particles' weights in shared memory as well costs the
// resident CTAs that hide that latency: 85 ms), the
// gather takes the tile kernel of pmb_tile.cuh (12.8 ms one field, 32 ms three; PMB_BIN_TILE_READOUT=0: bulk-copy ring
// 21 / 69 ms; PMB_BIN_READOUT=1: plain walk 24 ms).
static int pmb_traversal(pmb_ctx *ctx, const PmbGeom &g, const PmbParticles &p, int64_t npart, PmbBinned *bn,
                         const uint32_t **perm, bool *walk, bool is_paint)
{
    bn->pos = NULL; bn->dest = NULL; bn->verdict = 2;
    *perm = NULL;
    *walk = false;
    if (ctx->bin_bypass) {
        *walk = is_paint ? pmb_env_flag("PMB_BIN_PAINT", 2) != 0 : pmb_env_flag("PMB_BIN_READOUT", 0) != 0;
        return PMB_OK;
    }
    PMB_CHECK(pmb_bin_prepare(ctx, g, p, npart, bn));
    if (bn->verdict == 2) {
        PMB_CHECK(pmb_perm_prepare(ctx, g, p, npart, perm));
        *walk = *perm != NULL;
    }
    return PMB_OK;
}

// the arguments of a call on the sorted copy
static void pmb_binned_args(const pmb_resample_args *a, const PmbBinned &bn, pmb_resample_args *b)
{
    *b = *a;
    b->pos = bn.pos; b->pos_elsize = 8; b->pos_stride0 = 32; b->pos_stride1 = 8;      // (x, y, z, particle number) records
}

template <typename MeshT>
static int paint_atomic(pmb_ctx *ctx, const pmb_resample_args *a, const PmbGeom &g, const PmbWindow &w,
                        const PmbParticles &p)
{
    const int fam = fixed_family(w, a);
    char *mesh = (char *) a->mesh;
    if (fam && a->ndim == 3 && a->npart >= ((int64_t) 1 << 18) && pmb_env_flag("PMB_SCHED_KERNELS", 1)) {
        // large 3-D problems: locality-scheduled chunks + warp-aggregated atomics (pmb_sched.cuh)
        const uint32_t *order;
        int64_t nchunks;
        const bool chk = pmb_geom_needs_check(g);
        const int unit = pmb_env_flag("PMB_CARRY_UNIT", 128);
        // particle arrays without spatial order: a tile-sorted copy of the records (pmb_bin.cuh), else a walk
        // through a tile-binned permutation (pmb_perm.cuh)
        const uint32_t *perm;
        PmbBinned bn;
        bool walk;
        PMB_CHECK(pmb_traversal(ctx, g, p, a->npart, &bn, &perm, &walk, true));
        if (bn.pos) {
            pmb_resample_args b;
            pmb_binned_args(a, bn, &b);
            if (a->mass) {
                const double *smass;
                PMB_CHECK(pmb_bin_column(ctx, bn, a->mass, a->mass_elsize, a->mass_stride, a->npart, &smass));
                b.mass = smass; b.mass_elsize = 8; b.mass_stride = 8;
            }
            PmbParticles p2;
            fill_particles(&b, &p2);
            ctx->bin_bypass = 1;
            const int r = paint_atomic<MeshT>(ctx, &b, g, w, p2);
            ctx->bin_bypass = 0;
            return r;
        }
        if (ctx->bin_bypass && pmb_env_flag("PMB_BIN_PAINT", 2) >= 3 && fam == 2 && !(a->order[0] | a->order[1] | a->order[2])
            && pmb_pos_is_f8_rec4(p) && (!p.mass || (p.mass_elsize == 8 && p.ms == 8))) {
            // the tile-sorted copy: one CTA per tile, conflict-free deposit into a shared-memory tile (pmb_tile.cuh)
            PmbGeom32 g32;
            PmbTileGeom t;
            t.s[0] = ctx->bin_tiling[0]; t.s[1] = ctx->bin_tiling[1]; t.s[2] = ctx->bin_tiling[2];
            t.n1 = ctx->bin_tiling[3]; t.n2 = ctx->bin_tiling[4]; t.ntiles = ctx->bin_tiling[5];
            const size_t nhalo = (size_t) ((1 << t.s[0]) + 1) * ((1 << t.s[1]) + 1) * ((1 << t.s[2]) + 1);
            const size_t ncells = (size_t) 1 << (t.s[0] + t.s[1] + t.s[2]);
            const size_t smem = nhalo * sizeof(double) + (2 * ncells + 1) * sizeof(uint32_t) + 2 * PMB_TILE_BATCH * sizeof(uint16_t);
            if (geom32<MeshT>(a, g, &g32) && ncells <= 16384 && smem <= (size_t) 200 * 1024) {
                unsigned long long *ticket;
                PMB_CHECK(pmb_sched_ticket(ctx, &ticket));
                const uint32_t *counts = (const uint32_t *) ((const char *) ctx->bin_small + 256);
                const uint32_t *ends = counts + PMB_BIN_HARDTILES;
                int per_sm = (int) ((size_t) 220 * 1024 / (smem + 1024));
                if (per_sm > 6) per_sm = 6;
                if (per_sm < 1) per_sm = 1;
                const int gridt = (int) (t.ntiles < ctx->sm_count * per_sm ? t.ntiles : ctx->sm_count * per_sm);
                if (chk) {
                    PMB_CUDA(cudaFuncSetAttribute(pmb_k_paint_cic_tile<MeshT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
                    pmb_k_paint_cic_tile<MeshT, true><<<gridt, 256, smem, ctx->stream>>>(g32, t, counts, ends, (const double *) p.pos,
                        (const double *) p.mass, p.mass_scalar, (MeshT *) mesh, ticket);
                } else {
                    PMB_CUDA(cudaFuncSetAttribute(pmb_k_paint_cic_tile<MeshT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
                    pmb_k_paint_cic_tile<MeshT, false><<<gridt, 256, smem, ctx->stream>>>(g32, t, counts, ends, (const double *) p.pos,
                        (const double *) p.mass, p.mass_scalar, (MeshT *) mesh, ticket);
                }
                PMB_LAUNCH_CHECK(ctx);
                return PMB_OK;
            }
        }
        if (walk) {
            const int gridp = pmb_grid(ctx, a->npart, 256, 8);
            PmbGeom32 g32;
            unsigned long long *ticket = NULL;
            if (!perm && pmb_env_flag("PMB_BIN_PAINT", 2) >= 2) PMB_CHECK(pmb_sched_ticket(ctx, &ticket));
            if (fam == 2 && !(a->order[0] | a->order[1] | a->order[2]) && geom32<MeshT>(a, g, &g32)) {
                PMB_DISPATCH_CHECK(chk, (pmb_k_paint_cic32_perm<MeshT, CHECK><<<gridp, 256, 0, ctx->stream>>>(g32, p, (MeshT *) mesh, a->npart, perm, ticket)));
            } else {
                PMB_DISPATCH_CHECK(chk, PMB_DISPATCH_FAM(fam, 3,
                    (pmb_k_paint_perm<MeshT, FAM, CHECK><<<gridp, 256, 0, ctx->stream>>>(g, p, mesh, a->npart, a->pcs_gradient_scale_fix, perm, ticket))));
            }
            PMB_LAUNCH_CHECK(ctx);
            return PMB_OK;
        }
        if (fam == 2 && unit > 0 && !(a->order[0] | a->order[1] | a->order[2])) {
            // CIC: y-carry in registers + z aggregation in the warp (pmb_k_paint_cic_carry)
            PMB_CHECK(pmb_sched_prepare(ctx, g, p, a->npart, &order, &nchunks, 1));
            const int64_t nunits = (nchunks + unit - 1) / unit;
            const int64_t capu = (int64_t) ctx->sm_count * pmb_env_flag("PMB_GRID_MULT", 8);
            const int gridu = (int) (nunits < capu ? nunits : capu);
            // 32-bit element indices when the canvas spans < 2^31 elements and strides are whole elements
            bool idx32 = pmb_env_flag("PMB_CARRY32", 1) != 0;
            int64_t span = 0;
            for (int d = 0; d < 3; d++) {
                if (a->strides[d] < 0 || a->strides[d] % (int64_t) sizeof(MeshT)) idx32 = false;
                span += (a->size[d] - 1) * (a->strides[d] / (int64_t) sizeof(MeshT));
            }
            if (span >= ((int64_t) 1 << 31) - 1) idx32 = false;
            if (idx32) {
                PmbGeom32 g32;
                for (int d = 0; d < 3; d++) {
                    g32.scale[d] = g.scale[d]; g32.translate[d] = g.translate[d];
                    g32.period[d] = (int) g.period[d]; g32.size[d] = (int) g.size[d];
                    g32.estride[d] = (int) (a->strides[d] / (int64_t) sizeof(MeshT));
                }
                const int pminb = pmb_env_flag("PMB_PAINT_MINB", 4);
                // measured at 1024^3 (B200, ms; profiles/r2_kernels_ring.jsonl): the ring does not help the
                // scatter (11.7 vs 11.3 lattice order, 16.2 vs 15.7 Gaussian Zel'dovich): it is bound by the
                // red.global issue rate and the L2's read-for-update traffic, not by the position latency.
                // PMB_RING_PAINT=1 selects it for A/B runs.
                if (pmb_env_flag("PMB_RING_PAINT", 0) && pmb_pos_is_f8_rows(p) && ((uintptr_t) p.pos & 15) == 0) {
                    // particle stream through the bulk-copy ring (pmb_ring.cuh)
                    const int rminb = pmb_env_flag("PMB_RING_PAINT_MINB", 4);
                    const int64_t capr = (int64_t) ctx->sm_count * rminb;
                    const int gridr = (int) (nunits < capr ? nunits : capr);
#define PMB_RING_PAINT(MB) PMB_DISPATCH_CHECK(chk, (pmb_k_paint_cic_carry32_ring<MeshT, CHECK, MB><<<gridr, PMB_RING_THREADS, 0, ctx->stream>>>( \
                        g32, (const double *) p.pos, p, (MeshT *) mesh, a->npart, order, nchunks, unit)))
                    if (rminb <= 3) { PMB_RING_PAINT(3); } else if (rminb == 4) { PMB_RING_PAINT(4); } else { PMB_RING_PAINT(5); }
#undef PMB_RING_PAINT
                    PMB_LAUNCH_CHECK(ctx);
                    return PMB_OK;
                }
                if (pminb >= 5 && pmb_pos_is_f8_rows(p)) {
                    if (pminb == 5) {
                        PMB_DISPATCH_CHECK(chk, (pmb_k_paint_cic_carry32<MeshT, CHECK, true, 5><<<gridu, PMB_CHUNK, 0, ctx->stream>>>(
                            g32, p, (MeshT *) mesh, a->npart, order, nchunks, unit)));
                    } else {
                        PMB_DISPATCH_CHECK(chk, (pmb_k_paint_cic_carry32<MeshT, CHECK, true, 6><<<gridu, PMB_CHUNK, 0, ctx->stream>>>(
                            g32, p, (MeshT *) mesh, a->npart, order, nchunks, unit)));
                    }
                } else if (pmb_pos_is_f8_rows(p)) {
                    PMB_DISPATCH_CHECK(chk, (pmb_k_paint_cic_carry32<MeshT, CHECK, true><<<gridu, PMB_CHUNK, 0, ctx->stream>>>(
                        g32, p, (MeshT *) mesh, a->npart, order, nchunks, unit)));
                } else {
                    PMB_DISPATCH_CHECK(chk, (pmb_k_paint_cic_carry32<MeshT, CHECK, false><<<gridu, PMB_CHUNK, 0, ctx->stream>>>(
                        g32, p, (MeshT *) mesh, a->npart, order, nchunks, unit)));
                }
                PMB_LAUNCH_CHECK(ctx);
                return PMB_OK;
            }
            if (pmb_env_flag("PMB_PAINT_PREFETCH", 1)) {
                PMB_DISPATCH_CHECK(chk, (pmb_k_paint_cic_carry<MeshT, CHECK, true><<<gridu, PMB_CHUNK, 0, ctx->stream>>>(
                    g, p, mesh, a->npart, order, nchunks, unit)));
            } else {
                PMB_DISPATCH_CHECK(chk, (pmb_k_paint_cic_carry<MeshT, CHECK, false><<<gridu, PMB_CHUNK, 0, ctx->stream>>>(
                    g, p, mesh, a->npart, order, nchunks, unit)));
            }
            PMB_LAUNCH_CHECK(ctx);
            return PMB_OK;
        }
        if (fam >= 2 && unit > 0 && pmb_env_flag("PMB_CARRY_FAM", 1)) {
            // TSC / PCS and every gradient window: the same y-carry with FAM - 1 carried rows
            // (pmb_k_paint_carry32), when the canvas can be addressed with 32-bit element indices
            bool idx32 = true;
            int64_t span = 0;
            for (int d = 0; d < 3; d++) {
                if (a->strides[d] < 0 || a->strides[d] % (int64_t) sizeof(MeshT)) idx32 = false;
                span += (a->size[d] - 1) * (a->strides[d] / (int64_t) sizeof(MeshT));
            }
            if (span >= ((int64_t) 1 << 31) - 1) idx32 = false;
            if (idx32) {
                PMB_CHECK(pmb_sched_prepare(ctx, g, p, a->npart, &order, &nchunks, 1));
                const int64_t nunits = (nchunks + unit - 1) / unit;
                const int64_t capu = (int64_t) ctx->sm_count * pmb_env_flag("PMB_GRID_MULT", 8);
                const int gridu = (int) (nunits < capu ? nunits : capu);
                PmbGeom32o go;
                for (int d = 0; d < 3; d++) {
                    go.g.scale[d] = g.scale[d]; go.g.translate[d] = g.translate[d];
                    go.g.period[d] = (int) g.period[d]; go.g.size[d] = (int) g.size[d];
                    go.g.estride[d] = (int) (a->strides[d] / (int64_t) sizeof(MeshT));
                    go.order[d] = a->order[d];
                }
                go.pcsfix = a->pcs_gradient_scale_fix;
                const bool pos8 = pmb_pos_is_f8_rows(p);
#define PMB_CARRY_PAINT(FAMV)                                                                                  \
                do {                                                                                           \
                    if (pos8) { PMB_DISPATCH_CHECK(chk, (pmb_k_paint_carry32<MeshT, FAMV, CHECK, true><<<gridu, PMB_CHUNK, 0, ctx->stream>>>( \
                                    go, p, (MeshT *) mesh, a->npart, order, nchunks, unit))); }                \
                    else { PMB_DISPATCH_CHECK(chk, (pmb_k_paint_carry32<MeshT, FAMV, CHECK, false><<<gridu, PMB_CHUNK, 0, ctx->stream>>>( \
                                    go, p, (MeshT *) mesh, a->npart, order, nchunks, unit))); }                \
                } while (0)
                if (fam == 2) PMB_CARRY_PAINT(2); else if (fam == 3) PMB_CARRY_PAINT(3); else PMB_CARRY_PAINT(4);
#undef PMB_CARRY_PAINT
                PMB_LAUNCH_CHECK(ctx);
                return PMB_OK;
            }
        }
        PMB_CHECK(pmb_sched_prepare(ctx, g, p, a->npart, &order, &nchunks));
        unsigned long long *ticket;
        PMB_CHECK(pmb_sched_ticket(ctx, &ticket));
        const int64_t cap = (int64_t) ctx->sm_count * pmb_env_flag("PMB_GRID_MULT", 8);
        const int grid = (int) (nchunks < cap ? nchunks : cap);
        const bool merge = fam > 1 && pmb_env_flag("PMB_MERGE", 1);
        // bits 0-1: mesh L2 policy (1 evict_last), bit 2: streaming pos loads, bit 3: dynamic tickets
        const int dbg = pmb_env_flag("PMB_DBG", 1 | 4);
#define PMB_SCHED_PAINT(MERGEV) PMB_DISPATCH_CHECK(chk, PMB_DISPATCH_FAM(fam, 3, \
            (pmb_k_paint_sched<MeshT, FAM, CHECK, MERGEV><<<grid, PMB_CHUNK, 0, ctx->stream>>>(g, p, mesh, a->npart, a->pcs_gradient_scale_fix, order, nchunks, dbg, ticket))))
        if (merge) { PMB_SCHED_PAINT(true); } else { PMB_SCHED_PAINT(false); }
#undef PMB_SCHED_PAINT
        PMB_LAUNCH_CHECK(ctx);
        return PMB_OK;
    }
    if (fam) {
        int grid = pmb_grid(ctx, a->npart, 256, 8);
        const bool chk = pmb_geom_needs_check(g);
        PMB_DISPATCH_CHECK(chk, PMB_DISPATCH_NDIM(a->ndim, PMB_DISPATCH_FAM(fam, NDIM,
            (pmb_k_paint_tuned<MeshT, NDIM, FAM, CHECK><<<grid, 256, 0, ctx->stream>>>(g, p, mesh, a->npart, a->pcs_gradient_scale_fix, pmb_cache_mode() & 1)))));
    } else {
        int grid = pmb_grid(ctx, a->npart, 128, 8);
        PMB_DISPATCH_NDIM(a->ndim,
            (pmb_k_paint_dyn<MeshT, NDIM><<<grid, 128, 0, ctx->stream>>>(g, w, p, mesh, a->npart, a->pcs_gradient_scale_fix)));
    }
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}

template <typename KeyT, typename MeshT>
static int paint_deterministic(pmb_ctx *ctx, const pmb_resample_args *a, const PmbGeom &g_in, const PmbWindow &w,
                               const PmbParticles &p)
{
    // dense C-order element strides so the stencil walk yields linear cell numbers
    PmbGeom g = g_in;
    int64_t acc = 1;
    for (int d = a->ndim - 1; d >= 0; d--) { g.strides[d] = acc; acc *= g.size[d]; }
    int64_t ncell = acc;
    if (ncell == 0) return PMB_OK;
    const int fam = fixed_family(w, a);
    {
        // tuned windows on 3-D canvases: sort the PARTICLES by cell, every cell sums its own contributions (pmb_pull.cuh)
        bool done = false;
        PMB_CHECK((pmb_pull_paint<MeshT>(ctx, a, g_in, p, fam, &done)));
        if (done) return PMB_OK;
    }

    // widest per-particle stencil
    int64_t smax;
    if (fam) smax = fam;
    else {
        double hmax = a->hsml_scalar;
        if (a->hsml) {
            int grid = pmb_grid(ctx, a->npart, 256, 4);
            void *part;
            PMB_CHECK(pmb_scratch(ctx, sizeof(double) * grid, &part));
            pmb_k_max_real<double><<<grid, 256, 0, ctx->stream>>>(a->hsml, a->hsml_elsize, a->hsml_stride, a->npart, (double *) part);
            PMB_LAUNCH_CHECK(ctx);
            double *h = (double *) malloc(sizeof(double) * grid);
            if (!h) return PMB_ENOMEM;
            cudaError_t e = cudaMemcpyAsync(h, part, sizeof(double) * grid, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { free(h); return pmb_cuda_fail(e, "hsml max", __FILE__, __LINE__); }
            hmax = h[0];
            for (int i = 1; i < grid; i++) hmax = hmax > h[i] ? hmax : h[i];
            free(h);
        }
        PmbWinInfo info;
        pmb_window_info(w.nativesupport, w.support * hmax, &info);
        smax = info.support;
    }
    int64_t npts = 1;
    for (int d = 0; d < a->ndim; d++) npts *= smax;

    // bits of the sort key actually used: the width of ncell.  Every real key is < ncell < 2^key_bits and the
    // sentinel (all ones) truncated to key_bits bits is 2^key_bits - 1 >= ncell, so it still sorts last.
    int key_bits = 1;
    while (key_bits < (int) sizeof(KeyT) * 8 && ((int64_t) 1 << key_bits) <= ncell) key_bits++;

    // chunk size from the workspace budget: 2x keys + 2x values + cub temp
    const size_t per_pair = 2 * sizeof(KeyT) + 2 * sizeof(double);
    int64_t chunk = (int64_t) (ctx->det_chunk_bytes / (per_pair * (size_t) npts));
    if (chunk < 1) chunk = 1;
    if (chunk > a->npart) chunk = a->npart;
    const int64_t max_pairs = (int64_t) 1 << 30;   // keep cub item counts well inside int32
    if (chunk * npts > max_pairs) chunk = max_pairs / npts > 0 ? max_pairs / npts : 1;
    const int64_t pairs = chunk * npts;
    PMB_REQUIRE(pairs < ((int64_t) 1 << 31), "stencil of %lld points is too wide for the deterministic path", (long long) npts);

    size_t temp_bytes = 0;
    {
        cub::DoubleBuffer<KeyT> kb((KeyT *) NULL, (KeyT *) NULL);
        cub::DoubleBuffer<double> vb((double *) NULL, (double *) NULL);
        PMB_CUDA(cub::DeviceRadixSort::SortPairs(NULL, temp_bytes, kb, vb, (int) pairs, 0, key_bits, ctx->stream));
    }
    size_t off_k0 = 0;
    size_t off_k1 = off_k0 + ((sizeof(KeyT) * pairs + 255) & ~(size_t) 255);
    size_t off_v0 = off_k1 + ((sizeof(KeyT) * pairs + 255) & ~(size_t) 255);
    size_t off_v1 = off_v0 + ((sizeof(double) * pairs + 255) & ~(size_t) 255);
    size_t off_t = off_v1 + ((sizeof(double) * pairs + 255) & ~(size_t) 255);
    void *ws;
    PMB_CHECK(pmb_scratch(ctx, off_t + temp_bytes + 256, &ws));
    char *wsb = (char *) ws;
    KeyT *k0 = (KeyT *) (wsb + off_k0), *k1 = (KeyT *) (wsb + off_k1);
    double *v0 = (double *) (wsb + off_v0), *v1 = (double *) (wsb + off_v1);

    for (int64_t first = 0; first < a->npart; first += chunk) {
        const int64_t count = (a->npart - first) < chunk ? (a->npart - first) : chunk;
        const int64_t n = count * npts;
        if (fam) {
            int grid = pmb_grid(ctx, count, 256, 8);
            PMB_DISPATCH_NDIM(a->ndim, PMB_DISPATCH_FAM(fam, NDIM,
                (pmb_k_expand_tuned<KeyT, NDIM, FAM><<<grid, 256, 0, ctx->stream>>>(g, p, first, count, a->pcs_gradient_scale_fix, k0, v0))));
        } else {
            PMB_CUDA(cudaMemsetAsync(k0, 0xFF, sizeof(KeyT) * n, ctx->stream));
            PMB_CUDA(cudaMemsetAsync(v0, 0, sizeof(double) * n, ctx->stream));
            int grid = pmb_grid(ctx, count, 128, 8);
            PMB_DISPATCH_NDIM(a->ndim,
                (pmb_k_expand_dyn<KeyT, NDIM><<<grid, 128, 0, ctx->stream>>>(g, w, p, first, count, a->pcs_gradient_scale_fix, npts, k0, v0)));
        }
        PMB_LAUNCH_CHECK(ctx);
        cub::DoubleBuffer<KeyT> kb(k0, k1);
        cub::DoubleBuffer<double> vb(v0, v1);
        size_t tb = temp_bytes;
        PMB_CUDA(cub::DeviceRadixSort::SortPairs(wsb + off_t, tb, kb, vb, (int) n, 0, key_bits, ctx->stream));
        ctx->launches += 4;
        int grid = pmb_grid(ctx, n, 256, 8);
        pmb_k_segment_sum<KeyT, MeshT><<<grid, 256, 0, ctx->stream>>>(
            kb.Current(), vb.Current(), n, (char *) a->mesh,
            a->ndim > 1 ? g_in.size[a->ndim - 2] : 1, g_in.size[a->ndim - 1],
            a->ndim > 2 ? g_in.strides[a->ndim - 3] : 0,
            a->ndim > 1 ? g_in.strides[a->ndim - 2] : 0,
            g_in.strides[a->ndim - 1]);
        PMB_LAUNCH_CHECK(ctx);
    }
    return PMB_OK;
}

extern "C" int pmb_paint(pmb_ctx *ctx, const pmb_resample_args *a)
{
    PMB_CHECK(check_args(ctx, a, 0));
    PMB_REQUIRE(a->mode == PMB_MODE_ATOMIC || a->mode == PMB_MODE_DETERMINISTIC, "bad paint mode %d", a->mode);
    if (a->npart == 0) return PMB_OK;
    PmbGeom g;
    fill_geom(a, &g);
    PmbParticles p;
    fill_particles(a, &p);
    PmbWindow w;
    PMB_CHECK(pmb_resolve_window(ctx, a->kind, a->support, a->ndim, a->order, &w, 1));
    for (int d = 0; d < a->ndim; d++) if (a->size[d] == 0) return PMB_OK;
    if (a->mode == PMB_MODE_ATOMIC) {
        if (a->mesh_elsize == 8) return paint_atomic<double>(ctx, a, g, w, p);
        return paint_atomic<float>(ctx, a, g, w, p);
    }
    int64_t ncell = 1;
    for (int d = 0; d < a->ndim; d++) ncell *= a->size[d];
    const bool small = ncell < (int64_t) 0xFFFFFFFFll;
    if (a->mesh_elsize == 8)
        return small ? paint_deterministic<uint32_t, double>(ctx, a, g, w, p)
                     : paint_deterministic<uint64_t, double>(ctx, a, g, w, p);
    return small ? paint_deterministic<uint32_t, float>(ctx, a, g, w, p)
                 : paint_deterministic<uint64_t, float>(ctx, a, g, w, p);
}

template <typename MeshT, int NF>
static int launch_readout_tile(pmb_ctx *ctx, bool chk, int grid, size_t smem, const PmbGeom32 &g32, const PmbTileGeom &t,
                               const uint32_t *counts, const uint32_t *ends, const double *recs, const PmbFields &f,
                               unsigned long long *ticket)
{
    if (chk) {
        PMB_CUDA(cudaFuncSetAttribute(pmb_k_readout_cic_tile<MeshT, true, NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        pmb_k_readout_cic_tile<MeshT, true, NF><<<grid, 256, smem, ctx->stream>>>(g32, t, counts, ends, recs, f, ticket);
    } else {
        PMB_CUDA(cudaFuncSetAttribute(pmb_k_readout_cic_tile<MeshT, false, NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        pmb_k_readout_cic_tile<MeshT, false, NF><<<grid, 256, smem, ctx->stream>>>(g32, t, counts, ends, recs, f, ticket);
    }
    return PMB_OK;
}

// CIC gather of nf canvases in one sweep, particle stream through the bulk-copy ring (pmb_ring.cuh)
template <typename MeshT>
static int readout_ring(pmb_ctx *ctx, const PmbGeom32 &g32, const PmbParticles &p, const PmbFields &f, int nf,
                        int64_t npart, bool chk)
{
    const int64_t nchunks = (npart + PMB_CHUNK - 1) / PMB_CHUNK;
    unsigned long long *ticket;
    PMB_CHECK(pmb_sched_ticket(ctx, &ticket));
    // resident CTAs per SM, measured at 1024^3 (ms): one field 4 -> 8.99, 5 -> 9.34, 6 -> 9.57 (11.13 without
    // the ring); three fields 2 -> 21.0, 3 -> 20.0, 4 -> 28.6 (spills) (33.4 as three separate gathers)
    const int minb = pmb_pos_is_f8_rec4(p) ? (nf == 1 ? 4 : 3)
                                           : pmb_env_flag(nf == 1 ? "PMB_RING_READOUT_MINB" : "PMB_RING_READOUT3_MINB", nf == 1 ? 4 : 3);
    const int64_t cap = (int64_t) ctx->sm_count * minb;
    const int grid = (int) (nchunks < cap ? nchunks : cap);
    const double *pos = (const double *) p.pos;
    if (pmb_pos_is_f8_rec4(p) && ctx->bin_bypass && f.sel1 <= f.sel0 && pmb_env_flag("PMB_BIN_TILE_READOUT", 1)) {
        // the tile-sorted copy: one CTA per tile, the tile's mesh cells (+ halo) of every field in shared memory (pmb_tile.cuh)
        PmbTileGeom t;
        t.s[0] = ctx->bin_tiling[0]; t.s[1] = ctx->bin_tiling[1]; t.s[2] = ctx->bin_tiling[2];
        t.n1 = ctx->bin_tiling[3]; t.n2 = ctx->bin_tiling[4]; t.ntiles = ctx->bin_tiling[5];
        const size_t cells = (size_t) ((1 << t.s[0]) + 1) * ((1 << t.s[1]) + 1) * ((1 << t.s[2]) + 1);
        const size_t smem = cells * sizeof(double) * nf;
        if (smem <= (size_t) 200 * 1024) {
            const uint32_t *counts = (const uint32_t *) ((const char *) ctx->bin_small + 256);
            const uint32_t *ends = counts + PMB_BIN_HARDTILES;
            // resident CTAs per SM: what shared memory (227 KB) and 2048 threads allow; tiles in flight hide the latency
            // of a CTA's phases (tile load -> barrier -> particles)
            int per_sm = (int) ((size_t) 220 * 1024 / (smem + 1024));
            if (per_sm > 8) per_sm = 8;
            if (per_sm < 1) per_sm = 1;
            const int gridt = (int) (t.ntiles < ctx->sm_count * per_sm ? t.ntiles : ctx->sm_count * per_sm);
            if (nf == 1) PMB_CHECK((launch_readout_tile<MeshT, 1>(ctx, chk, gridt, smem, g32, t, counts, ends, pos, f, ticket)));
            else if (nf == 2) PMB_CHECK((launch_readout_tile<MeshT, 2>(ctx, chk, gridt, smem, g32, t, counts, ends, pos, f, ticket)));
            else PMB_CHECK((launch_readout_tile<MeshT, 3>(ctx, chk, gridt, smem, g32, t, counts, ends, pos, f, ticket)));
            PMB_LAUNCH_CHECK(ctx);
            return PMB_OK;
        }
    }
    if (pmb_pos_is_f8_rec4(p)) {
        // 32-byte records of the tile-sorted copy
#define PMB_RING_READ4(NFV, MB) PMB_DISPATCH_CHECK(chk, (pmb_k_readout_cic32_ring<MeshT, CHECK, NFV, MB, 4><<<grid, PMB_RING_THREADS, 0, ctx->stream>>>( \
            g32, pos, f, npart, nchunks, ticket)))
        if (nf == 1) { PMB_RING_READ4(1, 4); } else if (nf == 2) { PMB_RING_READ4(2, 3); } else { PMB_RING_READ4(3, 3); }
#undef PMB_RING_READ4
        PMB_LAUNCH_CHECK(ctx);
        return PMB_OK;
    }
    // PMB_RING_SCHED=1: chunks in the order of the spatial schedule (the cell of a chunk's middle particle) instead of
    // memory order -- for particles displaced by several cells the mesh rows in flight then stay in L2
    const uint32_t *order = NULL;
    if (pmb_env_flag("PMB_RING_SCHED", nf >= 2 ? 1 : 0) && !ctx->bin_bypass) {
        int64_t nch2;
        PmbGeom gg;
        memset(&gg, 0, sizeof(gg));
        gg.ndim = 3;
        for (int d = 0; d < 3; d++) { gg.scale[d] = g32.scale[d]; gg.translate[d] = g32.translate[d]; gg.period[d] = g32.period[d]; gg.size[d] = g32.size[d]; }
        PMB_CHECK(pmb_sched_prepare(ctx, gg, p, npart, &order, &nch2, 1));      // the scatter's schedule: built once per array
        PMB_CHECK(pmb_sched_ticket(ctx, &ticket));
    }
#define PMB_RING_READ(NFV, MB) PMB_DISPATCH_CHECK(chk, (pmb_k_readout_cic32_ring<MeshT, CHECK, NFV, MB><<<grid, PMB_RING_THREADS, 0, ctx->stream>>>( \
        g32, pos, f, npart, nchunks, ticket, order)))
    if (nf == 1) {
        if (minb <= 4) { PMB_RING_READ(1, 4); } else if (minb == 5) { PMB_RING_READ(1, 5); } else { PMB_RING_READ(1, 6); }
    } else if (nf == 2) {
        if (minb <= 3) { PMB_RING_READ(2, 3); } else { PMB_RING_READ(2, 4); }
    } else {
        if (minb <= 2) { PMB_RING_READ(3, 2); } else if (minb == 3) { PMB_RING_READ(3, 3); } else { PMB_RING_READ(3, 4); }
    }
#undef PMB_RING_READ
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}

template <typename MeshT>
static int readout_impl(pmb_ctx *ctx, const pmb_resample_args *a, const PmbGeom &g, const PmbWindow &w,
                        const PmbParticles &p)
{
    const int fam = fixed_family(w, a);
    const char *mesh = (const char *) a->mesh;
    // measured on B200 (1024^3 CIC, ms): grid-stride gather 14.4; chunked with dynamic tickets in
    // memory order 12.4 (DRAM traffic drops to the algorithmic 43 GB); tickets + spatial schedule
    // 13.6 (the schedule's indirection costs more than it saves for reads).  PMB_SCHED_READOUT=2
    // adds the spatial schedule, 0 falls back to the grid-stride kernel.
    const int sched_readout = pmb_env_flag("PMB_SCHED_READOUT", 1);
    if (fam && a->ndim == 3 && a->npart >= ((int64_t) 1 << 18) && sched_readout) {
        const uint32_t *order = NULL;
        int64_t nchunks = (a->npart + PMB_CHUNK - 1) / PMB_CHUNK;
        {
            // particle arrays without spatial order: tile-sorted copy (pmb_bin.cuh), else permutation walk (pmb_perm.cuh)
            const uint32_t *perm;
            PmbBinned bn;
            bool walk;
            PMB_CHECK(pmb_traversal(ctx, g, p, a->npart, &bn, &perm, &walk, false));
            double *stage = bn.pos ? (double *) pmb_bin_try_scratch(ctx, sizeof(double) * (size_t) a->npart) : NULL;
            if (bn.pos && !stage) {       // no room for the staging column: the permutation walk serves the array
                PMB_CHECK(pmb_perm_prepare(ctx, g, p, a->npart, &perm));
                walk = perm != NULL;
            }
            if (stage) {
                // gather on the sorted copy into the staging column, then back to the caller's order
                pmb_resample_args b;
                pmb_binned_args(a, bn, &b);
                b.out = stage; b.out_elsize = 8; b.out_stride = 8;
                PmbParticles p2;
                fill_particles(&b, &p2);
                ctx->bin_bypass = 1;
                const int r = readout_impl<MeshT>(ctx, &b, g, w, p2);
                ctx->bin_bypass = 0;
                PMB_CHECK(r);
                PmbFields f;
                memset(&f, 0, sizeof(f));
                f.out[0] = a->out; f.out_stride[0] = a->out_stride; f.out_elsize = a->out_elsize;
                return pmb_bin_unsort(ctx, bn, f, 1, stage, a->npart);
            }
            if (walk) {
                const int gridp = pmb_grid(ctx, a->npart, 256, 8);
                const bool chkp = pmb_geom_needs_check(g);
                PmbGeom32 g32;
                if (fam == 2 && !(a->order[0] | a->order[1] | a->order[2]) && geom32<MeshT>(a, g, &g32)) {
                    PmbFields f;
                    memset(&f, 0, sizeof(f));
                    f.mesh[0] = mesh; f.out[0] = a->out; f.out_stride[0] = a->out_stride; f.out_elsize = a->out_elsize;
                    PMB_DISPATCH_CHECK(chkp, (pmb_k_readout_cic32_perm<MeshT, CHECK, 1><<<gridp, 256, 0, ctx->stream>>>(g32, p, f, a->npart, perm)));
                } else {
                    PMB_DISPATCH_CHECK(chkp, PMB_DISPATCH_FAM(fam, 3,
                        (pmb_k_readout_perm<MeshT, FAM, CHECK><<<gridp, 256, 0, ctx->stream>>>(
                            g, p, mesh, a->npart, a->pcs_gradient_scale_fix, a->out, a->out_elsize, a->out_stride, perm))));
                }
                PMB_LAUNCH_CHECK(ctx);
                return PMB_OK;
            }
        }
        if (sched_readout >= 2) PMB_CHECK(pmb_sched_prepare(ctx, g, p, a->npart, &order, &nchunks));
        unsigned long long *ticket;
        PMB_CHECK(pmb_sched_ticket(ctx, &ticket));
        const int64_t cap = (int64_t) ctx->sm_count * pmb_env_flag("PMB_GRID_MULT", 8);
        const int grid = (int) (nchunks < cap ? nchunks : cap);
        const bool chk = pmb_geom_needs_check(g);
        if (fam == 2 && !(a->order[0] | a->order[1] | a->order[2]) && pmb_env_flag("PMB_READOUT32", 1)) {
            bool idx32 = true;
            int64_t span = 0;
            for (int d = 0; d < 3; d++) {
                if (a->strides[d] < 0 || a->strides[d] % (int64_t) sizeof(MeshT)) idx32 = false;
                span += (a->size[d] - 1) * (a->strides[d] / (int64_t) sizeof(MeshT));
            }
            if (span >= ((int64_t) 1 << 31) - 1) idx32 = false;
            if (idx32) {
                PmbGeom32 g32;
                for (int d = 0; d < 3; d++) {
                    g32.scale[d] = g.scale[d]; g32.translate[d] = g.translate[d];
                    g32.period[d] = (int) g.period[d]; g32.size[d] = (int) g.size[d];
                    g32.estride[d] = (int) (a->strides[d] / (int64_t) sizeof(MeshT));
                }
                if (pmb_env_flag("PMB_RING", 1) && ((pmb_pos_is_f8_rows(p) && ((uintptr_t) p.pos & 15) == 0) || pmb_pos_is_f8_rec4(p))) {
                    PmbFields f;
                    memset(&f, 0, sizeof(f));
                    f.mesh[0] = mesh; f.out[0] = a->out; f.out_stride[0] = a->out_stride; f.out_elsize = a->out_elsize;
                    return readout_ring<MeshT>(ctx, g32, p, f, 1, a->npart, chk);
                }
                // resident CTAs per SM: measured at 1024^3 (ms): 5 -> 11.44, 6 -> 10.73 (the gather is
                // latency-bound: more warps in flight beat more loads per warp, cf. the _pipe variant)
                const int minb = pmb_env_flag("PMB_READOUT_MINB", 6);
                if (minb >= 6 && pmb_pos_is_f8_rows(p)) {
                    const int64_t capb = (int64_t) ctx->sm_count * minb;
                    const int gridb = (int) (nchunks < capb ? nchunks : capb);
                    if (minb == 6) {
                        PMB_DISPATCH_CHECK(chk, (pmb_k_readout_cic32<MeshT, CHECK, true, 6><<<gridb, PMB_CHUNK, 0, ctx->stream>>>(
                            g32, p, (const MeshT *) mesh, a->npart, a->out, a->out_elsize, a->out_stride, order, nchunks, ticket)));
                    } else if (minb == 7) {
                        PMB_DISPATCH_CHECK(chk, (pmb_k_readout_cic32<MeshT, CHECK, true, 7><<<gridb, PMB_CHUNK, 0, ctx->stream>>>(
                            g32, p, (const MeshT *) mesh, a->npart, a->out, a->out_elsize, a->out_stride, order, nchunks, ticket)));
                    } else {
                        PMB_DISPATCH_CHECK(chk, (pmb_k_readout_cic32<MeshT, CHECK, true, 8><<<gridb, PMB_CHUNK, 0, ctx->stream>>>(
                            g32, p, (const MeshT *) mesh, a->npart, a->out, a->out_elsize, a->out_stride, order, nchunks, ticket)));
                    }
                    PMB_LAUNCH_CHECK(ctx);
                    return PMB_OK;
                }
                if (pmb_pos_is_f8_rows(p)) {
                    PMB_DISPATCH_CHECK(chk, (pmb_k_readout_cic32<MeshT, CHECK, true><<<grid, PMB_CHUNK, 0, ctx->stream>>>(
                        g32, p, (const MeshT *) mesh, a->npart, a->out, a->out_elsize, a->out_stride, order, nchunks, ticket)));
                } else {
                    PMB_DISPATCH_CHECK(chk, (pmb_k_readout_cic32<MeshT, CHECK, false><<<grid, PMB_CHUNK, 0, ctx->stream>>>(
                        g32, p, (const MeshT *) mesh, a->npart, a->out, a->out_elsize, a->out_stride, order, nchunks, ticket)));
                }
                PMB_LAUNCH_CHECK(ctx);
                return PMB_OK;
            }
        }
        const int variant = pmb_env_flag("PMB_READOUT_VARIANT", 2);
#define PMB_SCHED_READOUT(V) PMB_DISPATCH_CHECK(chk, PMB_DISPATCH_FAM(fam, 3, \
            (pmb_k_readout_sched<MeshT, FAM, CHECK, V><<<grid, PMB_CHUNK, 0, ctx->stream>>>( \
                g, p, mesh, a->npart, a->pcs_gradient_scale_fix, a->out, a->out_elsize, a->out_stride, order, nchunks, ticket))))
        if (variant == 0) { PMB_SCHED_READOUT(0); } else if (variant == 1) { PMB_SCHED_READOUT(1); } else { PMB_SCHED_READOUT(2); }
#undef PMB_SCHED_READOUT
        PMB_LAUNCH_CHECK(ctx);
        return PMB_OK;
    }
    if (fam) {
        int grid = pmb_grid(ctx, a->npart, 256, 8);
        const bool chk = pmb_geom_needs_check(g);
        PMB_DISPATCH_CHECK(chk, PMB_DISPATCH_NDIM(a->ndim, PMB_DISPATCH_FAM(fam, NDIM,
            (pmb_k_readout_tuned<MeshT, NDIM, FAM, CHECK><<<grid, 256, 0, ctx->stream>>>(
                g, p, mesh, a->npart, a->pcs_gradient_scale_fix, a->out, a->out_elsize, a->out_stride, pmb_cache_mode() & 1)))));
    } else {
        int grid = pmb_grid(ctx, a->npart, 128, 8);
        PMB_DISPATCH_NDIM(a->ndim,
            (pmb_k_readout_dyn<MeshT, NDIM><<<grid, 128, 0, ctx->stream>>>(
                g, w, p, mesh, a->npart, a->pcs_gradient_scale_fix, a->out, a->out_elsize, a->out_stride)));
    }
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}

extern "C" int pmb_readout(pmb_ctx *ctx, const pmb_resample_args *a)
{
    PMB_CHECK(check_args(ctx, a, 1));
    if (a->npart == 0) return PMB_OK;
    PmbGeom g;
    fill_geom(a, &g);
    PmbParticles p;
    fill_particles(a, &p);
    PmbWindow w;
    PMB_CHECK(pmb_resolve_window(ctx, a->kind, a->support, a->ndim, a->order, &w, 1));
    if (a->mesh_elsize == 8) return readout_impl<double>(ctx, a, g, w, p);
    return readout_impl<float>(ctx, a, g, w, p);
}

// readout of up to 3 canvases of identical geometry at the same positions
struct PmbGatherFuse {      // see PmbFields: results of the rank's own particles go straight to the gathered columns
    void *const *own_outs;
    const int32_t *own_index;
    int64_t own_begin, own_count;
};

template <typename MeshT>
static int readout_multi_impl(pmb_ctx *ctx, const pmb_resample_args *a, int nf, const void *const *meshes,
                              void *const *outs, const int64_t *out_strides, bool *done, const PmbGatherFuse *gf = NULL)
{
    *done = false;
    PmbGeom g;
    fill_geom(a, &g);
    PmbParticles p;
    fill_particles(a, &p);
    PmbWindow w;
    PMB_CHECK(pmb_resolve_window(ctx, a->kind, a->support, a->ndim, a->order, &w, 1));
    const int fam = fixed_family(w, a);
    if (!(fam == 2 && a->ndim == 3 && !(a->order[0] | a->order[1] | a->order[2]) && a->npart >= ((int64_t) 1 << 18))) return PMB_OK;
    PmbGeom32 g32;
    if (!geom32<MeshT>(a, g, &g32)) return PMB_OK;
    PmbFields f;
    memset(&f, 0, sizeof(f));
    for (int q = 0; q < nf; q++) { f.mesh[q] = meshes[q]; f.out[q] = outs[q]; f.out_stride[q] = out_strides[q]; }
    f.out_elsize = a->out_elsize;
    f.packed_rows = nf == 3 && ctx->bin_bypass && !gf && a->out_elsize == 8 && out_strides[0] == 32 && ((uintptr_t) outs[0] & 31) == 0
                    && (char *) outs[1] == (char *) outs[0] + 8 && (char *) outs[2] == (char *) outs[0] + 16;
    if (gf) {
        for (int q = 0; q < nf; q++) f.out2[q] = (double *) gf->own_outs[q];
        f.oidx = gf->own_index;
        f.sel0 = gf->own_begin;
        f.sel1 = gf->own_begin + gf->own_count;
    }
    const uint32_t *perm;
    PmbBinned bn;
    bool walk;
    PMB_CHECK(pmb_traversal(ctx, g, p, a->npart, &bn, &perm, &walk, false));
    const int rs = pmb_bin_rowwords(nf);
    double *stage = bn.pos ? (double *) pmb_bin_try_scratch(ctx, sizeof(double) * (size_t) rs * (size_t) a->npart) : NULL;
    if (bn.pos && !stage) {
        PMB_CHECK(pmb_perm_prepare(ctx, g, p, a->npart, &perm));
        walk = perm != NULL;
    }
    if (stage) {
        // the nf gathers on the sorted copy into (npart, nf) staging rows, then back to the caller's order -- through
        // pmb_store_result, so that the fused ghost sum sees the caller's particle numbers
        pmb_resample_args b;
        pmb_binned_args(a, bn, &b);
        b.out_elsize = 8;
        void *outs2[3];
        int64_t strides2[3];
        for (int q = 0; q < nf; q++) { outs2[q] = stage + q; strides2[q] = (int64_t) sizeof(double) * rs; }
        ctx->bin_bypass = 1;
        const int r = readout_multi_impl<MeshT>(ctx, &b, nf, meshes, outs2, strides2, done, NULL);
        ctx->bin_bypass = 0;
        PMB_CHECK(r);
        if (!*done) return PMB_OK;
        return pmb_bin_unsort(ctx, bn, f, nf, stage, a->npart);
    }
    if (walk) {
        const int gridp = pmb_grid(ctx, a->npart, 256, 8);
        const bool chk = pmb_geom_needs_check(g);
        if (nf == 1) { PMB_DISPATCH_CHECK(chk, (pmb_k_readout_cic32_perm<MeshT, CHECK, 1><<<gridp, 256, 0, ctx->stream>>>(g32, p, f, a->npart, perm))); }
        else if (nf == 2) { PMB_DISPATCH_CHECK(chk, (pmb_k_readout_cic32_perm<MeshT, CHECK, 2><<<gridp, 256, 0, ctx->stream>>>(g32, p, f, a->npart, perm))); }
        else { PMB_DISPATCH_CHECK(chk, (pmb_k_readout_cic32_perm<MeshT, CHECK, 3><<<gridp, 256, 0, ctx->stream>>>(g32, p, f, a->npart, perm))); }
        PMB_LAUNCH_CHECK(ctx);
        *done = true;
        return PMB_OK;
    }
    if (!pmb_env_flag("PMB_RING", 1) || !((pmb_pos_is_f8_rows(p) && ((uintptr_t) p.pos & 15) == 0) || pmb_pos_is_f8_rec4(p))) return PMB_OK;
    *done = true;
    return readout_ring<MeshT>(ctx, g32, p, f, nf, a->npart, pmb_geom_needs_check(g));
}

extern "C" int pmb_readout_multi(pmb_ctx *ctx, const pmb_resample_args *a, int nfields, const void *const *meshes_h,
                                 void *const *outs_h, const int64_t *out_strides_h)
{
    PMB_REQUIRE(ctx && a && meshes_h && outs_h && out_strides_h, "null argument");
    PMB_REQUIRE(nfields >= 1 && nfields <= 3, "1..3 fields");
    pmb_resample_args b = *a;
    b.mesh = (void *) meshes_h[0];
    b.out = outs_h[0];
    b.out_stride = out_strides_h[0];
    PMB_CHECK(check_args(ctx, &b, 1));
    for (int q = 0; q < nfields; q++) PMB_REQUIRE(a->npart == 0 || (meshes_h[q] && outs_h[q]), "null mesh / out %d", q);
    if (a->npart == 0) return PMB_OK;
    bool done = false;
    if (a->mesh_elsize == 8) PMB_CHECK(readout_multi_impl<double>(ctx, a, nfields, meshes_h, outs_h, out_strides_h, &done));
    else PMB_CHECK(readout_multi_impl<float>(ctx, a, nfields, meshes_h, outs_h, out_strides_h, &done));
    if (done) return PMB_OK;
    // any other window / geometry: one ordinary readout per field
    for (int q = 0; q < nfields; q++) {
        b = *a;
        b.mesh = (void *) meshes_h[q];
        b.out = outs_h[q];
        b.out_stride = out_strides_h[q];
        PMB_CHECK(pmb_readout(ctx, &b));
    }
    return PMB_OK;
}

extern "C" int pmb_readout_multi_gather(pmb_ctx *ctx, const pmb_resample_args *a, int nfields, const void *const *meshes_h,
                                        void *const *ghost_outs_h, void *const *own_outs_h, const int32_t *own_index,
                                        int64_t own_begin, int64_t own_count)
{
    PMB_REQUIRE(ctx && a && meshes_h && ghost_outs_h && own_outs_h, "null argument");
    PMB_REQUIRE(nfields >= 1 && nfields <= 3, "1..3 fields");
    PMB_REQUIRE(own_begin >= 0 && own_count >= 0 && own_begin + own_count <= a->npart, "bad range of own particles");
    PMB_REQUIRE(own_count == 0 || own_index, "null index of own particles");
    PMB_REQUIRE(a->out_elsize == 8, "the fused ghost sum produces float64 columns");
    pmb_resample_args b = *a;
    b.mesh = (void *) meshes_h[0];
    b.out = ghost_outs_h[0];
    b.out_stride = 8;
    PMB_CHECK(check_args(ctx, &b, 1));
    if (a->npart == 0) return PMB_OK;
    int64_t strides[3] = {8, 8, 8};
    PmbGatherFuse gf = {own_outs_h, own_index, own_begin, own_count};
    bool done = false;
    if (a->mesh_elsize == 8) PMB_CHECK(readout_multi_impl<double>(ctx, a, nfields, meshes_h, ghost_outs_h, strides, &done, &gf));
    else PMB_CHECK(readout_multi_impl<float>(ctx, a, nfields, meshes_h, ghost_outs_h, strides, &done, &gf));
    if (done) return PMB_OK;
    pmb_set_error("fused readout + ghost sum is available for the CIC window on 3-D meshes with >= 2^18 contiguous float64 positions");
    return PMB_EUNSUPPORTED;
}

extern "C" int pmb_readout_grad(pmb_ctx *ctx, const pmb_resample_args *a, void *out_grad, int64_t gs0, int64_t gs1)
{
    PMB_REQUIRE(ctx && a, "null argument");
    PMB_REQUIRE(a->ndim >= 1 && a->ndim <= 3, "ndim %d not supported (1..3)", a->ndim);
    PMB_REQUIRE(a->npart == 0 || out_grad, "null out_grad");
    PMB_REQUIRE(a->out_elsize == 4 || a->out_elsize == 8, "out must be float32 or float64");
    PMB_REQUIRE(a->mesh_elsize == 4 || a->mesh_elsize == 8, "canvas must be float32 or float64");
    for (int d = 0; d < a->ndim; d++)
        PMB_REQUIRE(a->order[d] == 0, "gradient of gradient is not supported (pm.py:822-823)");
    if (a->npart == 0) return PMB_OK;
    PmbGeom g;
    fill_geom(a, &g);
    PmbParticles p;
    fill_particles(a, &p);
    PmbWindow w;
    PMB_CHECK(pmb_resolve_window(ctx, a->kind, a->support, a->ndim, a->order, &w, 1));
    const int fam = fixed_family(w, a);
    if (!fam) {
        // generic windows whose stencil fits the per-thread arrays: one fused sweep (pmb_k_readout_grad_dyn)
        bool fits = !a->hsml;
        if (fits) {
            PmbWinInfo info;
            pmb_window_info(w.nativesupport, w.support * a->hsml_scalar, &info);
            fits = info.support <= PMB_MAX_SUPPORT;
        }
        if (fits && pmb_env_flag("PMB_GRAD_FUSED", 1)) {
            const char *mesh = (const char *) a->mesh;
            const int grid = pmb_grid(ctx, a->npart, 128, 8);
            if (a->mesh_elsize == 8) {
                PMB_DISPATCH_NDIM(a->ndim, (pmb_k_readout_grad_dyn<double, NDIM><<<grid, 128, 0, ctx->stream>>>(
                    g, w, p, mesh, a->npart, a->pcs_gradient_scale_fix, a->out, a->out_elsize, a->out_stride, out_grad, gs0, gs1)));
            } else {
                PMB_DISPATCH_NDIM(a->ndim, (pmb_k_readout_grad_dyn<float, NDIM><<<grid, 128, 0, ctx->stream>>>(
                    g, w, p, mesh, a->npart, a->pcs_gradient_scale_fix, a->out, a->out_elsize, a->out_stride, out_grad, gs0, gs1)));
            }
            PMB_LAUNCH_CHECK(ctx);
            return PMB_OK;
        }
        // per-particle hsml or very wide stencils: value pass + one pass per axis through the ordinary readout kernel
        pmb_resample_args b = *a;
        if (a->out) PMB_CHECK(pmb_readout(ctx, &b));
        for (int d = 0; d < a->ndim; d++) {
            b = *a;
            b.order[d] = 1;
            b.out = (char *) out_grad + d * gs1;
            b.out_stride = gs0;
            PMB_CHECK(pmb_readout(ctx, &b));
        }
        return PMB_OK;
    }
    const char *mesh = (const char *) a->mesh;
    int grid = pmb_grid(ctx, a->npart, 256, 8);
    const bool chk = pmb_geom_needs_check(g);
    if (a->mesh_elsize == 8) {
        PMB_DISPATCH_CHECK(chk, PMB_DISPATCH_NDIM(a->ndim, PMB_DISPATCH_FAM(fam, NDIM,
            (pmb_k_readout_grad_tuned<double, NDIM, FAM, CHECK><<<grid, 256, 0, ctx->stream>>>(
                g, p, mesh, a->npart, a->pcs_gradient_scale_fix, a->out, a->out_elsize, a->out_stride, out_grad, gs0, gs1)))));
    } else {
        PMB_DISPATCH_CHECK(chk, PMB_DISPATCH_NDIM(a->ndim, PMB_DISPATCH_FAM(fam, NDIM,
            (pmb_k_readout_grad_tuned<float, NDIM, FAM, CHECK><<<grid, 256, 0, ctx->stream>>>(
                g, p, mesh, a->npart, a->pcs_gradient_scale_fix, a->out, a->out_elsize, a->out_stride, out_grad, gs0, gs1)))));
    }
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}

extern "C" int pmb_bin_stats(pmb_ctx *ctx, int64_t *builds, int64_t *bytes_held)
{
    PMB_REQUIRE(ctx, "null context");
    if (builds) *builds = ctx->bin_builds;
    if (bytes_held) *bytes_held = (int64_t) (ctx->bin_pos_bytes + ctx->bin_dest_bytes + ctx->bin_col_bytes);
    return PMB_OK;
}

// everything the library caches for itself goes back to the device (the host allocator calls this before it gives up)
extern "C" int pmb_ctx_trim(pmb_ctx *ctx)
{
    PMB_REQUIRE(ctx, "null context");
    PMB_CUDA(cudaStreamSynchronize(ctx->stream));
    pmb_bin_free(ctx);
    if (ctx->scratch) { cudaFree(ctx->scratch); ctx->scratch = NULL; ctx->scratch_bytes = 0; }
    if (ctx->perm_ids) { cudaFree(ctx->perm_ids); ctx->perm_ids = NULL; ctx->perm_bytes = 0; ctx->perm_sig = 0; }
    cudaGetLastError();
    return PMB_OK;
}

extern "C" int pmb_bin_invalidate(pmb_ctx *ctx)
{
    PMB_REQUIRE(ctx, "null context");
    ctx->bin_sig = 0;
    ctx->bin_state = 0;
    ctx->bin_npart = -1;
    return PMB_OK;
}

extern "C" int pmb_bin_release(pmb_ctx *ctx)
{
    PMB_REQUIRE(ctx, "null context");
    PMB_CUDA(cudaStreamSynchronize(ctx->stream));
    pmb_bin_free(ctx);
    return PMB_OK;
}
