// pmb_whitenoise.cu -- Gaussian / unitary white noise in Fourier space, N-GenIC (Gadget) scheme.
//
// Replaces pmesh._whitenoise.generate -> pmesh_whitenoise_generator_fill
//   (pmesh/_whitenoise.pyx:25-45, pmesh/_whitenoise_imp.c:68-105, pmesh/_whitenoise_generics.h:29-232)
// and the RANLUX generator it draws from (gsl_rng_ranlxd1, pmesh/gsl/ranlxd.c:36-245).
//
// Structure of the reference: (1) ONE master stream, walked sequentially in a square spiral over
// (i, j), hands every Fourier column its own 31-bit seed; (2) every column (i, j, 0..N/2) is then
// filled from its own freshly seeded stream(s), independently of all others.  (1) is inherently
// serial and tiny (N0*N1 draws): it runs on the host, identically on every rank, like the reference
// does.  (2) is the work: one GPU thread per column, the 12-word generator state held in registers.
//
// The generator: words are k * 2^-48, z[n] = z[n-5] - z[n-12] - borrow (mod 1); all arithmetic on
// them is exact in doubles.  A refill advances the ring by 202 steps, then the 12 words of the ring
// are handed out oldest first.  On the device the ring is 12 registers with compile-time indices:
// a block of 12 steps updates the registers in place, 202 = 16 * 12 + 10, and handing out a word
// rotates the registers by one (so no dynamically indexed array ever spills to local memory).
//
// Only the compressed (half-complex, k_z <= N/2) block layout is produced -- the layout of every
// field of this engine; the reference's full-spectrum fill exists for c2c meshes (not built here).
#include <string.h>

#include "pmb_internal.h"
#include "pmb_wnrng.h"

// ---- device: the columns ---------------------------------------------------------------------------
#define WN_KTILE 16

template <typename T>   // float (complex64) or double (complex128)
__device__ __forceinline__ void wn_store(char *p, double re, double im)
{
    T *q = (T *) p;
    q[0] = (T) re;
    q[1] = (T) im;
}

// one thread per column; KCONTIG: k is the contiguous axis of the canvas -> values are staged in a
// (32 columns x 16 k) shared tile and written as 16-element rows; otherwise consecutive threads
// are consecutive along the contiguous column axis and store directly.
template <typename T, bool KCONTIG>
__global__ void __launch_bounds__(32)
pmb_k_whitenoise(WnArgs a, char *canvas)
{
    __shared__ double tile[KCONTIG ? 32 : 1][KCONTIG ? (2 * WN_KTILE + 2) : 2];
    const int lane = threadIdx.x;
    const int64_t ncol = a.size[0] * a.size[1];
    const int64_t col0 = (int64_t) blockIdx.x * 32;
    const int64_t col = col0 + lane;
    const bool valid = col < ncol;
    int64_t li = 0, lj = 0;
    if (valid) {
        if (a.fast_axis == 0) { li = col % a.size[0]; lj = col / a.size[0]; }
        else { lj = col % a.size[1]; li = col / a.size[1]; }
    }
    WnColumn c;
    if (valid) wn_column_init(c, a, li, lj);
    const int64_t kmax = a.N[2] / 2;
    char *base = canvas + li * a.strides[0] + lj * a.strides[1];
    for (int64_t k0 = 0; k0 <= kmax; k0 += WN_KTILE) {
        const int nk = (int) min((int64_t) WN_KTILE, kmax + 1 - k0);
        if (valid) {
            for (int kk = 0; kk < nk; kk++) {
                const int64_t k = k0 + kk;
                double re, im;
                if (!wn_column_mode<T>(c, a, k, re, im)) continue;
                if (KCONTIG) { tile[lane][2 * kk] = re; tile[lane][2 * kk + 1] = im; }
                else wn_store<T>(base + (k - a.start[2]) * a.strides[2], re, im);
            }
        }
        if (KCONTIG) {
            __syncwarp();
            // 32 lanes = 2 columns x 16 k: each half-warp writes one contiguous 16-element row
            const int kk = lane & (WN_KTILE - 1);
            const int64_t k = k0 + kk;
            for (int r = lane >> 4; r < 32; r += 2) {
                const int64_t cc = col0 + r;
                if (cc < ncol && kk < nk && k >= a.start[2] && k < a.start[2] + a.size[2]) {
                    int64_t ri, rj;
                    if (a.fast_axis == 0) { ri = cc % a.size[0]; rj = cc / a.size[0]; }
                    else { rj = cc % a.size[1]; ri = cc / a.size[1]; }
                    wn_store<T>(canvas + ri * a.strides[0] + rj * a.strides[1] + (k - a.start[2]) * a.strides[2],
                                tile[r][2 * kk], tile[r][2 * kk + 1]);
                }
            }
            __syncwarp();
        }
    }
}

extern "C" int pmb_whitenoise(pmb_ctx *ctx, void *cplx, int elsize, const int64_t *nmesh, const int64_t *start,
                              const int64_t *size, const int64_t *strides, unsigned int seed, int unitary)
{
    PMB_REQUIRE(ctx && nmesh && start && size && strides, "null argument");
    PMB_REQUIRE(elsize == 8 || elsize == 16, "white noise canvas must be complex64 or complex128");
    for (int d = 0; d < 3; d++) {
        PMB_REQUIRE(nmesh[d] >= 1 && start[d] >= 0 && size[d] >= 0 && start[d] + size[d] <= nmesh[d], "bad block on axis %d", d);
    }
    if (start[2] + size[2] > nmesh[2] / 2 + 1) {
        pmb_set_error("only the compressed (k_z <= N/2) half of the Fourier mesh can be generated");
        return PMB_EUNSUPPORTED;
    }
    const int64_t ncol = size[0] * size[1];
    if (ncol == 0 || size[2] == 0) return PMB_OK;
    PMB_REQUIRE(cplx, "null canvas");
    PMB_REQUIRE(ncol < ((int64_t) 1 << 36), "block too large");

    WnTables T;
    T.N0 = nmesh[0]; T.N1 = nmesh[1]; T.s0 = start[0]; T.s1 = start[1]; T.m0 = size[0]; T.m1 = size[1];
    T.t00 = (unsigned int *) calloc((size_t) ncol * 2, sizeof(unsigned int));
    if (!T.t00) return PMB_ENOMEM;
    T.t11 = T.t00 + ncol;
    wn_build_tables(T, seed);
    void *dev;
    int rc = pmb_scratch(ctx, sizeof(unsigned int) * 2 * (size_t) ncol, &dev);
    if (rc != PMB_OK) { free(T.t00); return rc; }
    cudaError_t e = cudaMemcpyAsync(dev, T.t00, sizeof(unsigned int) * 2 * (size_t) ncol, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);     // pageable source is freed next
    free(T.t00);
    if (e != cudaSuccess) return pmb_cuda_fail(e, "white noise seed tables", __FILE__, __LINE__);

    WnArgs a;
    for (int d = 0; d < 3; d++) { a.N[d] = nmesh[d]; a.start[d] = start[d]; a.size[d] = size[d]; a.strides[d] = strides[d]; }
    a.t00 = (const unsigned int *) dev;
    a.t11 = a.t00 + ncol;
    a.unitary = unitary ? 1 : 0;
    const bool kcontig = strides[2] == elsize;
    a.fast_axis = llabs((long long) strides[0]) < llabs((long long) strides[1]) ? 0 : 1;
    const int64_t grid = (ncol + 31) / 32;
    PMB_REQUIRE(grid < ((int64_t) 1 << 31), "block too large");
    if (elsize == 16) {
        if (kcontig) pmb_k_whitenoise<double, true><<<(int) grid, 32, 0, ctx->stream>>>(a, (char *) cplx);
        else pmb_k_whitenoise<double, false><<<(int) grid, 32, 0, ctx->stream>>>(a, (char *) cplx);
    } else {
        if (kcontig) pmb_k_whitenoise<float, true><<<(int) grid, 32, 0, ctx->stream>>>(a, (char *) cplx);
        else pmb_k_whitenoise<float, false><<<(int) grid, 32, 0, ctx->stream>>>(a, (char *) cplx);
    }
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}
