// pmb_fft.cu -- r2c / c2r on the padded pmesh layouts, slab-decomposed over the communicator,
// plus the k-space transfer kernels of the PM force step.
//
// Replaces pfft.Plan.execute as used by RealField.r2c / ComplexField.c2r (pmesh/pm.py:655-694,
// 987-1019): forward = unnormalised DFT times `scale` (the caller passes 1/prod(Nmesh), pm.py:692),
// backward = unnormalised inverse.  cuFFT does the local batched 1-D/2-D transforms (library time,
// accounted separately with events); the pack / global transpose / unpack(+scale) around the NCCL
// all-to-all and the transfer functions are this library's kernels.
//
// Layouts (elements of the real dtype T, complex = 2 T):
//   real  : C order (n0_local, n1, 2*(n2/2+1)), last axis padded -- PFFT_PADDED_R2C (pm.py:1335)
//   complex, 1 rank : C order (n0, n1, nc)
//   complex, P ranks: distributed along axis 1, memory order (1, 2, 0) = (n1_local, nc, n0)
//                     -- the "transposed out" representation (TransposedComplexField, pm.py:1078-1086)
#include <cufft.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "pmb_internal.h"
#include "pmb_ifft.cuh"

#define FFT_NEV 32

struct pmb_fft {
    pmb_ctx *ctx;
    int ndim;
    int64_t n[3];      // left-padded to 3-D: (1, 1, n) / (1, n0, n1) / (n0, n1, n2)
    int64_t nc;        // n[2] / 2 + 1
    int elsize;        // 4 or 8
    int P, rank;
    int64_t s0, m0, blk0;   // real-space slab of this rank along the first distributed axis
    int64_t s1, m1, blk1;   // complex-space slab along axis 1
    int dist_axis;          // index in n[] of the real-space distributed axis (P > 1): 0
    // process mesh P0 x P1 (rank = c0 * P1 + c1, C order like pfft.ProcMesh); slabs are P1 == 1.
    // Pencils (P1 > 1): real space split along axes (0, 1) over (P0, P1); complex space ("transposed out")
    // along axes (1, 2) over (P0, P1), memory order (1, 2, 0)
    int P0, P1, c0, c1;
    int64_t r1s, r1m;       // real space: my block of axis 1 (over P1); slabs: the whole axis
    int64_t s2, mc;         // complex space: my block of the half-complex axis 2 (over P1); slabs: 0, nc
    cufftHandle full_r2c, full_c2r;         // P == 1
    cufftHandle slab_r2c, slab_c2r, line;   // slabs
    cufftHandle slab_r2c_part;              // 2-D r2c over m0 / r2c_chunks planes (pipelined forward transform)
    int r2c_chunks;                         // 0: no partial plan
    cufftHandle pen_r2c, pen_c2r, pen_l1, pen_l0;   // pencils: 1-D transforms along axes 2, 1, 0
    bool have_full, have_slab, have_pen;
    void *work0, *work1;
    void *work2;                  // third line buffer of the one-launch fused backward pass (slabs), made on first use
    size_t work_bytes;
    // direct peer-memory global transpose (P > 1): every rank exposes two landing buffers through
    // CUDA IPC; the transpose kernels of the other ranks store straight into them over NVLink
    int p2p;                      // 1: peer path armed, 0: NCCL send/recv path
    int xcur;                     // landing buffer used by the next transform (they alternate)
    void *xbuf[2];                // my landing buffers (cudaMalloc, IPC-exported)
    void *peer_x[64][2];          // the same buffers of every rank, mapped into this process
    void *pool;                   // the pool entry that owns them
    cudaEvent_t ev[FFT_NEV][2];
    int ev_kind[FFT_NEV];         // 0: a cuFFT exec (library time), 1: a transpose kernel of this library, 2: the fused transfer + line transform
    int nev;
    float lib_ms;
    // second stream for the transposes of pmb_fft_c2r_multi (NVLink stores overlap the cuFFT of the other transforms)
    cudaStream_t xstream;
    cudaEvent_t xev[16];
    bool have_xstream;
    float xpose_ms;               // time inside the transpose (+ NVLink store) kernels
    double xpose_remote_bytes;    // bytes they stored into OTHER ranks' landing buffers
    // device copies of the per-axis tables of the transfer functions (wavenumbers | multipliers), kept per
    // (kind, direction, box, parameters): the force step applies the same few transfers every step
    struct TfCache { int kind, dir; double box[3], p[2]; void *dev; } tfc[8];
    int ntfc;
    // transfer + axis-0 inverse transform in one kernel (pmb_ifft.cuh): twiddle table, the 2-D c2r over all planes of
    // one rank (P == 1; P > 1 uses slab_c2r), time inside the fused kernel
    void *ifft_tw;
    cufftHandle plane_c2r, plane_r2c;
    bool have_plane, have_plane_r2c;
    float ifft_ms;
    int64_t ifft_launches;
};

// cached tables or NULL
static void *tf_cache_find(pmb_fft *f, int kind, int dir, const double *box, double p0, double p1)
{
    for (int i = 0; i < f->ntfc; i++) {
        const pmb_fft::TfCache &c = f->tfc[i];
        if (c.kind == kind && c.dir == dir && c.box[0] == box[0] && c.box[1] == box[1] && c.box[2] == box[2] &&
            c.p[0] == p0 && c.p[1] == p1) return c.dev;
    }
    return NULL;
}

// upload freshly built host tables (nbytes) and remember them; *dev is valid for kernels on the context's stream
static int tf_cache_store(pmb_fft *f, int kind, int dir, const double *box, double p0, double p1,
                          const void *host, size_t nbytes, void **dev)
{
    pmb_ctx *ctx = f->ctx;
    if (f->ntfc == 8) {
        // full: drop the oldest entry (kernels that read it are ordered before the free by the synchronisation)
        PMB_CUDA(cudaStreamSynchronize(ctx->stream));
        PMB_CUDA(cudaFree(f->tfc[0].dev));
        for (int i = 1; i < 8; i++) f->tfc[i - 1] = f->tfc[i];
        f->ntfc = 7;
    }
    PMB_CUDA(cudaMalloc(dev, nbytes));
    cudaError_t e = cudaMemcpyAsync(*dev, host, nbytes, cudaMemcpyHostToDevice, ctx->stream);   // pageable source: staged before return
    if (e != cudaSuccess) { cudaFree(*dev); return pmb_cuda_fail(e, "transfer tables", __FILE__, __LINE__); }
    pmb_fft::TfCache &c = f->tfc[f->ntfc++];
    c.kind = kind; c.dir = dir; c.p[0] = p0; c.p[1] = p1; c.dev = *dev;
    for (int d = 0; d < 3; d++) c.box[d] = box[d];
    return PMB_OK;
}

static int cufft_fail(cufftResult r, const char *what, int line)
{
    pmb_set_error("cuFFT error %d at %s:%d in %s", (int) r, __FILE__, line, what);
    return PMB_ECUDA;
}
#define PMB_CUFFT(call)                                                  \
    do {                                                                 \
        cufftResult _r = (call);                                         \
        if (_r != CUFFT_SUCCESS) return cufft_fail(_r, #call, __LINE__); \
    } while (0)

static void block_partition(int64_t n, int P, int rank, int64_t *blk, int64_t *start, int64_t *len)
{
    // FFTW/PFFT default block: ceil(n / P); trailing ranks may own fewer (or zero) rows
    int64_t b = (n + P - 1) / P;
    int64_t s = b * rank;
    if (s > n) s = n;
    int64_t e = s + b;
    if (e > n) e = n;
    *blk = b; *start = s; *len = e - s;
}

static int make_plan(pmb_fft *f, cufftHandle *h, int rank, long long *n, long long *inembed, long long idist,
                     long long *onembed, long long odist, cufftType type, long long batch);

// whole-mesh r2c / c2r plans of one rank, made on first use
static int ensure_full(pmb_fft *f)
{
    if (f->have_full) return PMB_OK;
    const int ndim = f->ndim;
    long long n[3], inr[3], inc[3];
    for (int d = 0; d < ndim; d++) {
        n[d] = f->n[3 - ndim + d];
        inr[d] = n[d];
        inc[d] = n[d];
    }
    inr[ndim - 1] = 2 * f->nc;
    inc[ndim - 1] = f->nc;
    long long rdist = 1, cdist = 1;
    for (int d = 0; d < ndim; d++) { rdist *= inr[d]; cdist *= inc[d]; }
    const bool dbl = f->elsize == 8;
    PMB_CHECK(make_plan(f, &f->full_r2c, ndim, n, inr, rdist, inc, cdist, dbl ? CUFFT_D2Z : CUFFT_R2C, 1));
    int rc = make_plan(f, &f->full_c2r, ndim, n, inc, cdist, inr, rdist, dbl ? CUFFT_Z2D : CUFFT_C2R, 1);
    if (rc != PMB_OK) { cufftDestroy(f->full_r2c); return rc; }
    f->have_full = true;
    return PMB_OK;
}

static int make_plan(pmb_fft *f, cufftHandle *h, int rank, long long *n, long long *inembed, long long idist,
                     long long *onembed, long long odist, cufftType type, long long batch)
{
    PMB_CUFFT(cufftCreate(h));
    size_t ws = 0;
    PMB_CUFFT(cufftMakePlanMany64(*h, rank, n, inembed, 1, idist, onembed, 1, odist, type, batch, &ws));
    PMB_CUFFT(cufftSetStream(*h, f->ctx->stream));
    return PMB_OK;
}

// ---- peer-memory landing buffers -------------------------------------------------------------------
// One process per GPU: the buffers are shared with cudaIpc handles, exchanged over the communicator.
// Any failure (IPC not permitted in the container, no peer access) leaves p2p = 0 on EVERY rank --
// the decision is agreed with an allgather -- and the transforms use NCCL send/recv.
//
// Lifetime: buffers and mappings live in a process-wide pool and are never unmapped or freed before
// the process exits.  CUDA leaves freeing an exported allocation that a peer still maps undefined,
// and plans are destroyed at different moments on different ranks (garbage collection), so a plan
// only RETURNS its pool entry when it dies; the next plan of the same mesh takes it again -- if every
// rank has one to take (agreed with an allgather; otherwise all ranks create a fresh entry together).
struct P2PEntry {
    pmb_ctx *ctx;
    int64_t n[3];
    int elsize, P, P1;
    size_t bytes;
    void *xbuf[2];
    void *peer_x[64][2];
    bool in_use;
    long long serial;     // creation number: entries are created collectively, so it is the same on every rank
    P2PEntry *next;
};
static P2PEntry *g_p2p_pool = NULL;
static long long g_p2p_serial = 0;
#define P2P_NCAND 16

static int p2p_setup(pmb_fft *f)
{
    pmb_ctx *ctx = f->ctx;
    f->p2p = 0;
    f->xcur = 0;
    f->pool = NULL;
    const char *env = getenv("PMB_FFT_P2P");
    const int want = (env ? atoi(env) : 1) && f->P <= 64;
    // 1. can everybody reuse THE SAME idle entry of this mesh?  Plans die at rank-dependent moments (garbage
    // collection), so the sets of idle entries differ between ranks: every rank publishes the serials of its
    // idle candidates and all take the smallest serial that is idle everywhere; otherwise a fresh entry is
    // created collectively.  (Agreeing only on "everybody has some idle entry" would let rank A store into
    // an entry that rank B has handed to another live plan.)
    struct Cand { int want; int n; long long serial[P2P_NCAND]; } mine_c, all_c[64];
    memset(&mine_c, 0, sizeof(mine_c));
    mine_c.want = want;
    for (P2PEntry *e = g_p2p_pool; e && want; e = e->next)
        if (!e->in_use && e->ctx == ctx && e->P == f->P && e->P1 == f->P1 && e->elsize == f->elsize && e->bytes >= f->work_bytes &&
            e->n[0] == f->n[0] && e->n[1] == f->n[1] && e->n[2] == f->n[2] && mine_c.n < P2P_NCAND)
            mine_c.serial[mine_c.n++] = e->serial;
    PMB_CHECK(pmb_allgather_host(ctx, &mine_c, all_c, sizeof(Cand)));
    int all_want = 1;
    for (int q = 0; q < f->P; q++) all_want = all_want && all_c[q].want;
    if (!all_want) return PMB_OK;
    long long common = -1;
    for (int i = 0; i < all_c[0].n; i++) {
        const long long cand = all_c[0].serial[i];
        bool everywhere = true;
        for (int q = 1; q < f->P && everywhere; q++) {
            bool found = false;
            for (int j = 0; j < all_c[q].n; j++) found = found || all_c[q].serial[j] == cand;
            everywhere = found;
        }
        if (everywhere && (common < 0 || cand < common)) common = cand;
    }
    P2PEntry *idle = NULL;
    for (P2PEntry *e = g_p2p_pool; e && common >= 0; e = e->next)
        if (e->serial == common) { idle = e; break; }
    const int all_idle = idle != NULL;
    if (all_idle) {
        idle->in_use = true;
        f->pool = idle;
    } else {
        // 2. a fresh entry, created by all ranks together (the serial counts attempts: same on every rank)
        g_p2p_serial++;
        P2PEntry *e = (P2PEntry *) calloc(1, sizeof(P2PEntry));
        if (!e) return PMB_ENOMEM;
        struct Rec { cudaIpcMemHandle_t h[2]; int ok; int pad; } mine, all[64];
        memset(&mine, 0, sizeof(mine));
        mine.ok = 1;
        for (int b = 0; b < 2 && mine.ok; b++) {
            if (cudaMalloc(&e->xbuf[b], f->work_bytes) != cudaSuccess) { e->xbuf[b] = NULL; mine.ok = 0; break; }
            if (cudaIpcGetMemHandle(&mine.h[b], e->xbuf[b]) != cudaSuccess) mine.ok = 0;
        }
        cudaGetLastError();
        PMB_CHECK(pmb_allgather_host(ctx, &mine, all, sizeof(Rec)));
        int ok = 1;
        for (int q = 0; q < f->P; q++) ok = ok && all[q].ok;
        for (int q = 0; q < f->P && ok; q++)
            for (int b = 0; b < 2 && ok; b++) {
                if (q == f->rank) { e->peer_x[q][b] = e->xbuf[b]; continue; }
                if (cudaIpcOpenMemHandle(&e->peer_x[q][b], all[q].h[b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    e->peer_x[q][b] = NULL;
                    ok = 0;
                }
            }
        cudaGetLastError();
        // every rank could map every buffer?
        int mine_ok = ok, all_ok[64];
        PMB_CHECK(pmb_allgather_host(ctx, &mine_ok, all_ok, sizeof(int)));
        for (int q = 0; q < f->P; q++) ok = ok && all_ok[q];
        if (!ok) {
            for (int q = 0; q < f->P; q++)
                for (int b = 0; b < 2; b++)
                    if (q != f->rank && e->peer_x[q][b]) cudaIpcCloseMemHandle(e->peer_x[q][b]);
            // peers may still be unmapping: nobody frees before everybody is done
            PMB_CHECK(pmb_allgather_host(ctx, &mine_ok, all_ok, sizeof(int)));
            for (int b = 0; b < 2; b++) if (e->xbuf[b]) cudaFree(e->xbuf[b]);
            cudaGetLastError();
            free(e);
            return PMB_OK;
        }
        e->ctx = ctx; e->P = f->P; e->P1 = f->P1; e->elsize = f->elsize; e->bytes = f->work_bytes;
        for (int d = 0; d < 3; d++) e->n[d] = f->n[d];
        e->in_use = true;
        e->serial = g_p2p_serial;
        e->next = g_p2p_pool;
        g_p2p_pool = e;
        f->pool = e;
    }
    P2PEntry *e = (P2PEntry *) f->pool;
    for (int b = 0; b < 2; b++) {
        f->xbuf[b] = e->xbuf[b];
        for (int q = 0; q < f->P; q++) f->peer_x[q][b] = e->peer_x[q][b];
    }
    f->p2p = 1;
    return PMB_OK;
}

static void p2p_teardown(pmb_fft *f)
{
    if (!f->p2p) return;
    // No collective and no unmapping here (see above): a peer stores into my buffers only inside a
    // transform, and I cannot have passed that transform's barrier before its stores landed; the
    // entry goes back to the pool for the next plan of this mesh.
    cudaStreamSynchronize(f->ctx->stream);
    ((P2PEntry *) f->pool)->in_use = false;
    f->pool = NULL;
    f->p2p = 0;
}

extern "C" int pmb_fft_create(pmb_ctx *ctx, int ndim, const int64_t *nmesh, int dtype_elsize, pmb_fft **out)
{
    const int np[2] = {ctx ? ctx->nranks : 1, 1};
    return pmb_fft_create_np(ctx, ndim, nmesh, dtype_elsize, np, out);
}

extern "C" int pmb_fft_create_np(pmb_ctx *ctx, int ndim, const int64_t *nmesh, int dtype_elsize, const int *np, pmb_fft **out)
{
    PMB_REQUIRE(ctx && nmesh && out && np, "null argument");
    PMB_REQUIRE(np[0] >= 1 && np[1] >= 1 && np[0] * np[1] == ctx->nranks, "process mesh %d x %d does not match %d ranks", np[0], np[1], ctx->nranks);
    PMB_REQUIRE(ndim >= 1 && ndim <= 3, "FFT supports 1..3 dimensions");
    PMB_REQUIRE(dtype_elsize == 4 || dtype_elsize == 8, "FFT dtype must be float32 or float64");
    for (int d = 0; d < ndim; d++) PMB_REQUIRE(nmesh[d] >= 1, "bad mesh size");
    PMB_REQUIRE(ctx->nranks == 1 || ndim == 3, "multi-rank FFT is implemented for 3-D meshes (slab decomposition)");
    pmb_fft *f = (pmb_fft *) calloc(1, sizeof(pmb_fft));
    if (!f) return PMB_ENOMEM;
    f->ctx = ctx;
    f->ndim = ndim;
    for (int d = 0; d < 3; d++) f->n[d] = 1;
    for (int d = 0; d < ndim; d++) f->n[3 - ndim + d] = nmesh[d];
    f->nc = f->n[2] / 2 + 1;
    f->elsize = dtype_elsize;
    f->P = ctx->nranks;
    f->rank = ctx->rank;
    f->P0 = np[0]; f->P1 = np[1];
    f->c0 = f->rank / f->P1; f->c1 = f->rank % f->P1;
    f->r1s = 0; f->r1m = f->n[1];
    f->s2 = 0; f->mc = f->nc;
    PMB_CUDA(cudaSetDevice(ctx->device));
    for (int i = 0; i < FFT_NEV; i++) {
        PMB_CUDA(cudaEventCreate(&f->ev[i][0]));
        PMB_CUDA(cudaEventCreate(&f->ev[i][1]));
    }
    const bool dbl = dtype_elsize == 8;
    if (f->P == 1) {
        f->s0 = 0; f->m0 = f->n[0]; f->s1 = 0; f->m1 = f->n[1];
        // the two whole-mesh plans (and their cuFFT work areas) are made when first used (ensure_full): a force step that
        // goes through pmb_fft_force3 never needs them
    } else if (f->P1 > 1) {
        // ---- pencils ----
        int64_t blk;
        block_partition(f->n[0], f->P0, f->c0, &f->blk0, &f->s0, &f->m0);     // real axis 0 over P0
        block_partition(f->n[1], f->P1, f->c1, &blk, &f->r1s, &f->r1m);        // real axis 1 over P1
        block_partition(f->n[1], f->P0, f->c0, &f->blk1, &f->s1, &f->m1);     // complex axis 1 over P0
        block_partition(f->nc, f->P1, f->c1, &blk, &f->s2, &f->mc);            // complex axis 2 over P1
        if (f->m0 * f->r1m > 0) {
            long long n1[1] = {f->n[2]};
            long long er[1] = {2 * f->nc}, ec[1] = {f->nc};
            PMB_CHECK(make_plan(f, &f->pen_r2c, 1, n1, er, 2 * f->nc, ec, f->nc, dbl ? CUFFT_D2Z : CUFFT_R2C, f->m0 * f->r1m));
            PMB_CHECK(make_plan(f, &f->pen_c2r, 1, n1, ec, f->nc, er, 2 * f->nc, dbl ? CUFFT_Z2D : CUFFT_C2R, f->m0 * f->r1m));
        }
        if (f->m0 * f->mc > 0) {
            long long n1[1] = {f->n[1]};
            PMB_CHECK(make_plan(f, &f->pen_l1, 1, n1, n1, f->n[1], n1, f->n[1], dbl ? CUFFT_Z2Z : CUFFT_C2C, f->m0 * f->mc));
        }
        if (f->m1 * f->mc > 0) {
            long long n1[1] = {f->n[0]};
            PMB_CHECK(make_plan(f, &f->pen_l0, 1, n1, n1, f->n[0], n1, f->n[0], dbl ? CUFFT_Z2Z : CUFFT_C2C, f->m1 * f->mc));
        }
        f->have_pen = true;
        int64_t celems = f->m0 * f->r1m * f->nc;
        if (f->m0 * f->mc * f->n[1] > celems) celems = f->m0 * f->mc * f->n[1];
        if (f->m1 * f->mc * f->n[0] > celems) celems = f->m1 * f->mc * f->n[0];
        // every rank allocates the largest block of any rank: peers store into my landing buffers with THEIR strides
        {
            int64_t mx[4] = {(f->n[0] + f->P0 - 1) / f->P0, (f->n[1] + f->P1 - 1) / f->P1, (f->n[1] + f->P0 - 1) / f->P0, (f->nc + f->P1 - 1) / f->P1};
            int64_t a = mx[0] * mx[1] * f->nc, b = mx[0] * mx[3] * f->n[1], c = mx[2] * mx[3] * f->n[0];
            celems = a > b ? a : b;
            if (c > celems) celems = c;
        }
        f->work_bytes = (size_t) celems * 2 * dtype_elsize + 256;
        PMB_CUDA(cudaMalloc(&f->work0, f->work_bytes));
        PMB_CHECK(p2p_setup(f));
        if (!f->p2p) {
            pmb_set_error("pencil (2-D process mesh) transforms need CUDA IPC peer access between the GPUs of the node; "
                          "it is not available here (or PMB_FFT_P2P=0): use np=[P] slabs");
            pmb_fft_destroy(f);
            return PMB_EUNSUPPORTED;
        }
    } else {
        block_partition(f->n[0], f->P, f->rank, &f->blk0, &f->s0, &f->m0);
        block_partition(f->n[1], f->P, f->rank, &f->blk1, &f->s1, &f->m1);
        if (f->m0 > 0) {
            long long n2[2] = {f->n[1], f->n[2]};
            long long inr[2] = {f->n[1], 2 * f->nc};
            long long inc[2] = {f->n[1], f->nc};
            PMB_CHECK(make_plan(f, &f->slab_r2c, 2, n2, inr, f->n[1] * 2 * f->nc, inc, f->n[1] * f->nc,
                                dbl ? CUFFT_D2Z : CUFFT_R2C, f->m0));
            PMB_CHECK(make_plan(f, &f->slab_c2r, 2, n2, inc, f->n[1] * f->nc, inr, f->n[1] * 2 * f->nc,
                                dbl ? CUFFT_Z2D : CUFFT_C2R, f->m0));
            // forward transform in plane chunks: the NVLink transpose of chunk k runs under the 2-D FFT of chunk k + 1.
            // Measured at 2 GPUs, 1024^3 (r2c, ms): whole 8.55; 2 chunks 8.09; 4 chunks 10.3 -- the 2-D cuFFT of a
            // quarter of the planes and a transpose held to 4 CTAs per SM (444 GB/s instead of 589) lose what the overlap
            // gains.  Off by default (PMB_FFT_CHUNKS=2 selects it).
            int chunks = 0;
            { const char *e = getenv("PMB_FFT_CHUNKS"); if (e) chunks = atoi(e); }
            if (chunks > 4) chunks = 4;
            f->r2c_chunks = 0;
            if (f->P > 1 && chunks >= 2 && f->m0 % chunks == 0 && f->m0 / chunks >= 8) {
                PMB_CHECK(make_plan(f, &f->slab_r2c_part, 2, n2, inr, f->n[1] * 2 * f->nc, inc, f->n[1] * f->nc,
                                    dbl ? CUFFT_D2Z : CUFFT_R2C, f->m0 / chunks));
                f->r2c_chunks = chunks;
            }
        }
        if (f->m1 > 0) {
            long long n1[1] = {f->n[0]};
            long long emb[1] = {f->n[0]};
            PMB_CHECK(make_plan(f, &f->line, 1, n1, emb, f->n[0], emb, f->n[0], dbl ? CUFFT_Z2Z : CUFFT_C2C, f->m1 * f->nc));
        }
        f->have_slab = true;
        int64_t celems = f->m0 * f->n[1] * f->nc;
        int64_t t = f->m1 * f->nc * f->n[0];
        if (t > celems) celems = t;
        f->work_bytes = (size_t) celems * 2 * dtype_elsize + 256;
        PMB_CUDA(cudaMalloc(&f->work0, f->work_bytes));
        PMB_CUDA(cudaMalloc(&f->work1, f->work_bytes));
        PMB_CHECK(p2p_setup(f));
    }
    *out = f;
    return PMB_OK;
}

extern "C" int pmb_fft_destroy(pmb_fft *f)
{
    if (!f) return PMB_OK;
    cudaSetDevice(f->ctx->device);
    cudaStreamSynchronize(f->ctx->stream);
    if (f->have_full) { cufftDestroy(f->full_r2c); cufftDestroy(f->full_c2r); }
    if (f->have_slab) {
        if (f->m0 > 0) { cufftDestroy(f->slab_r2c); cufftDestroy(f->slab_c2r); if (f->r2c_chunks) cufftDestroy(f->slab_r2c_part); }
        if (f->m1 > 0) cufftDestroy(f->line);
    }
    if (f->have_pen) {
        if (f->m0 * f->r1m > 0) { cufftDestroy(f->pen_r2c); cufftDestroy(f->pen_c2r); }
        if (f->m0 * f->mc > 0) cufftDestroy(f->pen_l1);
        if (f->m1 * f->mc > 0) cufftDestroy(f->pen_l0);
    }
    p2p_teardown(f);
    for (int i = 0; i < f->ntfc; i++) cudaFree(f->tfc[i].dev);
    if (f->ifft_tw) cudaFree(f->ifft_tw);
    if (f->have_plane) cufftDestroy(f->plane_c2r);
    if (f->have_plane_r2c) cufftDestroy(f->plane_r2c);
    if (f->have_xstream) {
        cudaStreamSynchronize(f->xstream);
        cudaStreamDestroy(f->xstream);
        for (int i = 0; i < 16; i++) cudaEventDestroy(f->xev[i]);
    }
    if (f->work0) cudaFree(f->work0);
    if (f->work1) cudaFree(f->work1);
    if (f->work2) cudaFree(f->work2);
    for (int i = 0; i < FFT_NEV; i++) { cudaEventDestroy(f->ev[i][0]); cudaEventDestroy(f->ev[i][1]); }
    free(f);
    return PMB_OK;
}

extern "C" int pmb_fft_layout(pmb_fft *f, int64_t *i_start, int64_t *i_shape, int64_t *i_strides,
                              int64_t *o_start, int64_t *o_shape, int64_t *o_strides,
                              int64_t *real_alloc_elems, int64_t *complex_alloc_elems)
{
    PMB_REQUIRE(f && i_start && i_shape && i_strides && o_start && o_shape && o_strides, "null argument");
    const int nd = f->ndim, pad = 3 - nd;
    int64_t is[3] = {f->s0, f->r1s, 0}, ish[3] = {f->m0, f->r1m, f->n[2]};
    int64_t ist[3] = {f->r1m * 2 * f->nc, 2 * f->nc, 1};
    int64_t os[3], osh[3], ost[3];
    if (f->P == 1) {
        os[0] = os[1] = os[2] = 0;
        osh[0] = f->n[0]; osh[1] = f->n[1]; osh[2] = f->nc;
        ost[0] = f->n[1] * f->nc; ost[1] = f->nc; ost[2] = 1;
    } else {
        os[0] = 0; os[1] = f->s1; os[2] = f->s2;
        osh[0] = f->n[0]; osh[1] = f->m1; osh[2] = f->mc;
        ost[0] = 1; ost[1] = f->mc * f->n[0]; ost[2] = f->n[0];
    }
    for (int d = 0; d < nd; d++) {
        i_start[d] = is[pad + d]; i_shape[d] = ish[pad + d]; i_strides[d] = ist[pad + d];
        o_start[d] = os[pad + d]; o_shape[d] = osh[pad + d]; o_strides[d] = ost[pad + d];
    }
    int64_t relems = f->m0 * f->r1m * 2 * f->nc;
    int64_t celems = f->P == 1 ? f->n[0] * f->n[1] * f->nc : f->m1 * f->mc * f->n[0];
    // either buffer may serve as the in-place partner of the other
    int64_t both = relems > 2 * celems ? relems : 2 * celems;
    if (real_alloc_elems) *real_alloc_elems = both > 0 ? both : 2;
    if (complex_alloc_elems) *complex_alloc_elems = both / 2 > 0 ? both / 2 : 1;
    return PMB_OK;
}

// ---- cuFFT time accounting ------------------------------------------------------
static int lib_flush(pmb_fft *f)
{
    if (f->nev == 0) return PMB_OK;
    for (int i = 0; i < f->nev; i++) {
        float ms = 0;
        PMB_CUDA(cudaEventSynchronize(f->ev[i][1]));       // the brackets sit on two streams
        PMB_CUDA(cudaEventElapsedTime(&ms, f->ev[i][0], f->ev[i][1]));
        if (f->ev_kind[i] == 0) f->lib_ms += ms; else if (f->ev_kind[i] == 1) f->xpose_ms += ms; else f->ifft_ms += ms;
    }
    f->nev = 0;
    return PMB_OK;
}
static int lib_begin(pmb_fft *f, int kind = 0, cudaStream_t stream = 0)
{
    if (f->nev == FFT_NEV) PMB_CHECK(lib_flush(f));
    f->ev_kind[f->nev] = kind;
    PMB_CUDA(cudaEventRecord(f->ev[f->nev][0], stream ? stream : f->ctx->stream));
    return PMB_OK;
}
static int lib_end(pmb_fft *f, cudaStream_t stream = 0)
{
    PMB_CUDA(cudaEventRecord(f->ev[f->nev][1], stream ? stream : f->ctx->stream));
    f->nev++;
    return PMB_OK;
}
// time spent in the transpose kernels and the bytes they stored over NVLink since the last reset:
// achieved link bandwidth of this rank = remote_bytes / ms
extern "C" int pmb_fft_transpose_stats(pmb_fft *f, float *ms, double *remote_bytes, int reset)
{
    PMB_REQUIRE(f && ms && remote_bytes, "null argument");
    PMB_CHECK(lib_flush(f));
    *ms = f->xpose_ms;
    *remote_bytes = f->xpose_remote_bytes;
    if (reset) { f->xpose_ms = 0; f->xpose_remote_bytes = 0; }
    return PMB_OK;
}

extern "C" int pmb_fft_library_ms(pmb_fft *f, float *ms, int reset)
{
    PMB_REQUIRE(f && ms, "null argument");
    PMB_CHECK(lib_flush(f));
    *ms = f->lib_ms;
    if (reset) f->lib_ms = 0;
    return PMB_OK;
}

// ---- kernels ------------------------------------------------------------------
template <typename C>   // C = float2 / double2
__global__ void pmb_k_cscale(C *a, int64_t n, double s)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        C v = a[i];
        v.x = (decltype(v.x)) (v.x * (decltype(v.x)) s);
        v.y = (decltype(v.y)) (v.y * (decltype(v.y)) s);
        a[i] = v;
    }
}

// out (C rows, R cols) = transpose(in (R rows, C cols)) * s ; 32x32 tiles through shared memory so
// that both the global reads and writes are coalesced 512-byte (double2) rows.
template <typename C>
__global__ void __launch_bounds__(256)
pmb_k_transpose(const C *__restrict__ in, C *__restrict__ out, int64_t R, int64_t Cc, double s,
                int64_t tiles_r, int64_t tiles_c)
{
    __shared__ C tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int64_t ntiles = tiles_r * tiles_c;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t tr = t / tiles_c, tc = t - tr * tiles_c;
        const int64_t r0 = tr * 32, c0 = tc * 32;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int64_t r = r0 + ty + 8 * k, c = c0 + tx;
            if (r < R && c < Cc) tile[ty + 8 * k][tx] = in[r * Cc + c];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int64_t c = c0 + ty + 8 * k, r = r0 + tx;
            if (r < R && c < Cc) {
                C v = tile[tx][ty + 8 * k];
                v.x = (decltype(v.x)) (v.x * (decltype(v.x)) s);
                v.y = (decltype(v.y)) (v.y * (decltype(v.y)) s);
                out[c * R + r] = v;
            }
        }
        __syncthreads();
    }
}

// forward pack: A (m0, n1, nc) -> per-destination blocks (m0, m1_q, nc) at offset m0 * s1_q * nc.
// backward unpack is the inverse copy (dir = 1).
template <typename C>
__global__ void pmb_k_slab_pack(const C *__restrict__ src, C *__restrict__ dst, int64_t m0, int64_t n1, int64_t nc,
                                int64_t blk1, int dir)
{
    int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int64_t total = m0 * n1 * nc;
    for (; t < total; t += stride) {
        const int64_t k = t % nc;
        const int64_t r = t / nc;
        const int64_t j = r % n1;
        const int64_t i = r / n1;
        const int64_t q = j / blk1;
        const int64_t s1q = q * blk1;
        int64_t m1q = n1 - s1q;
        if (m1q > blk1) m1q = blk1;
        const int64_t packed = m0 * s1q * nc + (i * m1q + (j - s1q)) * nc + k;
        if (dir == 0) dst[packed] = src[t]; else dst[t] = src[packed];
    }
}

template <typename C>
static int transpose(pmb_fft *f, const void *in, void *out, int64_t R, int64_t Cc, double s)
{
    if (R == 0 || Cc == 0) return PMB_OK;
    const int64_t tr = (R + 31) / 32, tc = (Cc + 31) / 32;
    int64_t grid = tr * tc;
    const int64_t cap = (int64_t) f->ctx->sm_count * 8;
    if (grid > cap) grid = cap;
    pmb_k_transpose<C><<<(int) grid, 256, 0, f->ctx->stream>>>((const C *) in, (C *) out, R, Cc, s, tr, tc);
    PMB_LAUNCH_CHECK(f->ctx);
    return PMB_OK;
}

template <typename C>
static int slab_pack(pmb_fft *f, const void *src, void *dst, int dir)
{
    const int64_t total = f->m0 * f->n[1] * f->nc;
    if (total == 0) return PMB_OK;
    pmb_k_slab_pack<C><<<pmb_grid(f->ctx, total, 256, 8), 256, 0, f->ctx->stream>>>(
        (const C *) src, (C *) dst, f->m0, f->n[1], f->nc, f->blk1, dir);
    PMB_LAUNCH_CHECK(f->ctx);
    return PMB_OK;
}

// Global transpose fused with the NVLink transfer: for every destination rank q the (R x ncols[q])
// block  in[r * in_ld + col0[q] + c]  is written transposed,  out[q][c * out_ld + r] (* s), straight
// into the landing buffer of rank q (peer memory; q == me is the local buffer).  32 x 32 tiles through
// shared memory: 512-byte coalesced reads locally, 512-byte contiguous stores over NVLink.  Tiles are
// dealt round-robin over the destinations, starting at my right neighbour, so that at any moment the
// traffic is spread over all links / peers.
struct XDest {
    void *out[64];        // landing buffer of destination q, already offset to my first output column
    int64_t col0[64];     // first source column of the block for destination q
    int64_t ncols[64];    // columns of that block
    int64_t obstride[64]; // elements between consecutive batches in the landing buffer of q
};

// The same kernel serves the pencil transposes: `nbatch` independent (R x columns) matrices, `in_bs`
// elements apart in the source and d.obstride[q] apart in the landing buffer of destination q; P is then
// the number of destinations (the ranks of my row or column of the process mesh) and `me` my index in it.
template <typename C>
__global__ void __launch_bounds__(256)
pmb_k_xpose_scatter(const C *__restrict__ in, int64_t in_ld, int64_t R, int64_t out_ld, XDest d, int P, int me,
                    double s, int64_t tiles_r, int64_t tiles_c_max, int64_t nbatch, int64_t in_bs)
{
    __shared__ C tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int64_t per_batch = tiles_r * tiles_c_max;
    const int64_t per_dest = per_batch * nbatch;
    const int64_t ntiles = per_dest * P;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int q = (int) (t % P) + me + 1;
        if (q >= P) q -= P;
        int64_t lt = t / P;
        const int64_t bt = lt / per_batch;
        lt -= bt * per_batch;
        const int64_t tc = lt / tiles_r, tr = lt - tc * tiles_r;
        const int64_t r0 = tr * 32, c0 = tc * 32;
        const int64_t nc = d.ncols[q];
        if (c0 >= nc) continue;          // uniform over the block
        const C *src = in + bt * in_bs + d.col0[q];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int64_t r = r0 + ty + 8 * k, c = c0 + tx;
            if (r < R && c < nc) tile[ty + 8 * k][tx] = src[r * in_ld + c];
        }
        __syncthreads();
        C *dst = (C *) d.out[q] + bt * d.obstride[q];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int64_t c = c0 + ty + 8 * k, r = r0 + tx;
            if (r < R && c < nc) {
                C v = tile[tx][ty + 8 * k];
                v.x = (decltype(v.x)) (v.x * (decltype(v.x)) s);
                v.y = (decltype(v.y)) (v.y * (decltype(v.y)) s);
                dst[c * out_ld + r] = v;
            }
        }
        __syncthreads();
    }
}

template <typename C>
static int xpose_scatter(pmb_fft *f, const void *in, int64_t in_ld, int64_t R, int64_t out_ld, const XDest &d, double s,
                         int ndest = -1, int me = -1, int64_t nbatch = 1, int64_t in_bs = 0, cudaStream_t stream = 0)
{
    const bool own_stream = stream != 0;        // a caller that runs the transposes on its own stream (c2r_multi)
    if (!own_stream) stream = f->ctx->stream;
    if (ndest < 0) { ndest = f->P; me = f->rank; }
    int64_t cmax = 0;
    for (int q = 0; q < ndest; q++) if (d.ncols[q] > cmax) cmax = d.ncols[q];
    if (R == 0 || cmax == 0 || nbatch == 0) return PMB_OK;
    const int64_t tr = (R + 31) / 32, tc = (cmax + 31) / 32;
    int64_t grid = tr * tc * ndest * nbatch;
    // on its own (high-priority) stream the transpose shares the SMs with the cuFFT kernels of the other transforms:
    // a few resident CTAs per SM keep the link busy (the kernel waits on NVLink stores) without starving them.
    // Measured at 2 GPUs, 1024^3, the three c2r of a step (ms): one by one 26.4; overlapped with 1 / 2 / 4 CTAs per
    // SM 33.0 / 24.6 / 23.0; 8 CTAs per SM at normal priority 39.6 (profiles/README.md)
    static int octas = -1;
    if (octas < 0) { const char *e = getenv("PMB_FFT_OVERLAP_CTAS"); octas = e ? atoi(e) : 4; if (octas < 1) octas = 1; }
    const int64_t cap = (int64_t) f->ctx->sm_count * (own_stream ? octas : 8);
    if (grid > cap) grid = cap;
    PMB_CHECK(lib_begin(f, 1, stream));       // events on the stream the kernel runs on (under other kernels when it is the transpose stream)
    pmb_k_xpose_scatter<C><<<(int) grid, 256, 0, stream>>>((const C *) in, in_ld, R, out_ld, d, ndest, me, s, tr, tc, nbatch, in_bs);
    PMB_LAUNCH_CHECK(f->ctx);
    PMB_CHECK(lib_end(f, stream));
    for (int q = 0; q < ndest; q++)
        if (q != me) f->xpose_remote_bytes += (double) R * (double) d.ncols[q] * (double) nbatch * (double) sizeof(C);
    return PMB_OK;
}
template <typename C>
static int xpose_scatter_e(pmb_fft *f, const void *in, int64_t in_ld, int64_t R, int64_t out_ld, const XDest &d, double s,
                           int ndest, int me, int64_t nbatch, int64_t in_bs)
{
    return xpose_scatter<C>(f, in, in_ld, R, out_ld, d, s, ndest, me, nbatch, in_bs);
}

// forward: work0 (m0, n1*nc) -> rank q gets the columns of its j-range, transposed, as rows
// (m1_q*nc) x n0 with my planes at columns [s0, s0 + m0)
static void xdest_fwd(const pmb_fft *f, int b, XDest *d)
{
    const size_t csz = 2 * (size_t) f->elsize;
    for (int q = 0; q < f->P; q++) {
        int64_t blk, s1q, m1q;
        block_partition(f->n[1], f->P, q, &blk, &s1q, &m1q);
        d->out[q] = (char *) f->peer_x[q][b] + (size_t) f->s0 * csz;
        d->col0[q] = s1q * f->nc;
        d->ncols[q] = m1q * f->nc;
        d->obstride[q] = 0;
    }
}
// backward: work0 (m1*nc, n0) -> rank p gets the columns of its plane range, transposed, as rows
// m0_p x (n1*nc) with my (j, k) lines at columns [s1*nc, (s1 + m1)*nc)
static void xdest_bwd(const pmb_fft *f, int b, XDest *d)
{
    const size_t csz = 2 * (size_t) f->elsize;
    for (int p = 0; p < f->P; p++) {
        int64_t blk, s0p, m0p;
        block_partition(f->n[0], f->P, p, &blk, &s0p, &m0p);
        d->out[p] = (char *) f->peer_x[p][b] + (size_t) (f->s1 * f->nc) * csz;
        d->col0[p] = s0p;
        d->ncols[p] = m0p;
        d->obstride[p] = 0;
    }
}

static int exec_r2c(pmb_fft *f, cufftHandle h, const void *in, void *out);
static int exec_c2r(pmb_fft *f, cufftHandle h, void *in, void *out);
static int exec_c2c(pmb_fft *f, cufftHandle h, void *in, void *out, int dir);

// ---- pencil transposes (P1 > 1) -----------------------------------------------------------------------
// Landing buffers: X0 = xbuf[0] holds (m0, mc, n1) -- axis 1 contiguous, X1 = xbuf[1] holds (m1, mc, n0) --
// axis 0 contiguous -- on the way forward, and (m0, r1m, nc) -- real-space planes -- on the way back.
// Forward:  A r2c along 2 -> work0 (m0, r1m, nc)
//           B scatter inside my ROW of the process mesh (fixed c0): rank (c0, c1') gets k in its block,
//             stored as X0[(i * mc' + kl) * n1 + r1s + jl]                         -> barrier
//           C c2c along 1, in place in X0
//           D scatter inside my COLUMN (fixed c1): rank (c0', c1) gets j in its block,
//             stored as X1[((j - s1') * mc + kl) * n0 + s0 + il]  (x scale)          -> barrier
//           E c2c along 0: X1 -> result (m1, mc, n0)
// Backward: the mirror image (E' into work0, D' into X0, C' in place, B' into X1, A' c2r X1 -> real).
// X0 is last read before the second barrier of a transform and X1 after it, and a rank can only store into
// a peer's X0 (X1) after the second barrier of the previous transform (the first barrier of this one), so
// the two buffers never need to alternate.
static void pen_dest_row(const pmb_fft *f, int b, int back, XDest *d)
{
    const size_t csz = 2 * (size_t) f->elsize;
    for (int q = 0; q < f->P1; q++) {
        const int rank = f->c0 * f->P1 + q;
        int64_t blk, s, m;
        if (!back) {
            // forward B: columns k of destination q's block; my rows j land at column offset r1s
            block_partition(f->nc, f->P1, q, &blk, &s, &m);
            d->out[q] = (char *) f->peer_x[rank][b] + (size_t) f->r1s * csz;
            d->col0[q] = s; d->ncols[q] = m; d->obstride[q] = m * f->n[1];
        } else {
            // backward B': columns j of destination q's real-space block; my rows kl land at column offset s2
            block_partition(f->n[1], f->P1, q, &blk, &s, &m);
            d->out[q] = (char *) f->peer_x[rank][b] + (size_t) f->s2 * csz;
            d->col0[q] = s; d->ncols[q] = m; d->obstride[q] = m * f->nc;
        }
    }
}
static void pen_dest_col(const pmb_fft *f, int b, int back, XDest *d)
{
    const size_t csz = 2 * (size_t) f->elsize;
    for (int q = 0; q < f->P0; q++) {
        const int rank = q * f->P1 + f->c1;
        int64_t blk, s, m;
        if (!back) {
            // forward D: columns j of destination q's complex block; my rows il land at column offset s0
            block_partition(f->n[1], f->P0, q, &blk, &s, &m);
            d->out[q] = (char *) f->peer_x[rank][b] + (size_t) f->s0 * csz;
            d->col0[q] = s; d->ncols[q] = m; d->obstride[q] = f->n[0];
        } else {
            // backward D': columns i of destination q's plane block; my rows jl land at column offset s1
            block_partition(f->n[0], f->P0, q, &blk, &s, &m);
            d->out[q] = (char *) f->peer_x[rank][b] + (size_t) f->s1 * csz;
            d->col0[q] = s; d->ncols[q] = m; d->obstride[q] = f->n[1];
        }
    }
}

template <typename C>
static int pencil_r2c(pmb_fft *f, const void *real, void *cplx, double scale)
{
    pmb_ctx *ctx = f->ctx;
    XDest d;
    // A
    if (f->m0 * f->r1m > 0) PMB_CHECK(exec_r2c(f, f->pen_r2c, real, f->work0));
    // B: per plane i a (r1m x nc) matrix; destination q takes columns [s2_q, s2_q + mc_q)
    pen_dest_row(f, 0, 0, &d);
    PMB_CHECK(xpose_scatter_e<C>(f, f->work0, f->nc, f->r1m, f->n[1], d, 1.0, f->P1, f->c1, f->m0, f->r1m * f->nc));
    PMB_CHECK(pmb_stream_barrier(ctx));
    // C
    if (f->m0 * f->mc > 0) PMB_CHECK(exec_c2c(f, f->pen_l1, f->xbuf[0], f->xbuf[0], CUFFT_FORWARD));
    // D: per kl a (m0 x n1) matrix with row stride mc * n1; destination q takes columns [s1_q, s1_q + m1_q)
    pen_dest_col(f, 1, 0, &d);
    PMB_CHECK(xpose_scatter_e<C>(f, f->xbuf[0], f->mc * f->n[1], f->m0, f->mc * f->n[0], d, scale, f->P0, f->c0, f->mc, f->n[1]));
    PMB_CHECK(pmb_stream_barrier(ctx));
    // E
    if (f->m1 * f->mc > 0) PMB_CHECK(exec_c2c(f, f->pen_l0, f->xbuf[1], cplx, CUFFT_FORWARD));
    return PMB_OK;
}

template <typename C>
static int pencil_c2r(pmb_fft *f, const void *cplx, void *real)
{
    pmb_ctx *ctx = f->ctx;
    XDest d;
    // E': (m1, mc, n0) lines along 0 -> work0 (the input is preserved)
    if (f->m1 * f->mc > 0) PMB_CHECK(exec_c2c(f, f->pen_l0, (void *) cplx, f->work0, CUFFT_INVERSE));
    // D': per kl a (m1 x n0) matrix with row stride mc * n0; destination q takes columns [s0_q, s0_q + m0_q)
    pen_dest_col(f, 0, 1, &d);
    PMB_CHECK(xpose_scatter_e<C>(f, f->work0, f->mc * f->n[0], f->m1, f->mc * f->n[1], d, 1.0, f->P0, f->c0, f->mc, f->n[0]));
    PMB_CHECK(pmb_stream_barrier(ctx));
    // C'
    if (f->m0 * f->mc > 0) PMB_CHECK(exec_c2c(f, f->pen_l1, f->xbuf[0], f->xbuf[0], CUFFT_INVERSE));
    // B': per plane i a (mc x n1) matrix; destination q takes columns [r1s_q, r1s_q + r1m_q)
    pen_dest_row(f, 1, 1, &d);
    PMB_CHECK(xpose_scatter_e<C>(f, f->xbuf[0], f->n[1], f->mc, f->nc, d, 1.0, f->P1, f->c1, f->m0, f->mc * f->n[1]));
    PMB_CHECK(pmb_stream_barrier(ctx));
    // A'
    if (f->m0 * f->r1m > 0) PMB_CHECK(exec_c2r(f, f->pen_c2r, f->xbuf[1], real));
    return PMB_OK;
}

// counts / offsets (in complex elements) of the global transpose.
// fwd: send to q the block (m0, m1_q, nc); receive from p the block (m0_p, m1, nc).
static void slab_counts(const pmb_fft *f, int fwd, int64_t *sc, int64_t *so, int64_t *rc, int64_t *ro)
{
    for (int q = 0; q < f->P; q++) {
        int64_t b, s0q, m0q, s1q, m1q;
        block_partition(f->n[0], f->P, q, &b, &s0q, &m0q);
        block_partition(f->n[1], f->P, q, &b, &s1q, &m1q);
        const int64_t a_cnt = f->m0 * m1q * f->nc, a_off = f->m0 * s1q * f->nc;   // (m0, m1_q, nc) blocks
        const int64_t b_cnt = m0q * f->m1 * f->nc, b_off = s0q * f->m1 * f->nc;   // (m0_q, m1, nc) blocks
        if (fwd) { sc[q] = a_cnt; so[q] = a_off; rc[q] = b_cnt; ro[q] = b_off; }
        else { sc[q] = b_cnt; so[q] = b_off; rc[q] = a_cnt; ro[q] = a_off; }
    }
}

static int exec_r2c(pmb_fft *f, cufftHandle h, const void *in, void *out)
{
    PMB_CHECK(lib_begin(f));
    if (f->elsize == 8) PMB_CUFFT(cufftExecD2Z(h, (cufftDoubleReal *) in, (cufftDoubleComplex *) out));
    else PMB_CUFFT(cufftExecR2C(h, (cufftReal *) in, (cufftComplex *) out));
    PMB_CHECK(lib_end(f));
    return PMB_OK;
}
static int exec_c2r(pmb_fft *f, cufftHandle h, void *in, void *out)
{
    PMB_CHECK(lib_begin(f));
    if (f->elsize == 8) PMB_CUFFT(cufftExecZ2D(h, (cufftDoubleComplex *) in, (cufftDoubleReal *) out));
    else PMB_CUFFT(cufftExecC2R(h, (cufftComplex *) in, (cufftReal *) out));
    PMB_CHECK(lib_end(f));
    return PMB_OK;
}
static int exec_c2c(pmb_fft *f, cufftHandle h, void *in, void *out, int dir)
{
    PMB_CHECK(lib_begin(f));
    if (f->elsize == 8) PMB_CUFFT(cufftExecZ2Z(h, (cufftDoubleComplex *) in, (cufftDoubleComplex *) out, dir));
    else PMB_CUFFT(cufftExecC2C(h, (cufftComplex *) in, (cufftComplex *) out, dir));
    PMB_CHECK(lib_end(f));
    return PMB_OK;
}

// second, high-priority stream of the plan: transposes that run under cuFFT kernels (pipelined r2c, c2r_multi)
static int ensure_xstream(pmb_fft *f)
{
    if (f->have_xstream) return PMB_OK;
    int prio_lo = 0, prio_hi = 0;
    PMB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    PMB_CUDA(cudaStreamCreateWithPriority(&f->xstream, cudaStreamNonBlocking, prio_hi));
    for (int i = 0; i < 16; i++) PMB_CUDA(cudaEventCreateWithFlags(&f->xev[i], cudaEventDisableTiming));
    f->have_xstream = true;
    return PMB_OK;
}

extern "C" int pmb_fft_r2c(pmb_fft *f, const void *real, void *cplx, double scale)
{
    PMB_REQUIRE(f && real && cplx, "null argument");
    pmb_ctx *ctx = f->ctx;
    const size_t csz = 2 * (size_t) f->elsize;
    if (f->P == 1) {
        PMB_CHECK(ensure_full(f));
        PMB_CHECK(exec_r2c(f, f->full_r2c, real, cplx));
        if (scale != 1.0) {
            const int64_t n = f->n[0] * f->n[1] * f->nc;
            if (f->elsize == 8) pmb_k_cscale<double2><<<pmb_grid(ctx, n, 256, 8), 256, 0, ctx->stream>>>((double2 *) cplx, n, scale);
            else pmb_k_cscale<float2><<<pmb_grid(ctx, n, 256, 8), 256, 0, ctx->stream>>>((float2 *) cplx, n, scale);
            PMB_LAUNCH_CHECK(ctx);
        }
        return PMB_OK;
    }
    if (f->P1 > 1) return f->elsize == 8 ? pencil_r2c<double2>(f, real, cplx, scale) : pencil_r2c<float2>(f, real, cplx, scale);
    static int overlap = -1;
    if (overlap < 0) { const char *e = getenv("PMB_FFT_OVERLAP"); overlap = e ? atoi(e) : 1; }
    if (f->p2p && f->r2c_chunks && overlap) {
        // Pipelined: the planes go through the 2-D r2c in r2c_chunks batches on the compute stream; the transpose of
        // every batch (NVLink stores into the owners' landing buffers) runs on the transpose stream under the FFT of
        // the next batch.  One barrier after the last batch, issued on the transpose stream; the compute stream
        // waits for it before the lines along axis 0.  Entry: the transpose stream starts after everything the
        // compute stream has queued so far (the reads of this landing buffer two transforms ago included).
        PMB_CHECK(ensure_xstream(f));
        cudaStream_t ms = ctx->stream, xs = f->xstream;
        const int b = f->xcur;
        f->xcur ^= 1;
        const int K = f->r2c_chunks;
        const int64_t Rc = f->m0 / K;
        PMB_CUDA(cudaEventRecord(f->xev[12], ms));
        PMB_CUDA(cudaStreamWaitEvent(xs, f->xev[12], 0));
        for (int k = 0; k < K; k++) {
            const char *in = (const char *) real + (size_t) k * Rc * f->n[1] * 2 * f->nc * f->elsize;
            char *wk = (char *) f->work0 + (size_t) k * Rc * f->n[1] * f->nc * csz;
            PMB_CHECK(exec_r2c(f, f->slab_r2c_part, in, wk));
            PMB_CUDA(cudaEventRecord(f->xev[k], ms));
            PMB_CUDA(cudaStreamWaitEvent(xs, f->xev[k], 0));
            XDest d;
            xdest_fwd(f, b, &d);
            for (int q = 0; q < f->P; q++) d.out[q] = (char *) d.out[q] + (size_t) k * Rc * csz;     // my planes [k Rc, (k + 1) Rc)
            if (f->elsize == 8) PMB_CHECK(xpose_scatter<double2>(f, wk, f->n[1] * f->nc, Rc, f->n[0], d, scale, -1, -1, 1, 0, xs));
            else PMB_CHECK(xpose_scatter<float2>(f, wk, f->n[1] * f->nc, Rc, f->n[0], d, scale, -1, -1, 1, 0, xs));
        }
        PMB_CHECK(pmb_stream_barrier_on(ctx, xs));
        PMB_CUDA(cudaEventRecord(f->xev[13], xs));
        PMB_CUDA(cudaStreamWaitEvent(ms, f->xev[13], 0));
        if (f->m1 > 0) PMB_CHECK(exec_c2c(f, f->line, f->xbuf[b], cplx, CUFFT_FORWARD));
        return PMB_OK;
    }
    // 1. local planes: 2-D r2c over (n1, n2):  real (m0, n1, 2nc) -> work0 (m0, n1, nc)
    if (f->m0 > 0) PMB_CHECK(exec_r2c(f, f->slab_r2c, real, f->work0));
    if (f->p2p) {
        // 2. transpose straight into the landing buffers of the owners of each j-range (NVLink
        //    stores; normalisation folded in), 3. in-stream barrier, 4. lines along axis 0 out of
        //    the landing buffer into the result.  Landing buffers alternate between transforms, so
        //    a rank that runs ahead never overwrites a buffer a peer is still reading: it cannot
        //    start transform t + 2 before everybody has passed the barrier of transform t + 1.
        const int b = f->xcur;
        f->xcur ^= 1;
        XDest d;
        xdest_fwd(f, b, &d);
        if (f->elsize == 8) PMB_CHECK(xpose_scatter<double2>(f, f->work0, f->n[1] * f->nc, f->m0, f->n[0], d, scale));
        else PMB_CHECK(xpose_scatter<float2>(f, f->work0, f->n[1] * f->nc, f->m0, f->n[0], d, scale));
        PMB_CHECK(pmb_stream_barrier(ctx));
        if (f->m1 > 0) PMB_CHECK(exec_c2c(f, f->line, f->xbuf[b], cplx, CUFFT_FORWARD));
        return PMB_OK;
    }
    // 2. pack per destination
    if (f->elsize == 8) PMB_CHECK(slab_pack<double2>(f, f->work0, f->work1, 0));
    else PMB_CHECK(slab_pack<float2>(f, f->work0, f->work1, 0));
    // 3. global transpose
    int64_t sc[64], so[64], rc[64], ro[64];
    PMB_REQUIRE(f->P <= 64, "at most 64 ranks");
    slab_counts(f, 1, sc, so, rc, ro);
    PMB_CHECK(pmb_alltoallv(ctx, f->work1, sc, so, f->work0, rc, ro, (int64_t) csz));
    // 4. work0 is now (n0, m1*nc): transpose to (m1*nc, n0) with the normalisation folded in
    if (f->elsize == 8) PMB_CHECK(transpose<double2>(f, f->work0, cplx, f->n[0], f->m1 * f->nc, scale));
    else PMB_CHECK(transpose<float2>(f, f->work0, cplx, f->n[0], f->m1 * f->nc, scale));
    // 5. lines along axis 0
    if (f->m1 > 0) PMB_CHECK(exec_c2c(f, f->line, cplx, cplx, CUFFT_FORWARD));
    return PMB_OK;
}

static int c2r_from_work(pmb_fft *f, void *real);
static int slab_c2r_after_lines(pmb_fft *f, void *real);
struct PmbIfftArgs;
static int ifft_launch(pmb_fft *f, const PmbIfftArgs &a, cudaStream_t stream);

extern "C" int pmb_fft_c2r(pmb_fft *f, const void *cplx, void *real)
{
    PMB_REQUIRE(f && real && cplx, "null argument");
    pmb_ctx *ctx = f->ctx;
    const size_t csz = 2 * (size_t) f->elsize;
    if (f->P == 1) {
        // cuFFT's multi-dimensional C2R may overwrite its input: go through the output buffer
        if (cplx != real)
            PMB_CUDA(cudaMemcpyAsync(real, cplx, (size_t) (f->n[0] * f->n[1] * f->nc) * csz, cudaMemcpyDeviceToDevice, ctx->stream));
        PMB_CHECK(ensure_full(f));
        PMB_CHECK(exec_c2r(f, f->full_c2r, real, real));
        return PMB_OK;
    }
    if (f->P1 > 1) return f->elsize == 8 ? pencil_c2r<double2>(f, cplx, real) : pencil_c2r<float2>(f, cplx, real);
    // 1. inverse lines along axis 0: cplx (m1*nc, n0) -> work0 (input preserved)
    if (f->m1 > 0) PMB_CHECK(exec_c2c(f, f->line, (void *) cplx, f->work0, CUFFT_INVERSE));
    return slab_c2r_after_lines(f, real);
}

// slabs: work0 holds the axis-0 inverse-transformed lines (m1*nc, n0)
static int slab_c2r_after_lines(pmb_fft *f, void *real)
{
    pmb_ctx *ctx = f->ctx;
    if (f->p2p) {
        // 2. transpose straight into the (m0_p, n1, nc) plane layout of every owner p, 3. barrier,
        // 4. 2-D c2r of my planes out of the landing buffer
        const int b = f->xcur;
        f->xcur ^= 1;
        XDest d;
        xdest_bwd(f, b, &d);
        if (f->elsize == 8) PMB_CHECK(xpose_scatter<double2>(f, f->work0, f->n[0], f->m1 * f->nc, f->n[1] * f->nc, d, 1.0));
        else PMB_CHECK(xpose_scatter<float2>(f, f->work0, f->n[0], f->m1 * f->nc, f->n[1] * f->nc, d, 1.0));
        PMB_CHECK(pmb_stream_barrier(ctx));
        if (f->m0 > 0) PMB_CHECK(exec_c2r(f, f->slab_c2r, f->xbuf[b], real));
        return PMB_OK;
    }
    return c2r_from_work(f, real);
}

// ---- several backward transforms with their transposes overlapped ------------------------------------
// The three c2r of a force evaluation are independent; run one after the other each is
//   L (lines along 0, cuFFT, HBM-bound) -> S (transpose + NVLink stores) -> barrier -> P (planes, cuFFT).
// Here L and P of all transforms stay on the compute stream and every S (+ its barrier) goes to a second
// stream: the NVLink stores of transform d run under the cuFFT kernels of transforms d - 1 and d + 1.
// Buffers: work0 / work1 alternate as the source of S_d; the two landing buffers alternate as its target.
// Ordering (events E; every barrier is issued on the transpose stream, so the NCCL calls of this communicator
// never run concurrently):
//   transpose stream: wait(entry) barrier | wait(L_d) [d >= 2: wait(P_{d-2}) barrier] S_d barrier rec(S_d) | ...
//                     | wait(P_last) barrier rec(exit)
//   compute stream:   rec(entry) L_0 rec L_1 rec | wait(S_d) P_d rec(P_d) ; L_{d+2} rec ... | wait(exit)
// - the entry barrier: every rank has finished all earlier work (reads of both landing buffers included)
//   before anybody stores into a peer;
// - S_d (d >= 2) reuses the landing buffer of transform d - 2: it waits for P_{d-2} here AND, through the
//   extra barrier, on every other rank;  L_{d+2} reuses the work buffer of S_d: it follows wait(S_d);
// - the exit barrier: every rank has finished its last P before a later transform may store into it.
static int c2r_multi_impl(pmb_fft *f, int n, const void *const *cplx_h, void *const *real_h, PmbIfftArgs *fused);

extern "C" int pmb_fft_c2r_multi(pmb_fft *f, int n, const void *const *cplx_h, void *const *real_h)
{
    PMB_REQUIRE(f && cplx_h && real_h && n >= 1 && n <= 4, "bad arguments");
    for (int d = 0; d < n; d++) PMB_REQUIRE(cplx_h[d] && real_h[d], "null field %d", d);
    return c2r_multi_impl(f, n, cplx_h, real_h, NULL);
}

// fused != NULL (slabs, n == 3): the lines of transform d come from the fused transfer + line transform of the common
// input (pmb_ifft.cuh) instead of cuFFT on cplx_h[d] (which is then not read)
static int c2r_multi_impl(pmb_fft *f, int n, const void *const *cplx_h, void *const *real_h, PmbIfftArgs *fused)
{
    static int overlap = -1;
    if (overlap < 0) { const char *e = getenv("PMB_FFT_OVERLAP"); overlap = e ? atoi(e) : 1; }
    auto fused_lines = [&](int d, void *dst) -> int {
        PmbIfftArgs a = *fused;
        a.ntr = 1;
        a.tr[0].axis0mul = d == 0; a.tr[0].nout = 1; a.tr[0].out[0] = dst; a.tr[0].linemul[0] = d;
        a.tr[0].sten_out = NULL; a.tr[0].sten_c = 0.0;
        return ifft_launch(f, a, f->ctx->stream);
    };
    if (!(f->P > 1 && f->P1 == 1 && f->p2p && n >= 2 && overlap && f->work1)) {
        for (int d = 0; d < n; d++) {
            if (fused) {
                if (f->m1 > 0) PMB_CHECK(fused_lines(d, f->work0));
                PMB_CHECK(slab_c2r_after_lines(f, real_h[d]));
            } else {
                PMB_CHECK(pmb_fft_c2r(f, cplx_h[d], real_h[d]));
            }
        }
        return PMB_OK;
    }
    pmb_ctx *ctx = f->ctx;
    PMB_CHECK(ensure_xstream(f));
    cudaStream_t ms = ctx->stream, xs = f->xstream;
    cudaEvent_t *evL = f->xev, *evS = f->xev + 4, *evP = f->xev + 8, evEntry = f->xev[12], evExit = f->xev[13];
    void *wk[2] = {f->work0, f->work1};
    // fused lines with the finite-difference stencil: ONE launch transforms every line once and writes the lines of
    // all three directions (three line buffers); otherwise one launch (or cuFFT call) per direction, two buffers
    bool all3 = fused && n == 3 && fused->tr[0].sten_c != 0.0;
    if (all3 && !f->work2 && cudaMalloc(&f->work2, f->work_bytes) != cudaSuccess) { cudaGetLastError(); f->work2 = NULL; all3 = false; }
    void *wk3[3] = {f->work0, f->work1, f->work2};
    auto wbuf = [&](int d) -> void * { return all3 ? wk3[d] : wk[d & 1]; };
    int xb[4];
    for (int d = 0; d < n; d++) { xb[d] = f->xcur; f->xcur ^= 1; }
    // entry
    PMB_CUDA(cudaEventRecord(evEntry, ms));
    PMB_CUDA(cudaStreamWaitEvent(xs, evEntry, 0));
    PMB_CHECK(pmb_stream_barrier_on(ctx, xs));
    auto line = [&](int d) -> int {          // L_d on the compute stream
        if (all3) {
            if (d > 0) return PMB_OK;         // written by the launch of d == 0
            if (f->m1 > 0) {
                PmbIfftArgs a = *fused;
                a.ntr = 1;
                a.tr[0].axis0mul = 0; a.tr[0].nout = 2;
                a.tr[0].out[0] = wk3[1]; a.tr[0].linemul[0] = 1;
                a.tr[0].out[1] = wk3[2]; a.tr[0].linemul[1] = 2;
                a.tr[0].sten_out = wk3[0];
                PMB_CHECK(ifft_launch(f, a, ms));
            }
            for (int e = 0; e < n; e++) PMB_CUDA(cudaEventRecord(evL[e], ms));
            return PMB_OK;
        }
        if (f->m1 > 0) {
            if (fused) PMB_CHECK(fused_lines(d, wk[d & 1]));
            else PMB_CHECK(exec_c2c(f, f->line, (void *) cplx_h[d], wk[d & 1], CUFFT_INVERSE));
        }
        PMB_CUDA(cudaEventRecord(evL[d], ms));
        return PMB_OK;
    };
    auto scatter = [&](int d) -> int {       // S_d + barrier on the transpose stream
        PMB_CUDA(cudaStreamWaitEvent(xs, evL[d], 0));
        if (d >= 2) {
            PMB_CUDA(cudaStreamWaitEvent(xs, evP[d - 2], 0));
            PMB_CHECK(pmb_stream_barrier_on(ctx, xs));
        }
        XDest dst;
        xdest_bwd(f, xb[d], &dst);
        if (f->elsize == 8) PMB_CHECK(xpose_scatter<double2>(f, wbuf(d), f->n[0], f->m1 * f->nc, f->n[1] * f->nc, dst, 1.0, -1, -1, 1, 0, xs));
        else PMB_CHECK(xpose_scatter<float2>(f, wbuf(d), f->n[0], f->m1 * f->nc, f->n[1] * f->nc, dst, 1.0, -1, -1, 1, 0, xs));
        PMB_CHECK(pmb_stream_barrier_on(ctx, xs));
        PMB_CUDA(cudaEventRecord(evS[d], xs));
        return PMB_OK;
    };
    auto planes = [&](int d) -> int {        // P_d on the compute stream
        PMB_CUDA(cudaStreamWaitEvent(ms, evS[d], 0));
        if (f->m0 > 0) PMB_CHECK(exec_c2r(f, f->slab_c2r, f->xbuf[xb[d]], real_h[d]));
        PMB_CUDA(cudaEventRecord(evP[d], ms));
        return PMB_OK;
    };
    PMB_CHECK(line(0));
    PMB_CHECK(scatter(0));
    if (n > 1) PMB_CHECK(line(1));
    for (int d = 0; d < n; d++) {
        if (d + 1 < n) PMB_CHECK(scatter(d + 1));     // S_{d+1} is queued before P_d: it runs under it
        PMB_CHECK(planes(d));
        if (d + 2 < n) PMB_CHECK(line(d + 2));        // reuses the work buffer of S_d, which P_d has waited for
    }
    // exit
    PMB_CUDA(cudaStreamWaitEvent(xs, evP[n - 1], 0));
    PMB_CHECK(pmb_stream_barrier_on(ctx, xs));
    PMB_CUDA(cudaEventRecord(evExit, xs));
    PMB_CUDA(cudaStreamWaitEvent(ms, evExit, 0));
    return PMB_OK;
}

// steps 2-5 of the backward transform; work0 holds the axis-0 inverse-transformed lines (m1*nc, n0)
static int c2r_from_work(pmb_fft *f, void *real)
{
    pmb_ctx *ctx = f->ctx;
    const size_t csz = 2 * (size_t) f->elsize;
    // 2. (m1*nc, n0) -> (n0, m1*nc): rows of destination p are contiguous
    if (f->elsize == 8) PMB_CHECK(transpose<double2>(f, f->work0, f->work1, f->m1 * f->nc, f->n[0], 1.0));
    else PMB_CHECK(transpose<float2>(f, f->work0, f->work1, f->m1 * f->nc, f->n[0], 1.0));
    // 3. global transpose back
    int64_t sc[64], so[64], rc[64], ro[64];
    slab_counts(f, 0, sc, so, rc, ro);
    PMB_CHECK(pmb_alltoallv(ctx, f->work1, sc, so, f->work0, rc, ro, (int64_t) csz));
    // 4. unpack blocks (m0, m1_q, nc) into the in-place complex view (m0, n1, nc) of the real buffer
    if (f->elsize == 8) PMB_CHECK(slab_pack<double2>(f, f->work0, real, 1));
    else PMB_CHECK(slab_pack<float2>(f, f->work0, real, 1));
    // 5. planes: 2-D c2r in place
    if (f->m0 > 0) PMB_CHECK(exec_c2r(f, f->slab_c2r, real, real));
    return PMB_OK;
}

// ---- transfer functions -----------------------------------------------------------
// Per-axis factors are tabulated on the host once per call (3 short tables: wavenumbers k_d[i], and
// the direction multiplier m[i] = kfinite(k_dir[i]) / k_dir[i] / window factor ...), so the kernel is
// a pure streaming multiply: one complex load, one complex store, no sin/div-mod per cell.
struct TfArgs {
    int kind, ndim, P;
    int64_t n[3], nc, s1, m1, s2, mc;
    const double *ktab[3];   // wavenumber per global index of each (left-padded) axis
    const double *mtab;      // per-index multiplier along the direction axis (or window factors, 3 axes)
    int dd;                  // direction axis (left-padded index)
    double pre;              // scalar folded into the multiplier (a pending normalisation of the input)
    double p0;
};

// wavenumber of global index i on an axis of n points, box length L (ref: pm.py:1213-1219):
// Nyquist and above map to negative frequencies.
static double host_wavenumber(int64_t i, int64_t n, double L)
{
    double w = (double) (i >= n / 2 ? i - n : i);
    w = w * (2 * 3.141592653589793 / (double) n);
    return w * (double) n / L;
}

static double host_sinc_unnormed(double x)
{
    if (x < 1e-5 && x > -1e-5) {
        double x2 = x * x;
        return 1.0 - x2 / 6. + x2 * x2 / 120.;
    }
    return sin(x) / x;
}

template <typename C>
__device__ __forceinline__ void pmb_tf_apply(const C *__restrict__ in, C *__restrict__ out, int64_t t,
                                             const TfArgs &a, double k0, double k1, double k2v, int64_t idir,
                                             int64_t i0, int64_t i1, int64_t i2)
{
    double k2 = 0;
    if (a.ndim > 2) k2 = k2 + k0 * k0;
    if (a.ndim > 1) k2 = k2 + k1 * k1;
    k2 = k2 + k2v * k2v;
    double re = 1.0, im = 0.0;
    switch (a.kind) {
    case PMB_TF_SCALE: re = a.p0; break;
    case PMB_TF_GRAVITY_FD4:
    case PMB_TF_GRADIENT_K:
        if (k2 == 0) k2 = 1.0;
        re = 0; im = a.mtab[idir] / k2;
        break;
    case PMB_TF_INV_LAPLACE:
        if (k2 == 0) k2 = 1.0;
        re = -1. / k2;
        break;
    case PMB_TF_GAUSS_LOWPASS:
        re = exp(-0.5 * k2 * (a.p0 * a.p0));
        break;
    case PMB_TF_COMPENSATE: {
        double tf = 1.0;
        if (a.ndim > 2) tf = tf * a.mtab[i0];
        if (a.ndim > 1) tf = tf * a.mtab[a.n[0] + i1];
        tf = tf * a.mtab[a.n[0] + a.n[1] + i2];
        re = 1.0 / tf;
    } break;
    case PMB_TF_IK:
        re = 0; im = a.mtab[idir];
        break;
    case PMB_TF_POWERLAW:
        re = k2 == 0 ? 0.0 : pow(k2, 0.5 * a.p0);
        break;
    }
    re = re * a.pre;
    im = im * a.pre;
    const C v = in[t];
    C o;
    o.x = (decltype(o.x)) ((double) v.x * re - (double) v.y * im);
    o.y = (decltype(o.y)) ((double) v.x * im + (double) v.y * re);
    out[t] = o;
}

// one WARP walks a whole row of the contiguous axis (512-byte coalesced accesses, at most one
// partially filled iteration per row); the row -> (outer indices) split costs one division per row
template <typename C>
__global__ void __launch_bounds__(256)
pmb_k_transfer(const C *__restrict__ in, C *__restrict__ out, int64_t nrows, int64_t rowlen, TfArgs a)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t row = (int64_t) blockIdx.x * 8 + warp; row < nrows; row += (int64_t) gridDim.x * 8) {
        int64_t i0, i1, i2 = 0;
        if (a.P == 1) {           // (n0, n1, nc): rows = (i0, i1), inner = i2
            i1 = row % a.n[1];
            i0 = row / a.n[1];
        } else {                  // (m1, mc, n0): rows = (j, i2), inner = i0
            i2 = a.s2 + row % a.mc;
            i1 = a.s1 + row / a.mc;
            i0 = 0;
        }
        const double kA = a.P == 1 ? a.ktab[0][i0] : a.ktab[2][i2];
        const double kB = a.ktab[1][i1];
        for (int64_t c = lane; c < rowlen; c += 32) {
            const int64_t t = row * rowlen + c;
            if (a.P == 1) {
                const int64_t idir = a.dd == 0 ? i0 : (a.dd == 1 ? i1 : c);
                pmb_tf_apply<C>(in, out, t, a, kA, kB, a.ktab[2][c], idir, i0, i1, c);
            } else {
                const int64_t idir = a.dd == 0 ? c : (a.dd == 1 ? i1 : i2);
                pmb_tf_apply<C>(in, out, t, a, a.ktab[0][c], kB, kA, idir, c, i1, i2);
            }
        }
    }
}

// The three gradient transfers of the force step in one pass: out_d = i m_d(k_d) / k^2 * in for d = 0, 1, 2
// (m_d = kfinite_d for PMB_TF_GRAVITY_FD4, k_d for PMB_TF_GRADIENT_K).  The density modes are read once and
// 1 / k^2 is formed once: 16 + 48 bytes per cell instead of 3 x 32.
template <typename C>
__global__ void __launch_bounds__(256)
pmb_k_transfer_grad3(const C *__restrict__ in, C *__restrict__ o0, C *__restrict__ o1, C *__restrict__ o2,
                     int64_t nrows, int64_t rowlen, TfArgs a)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double *m0 = a.mtab, *m1 = a.mtab + a.n[0], *m2 = a.mtab + a.n[0] + a.n[1];
    for (int64_t row = (int64_t) blockIdx.x * 8 + warp; row < nrows; row += (int64_t) gridDim.x * 8) {
        int64_t i0 = 0, i1, i2 = 0;
        if (a.P == 1) { i1 = row % a.n[1]; i0 = row / a.n[1]; }
        else { i2 = a.s2 + row % a.mc; i1 = a.s1 + row / a.mc; }
        const double kB = a.ktab[1][i1];
        for (int64_t c = lane; c < rowlen; c += 32) {
            const int64_t t = row * rowlen + c;
            const int64_t j0 = a.P == 1 ? i0 : c, j2 = a.P == 1 ? c : i2;
            const double k0 = a.ktab[0][j0], k2v = a.ktab[2][j2];
            // the same order of additions as pmb_tf_apply: ((k0^2) + k1^2) + k2^2
            double k2 = 0;
            k2 = k2 + k0 * k0;
            k2 = k2 + kB * kB;
            k2 = k2 + k2v * k2v;
            if (k2 == 0) k2 = 1.0;
            const C v = in[t];
            const double im[3] = {(m0[j0] / k2) * a.pre, (m1[i1] / k2) * a.pre, (m2[j2] / k2) * a.pre};
            C r;
            r.x = (decltype(r.x)) (0.0 - (double) v.y * im[0]); r.y = (decltype(r.y)) ((double) v.x * im[0]); o0[t] = r;
            r.x = (decltype(r.x)) (0.0 - (double) v.y * im[1]); r.y = (decltype(r.y)) ((double) v.x * im[1]); o1[t] = r;
            r.x = (decltype(r.x)) (0.0 - (double) v.y * im[2]); r.y = (decltype(r.y)) ((double) v.x * im[2]); o2[t] = r;
        }
    }
}

// wavenumbers (3 axes) | gradient multipliers (3 axes) of the two gradient transfers, cached on the plan
static int grad3_tables(pmb_fft *f, int kind, const double *boxsize_h, void **out)
{
    const int64_t ntab = f->n[0] + f->n[1] + f->n[2];
    int64_t off[3] = {0, f->n[0], f->n[0] + f->n[1]};
    void *dev = tf_cache_find(f, kind, -3, boxsize_h, 0.0, 0.0);
    if (!dev) {

        double *h = (double *) malloc(sizeof(double) * 2 * ntab);
        if (!h) return PMB_ENOMEM;
        double *m = h + ntab;
        for (int d = 0; d < 3; d++)
            for (int64_t i = 0; i < f->n[d]; i++) {
                const double k = host_wavenumber(i, f->n[d], boxsize_h[d]);
                h[off[d] + i] = k;
                if (kind == PMB_TF_GRAVITY_FD4) {
                    const double Cc = boxsize_h[d] / (double) f->n[d];
                    const double w = k * Cc;
                    m[off[d] + i] = 1.0 / Cc * 1 / 6.0 * (8 * sin(w) - sin(2 * w));
                } else {
                    m[off[d] + i] = k;
                }
            }
        const int rc = tf_cache_store(f, kind, -3, boxsize_h, 0.0, 0.0, h, sizeof(double) * 2 * ntab, &dev);
        free(h);
        if (rc != PMB_OK) return rc;
    }
    *out = dev;
    return PMB_OK;
}

extern "C" int pmb_transfer_grad3(pmb_fft *f, int kind, const double *boxsize_h, double prefactor, const void *in,
                                  void *const *outs_h)
{
    PMB_REQUIRE(f && boxsize_h && in && outs_h && outs_h[0] && outs_h[1] && outs_h[2], "null argument");
    PMB_REQUIRE(kind == PMB_TF_GRAVITY_FD4 || kind == PMB_TF_GRADIENT_K, "grad3 serves the two gradient transfers");
    PMB_REQUIRE(f->ndim == 3, "3-D meshes only");
    pmb_ctx *ctx = f->ctx;
    const int64_t ntab = f->n[0] + f->n[1] + f->n[2];
    int64_t off[3] = {0, f->n[0], f->n[0] + f->n[1]};
    void *dev = NULL;
    PMB_CHECK(grad3_tables(f, kind, boxsize_h, &dev));
    TfArgs a;
    memset(&a, 0, sizeof(a));
    a.kind = kind; a.ndim = 3; a.P = f->P;
    for (int d = 0; d < 3; d++) { a.n[d] = f->n[d]; a.ktab[d] = (const double *) dev + off[d]; }
    a.mtab = (const double *) dev + ntab;
    a.nc = f->nc; a.s1 = f->s1; a.m1 = f->m1; a.s2 = f->s2; a.mc = f->mc;
    a.pre = prefactor;
    int64_t nrows, rowlen;
    if (f->P == 1) { nrows = f->n[0] * f->n[1]; rowlen = f->nc; }
    else { nrows = f->m1 * f->mc; rowlen = f->n[0]; }
    if (nrows == 0 || rowlen == 0) return PMB_OK;
    int64_t grid = (nrows + 7) / 8;
    const int64_t cap = (int64_t) ctx->sm_count * 8;
    if (grid > cap) grid = cap;
    if (f->elsize == 8)
        pmb_k_transfer_grad3<double2><<<(int) grid, 256, 0, ctx->stream>>>((const double2 *) in, (double2 *) outs_h[0], (double2 *) outs_h[1],
                                                                            (double2 *) outs_h[2], nrows, rowlen, a);
    else
        pmb_k_transfer_grad3<float2><<<(int) grid, 256, 0, ctx->stream>>>((const float2 *) in, (float2 *) outs_h[0], (float2 *) outs_h[1],
                                                                          (float2 *) outs_h[2], nrows, rowlen, a);
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}

// ---- the three backward transforms of a force evaluation with the transfers fused into their first pass -----------
template <typename C, int N, bool CONTIG, bool WIDE = true>
static int ifft_launch_n(pmb_fft *f, const PmbIfftArgs &a, cudaStream_t stream)
{
    constexpr int B = pmb_ifft_bundle<C, N, CONTIG, WIDE>::B;
    typedef pmb_ifft_line<C, N, CONTIG, B> F;
    constexpr int threads = B * (N / 16);
    constexpr int MINB = threads <= 256 ? 2 : 1;
    auto kern = pmb_k_ifft_grad<C, N, CONTIG, B, MINB>;
    const size_t smem = (size_t) F::SMEM_ELEMS * sizeof(C) + pmb_ifft_tables<C, N>::BYTES;
    static int occ = 0;      // per instantiation; one device per process
    if (occ == 0) {
        PMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        int o = 0;
        PMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, threads, smem));
        if (o < 1) { pmb_set_error("fused line transform of %d points does not fit on an SM", N); return PMB_EUNSUPPORTED; }
        const char *e = getenv("PMB_IFFT_CTAS");
        if (e && atoi(e) > 0 && atoi(e) < o) o = atoi(e);
        occ = o;
    }
    const int64_t nbundles = (a.nlines + B - 1) / B;
    if (nbundles == 0) return PMB_OK;
    int64_t grid = (int64_t) f->ctx->sm_count * occ;
    if (grid > nbundles) grid = nbundles;
    PMB_CHECK(lib_begin(f, 2, stream));
    kern<<<(int) grid, threads, smem, stream>>>(a);
    PMB_CHECK(lib_end(f, stream));
    f->ifft_launches++;
    PMB_LAUNCH_CHECK(f->ctx);
    return PMB_OK;
}

template <typename C, bool CONTIG>
static int ifft_launch_c(pmb_fft *f, const PmbIfftArgs &a, cudaStream_t stream)
{
    switch (f->n[0]) {
    case 64: return ifft_launch_n<C, 64, CONTIG>(f, a, stream);
    case 128: return ifft_launch_n<C, 128, CONTIG>(f, a, stream);
    case 256: return ifft_launch_n<C, 256, CONTIG>(f, a, stream);
    case 512: return ifft_launch_n<C, 512, CONTIG>(f, a, stream);
    case 1024: {
        static int narrow = -1;
        if (narrow < 0) { const char *e = getenv("PMB_IFFT_NARROW"); narrow = e ? atoi(e) : 0; }
        if (!CONTIG && narrow) return ifft_launch_n<C, 1024, CONTIG, false>(f, a, stream);
        return ifft_launch_n<C, 1024, CONTIG>(f, a, stream);
    }
    case 2048: return ifft_launch_n<C, 2048, CONTIG>(f, a, stream);
    case 4096: return ifft_launch_n<C, 4096, CONTIG>(f, a, stream);
    }
    pmb_set_error("fused line transform: unsupported length %lld", (long long) f->n[0]);
    return PMB_EUNSUPPORTED;
}

static int ifft_launch(pmb_fft *f, const PmbIfftArgs &a, cudaStream_t stream)
{
    if (f->elsize == 8) return f->P == 1 ? ifft_launch_c<double2, false>(f, a, stream) : ifft_launch_c<double2, true>(f, a, stream);
    return f->P == 1 ? ifft_launch_c<float2, false>(f, a, stream) : ifft_launch_c<float2, true>(f, a, stream);
}

static bool ifft_supported(const pmb_fft *f, const double *box)
{
    static int on = -1;
    if (on < 0) { const char *e = getenv("PMB_IFFT"); on = e ? atoi(e) : 1; }
    if (!on || f->ndim != 3 || f->P1 != 1) return false;
    const int64_t n = f->n[0];
    if (!(n >= 64 && n <= 4096 && (n & (n - 1)) == 0)) return false;
    // the kernel's reciprocal of k^2 has no special cases: every k^2 of the mesh must be an ordinary number
    double lo = 1e300, hi = 0;
    for (int d = 0; d < 3; d++) {
        if (!(box[d] > 0)) return false;
        const double k1 = 2 * 3.141592653589793 / box[d], kn = k1 * (double) (f->n[d] / 2 + 1);
        if (f->n[d] > 1 && k1 * k1 < lo) lo = k1 * k1;
        hi += kn * kn;
    }
    return lo > 1e-280 && hi < 1e280 && hi == hi;
}

// twiddles exp(+2 pi i k / n0) in the field's precision, once per plan
static int ifft_twiddles(pmb_fft *f)
{
    if (f->ifft_tw) return PMB_OK;
    const int64_t n = f->n[0];
    const size_t csz = 2 * (size_t) f->elsize;
    void *h = malloc(csz * (size_t) n);
    if (!h) return PMB_ENOMEM;
    for (int64_t k = 0; k < n; k++) {
        // exact symmetries of the circle: octant reduction keeps cos / sin of the table consistent to the last bit
        const double ang = 2 * 3.14159265358979323846 * (double) k / (double) n;
        const double c = cos(ang), sn = sin(ang);
        if (f->elsize == 8) { ((double *) h)[2 * k] = c; ((double *) h)[2 * k + 1] = sn; }
        else { ((float *) h)[2 * k] = (float) c; ((float *) h)[2 * k + 1] = (float) sn; }
    }
    cudaError_t e = cudaMalloc(&f->ifft_tw, csz * (size_t) n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(f->ifft_tw, h, csz * (size_t) n, cudaMemcpyHostToDevice, f->ctx->stream);   // pageable: staged before return
    free(h);
    if (e != cudaSuccess) { if (f->ifft_tw) { cudaFree(f->ifft_tw); f->ifft_tw = NULL; } return pmb_cuda_fail(e, "twiddle table", __FILE__, __LINE__); }
    return PMB_OK;
}

static int c2r_grad3_impl(pmb_fft *f, int kind, const double *boxsize_h, double prefactor, const void *in,
                         void *const *reals_h, int forward);

extern "C" int pmb_fft_c2r_grad3(pmb_fft *f, int kind, const double *boxsize_h, double prefactor, const void *in,
                                 void *const *reals_h)
{
    return c2r_grad3_impl(f, kind, boxsize_h, prefactor, in, reals_h, 0);
}

// real_in -> 2-D r2c of every plane (cuFFT) -> [forward axis-0 transform, transfers, inverse axis-0 transform] in one
// kernel -> 2-D c2r of every plane (cuFFT), three times
extern "C" int pmb_fft_force3(pmb_fft *f, int kind, const double *boxsize_h, double prefactor, const void *real_in,
                              void *const *reals_h)
{
    PMB_REQUIRE(f && boxsize_h && real_in && reals_h && reals_h[0] && reals_h[1] && reals_h[2], "null argument");
    static int on = -1;
    if (on < 0) { const char *e = getenv("PMB_IFFT_FORWARD"); on = e ? atoi(e) : 1; }
    if (!on || f->P != 1 || !ifft_supported(f, boxsize_h)) {
        pmb_set_error("pmb_fft_force3 serves one rank and axis-0 lengths 64 .. 4096 (powers of two): use pmb_fft_r2c + pmb_fft_c2r_grad3");
        return PMB_EUNSUPPORTED;
    }
    for (int d = 0; d < 3; d++) PMB_REQUIRE(reals_h[d] != real_in, "the input field must not be one of the outputs");
    if (!f->have_plane_r2c) {
        long long n2[2] = {f->n[1], f->n[2]};
        long long inr[2] = {f->n[1], 2 * f->nc};
        long long inc[2] = {f->n[1], f->nc};
        PMB_CHECK(make_plan(f, &f->plane_r2c, 2, n2, inr, f->n[1] * 2 * f->nc, inc, f->n[1] * f->nc,
                            f->elsize == 8 ? CUFFT_D2Z : CUFFT_R2C, f->n[0]));
        f->have_plane_r2c = true;
    }
    // the planes' modes land in the buffer of output 2, which the kernel then overwrites in place (every thread writes
    // the points of the line it read)
    PMB_CHECK(exec_r2c(f, f->plane_r2c, real_in, reals_h[2]));
    return c2r_grad3_impl(f, kind, boxsize_h, prefactor, reals_h[2], reals_h, 1);
}

static int c2r_grad3_impl(pmb_fft *f, int kind, const double *boxsize_h, double prefactor, const void *in,
                         void *const *reals_h, int forward)
{
    PMB_REQUIRE(f && boxsize_h && in && reals_h && reals_h[0] && reals_h[1] && reals_h[2], "null argument");
    PMB_REQUIRE(kind == PMB_TF_GRAVITY_FD4 || kind == PMB_TF_GRADIENT_K, "grad3 serves the two gradient transfers");
    PMB_REQUIRE(f->ndim == 3, "3-D meshes only");
    for (int d = 0; d < 3; d++) {
        PMB_REQUIRE(forward || reals_h[d] != in, "the input modes must not be one of the outputs");
        for (int e = 0; e < d; e++) PMB_REQUIRE(reals_h[d] != reals_h[e], "outputs %d and %d are the same buffer", e, d);
    }
    if (!ifft_supported(f, boxsize_h)) {
        // three transfers in one pass into the outputs' in-place complex partners, then the transforms in place
        PMB_CHECK(pmb_transfer_grad3(f, kind, boxsize_h, prefactor, in, reals_h));
        const void *c[3] = {reals_h[0], reals_h[1], reals_h[2]};
        return pmb_fft_c2r_multi(f, 3, c, reals_h);
    }
    void *dev = NULL;
    PMB_CHECK(grad3_tables(f, kind, boxsize_h, &dev));
    PMB_CHECK(ifft_twiddles(f));
    const int64_t ntab = f->n[0] + f->n[1] + f->n[2];
    const int64_t off[3] = {0, f->n[0], f->n[0] + f->n[1]};
    PmbIfftArgs a;
    memset(&a, 0, sizeof(a));
    a.in = in;
    a.P = f->P;
    a.nc = f->nc; a.s1 = f->s1; a.mc = f->mc; a.s2 = f->s2;
    for (int d = 0; d < 3; d++) { a.ktab[d] = (const double *) dev + off[d]; a.mtab[d] = (const double *) dev + ntab + off[d]; }
    a.tw = f->ifft_tw;
    a.pre = prefactor;
    a.forward = forward;
    if (f->P == 1) {
        if (!f->have_plane) {
            long long n2[2] = {f->n[1], f->n[2]};
            long long inr[2] = {f->n[1], 2 * f->nc};
            long long inc[2] = {f->n[1], f->nc};
            PMB_CHECK(make_plan(f, &f->plane_c2r, 2, n2, inc, f->n[1] * f->nc, inr, f->n[1] * 2 * f->nc,
                                f->elsize == 8 ? CUFFT_Z2D : CUFFT_C2R, f->n[0]));
            f->have_plane = true;
        }
        a.nlines = f->n[1] * f->nc;
        // direction 0 has a multiplier that varies along the line; directions 1 and 2 are the SAME transform times a
        // constant of the line: two transforms, three outputs
        static int use_stencil = -1;
        if (use_stencil < 0) { const char *e = getenv("PMB_IFFT_STENCIL"); use_stencil = e ? atoi(e) : 1; }
        // (float64 fields only: the differences of neighbouring phi values cancel ~ 1 / (k C) leading digits on the
        // longest waves -- nothing against 1e-16, too much of float32's 6e-8 on long axes)
        if (kind == PMB_TF_GRAVITY_FD4 && use_stencil && f->elsize == 8) {
            // ... and with the finite-difference gradient direction 0 is a 5-point stencil along the line of the very
            // same transform: ONE transform, three outputs
            a.ntr = 1;
            a.tr[0].axis0mul = 0; a.tr[0].nout = 2; a.tr[0].out[0] = reals_h[1]; a.tr[0].linemul[0] = 1;
            a.tr[0].out[1] = reals_h[2]; a.tr[0].linemul[1] = 2;
            a.tr[0].sten_out = reals_h[0];
            a.tr[0].sten_c = 1.0 / (12.0 * (boxsize_h[0] / (double) f->n[0]));
        } else {
            a.ntr = 2;
            a.tr[0].axis0mul = 1; a.tr[0].nout = 1; a.tr[0].out[0] = reals_h[0]; a.tr[0].linemul[0] = 0;
            a.tr[1].axis0mul = 0; a.tr[1].nout = 2; a.tr[1].out[0] = reals_h[1]; a.tr[1].linemul[0] = 1;
            a.tr[1].out[1] = reals_h[2]; a.tr[1].linemul[1] = 2;
        }
        PMB_CHECK(ifft_launch(f, a, f->ctx->stream));
        for (int d = 0; d < 3; d++) PMB_CHECK(exec_c2r(f, f->plane_c2r, reals_h[d], reals_h[d]));
        return PMB_OK;
    }
    a.nlines = f->m1 * f->mc;
    // slabs: the stencil coefficient tells c2r_multi_impl that one launch can serve the three directions (float64
    // finite-difference gradient, see above); the per-direction launches ignore it
    {
        static int use_stencil = -1;
        if (use_stencil < 0) { const char *e = getenv("PMB_IFFT_STENCIL"); use_stencil = e ? atoi(e) : 1; }
        if (kind == PMB_TF_GRAVITY_FD4 && use_stencil && f->elsize == 8)
            a.tr[0].sten_c = 1.0 / (12.0 * (boxsize_h[0] / (double) f->n[0]));
    }
    const void *c[3] = {in, in, in};
    return c2r_multi_impl(f, 3, c, reals_h, &a);
}

// time inside the fused transfer + line transform kernels and their number since the last reset
extern "C" int pmb_fft_fused_stats(pmb_fft *f, float *ms, int64_t *launches, int reset)
{
    PMB_REQUIRE(f && ms && launches, "null argument");
    PMB_CHECK(lib_flush(f));
    *ms = f->ifft_ms;
    *launches = f->ifft_launches;
    if (reset) { f->ifft_ms = 0; f->ifft_launches = 0; }
    return PMB_OK;
}

// ---- collective reductions over the independent modes ------------------------------------------------
// sum over the stored half-complex modes of conj(b) * a * w, w = 2 for 0 < k_last < N/2 (the mode stands
// for itself and its Hermitian conjugate), 1 for k_last = 0 and N/2 (BaseComplexField._expand_hermitian,
// pm.py:911-918; cdot pm.py:945-974, cnorm pm.py:920-943 with the default norm).  One warp per row of the
// contiguous axis as in pmb_k_transfer; float64 accumulation; per-block partial sums, added on the host.
template <typename C>
__global__ void __launch_bounds__(256)
pmb_k_cdot(const C *__restrict__ a, const C *__restrict__ b, int64_t nrows, int64_t rowlen, int P, int64_t mc, int64_t s2,
           int64_t nlast, double2 *partial)
{
    __shared__ double sh[2][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double re = 0, im = 0;
    const int64_t nyq = nlast / 2;
    for (int64_t row = (int64_t) blockIdx.x * 8 + warp; row < nrows; row += (int64_t) gridDim.x * 8) {
        const int64_t i2row = P == 1 ? 0 : s2 + row % mc;
        for (int64_t c = lane; c < rowlen; c += 32) {
            const int64_t i2 = P == 1 ? c : i2row;
            const double w = (i2 != 0 && i2 != nyq) ? 2.0 : 1.0;
            const C x = a[row * rowlen + c], y = b[row * rowlen + c];
            // conj(y) * x
            re += w * ((double) y.x * (double) x.x + (double) y.y * (double) x.y);
            im += w * ((double) y.x * (double) x.y - (double) y.y * (double) x.x);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, o);
        im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    if (lane == 0) { sh[0][warp] = re; sh[1][warp] = im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0, i = 0;
        for (int k = 0; k < 8; k++) { r += sh[0][k]; i += sh[1][k]; }
        partial[blockIdx.x] = make_double2(r, i);
    }
}

extern "C" int pmb_cdot(pmb_fft *f, const void *a, const void *b, double *result_h)
{
    PMB_REQUIRE(f && a && b && result_h, "null argument");
    pmb_ctx *ctx = f->ctx;
    result_h[0] = result_h[1] = 0;
    int64_t nrows, rowlen;
    if (f->P == 1) { nrows = f->n[0] * f->n[1]; rowlen = f->nc; }
    else { nrows = f->m1 * f->mc; rowlen = f->n[0]; }
    if (nrows == 0 || rowlen == 0) return PMB_OK;
    int64_t grid = (nrows + 7) / 8;
    const int64_t cap = (int64_t) ctx->sm_count * 8;
    if (grid > cap) grid = cap;
    void *partial;
    PMB_CHECK(pmb_scratch(ctx, sizeof(double2) * grid, &partial));
    if (f->elsize == 8)
        pmb_k_cdot<double2><<<(int) grid, 256, 0, ctx->stream>>>((const double2 *) a, (const double2 *) b, nrows, rowlen, f->P, f->mc, f->s2, f->n[2], (double2 *) partial);
    else
        pmb_k_cdot<float2><<<(int) grid, 256, 0, ctx->stream>>>((const float2 *) a, (const float2 *) b, nrows, rowlen, f->P, f->mc, f->s2, f->n[2], (double2 *) partial);
    PMB_LAUNCH_CHECK(ctx);
    double2 *h = (double2 *) malloc(sizeof(double2) * grid);
    if (!h) return PMB_ENOMEM;
    cudaError_t e = cudaMemcpyAsync(h, partial, sizeof(double2) * grid, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { free(h); return pmb_cuda_fail(e, "cdot copy", __FILE__, __LINE__); }
    for (int64_t i = 0; i < grid; i++) { result_h[0] += h[i].x; result_h[1] += h[i].y; }
    free(h);
    return PMB_OK;
}

extern "C" int pmb_transfer(pmb_fft *f, int kind, int dir, const double *params_h, const double *boxsize_h,
                            const void *in, void *out)
{
    return pmb_transfer_scaled(f, kind, dir, params_h, boxsize_h, 1.0, in, out);
}

extern "C" int pmb_transfer_scaled(pmb_fft *f, int kind, int dir, const double *params_h, const double *boxsize_h,
                                   double prefactor, const void *in, void *out)
{
    PMB_REQUIRE(f && boxsize_h && in && out, "null argument");
    PMB_REQUIRE(kind >= PMB_TF_SCALE && kind <= PMB_TF_POWERLAW, "unknown transfer kind %d", kind);
    PMB_REQUIRE(dir >= 0 && dir < f->ndim, "bad direction %d", dir);
    pmb_ctx *ctx = f->ctx;
    const int pad = 3 - f->ndim;
    double box[3] = {1.0, 1.0, 1.0};
    for (int d = 0; d < f->ndim; d++) box[pad + d] = boxsize_h[d];
    const int64_t ntab = f->n[0] + f->n[1] + f->n[2];
    int64_t off[3] = {0, f->n[0], f->n[0] + f->n[1]};
    const int dd = dir + pad;
    int rc = PMB_OK;
    // tables that depend on parameters: the compensation (window kind, support); others only on kind / dir / box
    const double cp0 = (kind == PMB_TF_COMPENSATE && params_h) ? params_h[0] : 0.0;
    const double cp1 = (kind == PMB_TF_COMPENSATE && params_h) ? params_h[1] : 0.0;
    void *dev = tf_cache_find(f, kind, dir, box, cp0, cp1);
    if (!dev) {
    // host tables: [k0 | k1 | k2 | mult]
    double *h = (double *) malloc(sizeof(double) * 2 * ntab);
    if (!h) return PMB_ENOMEM;
    for (int d = 0; d < 3; d++)
        for (int64_t i = 0; i < f->n[d]; i++) h[off[d] + i] = d < pad ? 0.0 : host_wavenumber(i, f->n[d], box[d]);
    double *m = h + ntab;
    for (int64_t i = 0; i < ntab; i++) m[i] = 0.0;
    if (kind == PMB_TF_GRAVITY_FD4) {
        // kfinite = 1/C * 1/6 * (8 sin w - sin 2w), w = k C   (examples/nbody.py:166-168)
        const double Cc = box[dd] / (double) f->n[dd];
        for (int64_t i = 0; i < f->n[dd]; i++) {
            const double w = h[off[dd] + i] * Cc;
            m[i] = 1.0 / Cc * 1 / 6.0 * (8 * sin(w) - sin(2 * w));
        }
    } else if (kind == PMB_TF_GRADIENT_K || kind == PMB_TF_IK) {
        for (int64_t i = 0; i < f->n[dd]; i++) m[i] = h[off[dd] + i];
    } else if (kind == PMB_TF_COMPENSATE) {
        if (!params_h) { free(h); pmb_set_error("compensation needs params = {kind, support}"); return PMB_EINVAL; }
        PmbWindow w;
        rc = pmb_resolve_window(NULL, (int) params_h[0], (int) params_h[1], 0, NULL, &w, 0);
        if (rc != PMB_OK) { free(h); return rc; }
        PmbWinInfo info;
        pmb_window_info(w.nativesupport, (double) (int) params_h[1], &info);
        const int p = w.family == PMB_FAM_NEAREST ? 1 : w.family == PMB_FAM_LINEAR ? 2
                      : w.family == PMB_FAM_QUADRATIC ? 3 : w.family == PMB_FAM_CUBIC ? 4 : 0;
        for (int d = 0; d < 3; d++)
            for (int64_t i = 0; i < f->n[d]; i++) {
                double s = 1.0;
                if (p && d >= pad) {
                    const double wv = h[off[d] + i] * box[d] / (double) f->n[d];
                    const double b = host_sinc_unnormed(0.5 * (wv / info.vfactor));
                    s = b;
                    for (int j = 1; j < p; j++) s = s * b;
                }
                m[off[d] + i] = s;
            }
    }
    rc = tf_cache_store(f, kind, dir, box, cp0, cp1, h, sizeof(double) * 2 * ntab, &dev);
    free(h);
    if (rc != PMB_OK) return rc;
    }

    TfArgs a;
    memset(&a, 0, sizeof(a));
    a.kind = kind; a.ndim = f->ndim; a.P = f->P; a.dd = dd;
    for (int d = 0; d < 3; d++) { a.n[d] = f->n[d]; a.ktab[d] = (const double *) dev + off[d]; }
    a.mtab = (const double *) dev + ntab;
    a.nc = f->nc; a.s1 = f->s1; a.m1 = f->m1; a.s2 = f->s2; a.mc = f->mc;
    a.p0 = params_h ? params_h[0] : 0.0;
    a.pre = prefactor;
    int64_t nrows, rowlen;
    if (f->P == 1) { nrows = f->n[0] * f->n[1]; rowlen = f->nc; }
    else { nrows = f->m1 * f->mc; rowlen = f->n[0]; }
    if (nrows == 0 || rowlen == 0) return PMB_OK;
    int64_t grid = (nrows + 7) / 8;
    const int64_t cap = (int64_t) ctx->sm_count * 8;
    if (grid > cap) grid = cap;
    if (f->elsize == 8)
        pmb_k_transfer<double2><<<(int) grid, 256, 0, ctx->stream>>>((const double2 *) in, (double2 *) out, nrows, rowlen, a);
    else
        pmb_k_transfer<float2><<<(int) grid, 256, 0, ctx->stream>>>((const float2 *) in, (float2 *) out, nrows, rowlen, a);
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}
