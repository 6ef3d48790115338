/* pmesh_b200.h -- C ABI of libpmesh_b200.so, the B200 (sm_100a) engine behind
 * the pmesh Python API.
 *
 * Every entry point replaces one native boundary of the reference pmesh
 * (citations are file:line under the reference tree):
 *
 *   pmb_window_query / pmb_window_fwindow
 *        <- pmesh_painter_init, pmesh_painter_get_fwindow     pmesh/_window_imp.h:76-86
 *   pmb_paint / pmb_readout
 *        <- ResampleWindow.paint / .readout (Cython loops)     pmesh/_window.pyx:128-205
 *           pmesh_painter_paint / pmesh_painter_readout        pmesh/_window_imp.c:461-471
 *   pmb_decompose
 *        <- GridND.decompose chunk loop + gridnd_fill          pmesh/domain.py:561-652, pmesh/_domain.pyx:9-122
 *   pmb_take / pmb_exchange / pmb_gather_sum
 *        <- Layout._exchange (take + MPI Alltoallv), Layout.gather('sum')
 *                                                            pmesh/domain.py:173-206, 208-318
 *   pmb_fft_*
 *        <- pfft.Partition / pfft.Plan.execute call sites      pmesh/pm.py:226-242, 655-694, 987-1019, 1406-1441
 *   pmb_transfer
 *        <- Field.apply with the force-step transfer functions pmesh/pm.py:617-648, examples/nbody.py:154-181
 *   pmb_comm_*
 *        <- mpi4py communicator used by domain.py / pfft       pmesh/domain.py:112-114,199-205
 *   pmb_whitenoise
 *        <- _whitenoise.generate (N-GenIC scheme on ranlxd1)    pmesh/_whitenoise.pyx:25-45, pm.py:1656-1696
 *
 * Conventions: plain C types only.  Pointers are DEVICE pointers unless the
 * name ends in _h.  Every function returns 0 on success or a negative
 * PMB_E* code; pmb_last_error() gives the message of the last failure on the
 * calling thread.  The library never frees caller memory.  One pmb_ctx per
 * process/GPU; all work of a context is ordered on its one CUDA stream.
 */
#ifndef PMESH_B200_H
#define PMESH_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMB_OK 0
#define PMB_EINVAL (-1)   /* bad argument */
#define PMB_ECUDA (-2)    /* CUDA runtime / cuFFT failure */
#define PMB_ENCCL (-3)    /* NCCL failure */
#define PMB_ENOMEM (-4)
#define PMB_EUNSUPPORTED (-5)

#define PMB_MODE_ATOMIC 0         /* red.global.add scatter, order of additions undefined */
#define PMB_MODE_DETERMINISTIC 1  /* sort-by-cell + sequential segmented sum: bit-equal to the reference */

typedef struct pmb_ctx pmb_ctx;
typedef struct pmb_fft pmb_fft;

/* ---- context, memory, timing ------------------------------------------------ */
const char *pmb_last_error(void);
int pmb_version(void);
int pmb_device_count(int *n);
int pmb_ctx_create(int device, pmb_ctx **out);
int pmb_ctx_destroy(pmb_ctx *ctx);
int pmb_ctx_sync(pmb_ctx *ctx);
int pmb_malloc(pmb_ctx *ctx, size_t nbytes, void **out);
int pmb_free(pmb_ctx *ctx, void *ptr);
int pmb_malloc_host(pmb_ctx *ctx, size_t nbytes, void **out_h);  /* pinned */
int pmb_free_host(pmb_ctx *ctx, void *ptr_h);
int pmb_memcpy_h2d(pmb_ctx *ctx, void *dst, const void *src_h, size_t nbytes);
int pmb_memcpy_d2h(pmb_ctx *ctx, void *dst_h, const void *src, size_t nbytes);
int pmb_memcpy_d2d(pmb_ctx *ctx, void *dst, const void *src, size_t nbytes);
int pmb_memset(pmb_ctx *ctx, void *dst, int byte, size_t nbytes);
/* Copies on the context's two COPY streams (1: host -> device, 2: device -> host; pinned host memory),
 * which run concurrently with each other and with the kernels of the compute stream (0).  Ordering is
 * explicit: pmb_stream_record(stream, e) marks a point of a stream in event slot e (0..15),
 * pmb_stream_wait(stream, e) makes everything submitted to `stream` afterwards wait for it.  The host
 * never blocks except in pmb_stream_sync / pmb_ctx_sync. */
int pmb_memcpy_h2d_async(pmb_ctx *ctx, void *dst, const void *src_h, size_t nbytes);
int pmb_memcpy_d2h_async(pmb_ctx *ctx, void *dst_h, const void *src, size_t nbytes);
int pmb_stream_record(pmb_ctx *ctx, int stream_id, int event);
int pmb_stream_wait(pmb_ctx *ctx, int stream_id, int event);
int pmb_stream_sync(pmb_ctx *ctx, int stream_id);
int pmb_mem_info(pmb_ctx *ctx, size_t *free_bytes, size_t *total_bytes);
/* CUDA-event stopwatch on the context's stream (slots 0..15) */
int pmb_timer_start(pmb_ctx *ctx, int slot);
int pmb_timer_stop(pmb_ctx *ctx, int slot, float *ms);       /* records stop, synchronises, returns elapsed */
int pmb_launch_count(pmb_ctx *ctx, int64_t *n, int reset);   /* kernels this library launched so far */
int pmb_flush_l2(pmb_ctx *ctx);                              /* overwrite a 256 MiB scratch buffer */
/* workspace budget (bytes) of the deterministic paint: particles are processed in chunks that fit */
int pmb_set_workspace_limit(pmb_ctx *ctx, size_t nbytes);

/* ---- windows ------------------------------------------------------------------ */
/* upload one lookup table (lanczosN / acgN / dbN / symN); values are host doubles */
int pmb_window_set_table(pmb_ctx *ctx, int kind, const double *values_h, int n,
                         double step, double nativesupport, double hsupport);
/* <- pmesh_painter_init: resolve support. support_req <= 0 means native. */
int pmb_window_query(int kind, int support_req, int *support, int *nativesupport);
/* <- pmesh_painter_get_fwindow, vectorised over n host values */
int pmb_window_fwindow(int kind, int support, const double *w_h, double *out_h, int64_t n);

typedef struct pmb_resample_args {
    int kind;                 /* window kind (enum of pmesh/_window_imp.h:4-28) */
    int support;              /* requested integer support, <= 0: native */
    int ndim;                 /* 1..3 */
    int order[3];             /* 1 on the axis of differentiation (diffdir), else 0 */
    double scale[3];          /* Affine.scale */
    double translate[3];      /* Affine.translate */
    int64_t period[3];        /* Affine.period, 0 = not periodic */
    void *mesh;               /* canvas */
    int mesh_elsize;          /* 4 or 8 */
    int64_t size[3];          /* local canvas shape */
    int64_t strides[3];       /* canvas strides in BYTES */
    const void *pos;          /* (npart, >=ndim) */
    int pos_elsize;           /* 4 or 8 */
    int64_t npart;
    int64_t pos_stride0;      /* bytes between particles */
    int64_t pos_stride1;      /* bytes between coordinates */
    const void *mass;         /* paint: per-particle weight, NULL => mass_scalar.  */
    int mass_elsize;
    int64_t mass_stride;      /* bytes */
    double mass_scalar;
    const void *hsml;         /* per-particle support scaling, NULL => hsml_scalar */
    int hsml_elsize;
    int64_t hsml_stride;
    double hsml_scalar;       /* 1.0 for the native support */
    void *out;                /* readout: (npart,) result */
    int out_elsize;
    int64_t out_stride;       /* bytes */
    int mode;                 /* paint: PMB_MODE_* */
    int pcs_gradient_scale_fix; /* 0: bug-compatible tuned-PCS derivative (no scale[d] factor) */
} pmb_resample_args;

int pmb_paint(pmb_ctx *ctx, const pmb_resample_args *a);
int pmb_readout(pmb_ctx *ctx, const pmb_resample_args *a);
/* readout of nfields (1..3) canvases of IDENTICAL geometry (a->size, strides, dtype) at the same positions
 * in one sweep over the particles: cell indices and weights are computed once and used for every field
 * (the three force components of the PM step, examples/nbody.py:211-216, share one pass over the
 * positions).  a->mesh / a->out / a->out_stride are ignored; results are those of nfields pmb_readout calls. */
int pmb_readout_multi(pmb_ctx *ctx, const pmb_resample_args *a, int nfields, const void *const *meshes_h,
                      void *const *outs_h, const int64_t *out_strides_h);
/* pmb_readout_multi fused with the ghost sum of Layout.gather('sum') (domain.py:208-318): of the npart local
 * particles, [own_begin, own_begin + own_count) are the block this rank sent to itself; the value of own
 * particle k goes straight to row own_index[k] of the float64 column own_outs_h[q] (the original particle
 * order; the columns must be zero-filled: rows without an own copy stay 0), the values of the ghosts held
 * for other ranks go to the compact float64 column ghost_outs_h[q] (npart - own_count rows, own block cut
 * out) from which the reverse alltoallv takes them; pmb_gather_add_segments then adds what comes back.
 * PMB_EUNSUPPORTED when the window / geometry has no fused kernel (callers use pmb_readout + pmb_gather_sum). */
int pmb_readout_multi_gather(pmb_ctx *ctx, const pmb_resample_args *a, int nfields, const void *const *meshes_h,
                             void *const *ghost_outs_h, void *const *own_outs_h, const int32_t *own_index,
                             int64_t own_begin, int64_t own_count);
/* fused value + ndim gradients in one neighbour sweep (paint_vjp / readout_vjp helper).
 * out_value may be NULL; out_grad is (npart, ndim) with byte strides gs0, gs1, element size out_elsize. */
int pmb_readout_grad(pmb_ctx *ctx, const pmb_resample_args *a, void *out_grad, int64_t gs0, int64_t gs1);
/* Particle arrays WITHOUT spatial order (the reference makes no ordering assumption, _window.pyx:157-165): paint
 * and readout of >= 2^18 particles on 3-D meshes work on a tile-sorted COPY of the position records kept in the
 * context (a one-pass counting sort; reused only while a content hash of the caller's array is unchanged).
 * pmb_bin_stats: reorders done so far and bytes held; pmb_bin_release frees the copy; pmb_bin_invalidate keeps its
 * memory but forgets its content (the next call sorts again).  PMB_BIN=0 switches the
 * reorder off (the kernels then walk the array through a permutation), PMB_BIN=2 reorders every large array. */
int pmb_bin_stats(pmb_ctx *ctx, int64_t *builds, int64_t *bytes_held);
/* Memory-pressure hook: a host side that caches device blocks (the caching allocator of pmesh_b200/_lib.py; the
 * reference's numpy arrays have no equivalent) registers a function that gives them back; the library calls it when
 * an allocation of its own work space (scratch, sorted particle copies) fails and then tries once more. */
int pmb_set_trim_callback(pmb_ctx *ctx, void (*cb)(void *), void *arg);
/* the other direction: the library gives back what it caches for itself (scratch, sorted particle copies, permutations) */
int pmb_ctx_trim(pmb_ctx *ctx);
int pmb_bin_release(pmb_ctx *ctx);
int pmb_bin_invalidate(pmb_ctx *ctx);

/* elementwise helpers on (strided, up to 3-D) fields */
int pmb_field_fill(pmb_ctx *ctx, void *mesh, int elsize, int ndim, const int64_t *size,
                   const int64_t *strides, double value);
int pmb_field_scale(pmb_ctx *ctx, void *mesh, int elsize, int is_complex, int ndim, const int64_t *size,
                    const int64_t *strides, double factor);
int pmb_field_sum(pmb_ctx *ctx, const void *mesh, int elsize, int ndim, const int64_t *size,
                  const int64_t *strides, double *sum_h);

/* sum of a[i] * b[i] over two (strided, up to 3-D) views of identical shape and strides, float64 accumulation:
 * the rank-local term of RealField.cdot / cnorm (pm.py:897-905) */
int pmb_field_dot(pmb_ctx *ctx, const void *a, const void *b, int elsize, int ndim, const int64_t *size,
                  const int64_t *strides, double *dot_h);

/* ---- particle columns: the element-wise updates of a KDK step -------------------------------------
 * <- the numpy in-place arithmetic of the reference's integrator, examples/nbody.py:84-102 (symp2:
 *    V += F * K; S += V * D) and :200 (X = S + Q).  Columns are float32 / float64 (elsize) with byte
 *    strides; one multiply and one add per element, rounded like numpy (no FMA). */
/* y += a * x */
int pmb_axpy(pmb_ctx *ctx, void *y, int64_t y_stride, const void *x, int64_t x_stride, double a, int elsize, int64_t n);
/* out = a * x + b * y   (y == NULL: out = a * x) */
int pmb_lincomb(pmb_ctx *ctx, void *out, int64_t out_stride, const void *x, int64_t x_stride, double a,
                const void *y, int64_t y_stride, double b, int elsize, int64_t n);
/* x = x mod period with numpy's floored modulo: the `X % BoxSize` wrap of a driver */
int pmb_column_mod(pmb_ctx *ctx, void *x, int64_t x_stride, double period, int elsize, int64_t n);
/* fused kick + drift in one pass: V += F * kick; S += V * drift.  V, S are contiguous (npart, ncol)
 * rows; the force is column-wise, F_cols_h[d] = dense device column (npart,) as readout / gather
 * produce it (host array of ncol device pointers).  S == NULL: kick only. */
int pmb_kick_drift(pmb_ctx *ctx, void *V, void *S, const void *const *F_cols_h, int ncol,
                   double kick, double drift, int elsize, int64_t npart);

/* sum_i x[i] * y[i] accumulated in float64 (strided float32 / float64 columns): diagnostics such as the
 * rms of a force column or the particle-side term of cdot (pm.py:897-902) */
int pmb_dot(pmb_ctx *ctx, const void *x, int64_t x_stride, const void *y, int64_t y_stride, int elsize,
            int64_t n, double *dot_h);

/* ---- synthetic particles (bench / tests): counter-based, reproducible ------------- */
/* uniform in [0, box) per axis: pos[i,d] = box[d] * u(seed, i + first, d) */
int pmb_particles_uniform(pmb_ctx *ctx, void *pos, int pos_elsize, int64_t npart, int ndim,
                          const double *box, uint64_t seed, int64_t first);
/* lattice of shape n[0..ndim) offset by `shift` cells plus a smooth sinusoidal displacement of
 * amplitude `amp` cells (Zel'dovich-like clustering); rows [first, first+npart) of the C-order lattice */
int pmb_particles_lattice(pmb_ctx *ctx, void *pos, int pos_elsize, int64_t npart, int ndim,
                          const int64_t *n, const double *box, double shift, double amp,
                          uint64_t seed, int64_t first);

/* periodic replicas of a small particle set (bench / tests): rows of block b = (b0, b1, b2), C order over
 * nrep[0..ndim), are small[j] + b * period; blocks [first_block, first_block + nblocks) are written. */
int pmb_particles_replicate(pmb_ctx *ctx, void *pos, int pos_elsize, const void *small_pos, int64_t nsmall,
                            int ndim, const int64_t *nrep, const double *period,
                            int64_t first_block, int64_t nblocks);

/* ---- domain routing ------------------------------------------------------------- */
typedef struct pmb_decompose_args {
    const void *pos;
    int pos_elsize;
    int64_t npart;
    int64_t pos_stride0, pos_stride1;  /* bytes */
    int ndim;                   /* dimensions of the domain grid (<= 3, <= columns of pos) */
    double scale[3];            /* transform: x -> scale * x (pm.py:1786-1790) */
    double smoothing[3];
    const double *edges_h;      /* concatenated edges of every axis (host) */
    int nedges[3];              /* len(edges[d]) */
    int periodic;
    const int32_t *domain_assign_h;     /* [prod(nedges-1)] domain -> rank */
    const int16_t *domain_degenerate_h; /* [prod(nedges-1)] */
    int nranks;                 /* <= 64 */
} pmb_decompose_args;

/* Phase 1: computes the per-particle target sets and counts_h[nranks]; *ntotal = sum(counts).
 * Phase 2: pmb_decompose_fill writes indices[ntotal] (int32, grouped by rank asc, particle asc).
 * Both phases must be called with the same args on the same ctx, back to back. */
int pmb_decompose_count(pmb_ctx *ctx, const pmb_decompose_args *a, int32_t *counts_h, int64_t *ntotal);
int pmb_decompose_fill(pmb_ctx *ctx, const pmb_decompose_args *a, int32_t *indices);
/* *flag = 1 when the last pmb_decompose_count found the IDENTITY layout: a single periodic domain on
 * every axis, so every particle goes exactly once, in order, to one rank (indices = arange(npart)).
 * Callers may then skip pmb_decompose_fill and pass indices = NULL to pmb_take / pmb_gather_sum. */
int pmb_decompose_identity(pmb_ctx *ctx, int *flag);

/* out[j] = data[indices[j]]  (records of itemsize bytes) <- ndarray.take(indices, axis=0), domain.py:188
 * indices == NULL: the identity layout (a copy). */
int pmb_take(pmb_ctx *ctx, const void *data, int64_t itemsize, const int32_t *indices, int64_t n, void *out);
/* out[i] = sum_j { data[j] : indices[j] == i } accumulated in ascending j, in float64, cast to out
 * <- bincountv(indices, recv, minlength=sendlength), domain.py:26-48,300. ncomp values per record.
 * offsets_h[nranks+1] delimit the per-rank sorted segments of indices.
 * indices == NULL: the identity layout, out[i] = 0.0 + data[i]. */
int pmb_gather_sum(pmb_ctx *ctx, const void *data, int data_elsize, int ncomp, const int32_t *indices,
                   const int64_t *offsets_h, int nranks, int64_t nout, void *out, int out_elsize);
/* the same reduction with one device pointer per rank segment (segments_h[r] = first record of the
 * segment of rank r; host array of nranks device pointers): lets the caller leave its own ghosts where
 * they are instead of copying them through the reverse alltoallv (domain.py:274-281 self block). */
int pmb_gather_sum_segments(pmb_ctx *ctx, const void *const *segments_h, int data_elsize, int ncomp,
                            const int32_t *indices, const int64_t *offsets_h, int nranks, int64_t nout,
                            void *out, int out_elsize);

/* out[indices[j]] += data[j] over the segments of every rank but skip_rank, in rank order (float64 out) */
int pmb_gather_add_segments(pmb_ctx *ctx, const void *const *segments_h, int data_elsize, int ncomp,
                            const int32_t *indices, const int64_t *offsets_h, int nranks, int skip_rank,
                            int64_t nout, void *out);

/* ---- communicator (NCCL over NVLink), one rank per process ------------------------ */
int pmb_comm_unique_id(char *id128_h);                      /* 128 bytes */
int pmb_comm_init_rank(pmb_ctx *ctx, const char *id128_h, int rank, int nranks);
int pmb_comm_destroy(pmb_ctx *ctx);
int pmb_comm_rank(pmb_ctx *ctx, int *rank, int *nranks);
/* alltoallv of byte records: send/recv counts and offsets in records (host arrays, int64) */
int pmb_alltoallv(pmb_ctx *ctx, const void *send, const int64_t *sendcounts_h, const int64_t *sendoffsets_h,
                  void *recv, const int64_t *recvcounts_h, const int64_t *recvoffsets_h, int64_t itemsize);
int pmb_allreduce_f64(pmb_ctx *ctx, double *buf, int64_t n, int op /*0 sum, 1 max, 2 min*/);
int pmb_allgather_bytes(pmb_ctx *ctx, const void *send, void *recv, int64_t nbytes_per_rank);
int pmb_barrier(pmb_ctx *ctx);

/* ---- FFT (slab or pencil decomposition over the ctx communicator; cuFFT for the local 1-D/2-D FFTs) ---- */
/* real layout: padded, C order, local shape (n0_local, n1, 2*(n2/2+1)) [ndim 3];
 * complex "transposed" layout for nranks > 1: distributed along axis 1, memory order (1,2,0);
 * for nranks == 1: natural order (0,1,2).  dtype_elsize 4 (float) or 8 (double). */
int pmb_fft_create(pmb_ctx *ctx, int ndim, const int64_t *nmesh, int dtype_elsize, pmb_fft **out);
/* the same with an explicit process mesh np[0] x np[1] (np[0] * np[1] = ranks; rank = c0 * np[1] + c1, the C
 * order of pfft.ProcMesh, pm.py:1319-1327).  np[1] == 1: slabs (pmb_fft_create).  np[1] > 1: pencils --
 * real space split along axes (0, 1), complex space along axes (1, 2) with memory order (1, 2, 0)
 * (PFFT_TRANSPOSED_OUT on a 2-D process mesh); two global transposes per transform, each inside one row /
 * column of the process mesh, stored straight into peer memory (needs CUDA IPC between the GPUs). */
int pmb_fft_create_np(pmb_ctx *ctx, int ndim, const int64_t *nmesh, int dtype_elsize, const int *np, pmb_fft **out);
int pmb_fft_destroy(pmb_fft *plan);
/* real-space and transposed complex-space partition of THIS rank (starts/shapes per logical axis,
 * complex strides in elements) */
int pmb_fft_layout(pmb_fft *plan, int64_t *i_start, int64_t *i_shape, int64_t *i_strides,
                   int64_t *o_start, int64_t *o_shape, int64_t *o_strides,
                   int64_t *real_alloc_elems, int64_t *complex_alloc_elems);
/* forward: complex = FFT(real) * scale  (pm.py:689-692 uses scale = 1/prod(Nmesh)); real is preserved
 * unless real == complex (in place). backward: real = unnormalised inverse FFT(complex) (pm.py:1017);
 * complex is preserved unless in place. */
int pmb_fft_r2c(pmb_fft *plan, const void *real, void *cplx, double scale);
int pmb_fft_c2r(pmb_fft *plan, const void *cplx, void *real);
/* n (<= 4) backward transforms, results equal to n pmb_fft_c2r calls.  On slab decompositions with peer-memory
 * transposes the NVLink stores of every transform run on a second stream, under the cuFFT kernels of the others
 * (the three c2r of a force evaluation, examples/nbody.py:211-213); PMB_FFT_OVERLAP=0 runs them one by one. */
int pmb_fft_c2r_multi(pmb_fft *plan, int n, const void *const *cplx_h, void *const *real_h);
/* milliseconds spent inside the transpose kernels of the distributed transforms (events on the stream) and
 * the bytes they stored into other ranks' landing buffers over NVLink since the last reset */
int pmb_fft_transpose_stats(pmb_fft *plan, float *ms, double *remote_bytes, int reset);
/* seconds spent inside cuFFT exec calls since the last reset, measured with events (library time) */
int pmb_fft_library_ms(pmb_fft *plan, float *ms, int reset);

/* out = T(k) * in on the transposed complex layout of `plan`.  kinds: */
#define PMB_TF_SCALE 0            /* params[0] */
#define PMB_TF_GRAVITY_FD4 1      /* i*kfinite_d/k^2, examples/nbody.py:162-170; dir = d */
#define PMB_TF_GRADIENT_K 2       /* i*k_d/k^2,       examples/nbody.py:154-160 */
#define PMB_TF_INV_LAPLACE 3      /* -1/k^2,          examples/nbody.py:172-175 */
#define PMB_TF_GAUSS_LOWPASS 4    /* exp(-0.5 k^2 r^2), params[0] = r, examples/nbody.py:177-181 */
#define PMB_TF_COMPENSATE 5       /* 1/prod_d fwindow(w_d), window.py:65-80; params[0] = kind, params[1] = support */
#define PMB_TF_IK 6               /* i*k_d (plain gradient) */
#define PMB_TF_POWERLAW 7         /* |k|^p, 0 at k = 0; params[0] = p (shapes white noise into P(k) ~ k^(2p)) */
int pmb_transfer(pmb_fft *plan, int kind, int dir, const double *params_h, const double *boxsize_h,
                 const void *in, void *out);
/* out = prefactor * T(k) * in: the same kernel with a scalar folded into the multiplier, so that a
 * pending normalisation of `in` (the 1/prod(Nmesh) of r2c, pm.py:692, or a `rho *= fac` before it)
 * costs no pass of its own. */
int pmb_transfer_scaled(pmb_fft *plan, int kind, int dir, const double *params_h, const double *boxsize_h,
                        double prefactor, const void *in, void *out);

/* The three gradient transfers of the force step in ONE pass over the density modes:
 * outs_h[d] = prefactor * i m_d(k_d) / k^2 * in, d = 0, 1, 2, with m_d = kfinite_d (PMB_TF_GRAVITY_FD4,
 * examples/nbody.py:162-170) or k_d (PMB_TF_GRADIENT_K, :154-160).  Values equal three pmb_transfer_scaled calls. */
int pmb_transfer_grad3(pmb_fft *plan, int kind, const double *boxsize_h, double prefactor, const void *in,
                       void *const *outs_h);

/* The three backward transforms of a force evaluation with the gradient transfers FUSED into their first pass:
 * reals_h[d] = c2r(prefactor * i m_d(k_d) / k^2 * in), d = 0, 1, 2 -- what
 * `[rhok.apply(T_d).c2r() for d in range(3)]` computes (pm.py:617-648 apply, 987-1019 c2r; examples/nbody.py:162-170,
 * 211-213), equal to pmb_transfer_grad3 + pmb_fft_c2r_multi to rounding (the axis-0 transform is this library's own
 * kernel, pmb_ifft.cuh, instead of cuFFT: the density modes are multiplied on load, no transfer pass, no separate
 * axis-0 pass).  Axis 0 a power of two in 64 .. 4096 on one rank or slabs; anything else runs the two calls above.
 * `in` is preserved and must not be one of the outputs; reals_h[d] are real-field buffers (their in-place complex
 * partners are used as work space). */
int pmb_fft_c2r_grad3(pmb_fft *plan, int kind, const double *boxsize_h, double prefactor, const void *in,
                      void *const *reals_h);
/* The whole Fourier part of a force evaluation, reals_h[d] = c2r(i m_d(k_d) / k^2 * r2c(real_in) * prefactor), d = 0, 1, 2
 * -- `[rho.r2c().apply(T_d).c2r() for d in range(3)]` (pm.py:655-694, 617-648, 987-1019; examples/nbody.py:205-213) --
 * with the LAST pass of r2c, the transfers and the FIRST pass of the three c2r in one kernel of this library
 * (pmb_ifft.cuh): cuFFT transforms the planes (axes 1, 2), the kernel takes every axis-0 line forward, multiplies and
 * takes it back in registers / shared memory; the density modes are never written to memory.  prefactor carries the
 * 1 / prod(Nmesh) of r2c (pm.py:692).  One rank, axis 0 a power of two in 64 .. 4096; otherwise PMB_EUNSUPPORTED (the
 * caller composes pmb_fft_r2c + pmb_fft_c2r_grad3).  real_in is preserved. */
int pmb_fft_force3(pmb_fft *plan, int kind, const double *boxsize_h, double prefactor, const void *real_in,
                   void *const *reals_h);
/* milliseconds inside the fused transfer + line-transform kernels and their number since the last reset */
int pmb_fft_fused_stats(pmb_fft *plan, float *ms, int64_t *launches, int reset);

/* result_h[0..1] = (re, im) of sum over the LOCAL stored half-complex modes of conj(b) * a * w, w = 2 for
 * modes that stand for themselves and their Hermitian conjugate (0 < k_last < N/2), else 1 -- the rank-local
 * term of ComplexField.cdot / cnorm (pm.py:911-974, default metric and norm); float64 accumulation. */
int pmb_cdot(pmb_fft *plan, const void *a, const void *b, double *result_h);

/* ---- white noise (initial conditions) ----------------------------------------------------------- */
/* <- pmesh._whitenoise.generate (pmesh/_whitenoise.pyx:25-45) -> pmesh_whitenoise_generator_fill
 *    (pmesh/_whitenoise_imp.c:68-105, _whitenoise_generics.h:29-232) with gsl_rng_ranlxd1
 *    (pmesh/gsl/ranlxd.c).  Fills the local block [start, start + size) of the Hermitian
 *    half-spectrum (k_z <= N/2) of an nmesh[0..3) Fourier mesh; cplx is complex64 (elsize 8) or
 *    complex128 (elsize 16) with byte strides.  The random streams are bit-identical to the
 *    reference's; values agree to the last bits of libm's log / sin / cos.  Partition invariant. */
int pmb_whitenoise(pmb_ctx *ctx, void *cplx, int elsize, const int64_t *nmesh, const int64_t *start,
                   const int64_t *size, const int64_t *strides, unsigned int seed, int unitary);

#ifdef __cplusplus
}
#endif
#endif /* PMESH_B200_H */
