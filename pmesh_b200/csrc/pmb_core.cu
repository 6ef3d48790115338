// pmb_core.cu -- context, memory, timers, window registry and elementwise field helpers.
#include "pmb_internal.h"

#include <math.h>
#include <stdlib.h>

static thread_local char g_err[1024] = "";

void pmb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int pmb_cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    pmb_set_error("CUDA error %d (%s) at %s:%d in %s", (int) e, cudaGetErrorString(e), file, line, what);
    // the runtime remembers the last error until somebody reads it: a failed allocation that the caller recovers
    // from (the host allocator empties its pool and retries) must not surface at the next kernel-launch check
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? PMB_ENOMEM : PMB_ECUDA;
}

extern "C" const char *pmb_last_error(void) { return g_err; }
extern "C" int pmb_version(void) { return 100; }

extern "C" int pmb_device_count(int *n)
{
    PMB_REQUIRE(n, "null argument");
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) {
        *n = 0;
        return pmb_cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
    }
    return PMB_OK;
}

extern "C" int pmb_ctx_create(int device, pmb_ctx **out)
{
    PMB_REQUIRE(out, "null argument");
    int n = 0;
    PMB_CUDA(cudaGetDeviceCount(&n));
    PMB_REQUIRE(device >= 0 && device < n, "device %d out of range (have %d)", device, n);
    PMB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PMB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        pmb_set_error("device %d is sm_%d%d; this library ships sm_100a code only", device, prop.major, prop.minor);
        return PMB_EUNSUPPORTED;
    }
    pmb_ctx *ctx = (pmb_ctx *) calloc(1, sizeof(pmb_ctx));
    if (!ctx) return PMB_ENOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->nranks = 1;
    ctx->det_chunk_bytes = (size_t) 2 << 30;
    PMB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (int i = 0; i < PMB_NTIMERS; i++) {
        PMB_CUDA(cudaEventCreate(&ctx->t0[i]));
        PMB_CUDA(cudaEventCreate(&ctx->t1[i]));
    }
    *out = ctx;
    return PMB_OK;
}

extern "C" int pmb_ctx_destroy(pmb_ctx *ctx)
{
    if (!ctx) return PMB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    pmb_comm_destroy(ctx);
    for (int i = 0; i < PMB_NTIMERS; i++) {
        cudaEventDestroy(ctx->t0[i]);
        cudaEventDestroy(ctx->t1[i]);
    }
    for (int k = 0; k < PMB_NKINDS; k++)
        if (ctx->tables[k].d_values) cudaFree(ctx->tables[k].d_values);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->flush_buf) cudaFree(ctx->flush_buf);
    if (ctx->barrier_token) cudaFree(ctx->barrier_token);
    if (ctx->route_masks) cudaFree(ctx->route_masks);
    if (ctx->route_blockhist) cudaFree(ctx->route_blockhist);
    if (ctx->sched_buf) cudaFree(ctx->sched_buf);
    if (ctx->perm_ids) cudaFree(ctx->perm_ids);
    if (ctx->bin_pos) cudaFree(ctx->bin_pos);
    if (ctx->bin_dest) cudaFree(ctx->bin_dest);
    if (ctx->bin_col) cudaFree(ctx->bin_col);
    if (ctx->bin_small) cudaFree(ctx->bin_small);
    if (ctx->streams_ready) {
        for (int i = 0; i < 2; i++) { cudaStreamSynchronize(ctx->copy_stream[i]); cudaStreamDestroy(ctx->copy_stream[i]); }
        for (int i = 0; i < PMB_NTIMERS; i++) cudaEventDestroy(ctx->sev[i]);
    }
    cudaStreamDestroy(ctx->stream);
    free(ctx);
    return PMB_OK;
}

// ---- copy streams: overlap of PCIe transfers with compute ------------------------------------------
// Stream ids of this API: 0 = the compute stream every kernel of the context runs on, 1 = host -> device
// copies, 2 = device -> host copies (PCIe is full duplex: an upload and a download can run together and
// under a kernel).  Ordering between streams is explicit: record an event on one, wait for it on another.
static int streams_init(pmb_ctx *ctx)
{
    if (ctx->streams_ready) return PMB_OK;
    PMB_CUDA(cudaSetDevice(ctx->device));
    for (int i = 0; i < 2; i++) PMB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream[i], cudaStreamNonBlocking));
    for (int i = 0; i < PMB_NTIMERS; i++) PMB_CUDA(cudaEventCreateWithFlags(&ctx->sev[i], cudaEventDisableTiming));
    ctx->streams_ready = 1;
    return PMB_OK;
}
static int stream_of(pmb_ctx *ctx, int id, cudaStream_t *s)
{
    PMB_REQUIRE(id >= 0 && id <= 2, "stream id %d (0 compute, 1 upload, 2 download)", id);
    PMB_CHECK(streams_init(ctx));
    *s = id == 0 ? ctx->stream : ctx->copy_stream[id - 1];
    return PMB_OK;
}

extern "C" int pmb_memcpy_h2d_async(pmb_ctx *ctx, void *dst, const void *src_h, size_t nbytes)
{
    PMB_REQUIRE(ctx && (nbytes == 0 || (dst && src_h)), "null argument");
    cudaStream_t s;
    PMB_CHECK(stream_of(ctx, 1, &s));
    if (nbytes) PMB_CUDA(cudaMemcpyAsync(dst, src_h, nbytes, cudaMemcpyHostToDevice, s));
    return PMB_OK;
}

extern "C" int pmb_memcpy_d2h_async(pmb_ctx *ctx, void *dst_h, const void *src, size_t nbytes)
{
    PMB_REQUIRE(ctx && (nbytes == 0 || (dst_h && src)), "null argument");
    cudaStream_t s;
    PMB_CHECK(stream_of(ctx, 2, &s));
    if (nbytes) PMB_CUDA(cudaMemcpyAsync(dst_h, src, nbytes, cudaMemcpyDeviceToHost, s));
    return PMB_OK;
}

extern "C" int pmb_stream_record(pmb_ctx *ctx, int stream_id, int event)
{
    PMB_REQUIRE(ctx && event >= 0 && event < PMB_NTIMERS, "bad event slot");
    cudaStream_t s;
    PMB_CHECK(stream_of(ctx, stream_id, &s));
    PMB_CUDA(cudaEventRecord(ctx->sev[event], s));
    return PMB_OK;
}

extern "C" int pmb_stream_wait(pmb_ctx *ctx, int stream_id, int event)
{
    PMB_REQUIRE(ctx && event >= 0 && event < PMB_NTIMERS, "bad event slot");
    cudaStream_t s;
    PMB_CHECK(stream_of(ctx, stream_id, &s));
    PMB_CUDA(cudaStreamWaitEvent(s, ctx->sev[event], 0));
    return PMB_OK;
}

extern "C" int pmb_stream_sync(pmb_ctx *ctx, int stream_id)
{
    PMB_REQUIRE(ctx, "null context");
    cudaStream_t s;
    PMB_CHECK(stream_of(ctx, stream_id, &s));
    PMB_CUDA(cudaStreamSynchronize(s));
    return PMB_OK;
}

extern "C" int pmb_ctx_sync(pmb_ctx *ctx)
{
    PMB_REQUIRE(ctx, "null context");
    PMB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PMB_OK;
}

extern "C" int pmb_malloc(pmb_ctx *ctx, size_t nbytes, void **out)
{
    PMB_REQUIRE(ctx && out, "null argument");
    *out = NULL;
    if (nbytes == 0) nbytes = 16;
    PMB_CUDA(cudaMalloc(out, nbytes));
    return PMB_OK;
}

extern "C" int pmb_free(pmb_ctx *ctx, void *ptr)
{
    PMB_REQUIRE(ctx, "null context");
    if (ptr) {
        PMB_CUDA(cudaStreamSynchronize(ctx->stream));
        PMB_CUDA(cudaFree(ptr));
    }
    return PMB_OK;
}

extern "C" int pmb_malloc_host(pmb_ctx *ctx, size_t nbytes, void **out_h)
{
    PMB_REQUIRE(ctx && out_h, "null argument");
    if (nbytes == 0) nbytes = 16;
    PMB_CUDA(cudaMallocHost(out_h, nbytes));
    return PMB_OK;
}

extern "C" int pmb_free_host(pmb_ctx *ctx, void *ptr_h)
{
    PMB_REQUIRE(ctx, "null context");
    if (ptr_h) PMB_CUDA(cudaFreeHost(ptr_h));
    return PMB_OK;
}

extern "C" int pmb_memcpy_h2d(pmb_ctx *ctx, void *dst, const void *src_h, size_t nbytes)
{
    PMB_REQUIRE(ctx, "null context");
    if (!nbytes) return PMB_OK;
    PMB_CUDA(cudaMemcpyAsync(dst, src_h, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    PMB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PMB_OK;
}

extern "C" int pmb_memcpy_d2h(pmb_ctx *ctx, void *dst_h, const void *src, size_t nbytes)
{
    PMB_REQUIRE(ctx, "null context");
    if (!nbytes) return PMB_OK;
    PMB_CUDA(cudaMemcpyAsync(dst_h, src, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    PMB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PMB_OK;
}

extern "C" int pmb_memcpy_d2d(pmb_ctx *ctx, void *dst, const void *src, size_t nbytes)
{
    PMB_REQUIRE(ctx, "null context");
    if (!nbytes) return PMB_OK;
    PMB_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return PMB_OK;
}

extern "C" int pmb_memset(pmb_ctx *ctx, void *dst, int byte, size_t nbytes)
{
    PMB_REQUIRE(ctx, "null context");
    if (!nbytes) return PMB_OK;
    PMB_CUDA(cudaMemsetAsync(dst, byte, nbytes, ctx->stream));
    return PMB_OK;
}

extern "C" int pmb_mem_info(pmb_ctx *ctx, size_t *free_bytes, size_t *total_bytes)
{
    PMB_REQUIRE(ctx && free_bytes && total_bytes, "null argument");
    PMB_CUDA(cudaMemGetInfo(free_bytes, total_bytes));
    return PMB_OK;
}

extern "C" int pmb_timer_start(pmb_ctx *ctx, int slot)
{
    PMB_REQUIRE(ctx && slot >= 0 && slot < PMB_NTIMERS, "bad timer slot");
    PMB_CUDA(cudaEventRecord(ctx->t0[slot], ctx->stream));
    return PMB_OK;
}

extern "C" int pmb_timer_stop(pmb_ctx *ctx, int slot, float *ms)
{
    PMB_REQUIRE(ctx && ms && slot >= 0 && slot < PMB_NTIMERS, "bad timer slot");
    PMB_CUDA(cudaEventRecord(ctx->t1[slot], ctx->stream));
    PMB_CUDA(cudaEventSynchronize(ctx->t1[slot]));
    PMB_CUDA(cudaEventElapsedTime(ms, ctx->t0[slot], ctx->t1[slot]));
    return PMB_OK;
}

extern "C" int pmb_launch_count(pmb_ctx *ctx, int64_t *n, int reset)
{
    PMB_REQUIRE(ctx && n, "null argument");
    *n = ctx->launches;
    if (reset) ctx->launches = 0;
    return PMB_OK;
}

extern "C" int pmb_set_workspace_limit(pmb_ctx *ctx, size_t nbytes)
{
    PMB_REQUIRE(ctx && nbytes >= 4096, "workspace limit too small");
    ctx->det_chunk_bytes = nbytes;
    return PMB_OK;
}

// work space of the library itself (scratch, sorted particle copies): when the device is full the host side is
// asked to give cached blocks back (pmb_set_trim_callback), then the allocation is tried once more
cudaError_t pmb_work_alloc(pmb_ctx *ctx, void **out, size_t nbytes)
{
    cudaError_t e = cudaMalloc(out, nbytes);
    if (e == cudaErrorMemoryAllocation && ctx->trim_cb) {
        cudaGetLastError();
        ctx->trim_cb(ctx->trim_arg);
        e = cudaMalloc(out, nbytes);
    }
    return e;
}

extern "C" int pmb_set_trim_callback(pmb_ctx *ctx, void (*cb)(void *), void *arg)
{
    PMB_REQUIRE(ctx, "null context");
    ctx->trim_cb = cb;
    ctx->trim_arg = arg;
    return PMB_OK;
}

int pmb_scratch(pmb_ctx *ctx, size_t nbytes, void **out)
{
    if (nbytes > ctx->scratch_bytes) {
        if (ctx->scratch) {
            PMB_CUDA(cudaStreamSynchronize(ctx->stream));
            PMB_CUDA(cudaFree(ctx->scratch));
            ctx->scratch = NULL;
            ctx->scratch_bytes = 0;
        }
        size_t want = nbytes + (nbytes >> 3) + 256;
        PMB_CUDA(pmb_work_alloc(ctx, &ctx->scratch, want));
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return PMB_OK;
}

__global__ void pmb_k_flush(uint4 *buf, size_t n, unsigned v)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = make_uint4(v, v + 1, v + 2, v + 3);
}

extern "C" int pmb_flush_l2(pmb_ctx *ctx)
{
    PMB_REQUIRE(ctx, "null context");
    if (!ctx->flush_buf) {
        ctx->flush_bytes = (size_t) 256 << 20;
        PMB_CUDA(cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
    }
    static unsigned gen = 0;
    size_t n = ctx->flush_bytes / sizeof(uint4);
    pmb_k_flush<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((uint4 *) ctx->flush_buf, n, gen++);
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}

// ---------------------------------------------------------------- window registry
extern "C" int pmb_window_set_table(pmb_ctx *ctx, int kind, const double *values_h, int n,
                                    double step, double nativesupport, double hsupport)
{
    PMB_REQUIRE(ctx && values_h && n > 1, "bad table");
    int family, native, tuned;
    PMB_REQUIRE(pmb_kind_info(kind, &family, &native, &tuned) == 0, "unknown window kind %d", kind);
    PMB_REQUIRE(family == PMB_FAM_SYMTABLE || family == PMB_FAM_WAVELET, "kind %d is not table driven", kind);
    PMB_REQUIRE((int) nativesupport == native, "table support %g != %d for kind %d", nativesupport, native, kind);
    pmb_table *t = &ctx->tables[kind];
    if (t->d_values) {
        PMB_CUDA(cudaStreamSynchronize(ctx->stream));
        PMB_CUDA(cudaFree(t->d_values));
        t->d_values = NULL;
    }
    PMB_CUDA(cudaMalloc(&t->d_values, sizeof(double) * n));
    PMB_CUDA(cudaMemcpy(t->d_values, values_h, sizeof(double) * n, cudaMemcpyHostToDevice));
    t->n = n;
    t->step = step;
    t->nativesupport = nativesupport;
    t->hsupport = hsupport;
    return PMB_OK;
}

// <- pmesh_painter_init (pmesh/_window_imp.c:246-456); attaches the device lookup table
int pmb_resolve_window(pmb_ctx *ctx, int kind, int support_req, int ndim, const int *order,
                       PmbWindow *w, int for_device)
{
    PMB_REQUIRE(pmb_window_resolve(kind, support_req, ndim, order, w) == 0, "unknown window kind %d", kind);
    if ((w->family == PMB_FAM_SYMTABLE || w->family == PMB_FAM_WAVELET) && for_device) {
        PMB_REQUIRE(ctx && ctx->tables[kind].d_values, "lookup table of window kind %d was not uploaded", kind);
        w->table = ctx->tables[kind].d_values;
        w->tablesize = ctx->tables[kind].n;
        w->step = ctx->tables[kind].step;
        w->hsupport = ctx->tables[kind].hsupport;
    }
    return PMB_OK;
}

extern "C" int pmb_window_query(int kind, int support_req, int *support, int *nativesupport)
{
    PmbWindow w;
    PMB_CHECK(pmb_resolve_window(NULL, kind, support_req, 0, NULL, &w, 0));
    if (support) *support = w.support;
    if (nativesupport) *nativesupport = w.nativesupport;
    return PMB_OK;
}

static double sinc_unnormed(double x)
{
    if (x < 1e-5 && x > -1e-5) {
        double x2 = x * x;
        return 1.0 - x2 / 6. + x2 * x2 / 120.;
    }
    return sin(x) / x;
}

// <- pmesh_painter_get_fwindow (pmesh/_window_imp.c:473-484): sinc^p(w/2/vfactor), 1.0 for table windows
extern "C" int pmb_window_fwindow(int kind, int support, const double *w_h, double *out_h, int64_t n)
{
    PMB_REQUIRE(n == 0 || (w_h && out_h), "null argument");
    PmbWindow w;
    PMB_CHECK(pmb_resolve_window(NULL, kind, support, 0, NULL, &w, 0));
    PmbWinInfo info;
    pmb_window_info(w.nativesupport, (double) support, &info);
    int p = 0;
    if (w.family == PMB_FAM_NEAREST) p = 1;
    else if (w.family == PMB_FAM_LINEAR) p = 2;
    else if (w.family == PMB_FAM_QUADRATIC) p = 3;
    else if (w.family == PMB_FAM_CUBIC) p = 4;
    for (int64_t i = 0; i < n; i++) {
        if (!p) { out_h[i] = 1.0; continue; }
        double t = sinc_unnormed(0.5 * (w_h[i] / info.vfactor));
        double r = t;
        for (int j = 1; j < p; j++) r = r * t;
        out_h[i] = r;
    }
    return PMB_OK;
}

// ---------------------------------------------------------------- field helpers
struct FieldView {
    int64_t size[3];
    int64_t strides[3];
    int64_t n;
};

static int make_view(int ndim, const int64_t *size, const int64_t *strides, FieldView *v)
{
    PMB_REQUIRE(ndim >= 1 && ndim <= 3 && size && strides, "bad field view");
    // left-pad to 3-D
    int pad = 3 - ndim;
    v->n = 1;
    for (int d = 0; d < 3; d++) {
        if (d < pad) { v->size[d] = 1; v->strides[d] = 0; }
        else { v->size[d] = size[d - pad]; v->strides[d] = strides[d - pad]; }
        v->n *= v->size[d];
    }
    return PMB_OK;
}

__device__ __forceinline__ int64_t view_offset(const FieldView &v, int64_t i)
{
    int64_t k = i % v.size[2];
    int64_t r = i / v.size[2];
    int64_t j = r % v.size[1];
    int64_t a = r / v.size[1];
    return a * v.strides[0] + j * v.strides[1] + k * v.strides[2];
}

static bool view_is_dense(const FieldView &v, int elsize)
{
    int64_t acc = elsize;
    for (int d = 2; d >= 0; d--) {
        if (v.size[d] != 1 && v.strides[d] != acc) return false;
        acc *= v.size[d];
    }
    return true;
}

template <typename T>
__global__ void pmb_k_scale_flat(T *__restrict__ a, int64_t n, T factor)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) a[i] = (T) (a[i] * factor);
}

template <typename T>
__global__ void pmb_k_fill(char *mesh, FieldView v, T value)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; i < v.n; i += stride) *(T *) (mesh + view_offset(v, i)) = value;
}

template <typename T>
__global__ void pmb_k_scale(char *mesh, FieldView v, double factor)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; i < v.n; i += stride) {
        T *p = (T *) (mesh + view_offset(v, i));
        *p = (T) (*p * (T) factor);
    }
}

extern "C" int pmb_field_fill(pmb_ctx *ctx, void *mesh, int elsize, int ndim, const int64_t *size,
                              const int64_t *strides, double value)
{
    PMB_REQUIRE(ctx && mesh, "null argument");
    PMB_REQUIRE(elsize == 4 || elsize == 8, "elsize must be 4 or 8");
    FieldView v;
    PMB_CHECK(make_view(ndim, size, strides, &v));
    if (v.n == 0) return PMB_OK;
    if (value == 0.0 && view_is_dense(v, elsize)) {
        PMB_CUDA(cudaMemsetAsync(mesh, 0, (size_t) v.n * elsize, ctx->stream));
        return PMB_OK;
    }
    int grid = pmb_grid(ctx, v.n, 256, 8);
    if (elsize == 8) pmb_k_fill<double><<<grid, 256, 0, ctx->stream>>>((char *) mesh, v, value);
    else pmb_k_fill<float><<<grid, 256, 0, ctx->stream>>>((char *) mesh, v, (float) value);
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}

// numpy semantics of `value[...] *= factor`: real fields multiply in their own dtype
// (pm.py:692 multiplies a complex array by a python float: component-wise in the field dtype)
extern "C" int pmb_field_scale(pmb_ctx *ctx, void *mesh, int elsize, int is_complex, int ndim,
                               const int64_t *size, const int64_t *strides, double factor)
{
    PMB_REQUIRE(ctx && mesh, "null argument");
    PMB_REQUIRE(elsize == 4 || elsize == 8, "elsize must be 4 or 8 (per real component)");
    PMB_REQUIRE(ndim >= 1 && ndim <= 3, "bad ndim");
    int64_t sz[3], st[3];
    for (int d = 0; d < ndim; d++) { sz[d] = size[d]; st[d] = strides[d]; }
    if (is_complex) {
        // view the complex array as reals with the last axis doubled (requires a contiguous last axis)
        PMB_REQUIRE(st[ndim - 1] == 2 * elsize || ndim < 3,
                    "complex scale needs a contiguous last axis or ndim < 3");
        if (st[ndim - 1] == 2 * elsize) {
            sz[ndim - 1] *= 2;
            st[ndim - 1] = elsize;
        } else {
            // prepend nothing: add a trailing axis of 2 components
            sz[ndim] = 2; st[ndim] = elsize; ndim += 1;
        }
    }
    FieldView v;
    PMB_CHECK(make_view(ndim, sz, st, &v));
    if (v.n == 0) return PMB_OK;
    int grid = pmb_grid(ctx, v.n, 256, 8);
    if (view_is_dense(v, elsize)) {
        if (elsize == 8) pmb_k_scale_flat<double><<<grid, 256, 0, ctx->stream>>>((double *) mesh, v.n, factor);
        else pmb_k_scale_flat<float><<<grid, 256, 0, ctx->stream>>>((float *) mesh, v.n, (float) factor);
    } else if (elsize == 8) pmb_k_scale<double><<<grid, 256, 0, ctx->stream>>>((char *) mesh, v, factor);
    else pmb_k_scale<float><<<grid, 256, 0, ctx->stream>>>((char *) mesh, v, factor);
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}

template <typename T>
__global__ void pmb_k_sum(const char *mesh, FieldView v, double *out)
{
    __shared__ double sh[32];
    double acc = 0;
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; i < v.n; i += stride) acc += (double) *(const T *) (mesh + view_offset(v, i));
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        acc = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[blockIdx.x] = acc;
    }
}

// sum_i a[i] * b[i] over two views of the same shape and strides (RealField.cdot / cnorm, pm.py:897-905)
template <typename T>
__global__ void pmb_k_fdot(const char *a, const char *b, FieldView v, double *out)
{
    __shared__ double sh[32];
    double acc = 0;
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; i < v.n; i += stride) {
        const int64_t o = view_offset(v, i);
        acc += (double) *(const T *) (a + o) * (double) *(const T *) (b + o);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        acc = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[blockIdx.x] = acc;
    }
}

extern "C" int pmb_field_dot(pmb_ctx *ctx, const void *a, const void *b, int elsize, int ndim, const int64_t *size,
                             const int64_t *strides, double *dot_h)
{
    PMB_REQUIRE(ctx && a && b && dot_h, "null argument");
    PMB_REQUIRE(elsize == 4 || elsize == 8, "elsize must be 4 or 8");
    FieldView v;
    PMB_CHECK(make_view(ndim, size, strides, &v));
    *dot_h = 0;
    if (v.n == 0) return PMB_OK;
    int grid = pmb_grid(ctx, v.n, 256, 4);
    void *partial;
    PMB_CHECK(pmb_scratch(ctx, sizeof(double) * grid, &partial));
    if (elsize == 8) pmb_k_fdot<double><<<grid, 256, 0, ctx->stream>>>((const char *) a, (const char *) b, v, (double *) partial);
    else pmb_k_fdot<float><<<grid, 256, 0, ctx->stream>>>((const char *) a, (const char *) b, v, (double *) partial);
    PMB_LAUNCH_CHECK(ctx);
    double *h = (double *) malloc(sizeof(double) * grid);
    if (!h) return PMB_ENOMEM;
    cudaError_t e = cudaMemcpyAsync(h, partial, sizeof(double) * grid, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { free(h); return pmb_cuda_fail(e, "dot copy", __FILE__, __LINE__); }
    double sum = 0;
    for (int i = 0; i < grid; i++) sum += h[i];
    free(h);
    *dot_h = sum;
    return PMB_OK;
}

// not bit-reproducible against numpy's pairwise sum; used for csum/cmean style diagnostics (pm.py:725-743)
extern "C" int pmb_field_sum(pmb_ctx *ctx, const void *mesh, int elsize, int ndim, const int64_t *size,
                             const int64_t *strides, double *sum_h)
{
    PMB_REQUIRE(ctx && mesh && sum_h, "null argument");
    PMB_REQUIRE(elsize == 4 || elsize == 8, "elsize must be 4 or 8");
    FieldView v;
    PMB_CHECK(make_view(ndim, size, strides, &v));
    *sum_h = 0;
    if (v.n == 0) return PMB_OK;
    int grid = pmb_grid(ctx, v.n, 256, 4);
    void *partial;
    PMB_CHECK(pmb_scratch(ctx, sizeof(double) * grid, &partial));
    if (elsize == 8) pmb_k_sum<double><<<grid, 256, 0, ctx->stream>>>((const char *) mesh, v, (double *) partial);
    else pmb_k_sum<float><<<grid, 256, 0, ctx->stream>>>((const char *) mesh, v, (double *) partial);
    PMB_LAUNCH_CHECK(ctx);
    double *h = (double *) malloc(sizeof(double) * grid);
    if (!h) return PMB_ENOMEM;
    cudaError_t e = cudaMemcpyAsync(h, partial, sizeof(double) * grid, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { free(h); return pmb_cuda_fail(e, "sum copy", __FILE__, __LINE__); }
    double s = 0;
    for (int i = 0; i < grid; i++) s += h[i];
    free(h);
    *sum_h = s;
    return PMB_OK;
}

// ---------------------------------------------------------------- synthetic particles
// splitmix64 counter hash -> uniform double in [0,1): reproducible for any launch geometry.
__host__ __device__ static inline uint64_t pmb_mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ static inline double pmb_u01(uint64_t seed, uint64_t i, uint64_t d)
{
    uint64_t h = pmb_mix64(pmb_mix64(seed ^ (d * 0xD6E8FEB86659FD93ull)) + i);
    return (double) (h >> 11) * (1.0 / 9007199254740992.0);
}

__global__ void pmb_k_uniform(void *pos, int elsize, int64_t npart, int ndim, double b0, double b1, double b2,
                              uint64_t seed, int64_t first)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    int64_t stride = (int64_t) gridDim.x * blockDim.x;
    double box[3] = {b0, b1, b2};
    for (; i < npart; i += stride)
        for (int d = 0; d < ndim; d++) {
            double x = box[d] * pmb_u01(seed, (uint64_t) (i + first), d);
            if (elsize == 4) {
                float xf = (float) x;
                if (xf >= (float) box[d]) xf = 0.f;
                ((float *) pos)[i * ndim + d] = xf;
            } else ((double *) pos)[i * ndim + d] = x;
        }
}

extern "C" int pmb_particles_uniform(pmb_ctx *ctx, void *pos, int pos_elsize, int64_t npart, int ndim,
                                     const double *box, uint64_t seed, int64_t first)
{
    PMB_REQUIRE(ctx && pos && box, "null argument");
    PMB_REQUIRE(ndim >= 1 && ndim <= 3 && (pos_elsize == 4 || pos_elsize == 8), "bad ndim/elsize");
    if (!npart) return PMB_OK;
    double b[3] = {box[0], ndim > 1 ? box[1] : 0, ndim > 2 ? box[2] : 0};
    pmb_k_uniform<<<pmb_grid(ctx, npart, 256, 8), 256, 0, ctx->stream>>>(pos, pos_elsize, npart, ndim, b[0], b[1], b[2], seed, first);
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}

struct LatticeArgs {
    int ndim;
    int64_t n[3];
    double box[3];
    double shift, amp;
    double phase[3][3];
};

// x_d = (q_d + shift + amp * sum_e sin(2 pi m (q_e + shift)/n_e + phase_de) / ndim) * box_d / n_d  (mod box)
__global__ void pmb_k_lattice(void *pos, int elsize, int64_t npart, LatticeArgs a, int64_t first)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; i < npart; i += stride) {
        int64_t r = i + first;
        int64_t q[3] = {0, 0, 0};
        for (int d = a.ndim - 1; d >= 0; d--) { q[d] = r % a.n[d]; r /= a.n[d]; }
        for (int d = 0; d < a.ndim; d++) {
            double disp = 0;
            for (int e = 0; e < a.ndim; e++)
                disp += sin(6.283185307179586 * 4.0 * ((double) q[e] + a.shift) / (double) a.n[e] + a.phase[d][e]);
            double g = (double) q[d] + a.shift + a.amp * disp / a.ndim;
            double nn = (double) a.n[d];
            g = g - floor(g / nn) * nn;
            if (g >= nn) g = 0;
            double x = g * (a.box[d] / nn);
            if (elsize == 4) {
                float xf = (float) x;
                if (xf >= (float) a.box[d]) xf = 0.f;
                ((float *) pos)[i * a.ndim + d] = xf;
            } else ((double *) pos)[i * a.ndim + d] = x;
        }
    }
}

extern "C" int pmb_particles_lattice(pmb_ctx *ctx, void *pos, int pos_elsize, int64_t npart, int ndim,
                                     const int64_t *n, const double *box, double shift, double amp,
                                     uint64_t seed, int64_t first)
{
    PMB_REQUIRE(ctx && pos && box && n, "null argument");
    PMB_REQUIRE(ndim >= 1 && ndim <= 3 && (pos_elsize == 4 || pos_elsize == 8), "bad ndim/elsize");
    if (!npart) return PMB_OK;
    LatticeArgs a;
    memset(&a, 0, sizeof(a));
    a.ndim = ndim;
    for (int d = 0; d < ndim; d++) { a.n[d] = n[d]; a.box[d] = box[d]; }
    a.shift = shift;
    a.amp = amp;
    for (int d = 0; d < 3; d++)
        for (int e = 0; e < 3; e++) a.phase[d][e] = 6.283185307179586 * pmb_u01(seed, d * 3 + e, 7);
    pmb_k_lattice<<<pmb_grid(ctx, npart, 256, 8), 256, 0, ctx->stream>>>(pos, pos_elsize, npart, a, first);
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}

// periodic replicas of a small particle set: block b = (b0, b1, b2) (C order over nrep) holds
// small[j] + b * period, for the blocks [first_block, first_block + nblocks).  The field of the tiled
// problem is the periodic repetition of the small one, so the force on replica particle (b, j) equals
// the force on particle j of the small problem: bench.py checks the full-size pipeline against the CPU
// oracle of the small problem this way.
__global__ void pmb_k_replicate(void *pos, int elsize, const void *small, int64_t nsmall, int ndim,
                                int64_t r1, int64_t r2, double p0, double p1, double p2,
                                int64_t first_block, int64_t nblocks)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int64_t n = nblocks * nsmall;
    const double per[3] = {p0, p1, p2};
    for (; i < n; i += stride) {
        const int64_t bl = i / nsmall;
        const int64_t j = i - bl * nsmall;
        int64_t b = bl + first_block;
        int64_t bb[3];
        bb[2] = b % r2; b /= r2;
        bb[1] = b % r1; b /= r1;
        bb[0] = b;
        for (int d = 0; d < ndim; d++) {
            const int64_t bd = bb[3 - ndim + d];
            if (elsize == 8) ((double *) pos)[i * ndim + d] = ((const double *) small)[j * ndim + d] + (double) bd * per[d];
            else ((float *) pos)[i * ndim + d] = (float) ((double) ((const float *) small)[j * ndim + d] + (double) bd * per[d]);
        }
    }
}

extern "C" int pmb_particles_replicate(pmb_ctx *ctx, void *pos, int pos_elsize, const void *small, int64_t nsmall,
                                       int ndim, const int64_t *nrep, const double *period,
                                       int64_t first_block, int64_t nblocks)
{
    PMB_REQUIRE(ctx && pos && small && nrep && period, "null argument");
    PMB_REQUIRE(ndim >= 1 && ndim <= 3 && (pos_elsize == 4 || pos_elsize == 8), "bad ndim/elsize");
    if (nsmall <= 0 || nblocks <= 0) return PMB_OK;
    int64_t r[3] = {1, 1, 1};
    double p[3] = {0, 0, 0};
    for (int d = 0; d < ndim; d++) { r[3 - ndim + d] = nrep[d]; p[d] = period[d]; }
    pmb_k_replicate<<<pmb_grid(ctx, nsmall * nblocks, 256, 8), 256, 0, ctx->stream>>>(
        pos, pos_elsize, small, nsmall, ndim, r[1], r[2], p[0], p[1], p[2], first_block, nblocks);
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}

// ---- particle columns: the element-wise updates of a KDK (leap-frog) step ---------------------------
// The reference's integrator (examples/nbody.py:84-102, `symp2`) is numpy in place: V += F * K;
// S += V * D; X = S + Q.  Same arithmetic, same rounding order (one multiply, one add per element, no
// FMA: the library is built -fmad=false), on device-resident columns of any byte stride.
template <typename T>
__global__ void __launch_bounds__(256)
pmb_k_lincomb(char *out, int64_t so, const char *x, int64_t sx, const char *y, int64_t sy, double a, double b, int mode, int64_t n)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const T ta = (T) a, tb = (T) b;
    for (; i < n; i += stride) {
        const T xv = *(const T *) (x + i * sx);
        T r;
        if (mode == 0) r = *(T *) (out + i * so) + xv * ta;                       // out += a x
        else if (mode == 1) r = xv * ta + *(const T *) (y + i * sy) * tb;         // out = a x + b y
        else r = xv * ta;                                                         // out = a x
        *(T *) (out + i * so) = r;
    }
}

// V += F * kick; S += V * drift in one pass.  V, S: (npart, NCOL) rows (like positions); the force is
// held column-wise, F[d] dense (npart,), the way readout / gather produce it.  S may be NULL (kick only).
struct KdCols { const void *f[3]; };
template <typename T, int NCOL>
__global__ void __launch_bounds__(256)
pmb_k_kick_drift(T *__restrict__ V, T *__restrict__ S, KdCols F, double kick, double drift, int64_t npart)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int64_t n = npart * NCOL;
    const T tk = (T) kick, td = (T) drift;
    for (; i < n; i += stride) {
        const int64_t p = i / NCOL;
        const int d = (int) (i - p * NCOL);
        const T f = __ldcs((const T *) F.f[d] + p);
        const T v = V[i] + f * tk;
        V[i] = v;
        if (S) S[i] = S[i] + v * td;
    }
}

static int lincomb(pmb_ctx *ctx, void *out, int64_t so, const void *x, int64_t sx, const void *y, int64_t sy,
                   double a, double b, int mode, int elsize, int64_t n)
{
    PMB_REQUIRE(ctx && n >= 0, "bad arguments");
    PMB_REQUIRE(elsize == 4 || elsize == 8, "columns must be float32 or float64");
    if (n == 0) return PMB_OK;
    PMB_REQUIRE(out && x && (mode != 1 || y), "null column");
    const int grid = pmb_grid(ctx, n, 256, 8);
    if (elsize == 8) pmb_k_lincomb<double><<<grid, 256, 0, ctx->stream>>>((char *) out, so, (const char *) x, sx, (const char *) y, sy, a, b, mode, n);
    else pmb_k_lincomb<float><<<grid, 256, 0, ctx->stream>>>((char *) out, so, (const char *) x, sx, (const char *) y, sy, a, b, mode, n);
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}

extern "C" int pmb_axpy(pmb_ctx *ctx, void *y, int64_t y_stride, const void *x, int64_t x_stride, double a, int elsize, int64_t n)
{
    return lincomb(ctx, y, y_stride, x, x_stride, NULL, 0, a, 0.0, 0, elsize, n);
}

extern "C" int pmb_lincomb(pmb_ctx *ctx, void *out, int64_t out_stride, const void *x, int64_t x_stride, double a,
                           const void *y, int64_t y_stride, double b, int elsize, int64_t n)
{
    return lincomb(ctx, out, out_stride, x, x_stride, y, y_stride, a, b, y ? 1 : 2, elsize, n);
}

// x = x mod period, numpy's floored modulo (x - floor(x / period) * period, result in [0, period]):
// the periodic wrap `X % BoxSize` a driver applies to positions (examples/nbody.py:200 area)
template <typename T>
__global__ void __launch_bounds__(256)
pmb_k_column_mod(char *x, int64_t sx, double period, int64_t n)
{
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const double v = (double) *(T *) (x + i * sx);
        double r = fmod(v, period);
        if (r != 0 && ((r < 0) != (period < 0))) r += period;
        *(T *) (x + i * sx) = (T) r;
    }
}

extern "C" int pmb_column_mod(pmb_ctx *ctx, void *x, int64_t x_stride, double period, int elsize, int64_t n)
{
    PMB_REQUIRE(ctx && n >= 0, "bad arguments");
    PMB_REQUIRE(elsize == 4 || elsize == 8, "columns must be float32 or float64");
    PMB_REQUIRE(period != 0, "zero period");
    if (n == 0) return PMB_OK;
    PMB_REQUIRE(x, "null column");
    const int grid = pmb_grid(ctx, n, 256, 8);
    if (elsize == 8) pmb_k_column_mod<double><<<grid, 256, 0, ctx->stream>>>((char *) x, x_stride, period, n);
    else pmb_k_column_mod<float><<<grid, 256, 0, ctx->stream>>>((char *) x, x_stride, period, n);
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}

template <typename T>
static void kick_drift_launch(pmb_ctx *ctx, void *V, void *S, const KdCols &F, int ncol, double kick, double drift, int64_t npart)
{
    const int grid = pmb_grid(ctx, npart * ncol, 256, 8);
    if (ncol == 1) pmb_k_kick_drift<T, 1><<<grid, 256, 0, ctx->stream>>>((T *) V, (T *) S, F, kick, drift, npart);
    else if (ncol == 2) pmb_k_kick_drift<T, 2><<<grid, 256, 0, ctx->stream>>>((T *) V, (T *) S, F, kick, drift, npart);
    else pmb_k_kick_drift<T, 3><<<grid, 256, 0, ctx->stream>>>((T *) V, (T *) S, F, kick, drift, npart);
}

extern "C" int pmb_kick_drift(pmb_ctx *ctx, void *V, void *S, const void *const *F_cols_h, int ncol,
                              double kick, double drift, int elsize, int64_t npart)
{
    PMB_REQUIRE(ctx && npart >= 0, "bad arguments");
    PMB_REQUIRE(elsize == 4 || elsize == 8, "columns must be float32 or float64");
    PMB_REQUIRE(ncol >= 1 && ncol <= 3, "1..3 columns");
    if (npart == 0) return PMB_OK;
    PMB_REQUIRE(V && F_cols_h, "null column");
    KdCols F;
    for (int d = 0; d < 3; d++) F.f[d] = d < ncol ? F_cols_h[d] : NULL;
    for (int d = 0; d < ncol; d++) PMB_REQUIRE(F.f[d], "null force column %d", d);
    if (elsize == 8) kick_drift_launch<double>(ctx, V, S, F, ncol, kick, drift, npart);
    else kick_drift_launch<float>(ctx, V, S, F, ncol, kick, drift, npart);
    PMB_LAUNCH_CHECK(ctx);
    return PMB_OK;
}

// sum_i x[i] * y[i] in float64 (diagnostics: rms of a column, cdot-style reductions); strided columns
template <typename T>
__global__ void __launch_bounds__(256)
pmb_k_dot(const char *x, int64_t sx, const char *y, int64_t sy, int64_t n, double *out)
{
    __shared__ double sh[32];
    double acc = 0;
    int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) acc += (double) *(const T *) (x + i * sx) * (double) *(const T *) (y + i * sy);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        acc = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[blockIdx.x] = acc;
    }
}

extern "C" int pmb_dot(pmb_ctx *ctx, const void *x, int64_t x_stride, const void *y, int64_t y_stride, int elsize,
                       int64_t n, double *dot_h)
{
    PMB_REQUIRE(ctx && dot_h && n >= 0, "bad arguments");
    PMB_REQUIRE(elsize == 4 || elsize == 8, "columns must be float32 or float64");
    *dot_h = 0;
    if (n == 0) return PMB_OK;
    PMB_REQUIRE(x && y, "null column");
    const int grid = pmb_grid(ctx, n, 256, 4);
    void *partial;
    PMB_CHECK(pmb_scratch(ctx, sizeof(double) * grid, &partial));
    if (elsize == 8) pmb_k_dot<double><<<grid, 256, 0, ctx->stream>>>((const char *) x, x_stride, (const char *) y, y_stride, n, (double *) partial);
    else pmb_k_dot<float><<<grid, 256, 0, ctx->stream>>>((const char *) x, x_stride, (const char *) y, y_stride, n, (double *) partial);
    PMB_LAUNCH_CHECK(ctx);
    double *h = (double *) malloc(sizeof(double) * grid);
    if (!h) return PMB_ENOMEM;
    cudaError_t e = cudaMemcpyAsync(h, partial, sizeof(double) * grid, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { free(h); return pmb_cuda_fail(e, "dot copy", __FILE__, __LINE__); }
    double s = 0;
    for (int i = 0; i < grid; i++) s += h[i];
    free(h);
    *dot_h = s;
    return PMB_OK;
}
