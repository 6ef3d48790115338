// pmb_comm.cu -- the communicator: NCCL over NVLink 5 / NVSwitch, one rank per process.
//
// Replaces the mpi4py calls on the data path of the reference:
//   Alltoallv of packed particle records ... pmesh/domain.py:199-205 (exchange), :274-281 (gather)
//   PFFT's internal global transposes ...... pmesh/pm.py:689,1017 (plan.execute)
//   allreduce of scalars ................... pmesh/pm.py:296,739,899 (cgetitem, csum, cdot)
// On an NVSwitch node every peer is reachable at full bandwidth, so the alltoallv is one flat
// ncclGroup of P sends + P receives; the self block is a device memcpy.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include "pmb_internal.h"

// NCCL is bound at run time (dlopen), not at link time: a process that also imports torch must end
// up with ONE libnccl.so.2 -- whichever is already loaded (torch bundles its own, newer than the
// system one) -- and loading this library must not pin the system copy first.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)(void);
    ncclResult_t (*GroupEnd)(void);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl;
static int g_nccl_state = 0;   // 0 not tried, 1 ok, -1 failed

static int nccl_load(void)
{
    if (g_nccl_state == 1) return PMB_OK;
    if (g_nccl_state == -1) { pmb_set_error("NCCL library could not be loaded"); return PMB_ENCCL; }
    const char *env = getenv("PMESH_B200_NCCL_LIB");
    void *h = dlopen(env && *env ? env : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        g_nccl_state = -1;
        pmb_set_error("dlopen(libnccl.so.2) failed: %s", dlerror());
        return PMB_ENCCL;
    }
#define PMB_SYM(field, name)                                              \
    *(void **) (&g_nccl.field) = dlsym(h, name);                          \
    if (!g_nccl.field) { g_nccl_state = -1; pmb_set_error("NCCL symbol %s missing", name); return PMB_ENCCL; }
    PMB_SYM(GetUniqueId, "ncclGetUniqueId");
    PMB_SYM(CommInitRank, "ncclCommInitRank");
    PMB_SYM(CommDestroy, "ncclCommDestroy");
    PMB_SYM(Send, "ncclSend");
    PMB_SYM(Recv, "ncclRecv");
    PMB_SYM(GroupStart, "ncclGroupStart");
    PMB_SYM(GroupEnd, "ncclGroupEnd");
    PMB_SYM(AllReduce, "ncclAllReduce");
    PMB_SYM(AllGather, "ncclAllGather");
    PMB_SYM(GetErrorString, "ncclGetErrorString");
#undef PMB_SYM
    g_nccl_state = 1;
    return PMB_OK;
}

static int nccl_fail(ncclResult_t r, const char *what, int line)
{
    pmb_set_error("NCCL error %d (%s) at %s:%d in %s", (int) r,
                  g_nccl_state == 1 ? g_nccl.GetErrorString(r) : "?", __FILE__, line, what);
    return PMB_ENCCL;
}

#define PMB_NCCL(call)                                                   \
    do {                                                                 \
        ncclResult_t _r = (call);                                        \
        if (_r != ncclSuccess) return nccl_fail(_r, #call, __LINE__);    \
    } while (0)

extern "C" int pmb_comm_unique_id(char *id128_h)
{
    PMB_REQUIRE(id128_h, "null argument");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    PMB_CHECK(nccl_load());
    ncclUniqueId id;
    PMB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128_h, &id, sizeof(id));
    return PMB_OK;
}

extern "C" int pmb_comm_init_rank(pmb_ctx *ctx, const char *id128_h, int rank, int nranks)
{
    PMB_REQUIRE(ctx && id128_h, "null argument");
    PMB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d / %d", rank, nranks);
    PMB_REQUIRE(!ctx->comm, "communicator already initialised");
    PMB_CHECK(nccl_load());
    PMB_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128_h, sizeof(id));
    ncclComm_t comm;
    PMB_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->comm = (ncclComm *) comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    return PMB_OK;
}

extern "C" int pmb_comm_destroy(pmb_ctx *ctx)
{
    if (!ctx || !ctx->comm) return PMB_OK;
    cudaStreamSynchronize(ctx->stream);
    if (g_nccl_state == 1) g_nccl.CommDestroy((ncclComm_t) ctx->comm);
    ctx->comm = NULL;
    ctx->rank = 0;
    ctx->nranks = 1;
    return PMB_OK;
}

extern "C" int pmb_comm_rank(pmb_ctx *ctx, int *rank, int *nranks)
{
    PMB_REQUIRE(ctx, "null context");
    if (rank) *rank = ctx->rank;
    if (nranks) *nranks = ctx->nranks;
    return PMB_OK;
}

extern "C" int pmb_alltoallv(pmb_ctx *ctx, const void *send, const int64_t *sendcounts_h, const int64_t *sendoffsets_h,
                             void *recv, const int64_t *recvcounts_h, const int64_t *recvoffsets_h, int64_t itemsize)
{
    PMB_REQUIRE(ctx && sendcounts_h && sendoffsets_h && recvcounts_h && recvoffsets_h && itemsize > 0, "bad alltoallv arguments");
    const int P = ctx->nranks, me = ctx->rank;
    PMB_REQUIRE(sendcounts_h[me] == recvcounts_h[me], "self send/recv counts differ");
    if (sendcounts_h[me] > 0)
        PMB_CUDA(cudaMemcpyAsync((char *) recv + recvoffsets_h[me] * itemsize,
                                 (const char *) send + sendoffsets_h[me] * itemsize,
                                 (size_t) sendcounts_h[me] * itemsize, cudaMemcpyDeviceToDevice, ctx->stream));
    if (P == 1) return PMB_OK;
    PMB_REQUIRE(ctx->comm, "communicator not initialised");
    ncclComm_t comm = (ncclComm_t) ctx->comm;
    PMB_NCCL(g_nccl.GroupStart());
    for (int q = 0; q < P; q++) {
        if (q == me) continue;
        if (sendcounts_h[q] > 0)
            PMB_NCCL(g_nccl.Send((const char *) send + sendoffsets_h[q] * itemsize, (size_t) sendcounts_h[q] * itemsize,
                              ncclChar, q, comm, ctx->stream));
        if (recvcounts_h[q] > 0)
            PMB_NCCL(g_nccl.Recv((char *) recv + recvoffsets_h[q] * itemsize, (size_t) recvcounts_h[q] * itemsize,
                              ncclChar, q, comm, ctx->stream));
    }
    PMB_NCCL(g_nccl.GroupEnd());
    return PMB_OK;
}

extern "C" int pmb_allreduce_f64(pmb_ctx *ctx, double *buf, int64_t n, int op)
{
    PMB_REQUIRE(ctx && (buf || n == 0) && n >= 0, "bad allreduce arguments");
    PMB_REQUIRE(op >= 0 && op <= 2, "bad reduction op");
    if (ctx->nranks == 1 || n == 0) return PMB_OK;
    PMB_REQUIRE(ctx->comm, "communicator not initialised");
    ncclRedOp_t ops[3] = {ncclSum, ncclMax, ncclMin};
    PMB_NCCL(g_nccl.AllReduce(buf, buf, (size_t) n, ncclDouble, ops[op], (ncclComm_t) ctx->comm, ctx->stream));
    return PMB_OK;
}

extern "C" int pmb_allgather_bytes(pmb_ctx *ctx, const void *send, void *recv, int64_t nbytes_per_rank)
{
    PMB_REQUIRE(ctx && nbytes_per_rank >= 0, "bad allgather arguments");
    if (nbytes_per_rank == 0) return PMB_OK;
    PMB_REQUIRE(send && recv, "null buffer");
    if (ctx->nranks == 1) {
        if (send != recv)
            PMB_CUDA(cudaMemcpyAsync(recv, send, (size_t) nbytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
        return PMB_OK;
    }
    PMB_REQUIRE(ctx->comm, "communicator not initialised");
    PMB_NCCL(g_nccl.AllGather(send, recv, (size_t) nbytes_per_rank, ncclChar, (ncclComm_t) ctx->comm, ctx->stream));
    return PMB_OK;
}

extern "C" int pmb_barrier(pmb_ctx *ctx)
{
    PMB_REQUIRE(ctx, "null context");
    if (ctx->nranks > 1) {
        PMB_REQUIRE(ctx->comm, "communicator not initialised");
        void *token;
        PMB_CHECK(pmb_scratch(ctx, 256, &token));
        PMB_CUDA(cudaMemsetAsync(token, 0, 8, ctx->stream));
        PMB_NCCL(g_nccl.AllReduce(token, token, 1, ncclDouble, ncclSum, (ncclComm_t) ctx->comm, ctx->stream));
    }
    PMB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PMB_OK;
}

// ---- internal helpers (pmb_internal.h) ---------------------------------------------------------
// in-stream barrier: every rank's stream has reached this point when the allreduce completes; the
// host does not wait.  Used to order direct peer-memory writes (pmb_fft.cu) between GPUs.
int pmb_stream_barrier(pmb_ctx *ctx)
{
    if (ctx->nranks <= 1) return PMB_OK;
    PMB_REQUIRE(ctx->comm, "communicator not initialised");
    if (!ctx->barrier_token) {
        PMB_CUDA(cudaMalloc(&ctx->barrier_token, 256));
        PMB_CUDA(cudaMemsetAsync(ctx->barrier_token, 0, 256, ctx->stream));
    }
    PMB_NCCL(g_nccl.AllReduce(ctx->barrier_token, ctx->barrier_token, 1, ncclInt, ncclMax, (ncclComm_t) ctx->comm, ctx->stream));
    return PMB_OK;
}

// the same barrier on an explicit stream (the FFT's transpose stream); its own token so that it never shares
// a buffer with a barrier of the compute stream
int pmb_stream_barrier_on(pmb_ctx *ctx, cudaStream_t stream)
{
    if (ctx->nranks <= 1) return PMB_OK;
    PMB_REQUIRE(ctx->comm, "communicator not initialised");
    if (!ctx->barrier_token) {
        PMB_CUDA(cudaMalloc(&ctx->barrier_token, 256));
        PMB_CUDA(cudaMemsetAsync(ctx->barrier_token, 0, 256, ctx->stream));
        PMB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    char *tok = (char *) ctx->barrier_token + 128;
    PMB_NCCL(g_nccl.AllReduce(tok, tok, 1, ncclInt, ncclMax, (ncclComm_t) ctx->comm, stream));
    return PMB_OK;
}

// allgather of small host records (setup paths only): staged through device scratch, synchronous
int pmb_allgather_host(pmb_ctx *ctx, const void *send_h, void *recv_h, size_t nbytes)
{
    if (ctx->nranks <= 1) { memcpy(recv_h, send_h, nbytes); return PMB_OK; }
    PMB_REQUIRE(ctx->comm, "communicator not initialised");
    void *dev;
    PMB_CHECK(pmb_scratch(ctx, nbytes * (size_t) (ctx->nranks + 1), &dev));
    char *dsend = (char *) dev, *drecv = (char *) dev + nbytes;
    PMB_CUDA(cudaMemcpyAsync(dsend, send_h, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    PMB_NCCL(g_nccl.AllGather(dsend, drecv, nbytes, ncclChar, (ncclComm_t) ctx->comm, ctx->stream));
    PMB_CUDA(cudaMemcpyAsync(recv_h, drecv, nbytes * (size_t) ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream));
    PMB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PMB_OK;
}
