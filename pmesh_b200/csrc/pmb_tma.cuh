// pmb_tma.cuh -- bulk asynchronous copies (TMA, cp.async.bulk) and mbarrier helpers for sm_100a.
//
// The particle stream of the paint / readout kernels is a flat array that every kernel reads exactly
// once, in chunks of a few KB.  Instead of per-lane LDG (three stride-24-byte scalar loads per particle,
// whose latency one iteration of software prefetch does not hide -- profiles/r1_*_hotspots.txt), one
// elected thread hands whole chunks to the copy engine several iterations ahead; the bytes land in a
// shared-memory ring and completion is signalled on an mbarrier (transaction count), which the consumer
// threads wait on.  SASS: UBLKCP (the bulk copy), SYNCS (mbarrier arrive / try_wait).
#pragma once
#include <stdint.h>

__device__ __forceinline__ uint32_t pmb_smem_addr(const void *p)
{
    return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void pmb_mbar_init(uint64_t *bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pmb_smem_addr(bar)), "r"(arrivals) : "memory");
}

// make the initialised barriers visible to the async proxy (the copy engine) before the first copy
__device__ __forceinline__ void pmb_mbar_init_fence(void)
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// one arrival + `bytes` expected from bulk copies that name this barrier
__device__ __forceinline__ void pmb_mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pmb_smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void pmb_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pmb_smem_addr(bar)) : "memory");
}

// block until the phase with the given parity has completed (try_wait suspends the thread in hardware)
__device__ __forceinline__ void pmb_mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "PMB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra PMB_DONE;\n"
        "bra PMB_WAIT;\n"
        "PMB_DONE:\n"
        "}\n" ::"r"(pmb_smem_addr(bar)), "r"(parity) : "memory");
}

// global -> shared bulk copy of `bytes` (multiple of 16; both addresses 16-byte aligned), completion
// counted on `bar`; the stream is read once: L2 evict_first
__device__ __forceinline__ void pmb_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(pmb_smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(pmb_smem_addr(bar)), "l"(policy) : "memory");
}

__device__ __forceinline__ uint64_t pmb_policy_evict_first(void)
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
