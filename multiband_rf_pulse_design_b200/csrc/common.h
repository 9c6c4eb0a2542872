// Shared host/device helpers of libmbrf.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/mbrf.h"

namespace mbrf {

void set_error(const char *fmt, ...);
int require_device();  // MBRF_OK or MBRF_ENODEVICE (message set)
extern std::atomic<unsigned long long> g_launches;
int sm_count();         // SMs of the current device (cached per device)

#define MBRF_CUDA(call)                                                                      \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            ::mbrf::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),       \
                              __FILE__, __LINE__);                                           \
            return MBRF_ECUDA;                                                               \
        }                                                                                    \
    } while (0)

#define MBRF_LAUNCH_CHECK()                                                                  \
    do {                                                                                     \
        ::mbrf::g_launches.fetch_add(1, std::memory_order_relaxed);                          \
        MBRF_CUDA(cudaGetLastError());                                                       \
    } while (0)

// Growable device scratch owned by a host thread (one per thread: the MEX boundary is
// single-threaded, torch worker threads each get their own).  Freed at thread exit.
struct DeviceScratch {
    void *ptr = nullptr;
    size_t bytes = 0;
    int device = -1;
    int reserve(size_t need);  // MBRF_OK / MBRF_ECUDA
    ~DeviceScratch();
};

// ---------------------------------------------------------------------------
// device-side: mbarrier + TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP)
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    // make the initialised barrier visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
#endif

}  // namespace mbrf
