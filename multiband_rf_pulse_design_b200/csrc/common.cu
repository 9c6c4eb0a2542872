// libmbrf.so — library plumbing: error text, device selection, launch counter.
#include "common.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace mbrf {

static thread_local char t_err[1024] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
}

int require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); libmbrf has no CPU path",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return MBRF_ENODEVICE;
    }
    return MBRF_OK;
}

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int DeviceScratch::reserve(size_t need)
{
    int dev = 0;
    MBRF_CUDA(cudaGetDevice(&dev));
    if (ptr && dev == device && bytes >= need) return MBRF_OK;
    if (ptr) {
        cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
    size_t want = need + need / 4 + 4096;
    MBRF_CUDA(cudaMalloc(&ptr, want));
    bytes = want;
    device = dev;
    return MBRF_OK;
}

DeviceScratch::~DeviceScratch()
{
    // at process/thread teardown the context may already be gone; ignore errors
    if (ptr) cudaFree(ptr);
}

}  // namespace mbrf

extern "C" {

const char *mbrf_version(void) { return "mbrf-b200 0.1 (sm_100a)"; }
const char *mbrf_last_error(void) { return mbrf::t_err; }

int mbrf_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int mbrf_set_device(int device)
{
    if (int rc = mbrf::require_device()) return rc;
    MBRF_CUDA(cudaSetDevice(device));
    return MBRF_OK;
}

int mbrf_get_device(int *device)
{
    if (!device) return MBRF_EINVAL;
    if (int rc = mbrf::require_device()) return rc;
    MBRF_CUDA(cudaGetDevice(device));
    return MBRF_OK;
}

int mbrf_device_sm_count(int *out)
{
    if (!out) return MBRF_EINVAL;
    if (int rc = mbrf::require_device()) return rc;
    *out = mbrf::sm_count();
    return MBRF_OK;
}

unsigned long long mbrf_launch_count(void) { return mbrf::g_launches.load(); }

/* ---- peer-mapped result buffers: the gather of a sharded run without a collective ------------------------------------
 * The destination rank allocates the final [planes x S] result once and exports a CUDA IPC handle; every other rank (one
 * process per GPU) opens it and passes pointers INTO it as the output arrays of mbrf_bloch_device / mbrf_abr_device: the
 * kernels store their slice straight into the destination GPU's memory over NVLink (24 B per spin after 512 steps of
 * arithmetic: the link is idle), so "compute then gather" is one kernel per rank and no transfer step exists. */
int mbrf_peer_alloc(unsigned long long bytes, void **dptr, unsigned char handle[64])
{
    if (!dptr || !handle || bytes == 0) return MBRF_EINVAL;
    if (int rc = mbrf::require_device()) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    MBRF_CUDA(cudaMalloc(dptr, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *dptr);
    if (e != cudaSuccess) {
        cudaFree(*dptr);
        *dptr = nullptr;
        mbrf::set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return MBRF_ECUDA;
    }
    memcpy(handle, &h, 64);
    return MBRF_OK;
}

int mbrf_peer_open(const unsigned char handle[64], void **dptr)
{
    if (!dptr || !handle) return MBRF_EINVAL;
    if (int rc = mbrf::require_device()) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    MBRF_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return MBRF_OK;
}

int mbrf_peer_close(void *dptr)
{
    if (!dptr) return MBRF_OK;
    MBRF_CUDA(cudaIpcCloseMemHandle(dptr));
    return MBRF_OK;
}

int mbrf_peer_free(void *dptr)
{
    if (!dptr) return MBRF_OK;
    MBRF_CUDA(cudaFree(dptr));
    return MBRF_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// FP64 pipe peak: the Bloch / SLR kernels are bound by the FP64 CUDA-core pipe, whose
// peak is not in MEASURED_PEAKS.json.  A DFMA-only kernel with 8 independent chains per
// thread measures it on the box the bench runs on.
// ---------------------------------------------------------------------------
namespace mbrf {
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) out[0] = s;  // keep the chains alive
}
}  // namespace mbrf

extern "C" int mbrf_measure_fp64_peak(double *tflops, double *ms)
{
    using namespace mbrf;
    if (!tflops) return MBRF_EINVAL;
    if (int rc = require_device()) return rc;
    double *d = nullptr;
    MBRF_CUDA(cudaMalloc(&d, 8));
    cudaEvent_t e0, e1;
    MBRF_CUDA(cudaEventCreate(&e0));
    MBRF_CUDA(cudaEventCreate(&e1));
    const int iters = 4096, blocks = sm_count() * 8, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        MBRF_CUDA(cudaEventRecord(e0, 0));
        fp64_peak_kernel<<<blocks, threads>>>(d, iters, 1.0);
        MBRF_LAUNCH_CHECK();
        MBRF_CUDA(cudaEventRecord(e1, 0));
        MBRF_CUDA(cudaEventSynchronize(e1));
        float t = 0;
        MBRF_CUDA(cudaEventElapsedTime(&t, e0, e1));
        if (rep > 0 && t < best) best = t;
    }
    const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms) *ms = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return MBRF_OK;
}
