// Host-pointer pipeline shared by the drop-in entry points (mbrf_bloch / mbrf_blochsimfz / mbrf_abr): what a MEX gateway
// calls with the arrays MATLAB hands it.  Those arrays are PAGEABLE (mxCreateDoubleMatrix), so the result cannot be stored
// into them by the GPU; and a MATLAB host is one process, so the only way to use all GPUs of the box is inside the call.
//
//   * the item range (spins / positions) is cut into contiguous per-device ranges over `mbrf_set_fanout` devices, each range
//     into chunks;  every device runs  inputs H2D -> [kernel(chunk) -> D2H(chunk) into a pinned ring] x chunks  on its own stream;
//   * the calling thread waits for the chunks in order and a small pool of host threads copies each one from the pinned
//     ring into the caller's arrays while the GPUs work on the following chunks;
//   * page-locked result arrays skip the ring: one device stores straight into them (zero-copy), several devices DMA
//     their slices into them.
#pragma once
#include "common.h"

#include <functional>
#include <vector>

namespace mbrf {
namespace hostpipe {

struct Slot {                 // one device as seen by one host thread
    int device = -1;
    cudaStream_t st = nullptr;        // kernels and input copies
    cudaStream_t st_copy = nullptr;   // result copies, so that the kernels of the next chunk do not wait for them
    DeviceScratch dev;
    void *pin = nullptr;
    size_t pin_bytes = 0;
    std::vector<cudaEvent_t> ev, ev_k;
    int reserve_pin(size_t need);
    int event(size_t k, cudaEvent_t *out);
};

// this host thread's context on `device`; makes `device` current.  nullptr + error text on failure
Slot *slot(int device);
int fanout();                 // devices one host-pointer call may spread over (>= 1)

struct Desc {
    size_t in_bytes = 0;                               // packed small inputs, replicated to every device
    std::function<void(char *)> pack;                  // fills a host block of in_bytes
    size_t ws_bytes = 0;                               // device workspace per device (16-byte aligned)
    int ncomp = 0;                                     // output components (3: mx,my,mz; 4: alpha/beta planes)
    double *const *host_out = nullptr;                 // [ncomp]; item i of component c at host_out[c] + i * item_doubles
    size_t item_doubles = 1;
    long long items = 0;
    int ncomp_in = 0;                                  // optional per-item inputs (initial magnetisation)
    const double *const *host_in = nullptr;            // element i of component c at host_in[c][i * in_stride]
    long long in_stride = 1;
    // enqueue the kernels of items [item0, item0 + n) on `st`: d_in = the packed inputs on this device, d_item_in[c] = this
    // chunk's per-item inputs (contiguous, or null), d_out[c] = where component c of the chunk goes (item-major)
    std::function<int(cudaStream_t st, const char *d_in, char *d_ws, long long item0, long long n,
                      const double *const *d_item_in, double *const *d_out)> launch;
};

int run(const Desc &d);

}  // namespace hostpipe
}  // namespace mbrf
