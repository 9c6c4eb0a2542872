// In-place radix-2 FFT of a power-of-two complex fp64 sequence held in shared memory by one CTA, and the twiddle table
// it reads.  Shared by the spectral factorisation (fmp.cu) and the inverse SLR transform (islr.cu).
#pragma once
#include <cuda_runtime.h>

namespace mbrf {
namespace fftsm {

static __global__ void twiddle_kernel(double2 *tw, int half)   // tw[k] = exp(-2 pi i k / (2 half)), k < half
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= half) return;
    double s, c;
    sincospi(-(double)k / (double)half, &s, &c);
    tw[k] = make_double2(c, s);
}

// in-place FFT of s[0..N) (N = 2^lg), forward (inverse = false) or unnormalised inverse; all threads of the CTA.
// tw is the table of a transform of length N << tw_shift (a longer table serves shorter transforms).
__device__ __forceinline__ void fft_inplace(double2 *s, int lg, const double2 *__restrict__ tw, bool inverse, int tw_shift = 0)
{
    const int N = 1 << lg;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {            // bit-reversal permutation
        const int j = (int)(__brev((unsigned)i) >> (32 - lg));
        if (i < j) { const double2 t = s[i]; s[i] = s[j]; s[j] = t; }
    }
    __syncthreads();
    for (int st = 0; st < lg; ++st) {
        const int half = 1 << st;                                  // butterflies of span `half`
        const int tstep = ((N >> 1) >> st) << tw_shift;            // twiddle index stride: w = exp(-+2 pi i j / (2 half))
        for (int b = threadIdx.x; b < (N >> 1); b += blockDim.x) {
            const int j = b & (half - 1);
            const int i0 = ((b >> st) << (st + 1)) + j, i1 = i0 + half;
            double2 w = tw[j * tstep];
            if (inverse) w.y = -w.y;
            const double2 a = s[i0], c = s[i1];
            const double tr = fma(c.x, w.x, -c.y * w.y), ti = fma(c.x, w.y, c.y * w.x);
            s[i0] = make_double2(a.x + tr, a.y + ti);
            s[i1] = make_double2(a.x - tr, a.y - ti);
        }
        __syncthreads();
    }
}

}  // namespace fftsm
}  // namespace mbrf
