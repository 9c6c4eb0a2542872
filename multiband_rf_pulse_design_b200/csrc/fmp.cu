// Batched minimum-phase spectral factorisation -- fmp2 / mag2mp / fftc of fir_ap_cvx.m:253-304 (J. Pauly's
// fmp.m / mag2mp.m), the step that turns the solver's autocorrelation x into the function's return value h
// (fir_ap_cvx.m:185-202).  SURVEY.md 8(f) "next" row 1.
//
//   hp    = [zeros(ceil((lp-l)/2)), r, zeros(floor((lp-l)/2))],  lp = 8 * 2^ceil(log2(l)),  l = 2n-1        (:266-267)
//   hpf   = fftshift(fft(fftshift(hp)))                                                                      (:254, :268)
//   xl    = log(sqrt(abs(hpf)));  xlf = fft(xl);  keep DC and Nyquist, double 2..n/2, zero the rest          (:293-300)
//   hpfmp = exp(ifft(xlfp))                                                                                  (:301-302)
//   hmp   = ifft(fftshift(conj(hpfmp)))(1 : (l+1)/2)                                                         (:275-276)
//
// One CTA per design: the padded signal (4096 complex doubles = 64 KB at n = 256, 128 KB at n = 512) lives in shared memory
// for all four transforms; in-place radix-2 decimation-in-time FFT in fp64, twiddles from a table built once per call with
// sincospi (the table is 32 KB .. 64 KB and stays in L1/L2).  The fftshifts are index arithmetic folded into the stages
// around them.  Designs are independent: the grid is the batch.
#include "common.h"
#include "fft_smem.cuh"

#include <cmath>
#include <cstring>
#include <vector>

namespace mbrf {
namespace fmp {

using fftsm::fft_inplace;
using fftsm::twiddle_kernel;

constexpr int THREADS = 512;

// r: [B][l] complex (split planes), h: [B][n] complex (split planes); l = 2n-1, N = 8 * 2^ceil(log2 l) = 2^lg
__global__ void __launch_bounds__(THREADS) fmp2_kernel(const double *__restrict__ r_re, const double *__restrict__ r_im, int n,
                                                       int lg, const double2 *__restrict__ tw, double *__restrict__ h_re,
                                                       double *__restrict__ h_im)
{
    extern __shared__ double2 s[];
    const int N = 1 << lg, l = 2 * n - 1, H = N >> 1;
    const int b = blockIdx.x;
    const int padl = (N - l + 1) / 2;                              // ceil((lp - l) / 2) leading zeros
    // s = fftshift(hp): s[i] = hp[(i + N/2) mod N]
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int k = ((i + H) & (N - 1)) - padl;
        double2 v = make_double2(0.0, 0.0);
        if (k >= 0 && k < l) v = make_double2(r_re[(size_t)b * l + k], r_im ? r_im[(size_t)b * l + k] : 0.0);
        s[i] = v;
    }
    __syncthreads();
    fft_inplace(s, lg, tw, false);
    // hpf = fftshift(F); xl = log(sqrt(abs(hpf))): a permutation of F's magnitudes, so compute in place and remember that
    // index i of `s` now holds xl[(i + N/2) mod N]
    for (int i = threadIdx.x; i < N; i += blockDim.x) s[i] = make_double2(log(sqrt(hypot(s[i].x, s[i].y))), 0.0);
    __syncthreads();
    // xlf = fft(xl) with xl[m] = s[(m + N/2) mod N]: a circular shift by N/2 in time is the factor (-1)^k in frequency
    fft_inplace(s, lg, tw, false);
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        double2 v = s[k];
        if (k & 1) { v.x = -v.x; v.y = -v.y; }
        const double g = (k == 0 || k == H) ? 1.0 : (k < H ? 2.0 : 0.0);   // keep DC and Nyquist, double positive, zero negative
        s[k] = make_double2(v.x * g, v.y * g);
    }
    __syncthreads();
    fft_inplace(s, lg, tw, true);                                  // xlaf * N
    const double invN = 1.0 / (double)N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {            // a = exp(xlaf); then conj(a)
        const double mag = exp(s[i].x * invN);
        double sn, cs;
        sincos(s[i].y * invN, &sn, &cs);
        s[i] = make_double2(mag * cs, -mag * sn);
    }
    __syncthreads();
    // hpmp = ifft(fftshift(conj(a))): the shift by N/2 of the input is the factor (-1)^m on the output
    fft_inplace(s, lg, tw, true);
    for (int m = threadIdx.x; m < n; m += blockDim.x) {
        const double sg = (m & 1) ? -invN : invN;
        h_re[(size_t)b * n + m] = s[m].x * sg;
        h_im[(size_t)b * n + m] = s[m].y * sg;
    }
}

static int fft_log2(int n)   // lg of lp = 8 * 2^ceil(log2(2n-1))
{
    const int l = 2 * n - 1;
    int p = 0;
    while ((1 << p) < l) ++p;
    return p + 3;
}

struct Ctx {
    DeviceScratch dev;
};
static thread_local Ctx t_ctx;

}  // namespace fmp
}  // namespace mbrf

using namespace mbrf;
using namespace mbrf::fmp;

extern "C" {

/* largest n the shared-memory kernel takes (the padded length 8 * 2^ceil(log2(2n-1)) complex doubles must fit 227 KB) */
int mbrf_fmp2_max_taps(void) { return 512; }

/*
 * Batched fmp2 on device pointers: r_re, r_im [B x (2n-1)] (r_im may be NULL), h_re, h_im [B x n];
 * workspace: (4 << ceil(log2(2n-1))) * 16 bytes for the twiddle table.
 */
int mbrf_fmp2_batch_device(const double *r_re, const double *r_im, int n, int B, double *h_re, double *h_im, void *workspace,
                           void *stream)
{
    if (int rc = require_device()) return rc;
    if (!r_re || !h_re || !h_im || !workspace || n < 1 || B < 1) {
        set_error("fmp2: bad arguments (n=%d B=%d)", n, B);
        return MBRF_EINVAL;
    }
    if (n > mbrf_fmp2_max_taps()) {
        set_error("fmp2: n=%d exceeds the shared-memory kernel's limit of %d taps", n, mbrf_fmp2_max_taps());
        return MBRF_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int lg = fft_log2(n), N = 1 << lg;
    double2 *tw = (double2 *)workspace;
    twiddle_kernel<<<(N / 2 + 255) / 256, 256, 0, st>>>(tw, N / 2);
    MBRF_LAUNCH_CHECK();
    const size_t smem = (size_t)N * sizeof(double2);
    if (smem > 48 * 1024) {        // set on every call: the attribute is per device, a cache per process would miss a device change
        MBRF_CUDA(cudaFuncSetAttribute(fmp2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    fmp2_kernel<<<B, THREADS, smem, st>>>(r_re, r_im, n, lg, tw, h_re, h_im);
    MBRF_LAUNCH_CHECK();
    return MBRF_OK;
}

unsigned long long mbrf_fmp2_workspace_bytes(int n)
{
    if (n < 1) return 0;
    return (unsigned long long)(1 << fft_log2(n)) / 2 * sizeof(double2);
}

/* Host pointers: h = fmp2(r) for B autocorrelation sequences (fir_ap_cvx.m:262-283), one H2D / D2H per call. */
int mbrf_fmp2_batch(const double *r_re, const double *r_im, int n, int B, double *h_re, double *h_im)
{
    if (int rc = require_device()) return rc;
    if (!r_re || !h_re || !h_im || n < 1 || B < 1) {
        set_error("fmp2: bad arguments (n=%d B=%d)", n, B);
        return MBRF_EINVAL;
    }
    const size_t l = (size_t)(2 * n - 1), nr = l * B * 8, nh = (size_t)n * B * 8;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t wsb = al(mbrf_fmp2_workspace_bytes(n));
    Ctx &cx = t_ctx;
    if (int rc = cx.dev.reserve(2 * al(nr) + 2 * al(nh) + wsb)) return rc;
    char *d = (char *)cx.dev.ptr;
    double *drr = (double *)d, *dri = (double *)(d + al(nr)), *dhr = (double *)(d + 2 * al(nr)), *dhi = (double *)(d + 2 * al(nr) + al(nh));
    void *ws = d + 2 * al(nr) + 2 * al(nh);
    MBRF_CUDA(cudaMemcpyAsync(drr, r_re, nr, cudaMemcpyHostToDevice, 0));
    if (r_im) MBRF_CUDA(cudaMemcpyAsync(dri, r_im, nr, cudaMemcpyHostToDevice, 0));
    if (int rc = mbrf_fmp2_batch_device(drr, r_im ? dri : nullptr, n, B, dhr, dhi, ws, nullptr)) return rc;
    MBRF_CUDA(cudaMemcpyAsync(h_re, dhr, nh, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaMemcpyAsync(h_im, dhi, nh, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaStreamSynchronize(0));
    return MBRF_OK;
}

}  // extern "C"
