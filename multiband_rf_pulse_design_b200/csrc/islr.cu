// Batched inverse SLR transform -- rf_tools/b2a.m:13-28 (minimum-phase alpha polynomial of a beta polynomial) and
// rf_tools/ab2rf.m:12-26 (peel-off recursion alpha, beta -> complex RF), the step after the FIR design in
// dzrf_mb.m:239-240.  SURVEY.md 8(f) "next" row 2.  One CTA per pulse, everything in shared memory.
//
// b2a:    bf  = fft([bc, zeros(1, 7n)]);  if max|bf| >= 1: bf = bf / (1e-8 + max|bf|)                     (b2a.m:17-25)
//         afa = mag2mp(sqrt(1 - bf .* conj(bf)));  aca = fft(afa) / (8n);  aca = aca(n:-1:1)               (:26-28, mag2mp.m:25-33)
// ab2rf:  for i = n..1:  c = sqrt(1 / (1 + |b_i / a_i|^2)),  s = conj(c b_i / a_i),
//                        rf_i = 2 atan2(|s|, c) exp(j angle(s)),
//                        a <- (c a + s b)(2:i),  b <- (-conj(s) a + c b)(1:i-1)                            (ab2rf.m:16-26)
// The recursion is inherently serial in i (n steps, one CTA barrier each, i complex updates spread over the CTA):
// parallelism comes from the batch.  b2a needs a length-8n DFT: radix-2 for n = 2^k <= 1024, Bluestein for other n <= 512.
#include "common.h"
#include "fft_smem.cuh"

#include <cmath>

namespace mbrf {
namespace islr {

using fftsm::fft_inplace;
using fftsm::twiddle_kernel;

constexpr int THREADS = 512;

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cdiv(double2 a, double2 b)
{
    const double d = b.x * b.x + b.y * b.y;
    return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

// DFT of s[0..N) in place for ANY even N (all threads of the CTA).  N = 2^lg: radix-2.  Otherwise Bluestein's chirp-z
// identity  k m = (k^2 + m^2 - (k-m)^2) / 2:  X_k = c_k sum_m (x_m c_m) conj(c)_{k-m},  c_m = exp(-i pi m^2 / N), a circular
// convolution of length L = 2^lgL >= 2N-1 done with three radix-2 transforms (the transform of the chirp, Bhat, is computed
// once per call).  The buffer s must hold L entries; entries >= N are scratch.  Unnormalised inverse on request.
struct DftPlan {
    int N, lgL;                 // lgL = log2 N when N is a power of two (then chirp = Bhat = nullptr)
    const double2 *tw;          // twiddles of the length-2^lgL transform
    const double2 *chirp;       // [N]
    const double2 *Bhat;        // [2^lgL]
};
__device__ void dft_inplace(double2 *s, const DftPlan &pl, bool inverse)
{
    if (!pl.chirp) { fft_inplace(s, pl.lgL, pl.tw, inverse); return; }
    const int N = pl.N, L = 1 << pl.lgL;
    for (int m = threadIdx.x; m < L; m += blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        if (m < N) {
            v = s[m];
            if (inverse) v.y = -v.y;                      // idft(x) = conj(dft(conj(x)))
            v = cmul(v, pl.chirp[m]);
        }
        s[m] = v;
    }
    __syncthreads();
    fft_inplace(s, pl.lgL, pl.tw, false);
    for (int i = threadIdx.x; i < L; i += blockDim.x) s[i] = cmul(s[i], pl.Bhat[i]);
    __syncthreads();
    fft_inplace(s, pl.lgL, pl.tw, true);
    const double invL = 1.0 / (double)L;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        double2 v = cmul(pl.chirp[k], s[k]);
        v.x *= invL; v.y *= invL;
        if (inverse) v.y = -v.y;
        s[k] = v;
    }
    __syncthreads();
}
// one CTA: chirp[m] = exp(-i pi m^2 / N) (m^2 reduced mod 2N in integers) and Bhat = FFT_L of conj(chirp) wrapped to (-N, N)
__global__ void __launch_bounds__(THREADS) bluestein_setup_kernel(int N, int lgL, const double2 *__restrict__ tw, double2 *chirp,
                                                                  double2 *Bhat)
{
    extern __shared__ double2 s[];
    const int L = 1 << lgL;
    for (int i = threadIdx.x; i < L; i += blockDim.x) s[i] = make_double2(0.0, 0.0);
    __syncthreads();
    for (int m = threadIdx.x; m < N; m += blockDim.x) {
        const long long r = ((long long)m * m) % (2LL * N);
        double sn, cs;
        sincospi(-(double)r / (double)N, &sn, &cs);
        chirp[m] = make_double2(cs, sn);
        s[m] = make_double2(cs, -sn);
        if (m > 0) s[L - m] = make_double2(cs, -sn);
    }
    __syncthreads();
    fft_inplace(s, lgL, tw, false);
    for (int i = threadIdx.x; i < L; i += blockDim.x) Bhat[i] = s[i];
}

// b: [B][n] complex (split planes) -> a: [B][n]; N = 8n (any n: dft_inplace)
__global__ void __launch_bounds__(THREADS) b2a_kernel(const double *__restrict__ b_re, const double *__restrict__ b_im, int n,
                                                      const DftPlan pl, double *__restrict__ a_re, double *__restrict__ a_im)
{
    extern __shared__ double2 s[];
    __shared__ double red[THREADS];
    const int N = pl.N, H = N >> 1, q = blockIdx.x;
    for (int i = threadIdx.x; i < N; i += blockDim.x)
        s[i] = i < n ? make_double2(b_re[(size_t)q * n + i], b_im ? b_im[(size_t)q * n + i] : 0.0) : make_double2(0.0, 0.0);
    __syncthreads();
    dft_inplace(s, pl, false);                                      // bf
    double m = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) m = fmax(m, hypot(s[i].x, s[i].y));
    red[threadIdx.x] = m;
    __syncthreads();
    for (int k = THREADS / 2; k > 0; k >>= 1) { if ((int)threadIdx.x < k) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + k]); __syncthreads(); }
    const double bfmax = red[0];
    const double sc = bfmax >= 1.0 ? 1.0 / (1e-8 + bfmax) : 1.0;    // b2a.m:23-25
    for (int i = threadIdx.x; i < N; i += blockDim.x) {             // xl = log(sqrt(1 - |bf|^2))
        const double re = s[i].x * sc, im = s[i].y * sc;
        s[i] = make_double2(log(sqrt(1.0 - (re * re + im * im))), 0.0);
    }
    __syncthreads();
    dft_inplace(s, pl, false);                                      // xlf
    for (int k = threadIdx.x; k < N; k += blockDim.x) {             // mag2mp.m:28-31
        const double g = (k == 0 || k == H) ? 1.0 : (k < H ? 2.0 : 0.0);
        s[k] = make_double2(s[k].x * g, s[k].y * g);
    }
    __syncthreads();
    dft_inplace(s, pl, true);
    const double invN = 1.0 / (double)N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {             // afa = exp(xlaf)
        const double mag = exp(s[i].x * invN);
        double sn, cs;
        sincos(s[i].y * invN, &sn, &cs);
        s[i] = make_double2(mag * cs, mag * sn);
    }
    __syncthreads();
    dft_inplace(s, pl, false);                                      // aca = fft(afa) / blp, reversed first n
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        a_re[(size_t)q * n + i] = s[n - 1 - i].x * invN;
        a_im[(size_t)q * n + i] = s[n - 1 - i].y * invN;
    }
}

// a, b: [B][n] complex -> rf: [B][n] complex
__global__ void __launch_bounds__(THREADS) ab2rf_kernel(const double *__restrict__ a_re, const double *__restrict__ a_im,
                                                        const double *__restrict__ b_re, const double *__restrict__ b_im, int n,
                                                        double *__restrict__ rf_re, double *__restrict__ rf_im)
{
    extern __shared__ double2 sm[];                                 // [2][2][n]: ping-pong of (a, b)
    const int q = blockIdx.x;
    double2 *A[2] = {sm, sm + 2 * n}, *Bc[2] = {sm + n, sm + 3 * n};
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        A[0][j] = make_double2(a_re[(size_t)q * n + j], a_im ? a_im[(size_t)q * n + j] : 0.0);
        Bc[0][j] = make_double2(b_re[(size_t)q * n + j], b_im ? b_im[(size_t)q * n + j] : 0.0);
    }
    __syncthreads();
    int cur = 0;
    for (int i = n; i >= 1; --i) {
        const double2 *ac = A[cur], *bc = Bc[cur];
        const double2 ratio = cdiv(bc[i - 1], ac[i - 1]);           // every thread: the same two shared-memory words
        const double c = sqrt(1.0 / (1.0 + (ratio.x * ratio.x + ratio.y * ratio.y)));
        const double2 sv = make_double2(c * ratio.x, -c * ratio.y); // s = conj(c b/a)
        if (threadIdx.x == 0) {
            const double theta = atan2(hypot(sv.x, sv.y), c), psi = atan2(sv.y, sv.x);
            double sn, cs;
            sincos(psi, &sn, &cs);
            rf_re[(size_t)q * n + i - 1] = 2.0 * theta * cs;
            rf_im[(size_t)q * n + i - 1] = 2.0 * theta * sn;
        }
        double2 *an = A[cur ^ 1], *bn = Bc[cur ^ 1];
        const double2 ms = make_double2(-sv.x, sv.y);               // -conj(s)
        for (int j = threadIdx.x; j < i; j += blockDim.x) {
            const double2 aj = ac[j], bj = bc[j];
            const double2 sb = cmul(sv, bj), ma = cmul(ms, aj);
            if (j >= 1) an[j - 1] = make_double2(c * aj.x + sb.x, c * aj.y + sb.y);        // acn(2:i)
            if (j <= i - 2) bn[j] = make_double2(ma.x + c * bj.x, ma.y + c * bj.y);        // bcn(1:i-1)
        }
        __syncthreads();
        cur ^= 1;
    }
}

struct Ctx {
    DeviceScratch dev;
};
static thread_local Ctx t_ctx;

static int lg8n(int n)
{
    int p = 0;
    while ((1 << p) < n) ++p;
    return ((1 << p) == n) ? p + 3 : -1;
}

}  // namespace islr
}  // namespace mbrf

using namespace mbrf;
using namespace mbrf::islr;

extern "C" {

/* aca = b2a(bc) for B beta polynomials of n coefficients (rf_tools/b2a.m:13-28); host pointers, split planes, row-major [B x n];
 * b_im may be NULL.  n = 2^k <= 1024 runs a radix-2 transform of length 8n, any other n <= 512 Bluestein's algorithm. */
int mbrf_b2a_batch(const double *b_re, const double *b_im, int n, int B, double *a_re, double *a_im)
{
    if (int rc = require_device()) return rc;
    if (!b_re || !a_re || !a_im || n < 1 || B < 1) { set_error("b2a: bad arguments (n=%d B=%d)", n, B); return MBRF_EINVAL; }
    const int N = 8 * n;
    int lg = lg8n(n);
    const bool pow2 = lg >= 0;
    if (!pow2) { lg = 0; while ((1 << lg) < 2 * N - 1) ++lg; }
    const int L = 1 << lg;
    if ((size_t)L * sizeof(double2) > 200 * 1024) {
        set_error("b2a: n=%d too long for the shared-memory transform (n = 2^k <= 1024, other n <= 512)", n);
        return MBRF_EINVAL;
    }
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t nb = (size_t)n * B * 8;
    Ctx &cx = t_ctx;
    if (int rc = cx.dev.reserve(4 * al(nb) + al((size_t)L / 2 * 16) + al((size_t)N * 16) + al((size_t)L * 16))) return rc;
    char *d = (char *)cx.dev.ptr;
    double *dbr = (double *)d, *dbi = (double *)(d + al(nb)), *dar = (double *)(d + 2 * al(nb)), *dai = (double *)(d + 3 * al(nb));
    double2 *tw = (double2 *)(d + 4 * al(nb));
    double2 *chirp = (double2 *)((char *)tw + al((size_t)L / 2 * 16));
    double2 *Bhat = (double2 *)((char *)chirp + al((size_t)N * 16));
    MBRF_CUDA(cudaMemcpyAsync(dbr, b_re, nb, cudaMemcpyHostToDevice, 0));
    if (b_im) MBRF_CUDA(cudaMemcpyAsync(dbi, b_im, nb, cudaMemcpyHostToDevice, 0));
    twiddle_kernel<<<(L / 2 + 255) / 256, 256>>>(tw, L / 2);
    MBRF_LAUNCH_CHECK();
    const size_t smem = (size_t)L * sizeof(double2);
    if (smem > 40 * 1024) {        // set on every call: the attribute is per device
        MBRF_CUDA(cudaFuncSetAttribute(b2a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MBRF_CUDA(cudaFuncSetAttribute(bluestein_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    DftPlan pl;
    pl.N = N; pl.lgL = lg; pl.tw = tw; pl.chirp = nullptr; pl.Bhat = nullptr;
    if (!pow2) {
        bluestein_setup_kernel<<<1, THREADS, smem>>>(N, lg, tw, chirp, Bhat);
        MBRF_LAUNCH_CHECK();
        pl.chirp = chirp; pl.Bhat = Bhat;
    }
    b2a_kernel<<<B, THREADS, smem>>>(dbr, b_im ? dbi : nullptr, n, pl, dar, dai);
    MBRF_LAUNCH_CHECK();
    MBRF_CUDA(cudaMemcpyAsync(a_re, dar, nb, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaMemcpyAsync(a_im, dai, nb, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaStreamSynchronize(0));
    return MBRF_OK;
}

/* rf = ab2rf(ac, bc) for B polynomial pairs of n <= 2048 coefficients (rf_tools/ab2rf.m:12-26); host pointers, split planes,
 * row-major [B x n]; the imaginary inputs may be NULL. */
int mbrf_ab2rf_batch(const double *a_re, const double *a_im, const double *b_re, const double *b_im, int n, int B, double *rf_re,
                     double *rf_im)
{
    if (int rc = require_device()) return rc;
    if (!a_re || !b_re || !rf_re || !rf_im || n < 1 || B < 1) { set_error("ab2rf: bad arguments (n=%d B=%d)", n, B); return MBRF_EINVAL; }
    if (n > 2048) { set_error("ab2rf: n=%d exceeds 2048 coefficients (shared-memory recursion)", n); return MBRF_EINVAL; }
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t nb = (size_t)n * B * 8;
    Ctx &cx = t_ctx;
    if (int rc = cx.dev.reserve(6 * al(nb))) return rc;
    char *d = (char *)cx.dev.ptr;
    double *p[6];
    for (int k = 0; k < 6; ++k) p[k] = (double *)(d + k * al(nb));
    MBRF_CUDA(cudaMemcpyAsync(p[0], a_re, nb, cudaMemcpyHostToDevice, 0));
    if (a_im) MBRF_CUDA(cudaMemcpyAsync(p[1], a_im, nb, cudaMemcpyHostToDevice, 0));
    MBRF_CUDA(cudaMemcpyAsync(p[2], b_re, nb, cudaMemcpyHostToDevice, 0));
    if (b_im) MBRF_CUDA(cudaMemcpyAsync(p[3], b_im, nb, cudaMemcpyHostToDevice, 0));
    const size_t smem = (size_t)4 * n * sizeof(double2);
    if (smem > 40 * 1024) {        // set on every call: the attribute is per device
        MBRF_CUDA(cudaFuncSetAttribute(ab2rf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    ab2rf_kernel<<<B, THREADS, smem>>>(p[0], a_im ? p[1] : nullptr, p[2], b_im ? p[3] : nullptr, n, p[4], p[5]);
    MBRF_LAUNCH_CHECK();
    MBRF_CUDA(cudaMemcpyAsync(rf_re, p[4], nb, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaMemcpyAsync(rf_im, p[5], nb, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaStreamSynchronize(0));
    return MBRF_OK;
}

}  // extern "C"
