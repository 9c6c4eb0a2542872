// Batched zero flipping — the loop of fir_flip_zero.m:66-99.
//
// The reference takes the zeros Z = roots(h) of a filter, and for each of up to 2^12 flip patterns (columns of `mask`)
// reflects the selected passband zeros about the unit circle (`flip_zero`, :112-117), expands the polynomial again with
// poly() (:70), rescales it to the DC gain of h (:71), records power and peak (:74-75), then keeps the pattern with the
// smallest peak (:96-99).  That is `Num` independent O(N^2) recursions, one after the other in MATLAB.
//
// Here: one CTA per pattern.  The coefficients live in shared memory (ping-pong, one barrier per zero), the zeros are
// multiplied in in the ORDER of Z as poly() does (c(2:j+1) -= e(j) * c(1:j)), so the result agrees with the reference's
// to rounding; a second one-CTA kernel takes the first minimum of the peaks and copies that pattern's taps out.
#include <math.h>
#include "common.h"

namespace mbrf {
namespace flipzero {

// 1/|z| * exp(i*angle(z))  — fir_flip_zero.m:116
__global__ void reflect_kernel(const double *z_re, const double *z_im, const int *idx_pb, int n_pb, double2 *zf)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pb) return;
    const double zr = z_re[idx_pb[k]], zi = z_im ? z_im[idx_pb[k]] : 0.0;
    const double inv = 1.0 / hypot(zr, zi);
    double sn, cs;
    sincos(atan2(zi, zr), &sn, &cs);
    zf[k] = make_double2(inv * cs, inv * sn);
}

__device__ __forceinline__ double2 block_sum2(double2 v, double2 *scratch)
{
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_down_sync(0xffffffffu, v.x, o);
        v.y += __shfl_down_sync(0xffffffffu, v.y, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nw ? scratch[lane] : make_double2(0.0, 0.0);
        for (int o = 16; o > 0; o >>= 1) {
            v.x += __shfl_down_sync(0xffffffffu, v.x, o);
            v.y += __shfl_down_sync(0xffffffffu, v.y, o);
        }
        if (lane == 0) scratch[0] = v;
    }
    __syncthreads();
    return scratch[0];
}

// pb_pos[j]: position of zero j among the passband zeros, or -1; mask [nmask x n_pb] (1 = flipped)
__global__ void expand_kernel(const double *z_re, const double *z_im, int nroots, const int *pb_pos, const double2 *zf,
                              const unsigned char *mask, int n_pb, double hsum_re, double hsum_im, double *h_re, double *h_im,
                              double *peak, double *power)
{
    extern __shared__ double2 sm[];
    const int N = nroots + 1, m = blockIdx.x;
    double2 *buf[2] = {sm, sm + N};
    double2 *scratch = sm + 2 * N;                       // 32 entries
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const double2 v = make_double2(k == 0 ? 1.0 : 0.0, 0.0);
        buf[0][k] = v;
        buf[1][k] = v;
    }
    __syncthreads();
    int cur = 0;
    const unsigned char *mrow = mask + (size_t)m * n_pb;
    for (int j = 0; j < nroots; ++j) {
        const int p = pb_pos[j];
        double2 e;                                       // Z_each(j), :69
        if (p >= 0 && mrow[p]) e = zf[p];
        else e = make_double2(z_re[j], z_im ? z_im[j] : 0.0);
        const double2 *co = buf[cur];
        double2 *cn = buf[cur ^ 1];
        for (int k = threadIdx.x + 1; k <= j + 1; k += blockDim.x) {   // c(2:j+1) = c(2:j+1) - e(j)*c(1:j), poly.m
            const double2 a = co[k - 1], b = co[k];
            const double pr = __dsub_rn(__dmul_rn(e.x, a.x), __dmul_rn(e.y, a.y));
            const double pi = __dadd_rn(__dmul_rn(e.x, a.y), __dmul_rn(e.y, a.x));
            cn[k] = make_double2(__dsub_rn(b.x, pr), __dsub_rn(b.y, pi));
        }
        __syncthreads();
        cur ^= 1;
    }
    const double2 *c = buf[cur];
    double2 part = make_double2(0.0, 0.0);
    for (int k = threadIdx.x; k < N; k += blockDim.x) { part.x += c[k].x; part.y += c[k].y; }
    const double2 s = block_sum2(part, scratch);         // sum(h_each)
    // h_each * sum(h) / sum(h_each), :71 (evaluated left to right: the product first, then the complex division)
    const double den = s.x * s.x + s.y * s.y;
    double pw = 0.0, pk = 0.0;
    bool bad = false;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const double2 v = c[k];
        const double nr = v.x * hsum_re - v.y * hsum_im, ni = v.x * hsum_im + v.y * hsum_re;
        const double hr = (nr * s.x + ni * s.y) / den, hi = (ni * s.x - nr * s.y) / den;
        h_re[(size_t)m * N + k] = hr;
        h_im[(size_t)m * N + k] = hi;
        const double ab = hypot(hr, hi);
        pw += ab * ab;                                   // sum(abs(h_each).^2), :74
        if (ab != ab) bad = true;
        pk = fmax(pk, ab);                               // max(abs(h_each)), :75
    }
    const double2 r = block_sum2(make_double2(pw, bad ? 1.0 : 0.0), scratch);
    // block maximum through the same scratch
    for (int o = 16; o > 0; o >>= 1) pk = fmax(pk, __shfl_down_sync(0xffffffffu, pk, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5].x = pk;
    __syncthreads();
    if (threadIdx.x == 0) {
        double best = 0.0;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) best = fmax(best, scratch[w].x);
        peak[m] = r.y > 0.0 ? nan("") : best;            // MATLAB's max() skips NaN, min() below skips a NaN peak
        power[m] = r.x;
    }
}

// [~, min_idx] = min(peak), h_new = h_array(:, min_idx) — :96-99.  First index of the minimum; NaN peaks are skipped.
__global__ void pick_kernel(const double *peak, int nmask, const double *h_re, const double *h_im, int N, int *best_out,
                            double *out_re, double *out_im)
{
    __shared__ double sv[1024];
    __shared__ int si[1024];
    double bv = INFINITY;
    int bi = 0x7fffffff;
    for (int m = threadIdx.x; m < nmask; m += blockDim.x) {
        const double p = peak[m];
        if (p < bv) { bv = p; bi = m; }                  // strided scan keeps the smallest index within a thread
    }
    sv[threadIdx.x] = bv;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double v = sv[threadIdx.x + o];
            const int i = si[threadIdx.x + o];
            if (v < sv[threadIdx.x] || (v == sv[threadIdx.x] && i < si[threadIdx.x])) { sv[threadIdx.x] = v; si[threadIdx.x] = i; }
        }
        __syncthreads();
    }
    const int best = si[0] == 0x7fffffff ? 0 : si[0];    // all NaN: MATLAB returns index 1
    if (threadIdx.x == 0) *best_out = best;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        out_re[k] = h_re[(size_t)best * N + k];
        out_im[k] = h_im[(size_t)best * N + k];
    }
}

struct Ctx {
    DeviceScratch dev;
};
static thread_local Ctx t_ctx;

}  // namespace flipzero
}  // namespace mbrf

using namespace mbrf;
using namespace mbrf::flipzero;

extern "C" {

int mbrf_flip_zero_max_taps(void) { return 4096; }

int mbrf_flip_zero_batch(const double *z_re, const double *z_im, int nroots, const int *idx_pb, int n_pb,
                         const unsigned char *mask, int nmask, double hsum_re, double hsum_im, int *best, double *h_re,
                         double *h_im, double *peak, double *power, double *h_all_re, double *h_all_im)
{
    if (int rc = require_device()) return rc;
    const int N = nroots + 1;
    if (!z_re || !h_re || !h_im || nroots < 1 || nmask < 1 || n_pb < 0 || (n_pb > 0 && (!idx_pb || !mask))) {
        set_error("flip_zero: bad arguments (nroots=%d n_pb=%d nmask=%d)", nroots, n_pb, nmask);
        return MBRF_EINVAL;
    }
    if (N > mbrf_flip_zero_max_taps()) {
        set_error("flip_zero: %d taps exceed %d (shared-memory recursion)", N, mbrf_flip_zero_max_taps());
        return MBRF_EINVAL;
    }
    // position of every zero among the passband zeros (host: nroots integers)
    int *pb_pos = (int *)malloc(sizeof(int) * (size_t)nroots);
    if (!pb_pos) { set_error("flip_zero: out of host memory"); return MBRF_ENOMEM; }
    for (int j = 0; j < nroots; ++j) pb_pos[j] = -1;
    for (int k = 0; k < n_pb; ++k) {
        if (idx_pb[k] < 0 || idx_pb[k] >= nroots || pb_pos[idx_pb[k]] >= 0) {
            free(pb_pos);
            set_error("flip_zero: idx_pb[%d] = %d out of range or repeated", k, idx_pb[k]);
            return MBRF_EINVAL;
        }
        pb_pos[idx_pb[k]] = k;
    }
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t zb = al((size_t)nroots * 8), ib = al((size_t)(nroots > n_pb ? nroots : n_pb) * 4), fb = al((size_t)(n_pb + 1) * 16),
                 mb = al((size_t)nmask * (n_pb > 0 ? n_pb : 1)), hb = al((size_t)nmask * N * 8), pb = al((size_t)nmask * 8),
                 ob = al((size_t)N * 8);
    Ctx &cx = t_ctx;
    int rc = cx.dev.reserve(2 * zb + 2 * ib + fb + mb + 2 * hb + 2 * pb + 2 * ob + 256);
    if (rc) { free(pb_pos); return rc; }
    char *d = (char *)cx.dev.ptr;
    double *dzr = (double *)d; d += zb;
    double *dzi = (double *)d; d += zb;
    int *dpos = (int *)d; d += ib;
    int *didx = (int *)d; d += ib;
    double2 *dzf = (double2 *)d; d += fb;
    unsigned char *dmask = (unsigned char *)d; d += mb;
    double *dhr = (double *)d; d += hb;
    double *dhi = (double *)d; d += hb;
    double *dpk = (double *)d; d += pb;
    double *dpw = (double *)d; d += pb;
    double *dor = (double *)d; d += ob;
    double *doi = (double *)d; d += ob;
    int *dbest = (int *)d;
    cudaError_t e = cudaMemcpyAsync(dzr, z_re, (size_t)nroots * 8, cudaMemcpyHostToDevice, 0);
    if (e == cudaSuccess && z_im) e = cudaMemcpyAsync(dzi, z_im, (size_t)nroots * 8, cudaMemcpyHostToDevice, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dpos, pb_pos, (size_t)nroots * 4, cudaMemcpyHostToDevice, 0);
    if (e == cudaSuccess && n_pb) e = cudaMemcpyAsync(didx, idx_pb, (size_t)n_pb * 4, cudaMemcpyHostToDevice, 0);
    if (e == cudaSuccess && n_pb) e = cudaMemcpyAsync(dmask, mask, (size_t)nmask * n_pb, cudaMemcpyHostToDevice, 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);        // pb_pos is pageable and freed below
    free(pb_pos);
    MBRF_CUDA(e);
    if (n_pb) {
        reflect_kernel<<<(n_pb + 127) / 128, 128>>>(dzr, z_im ? dzi : nullptr, didx, n_pb, dzf);
        MBRF_LAUNCH_CHECK();
    }
    int threads = (N / 2 + 31) / 32 * 32;                  // the recursion updates j+1 <= N-1 coefficients, N/2 on average
    threads = threads < 64 ? 64 : (threads > 512 ? 512 : threads);
    const size_t smem = ((size_t)2 * N + 32) * sizeof(double2);
    if (smem > 40 * 1024) MBRF_CUDA(cudaFuncSetAttribute(expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    expand_kernel<<<nmask, threads, smem>>>(dzr, z_im ? dzi : nullptr, nroots, dpos, dzf, dmask, n_pb, hsum_re, hsum_im, dhr, dhi,
                                            dpk, dpw);
    MBRF_LAUNCH_CHECK();
    pick_kernel<<<1, 1024>>>(dpk, nmask, dhr, dhi, N, dbest, dor, doi);
    MBRF_LAUNCH_CHECK();
    MBRF_CUDA(cudaMemcpyAsync(h_re, dor, (size_t)N * 8, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaMemcpyAsync(h_im, doi, (size_t)N * 8, cudaMemcpyDeviceToHost, 0));
    if (best) MBRF_CUDA(cudaMemcpyAsync(best, dbest, 4, cudaMemcpyDeviceToHost, 0));
    if (peak) MBRF_CUDA(cudaMemcpyAsync(peak, dpk, (size_t)nmask * 8, cudaMemcpyDeviceToHost, 0));
    if (power) MBRF_CUDA(cudaMemcpyAsync(power, dpw, (size_t)nmask * 8, cudaMemcpyDeviceToHost, 0));
    if (h_all_re) MBRF_CUDA(cudaMemcpyAsync(h_all_re, dhr, (size_t)nmask * N * 8, cudaMemcpyDeviceToHost, 0));
    if (h_all_im) MBRF_CUDA(cudaMemcpyAsync(h_all_im, dhi, (size_t)nmask * N * 8, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaStreamSynchronize(0));
    return MBRF_OK;
}

}  // extern "C"
