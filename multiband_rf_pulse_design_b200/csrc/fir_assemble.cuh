// Device-side problem assembly of fir_ap_cvx: everything fir_ap_cvx.m:51-142 derives per design from the
// band specification -- band membership of every grid row, the linearly interpolated amplitude +- ripple, the transition
// bounds, the squared / floored power bounds, the stop rows, the peak radii -- written straight into the solver's padded
// [row x design] arrays.  The reference does this with MATLAB vector code per design; round 1/2 did it in numpy on the host
// (O(M) per design, then 2 x M x B doubles over PCIe).  Here one thread owns one (row, design) entry.
//
// Batches share one matrix through the UNION of the designs' grids (the base grid linspace(-pi, pi, 2 n oversamp) is common,
// the band-edge samples differ): `w` is that union, sorted and unique, built on the host (a few thousand doubles);
// `is_base[i]` says whether row i is a base sample.  Design b owns row i iff it is a base sample or equals one of b's edges.
// Arithmetic follows the numpy mirror expression by expression with explicit _rn intrinsics (no FMA contraction), so the
// bounds are BIT-identical to the host assembly (tests/test_fir_assemble_gpu.py).
#pragma once
#include <math.h>

namespace mbrf {
namespace assemble {

constexpr int MAX_BANDS = 16;

struct ApSpec {            // device pointers, design-major
    const double *f;       // [B x 2 nband] band edges in rad (already times pi, fir_ap_cvx.m:44)
    const double *a;       // [B x 2 nband] amplitudes at the edges
    const double *d;       // [B x nband] ripples
    const double *obj;     // [B] stop-band weight
    const double *peak;    // [B] Peak
    int nband, n, B;
};

struct ApRed {             // per design: reductions over its own rows
    double umax, lmin, minsq;
    int ntran, pad;
};

__device__ __forceinline__ bool owns_row(const ApSpec &s, int b, double w, bool base)
{
    if (base) return true;
    const double *f = s.f + (size_t)b * 2 * s.nband;
    for (int k = 0; k < 2 * s.nband; ++k)
        if (f[k] == w) return true;
    return false;
}

// amplitude of band k at w — fir_ap_cvx.m:57-60
__device__ __forceinline__ double band_amp(double a0, double a1, double lo, double hi, double w)
{
    if (lo == hi) return a0;
    return __dadd_rn(a0, __dmul_rn(__dsub_rn(a1, a0), __ddiv_rn(__dsub_rn(w, lo), __dsub_rn(hi, lo))));
}

// Pass 1: per design, over the rows it owns: max U, min L (transition bounds, :67-75), min sqrt(U_b) over band entries and the
// number of transition rows (stop rows, :125).  One CTA per design.
__global__ void ap_reduce_kernel(ApSpec s, const double *__restrict__ w, const unsigned char *__restrict__ is_base, int M1, ApRed *red)
{
    const int b = blockIdx.x;
    const double *f = s.f + (size_t)b * 2 * s.nband, *a = s.a + (size_t)b * 2 * s.nband, *d = s.d + (size_t)b * s.nband;
    double umax = -INFINITY, lmin = INFINITY, minsq = INFINITY;
    int ntran = 0;
    for (int i = threadIdx.x; i < M1; i += blockDim.x) {
        const double wi = w[i];
        if (!owns_row(s, b, wi, is_base[i] != 0)) continue;
        bool in_band = false;
        for (int k = 0; k < s.nband; ++k) {
            const double lo = f[2 * k], hi = f[2 * k + 1];
            if (wi >= lo && wi <= hi) {                                      // :54
                in_band = true;
                const double amp = band_amp(a[2 * k], a[2 * k + 1], lo, hi, wi);
                const double U = __dadd_rn(amp, d[k]), L = __dsub_rn(amp, d[k]);
                umax = fmax(umax, U);
                lmin = fmin(lmin, L);
                minsq = fmin(minsq, sqrt(__dmul_rn(U, U)));
            }
        }
        if (!in_band) ++ntran;
    }
    __shared__ double su[32], sl[32], sq[32];
    __shared__ int sn[32];
    for (int o = 16; o > 0; o >>= 1) {
        umax = fmax(umax, __shfl_down_sync(0xffffffffu, umax, o));
        lmin = fmin(lmin, __shfl_down_sync(0xffffffffu, lmin, o));
        minsq = fmin(minsq, __shfl_down_sync(0xffffffffu, minsq, o));
        ntran += __shfl_down_sync(0xffffffffu, ntran, o);
    }
    if ((threadIdx.x & 31) == 0) { su[threadIdx.x >> 5] = umax; sl[threadIdx.x >> 5] = lmin; sq[threadIdx.x >> 5] = minsq; sn[threadIdx.x >> 5] = ntran; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) { umax = fmax(umax, su[k]); lmin = fmin(lmin, sl[k]); minsq = fmin(minsq, sq[k]); ntran += sn[k]; }
        if (ntran > 0) minsq = fmin(minsq, sqrt(__dmul_rn(umax, umax)));     // transition rows carry U_tran = max U, :67-70
        ApRed r;
        r.umax = umax; r.lmin = lmin; r.minsq = minsq; r.ntran = ntran; r.pad = 0;
        red[b] = r;
    }
}

// Bounds and stop flag of design b at grid value wi (which b owns) — fir_ap_cvx.m:53-82,103-125.  A value that lies in two
// touching bands is one row here (the union grid is unique): the intersection of its bounds, a stop row if either entry is.
__device__ __forceinline__ void ap_row(const ApSpec &s, int b, const ApRed &r, double wi, double &lo_out, double &hi_out, bool &stop)
{
    const double *f = s.f + (size_t)b * 2 * s.nband, *a = s.a + (size_t)b * 2 * s.nband, *d = s.d + (size_t)b * s.nband;
    const double thr = __dadd_rn(r.minsq, 1e-2);                             // :125
    double lo = -INFINITY, hi = INFINITY;
    bool in_band = false;
    stop = false;
    auto entry = [&](double U, double L) {
        const double Ub = __dmul_rn(U, U);                                   // :103-106
        double Lc = L < 0.0 ? 0.0 : L;                                       // :110-112
        double Lb = __dmul_rn(Lc, Lc);
        if (Lb < 1e-20) Lb = 1e-20;                                          // :115-116
        hi = fmin(hi, Ub);
        lo = fmax(lo, Lb);
        if (sqrt(Ub) < thr) stop = true;
    };
    for (int k = 0; k < s.nband; ++k) {
        const double e0 = f[2 * k], e1 = f[2 * k + 1];
        if (wi >= e0 && wi <= e1) {
            in_band = true;
            const double amp = band_amp(a[2 * k], a[2 * k + 1], e0, e1, wi);
            entry(__dadd_rn(amp, d[k]), __dsub_rn(amp, d[k]));
        }
    }
    if (!in_band) entry(r.umax, fmin(0.0, r.lmin));                          // :67-75
    lo_out = lo;
    hi_out = hi;
}

// stop_any[i] = 1 iff row i is a stop row of at least one design (these rows are appended once more as the stop block).
// One warp per row.
__global__ void ap_stop_any_kernel(ApSpec s, const double *__restrict__ w, const unsigned char *__restrict__ is_base, int M1,
                                   const ApRed *__restrict__ red, unsigned char *__restrict__ stop_any)
{
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= M1) return;
    const double wi = w[i];
    const bool base = is_base[i] != 0;
    bool any = false;
    for (int b = lane; b < s.B; b += 32) {
        if (!owns_row(s, b, wi, base)) continue;
        double lo, hi;
        bool st;
        ap_row(s, b, red[b], wi, lo, hi, st);
        any |= st;
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) stop_any[i] = any ? 1 : 0;
}

// Pass 2: lo / hi [Mp x Bp] of the padded problem.  Rows [0, M1): the union grid; rows [M1, M1 + ns): the stop block, row k a
// copy of union row srows[k] with hi = 0 where it is a stop row of the design (membership flag of the block) and +inf
// elsewhere; rows beyond and designs b >= B: (-inf, +inf).
__global__ void ap_fill_rows_kernel(ApSpec s, const double *__restrict__ w, const unsigned char *__restrict__ is_base, int M1,
                                    const int *__restrict__ srows, int ns, const ApRed *__restrict__ red, int Mp, int Bp,
                                    double *__restrict__ lo, double *__restrict__ hi)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (b >= Bp || row >= Mp) return;
    double l = -INFINITY, h = INFINITY;
    if (b < s.B && row < M1 + ns) {
        const int i = row < M1 ? row : srows[row - M1];
        const double wi = w[i];
        if (owns_row(s, b, wi, is_base[i] != 0)) {
            double rl, rh;
            bool st;
            ap_row(s, b, red[b], wi, rl, rh, st);
            if (row < M1) { l = rl; h = rh; }
            else if (st) h = 0.0;
        }
    }
    lo[(size_t)row * Bp + b] = l;
    hi[(size_t)row * Bp + b] = h;
}

// Column-side arrays: c = e_1 (minimise x(1) + ..., :163), |x(1)| <= n Peak (:167, i = 1), radii (n - i + 1) Peak of the
// pairs (x_i, x_{n+i-1}), i = 2..n (:166-168), stop-band weight.  Padding as the host upload does it.
__global__ void ap_fill_cols_kernel(ApSpec s, int Np, int Bp, int npairs, double *__restrict__ c, double *__restrict__ bl,
                                    double *__restrict__ bu, double *__restrict__ rho, double *__restrict__ ct)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (b >= Bp) return;
    const bool live = b < s.B;
    if (j < Np) {
        const size_t o = (size_t)j * Bp + b;
        c[o] = (live && j == 0) ? 1.0 : 0.0;
        const double r0 = live ? __dmul_rn((double)s.n, s.peak[b]) : INFINITY;
        bl[o] = (j == 0 && live) ? -r0 : -INFINITY;
        bu[o] = (j == 0 && live) ? r0 : INFINITY;
    }
    if (j < npairs) rho[(size_t)j * Bp + b] = live ? __dmul_rn((double)(s.n - 1 - j), s.peak[b]) : 1.0;
    if (j == 0) ct[b] = live ? s.obj[b] : 0.0;
}

// r = [conj(r(end:-1:2)), r] with r = [x(1), x(2:n) + 1i*x(n+1:2n-1)]  (fir_ap_cvx.m:185-186): x [Np x Bp] column layout in,
// r_re / r_im [B x (2n-1)] row-major out — the input layout of the fmp2 kernel.
__global__ void ap_x_to_r_kernel(const double *__restrict__ x, int n, int B, int Bp, double *__restrict__ r_re, double *__restrict__ r_im,
                                 double *__restrict__ x_rows)
{
    const int b = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;     // t = 0 .. 2n-2 (output position)
    const int len = 2 * n - 1;
    if (b >= B || t >= len) return;
    const int k = t - (n - 1);                                               // lag -(n-1) .. n-1
    const int ak = k < 0 ? -k : k;
    const double re = x[(size_t)ak * Bp + b];
    const double im = ak == 0 ? 0.0 : x[(size_t)(n - 1 + ak) * Bp + b];
    r_re[(size_t)b * len + t] = re;
    r_im[(size_t)b * len + t] = k < 0 ? -im : im;
    x_rows[(size_t)b * len + t] = x[(size_t)t * Bp + b];                     // the solution itself, design-major
}

}  // namespace assemble
}  // namespace mbrf
