// Batched restarted PDHG for the convex FIR design step (fir_ap_cvx.m:160-169, ss/fir_linprog.m:246,
// and — through polygonal/disk rows — fir_qp_cvx.m) on B200.
//
// Canonical form, one dense frequency-sampled Fourier matrix K shared by a batch of B designs:
//
//        minimise  c^T z    subject to   lo <= K z <= hi ,   z in X
//        X = product of boxes  bl_j <= z_j <= bu_j  and 2-D disks  ||(z_i, z_j)|| <= rho
//
// Every per-design vector is stored [dim x Bp] with the design index fastest, so that
//   * the two products of an iteration, K * Zbar and K^T * Y, are dense fp64 GEMMs
//     [Mp x Np] x [Np x Bp] and [Np x Mp] x [Mp x Bp] (kernels dgemm_nn / dgemm_tn below), and
//   * all vector kernels (prox, projections, reductions) are coalesced over designs.
// Batches of 1..8 designs take matrix-vector kernels (the matrix then streams from L2).
//
// Iteration (Chambolle-Pock with PDLP-style restarts and primal weight, per design b):
//     g    = c + K^T y
//     z+   = P_X(z - tau_b g)            zbar = 2 z+ - z
//     v    = y + sigma_b K zbar          y+   = v - sigma_b clip(v / sigma_b, lo, hi)
//     tau_b = eta / omega_b, sigma_b = eta * omega_b, eta = 0.9 / ||K||_2
// Every `check_every` iterations both the running average and the current iterate are scored by
//     err = max(row violation, natural residual ||z - P_X(z - g)||_inf, |c^T z - (-h*(y) + g^T z)|)
// and the better one becomes the restart candidate (sufficient / necessary decay or 0.36 rule).
//
// Precision: fp64 state and fp64 convergence checks.  The two products of every iteration run on the tcgen05 tensor cores
// as a split-integer product (tc_gemm.cuh: int8 digit planes, exact int32 level sums in TMEM, ~1e-11 relative: the solver takes
// exactly the iterations it takes on fp64 tiles); the fp64 mma.sync / SIMT tile kernels below serve the checks, the power
// iteration and mbrf_pdhg_set_gemm(1 / 0).  See DESIGN.md section 6 for the measured numbers.
#include "common.h"
#include "tc_gemm.cuh"

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dgemm.cuh"

namespace mbrf {
namespace pdhg {

// ---------------------------------------------------------------------------
// solver state
// ---------------------------------------------------------------------------
struct Ctl {           // per-design control block (device), doubles for simplicity
    double tau, sigma, omega;
    double last_err, prev_err;
    double since, cnt;          // iterations since last restart / samples in the running sums
    double status;              // 0 running, 1 solved, 2 infeasible (certificate), 3 iteration limit
    double iters;               // iteration at which status was decided
    double obj, dual, pr, dr, rigorous, tmax;
    double restart, use_avg;    // decisions of the last check
};

struct Problem {
    int Mp, Np, Bp, B, npairs, P;       // padded sizes, live designs, pairs, split-K slabs
    int ldk;
    const double *K, *KT;                       // K [Mp x ldk] and its transpose [Np x Mp]
    double *c, *lo, *hi, *bl, *bu, *rho;        // [dim x Bp]  (compacted in place when designs finish)
    const int *pair_i, *pair_j;                 // [npairs]
    const int *pair_of;                         // [Np]: pair index of a coordinate or -1
    double *obj_upper;                          // [Bp] or null
    int srow0, ns;                              // simplex block: rows [srow0, srow0+ns) carry  w_b * max_i (K z)_i
    double *sw;                                 // [Bp] weights w_b of that block (null when ns == 0)
    int drow0, nd;                              // disk block: row pairs (drow0+2i, drow0+2i+1): ||K z - centre|| <= R,
                                                //   centre in lo[pair], R in hi[first row]          (fir_qp_cvx.m:148-157)
    int grow0, ng;                              // group block: row pairs carry  gw_b * max_i ||(K z)_pair_i||  (obj*Peak,
    double *gw;                                 //   fir_qp_cvx.m:147,158-160); multipliers live in the l1,2 ball of radius gw_b
    int grow2, ng2;                             // centred group block: row pairs carry  gw2_b * max_i ||(K z)_pair_i - (lo_r, lo_r+1)||
    double *gw2;                                //   (delta of the minimax form, fir_qp_cvx.m:170-177, rows pre-scaled by 1/D_i)
    int sp_lo, sp_hi;                           // rows outside [sp_lo, sp_hi) are plain interval rows (no block owns them)
    int nn;                                     // norm term: lam_b * ||z[0..nn)||_2 in the objective (E_total, :147,161)
    double *lam;                                // [Bp]
    double *nrm;                                // [4 x Bp] scratch: ||zhat||^2 per design (z-update), metrics sums
    double *z, *zbar, *zs, *z0, *zbest;         // [Np x Bp]
    double *y, *ys, *y0, *ybest;                // [Mp x Bp]
    double *S;                                  // [Mp x Bp]   K*zbar, K*z, K*zs
    double *G;                                  // [P x Np x Bp] split-K slabs of K^T y
    double *G2;                                 // [Np x Bp] reduced K^T (.) for metrics
    double *acc;                                // [2 cand][NACC][Bp] reduction targets
    double *mxy, *mxz;                          // [Bp] max |y|, max |zbar| per design for the tcgen05 digit planes (or null)
    Ctl *ctl;
    int *active;                                // number of designs still running
    int halpern, kk;                            // reflected Halpern iteration (mbrf_pdhg_set_halpern); kk: iteration index inside the block
    int *iter_dev;                              // iterations done (advanced by the check block, so the check can live in the graph)
    double eta, eps_pr, eps_dr, eps_gap;
    int check_every;
    double beta_suff, beta_nec, beta_art, omega_theta, min_since;   // restart rule constants (PDLP defaults 0.2, 0.8, 0.36, 0.5)
};

enum { A_PR = 0, A_DR, A_POBJ, A_GZ, A_HS, A_DZ2, A_DY2, A_RIGX, A_TMAX, A_GMAX, A_ZZ, A_ZV, A_VV, A_GMAX2, A_RIGN, A_GG, A_RR, NACC };

__device__ __forceinline__ void atomic_max_pos(double *addr, double v)
{
    if (!(v > 0.0)) return;
    atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

// Reflected Halpern PDHG (Lu & Yang 2024, "r2HPDHG"): with T the PDHG step and (z0, y0) the restart anchor,
//     (z, y) <- rho_k (2 T(z, y) - (z, y)) + (1 - rho_k) (z0, y0),   rho_k = (k + 1) / (k + 2),  k = iterations since the restart.
// The kernels keep the PDHG output T(z, y) in the arrays of the running sums (zs, ys with cnt = 1: it is the candidate of the
// convergence checks and the restart point) and store the Halpern iterate as the current (z, y).  On the numpy twin
// (tools/halpern_proto.py) this needs 1.1-2.8x fewer iterations than the averaged restarts on the fir_ap_cvx problems.
__device__ __forceinline__ double hp_rho(const Problem &p, int b)
{
    const double k = p.ctl[b].since + (double)p.kk;
    return (k + 1.0) / (k + 2.0);
}
// primal side: old iterate zo, PDHG output zn -> writes z, zbar (= 2 zn - zo, the extrapolation of the dual step) and zs
__device__ __forceinline__ double store_z(const Problem &p, size_t o, double zo, double zn, double rho)
{
    const double zb = 2.0 * zn - zo;
    p.zbar[o] = zb;
    if (p.halpern) { p.z[o] = fma(rho, zb - p.z0[o], p.z0[o]); p.zs[o] = zn; }
    else { p.z[o] = zn; p.zs[o] += zn; }
    return zb;
}
// dual side: old iterate yo, PDHG output yn -> writes y and ys, returns |y| as stored (for the digit-plane maxima)
__device__ __forceinline__ double store_y(const Problem &p, size_t o, double yo, double yn, double rho)
{
    if (p.halpern) {
        const double a0 = p.y0[o], y2 = fma(rho, (2.0 * yn - yo) - a0, a0);
        p.y[o] = y2; p.ys[o] = yn;
        return fabs(y2);
    }
    p.y[o] = yn; p.ys[o] += yn;
    return fabs(yn);
}

// Norm term (fir_qp_cvx.m: E_total with norm(x,2) <= E_total): the primal prox is the block soft-threshold
//   z+ = max(0, 1 - tau*lam/||zhat||) * zhat,   zhat = z - tau*(c + K^T y)
// phase 1 stores zhat in zbar and accumulates ||zhat||^2 per design; phase 2 applies the shrink.
__global__ void z_hat_kernel(Problem p)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.Bp) return;
    const double tau = p.ctl[b].tau;
    const size_t stride = (size_t)p.Np * p.Bp;
    double s2 = 0.0;
    for (int j = blockIdx.y; j < p.Np; j += gridDim.y) {
        const size_t o = (size_t)j * p.Bp + b;
        double g = p.c[o];
        for (int s = 0; s < p.P; ++s) g += p.G[s * stride + o];
        double zh = p.z[o] - tau * g;
        zh = fmin(fmax(zh, p.bl[o]), p.bu[o]);
        p.zbar[o] = zh;
        if (j < p.nn) s2 = fma(zh, zh, s2);
    }
    atomicAdd(p.nrm + b, s2);
}
// thin batches (Bp <= 8): one CTA per design, the coordinates spread over its threads
__global__ void __launch_bounds__(256) z_hat_thin_kernel(Problem p)
{
    __shared__ double sh[256];
    const int b = blockIdx.x;
    const double tau = p.ctl[b].tau;
    const size_t stride = (size_t)p.Np * p.Bp;
    double s2 = 0.0;
    for (int j = threadIdx.x; j < p.Np; j += 256) {
        const size_t o = (size_t)j * p.Bp + b;
        double g = p.c[o];
        for (int s = 0; s < p.P; ++s) g += p.G[s * stride + o];
        double zh = p.z[o] - tau * g;
        zh = fmin(fmax(zh, p.bl[o]), p.bu[o]);
        p.zbar[o] = zh;
        if (j < p.nn) s2 = fma(zh, zh, s2);
    }
    sh[threadIdx.x] = s2;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k]; __syncthreads(); }
    if (threadIdx.x == 0) p.nrm[b] = sh[0];
}
__global__ void z_shrink_kernel(Problem p)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)p.Np * p.Bp) return;
    const int b = (int)(idx % p.Bp);
    const int j = (int)(idx / p.Bp);
    const double nv = sqrt(p.nrm[b]);
    const double sh = (j < p.nn && nv > 0.0) ? fmax(0.0, 1.0 - p.ctl[b].tau * p.lam[b] / nv) : 1.0;
    const double zo = p.z[idx], zn = sh * p.zbar[idx];
    store_z(p, (size_t)idx, zo, zn, hp_rho(p, b));
}

// z+ = P_X(z - tau (c + sum_p G_p)), zbar = 2 z+ - z, zs += z+.   One thread per (coordinate, design);
// the thread of pair member i handles both members (j is skipped).  Returns max |zbar| written.
__device__ __forceinline__ double z_update_elem(const Problem &p, int j, int b)
{
    const int pr = p.pair_of[j];
    if (pr >= 0 && p.pair_j[pr] == j) return 0.0;  // second member: done by the first
    const double tau = p.ctl[b].tau;
    const size_t stride = (size_t)p.Np * p.Bp;
    auto grad = [&](int jj) {
        double g = p.c[(size_t)jj * p.Bp + b];
        for (int s = 0; s < p.P; ++s) g += p.G[s * stride + (size_t)jj * p.Bp + b];
        return g;
    };
    const size_t o = (size_t)j * p.Bp + b;
    if (pr < 0) {
        const double zo = p.z[o];
        double zn = zo - tau * grad(j);
        zn = fmin(fmax(zn, p.bl[o]), p.bu[o]);
        return fabs(store_z(p, o, zo, zn, hp_rho(p, b)));
    }
    const int j2 = p.pair_j[pr];
    const size_t o2 = (size_t)j2 * p.Bp + b;
    const double z1 = p.z[o], z2 = p.z[o2];
    double a = z1 - tau * grad(j), c2 = z2 - tau * grad(j2);
    const double r = hypot(a, c2), rho = p.rho[(size_t)pr * p.Bp + b];
    if (r > rho) {
        const double s = rho / r;
        a *= s;
        c2 *= s;
    }
    const double hr = hp_rho(p, b);
    const double zb1 = store_z(p, o, z1, a, hr), zb2 = store_z(p, o2, z2, c2, hr);
    return fmax(fabs(zb1), fabs(zb2));
}
__global__ void z_update_kernel(Problem p)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = (int)(idx % p.Bp);
    const int j = (int)(idx / p.Bp);
    if (j >= p.Np) return;
    z_update_elem(p, j, b);
}
// Thin batches (<= 8 designs): one warp per (coordinate, design); the lanes read the split-K slabs of the gradient (up to 32) in
// parallel and shuffle-reduce them.  With one thread per item the 25 dependent slab loads made this the longest kernel of a
// thin iteration (23 us of 34 at one design); the update itself is z_update_elem's.
__global__ void __launch_bounds__(256) z_update_thin_kernel(Problem p)
{
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (item >= p.Np * p.Bp) return;
    const int b = item % p.Bp, j = item / p.Bp;
    const int pr = p.pair_of[j];
    if (pr >= 0 && p.pair_j[pr] == j) return;      // second member: done by the first
    const size_t stride = (size_t)p.Np * p.Bp;
    auto grad = [&](int jj) {
        double g = 0.0;
        for (int s = lane; s < p.P; s += 32) g += p.G[s * stride + (size_t)jj * p.Bp + b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
        return g + p.c[(size_t)jj * p.Bp + b];
    };
    const double tau = p.ctl[b].tau;
    const size_t o = (size_t)j * p.Bp + b;
    if (pr < 0) {
        const double g = grad(j);
        if (lane) return;
        const double zo = p.z[o];
        double zn = zo - tau * g;
        zn = fmin(fmax(zn, p.bl[o]), p.bu[o]);
        store_z(p, o, zo, zn, hp_rho(p, b));
        return;
    }
    const int j2 = p.pair_j[pr];
    const double g1 = grad(j), g2 = grad(j2);
    if (lane) return;
    const size_t o2 = (size_t)j2 * p.Bp + b;
    const double z1 = p.z[o], z2 = p.z[o2];
    double a = z1 - tau * g1, c2 = z2 - tau * g2;
    const double r = hypot(a, c2), rho = p.rho[(size_t)pr * p.Bp + b];
    if (r > rho) {
        const double sc = rho / r;
        a *= sc;
        c2 *= sc;
    }
    const double hr = hp_rho(p, b);
    store_z(p, o, z1, a, hr);
    store_z(p, o2, z2, c2, hr);
}

// Batches of >= 64 designs: a CTA owns 64 designs (tx) and strides the coordinates (ty, blockIdx.y), so that the per-design
// max |zbar| the tcgen05 digit planes need costs one shared-memory reduction and one atomic per design and CTA.
__global__ void __launch_bounds__(256) z_update_wide_kernel(Problem p)
{
    __shared__ double sh[4][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int b = blockIdx.x * 64 + tx;
    double m = 0.0;
    for (int j = blockIdx.y * 4 + ty; j < p.Np; j += 4 * gridDim.y) m = fmax(m, z_update_elem(p, j, b));
    if (!p.mxz) return;
    sh[ty][tx] = m;
    __syncthreads();
    if (ty == 0) atomic_max_pos(p.mxz + b, fmax(fmax(sh[0][tx], sh[1][tx]), fmax(sh[2][tx], sh[3][tx])));
}

// v = y + sigma S;  y+ = v - sigma clip(v/sigma, lo, hi);  ys += y+.  Returns max |y+| written.
__device__ __forceinline__ double y_update_elem(const Problem &p, int row, int b, long long idx)
{
    if (row >= p.srow0 && row < p.srow0 + p.ns) return 0.0;   // simplex block: simplex_update_kernel
    if (row >= p.grow0 && row < p.grow0 + 2 * p.ng) return 0.0;   // group block: group_update_kernel
    if (row >= p.grow2 && row < p.grow2 + 2 * p.ng2) return 0.0;  // centred group block: group_update_kernel
    const double sig = p.ctl[b].sigma;
    if (row >= p.drow0 && row < p.drow0 + 2 * p.nd) {     // disk pair: y+ = v - sigma * P_disk(v / sigma)
        if ((row - p.drow0) & 1) return 0.0;              // the first row of the pair does both
        const long long i2 = idx + p.Bp;
        const double yo1 = p.y[idx], yo2 = p.y[i2];
        const double v1 = yo1 + sig * p.S[idx], v2 = yo2 + sig * p.S[i2];
        const double c1 = p.lo[idx], c2 = p.lo[i2], R = p.hi[idx];
        const double d1 = v1 / sig - c1, d2 = v2 / sig - c2;
        const double dn = hypot(d1, d2);
        double y1 = 0.0, y2 = 0.0;
        if (dn > R) {                                     // outside: P = c + R d/|d|  ->  y = sigma (d - R d/|d|)
            const double f = sig * (1.0 - R / dn);
            y1 = f * d1;
            y2 = f * d2;
        }
        const double hr = hp_rho(p, b);
        return fmax(store_y(p, (size_t)idx, yo1, y1, hr), store_y(p, (size_t)i2, yo2, y2, hr));
    }
    const double yo = p.y[idx];
    const double v = yo + sig * p.S[idx];
    const double w = v / sig, lo = p.lo[idx], hi = p.hi[idx];
    // exact zero inside the interval: v - sig*(v/sig) would leave rounding dust that h*(y) multiplies by +-inf
    const double yn = w > hi ? v - sig * hi : (w < lo ? v - sig * lo : 0.0);
    return store_y(p, (size_t)idx, yo, yn, hp_rho(p, b));
}
__global__ void y_update_kernel(Problem p)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)p.Mp * p.Bp) return;
    y_update_elem(p, (int)(idx / p.Bp), (int)(idx % p.Bp), idx);
}
// wide batches: see z_update_wide_kernel.  Four rows per thread and pass, all loads issued before the first store
// (the arrays never alias, which the compiler cannot know through the Problem struct), plain interval rows inline.
__global__ void __launch_bounds__(256) y_update_wide_kernel(Problem p)
{
    __shared__ double sh[4][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int b = blockIdx.x * 64 + tx;
    const int step = 4 * gridDim.y;
    const double sig = p.ctl[b].sigma, hr = hp_rho(p, b);
    double m = 0.0;
    for (int row = blockIdx.y * 4 + ty; row < p.Mp; row += 4 * step) {
        double y[4], S[4], lo[4], hi[4], ys[4];     // ys holds the restart anchor y0 in Halpern mode
        bool plain[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = row + u * step;
            plain[u] = r < p.Mp && (r < p.sp_lo || r >= p.sp_hi);
            if (plain[u]) {
                const size_t o = (size_t)r * p.Bp + b;
                y[u] = p.y[o]; S[u] = p.S[o]; lo[u] = p.lo[o]; hi[u] = p.hi[o]; ys[u] = p.halpern ? p.y0[o] : p.ys[o];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = row + u * step;
            if (plain[u]) {
                const size_t o = (size_t)r * p.Bp + b;
                const double v = y[u] + sig * S[u];
                const double w = v / sig;
                const double yn = w > hi[u] ? v - sig * hi[u] : (w < lo[u] ? v - sig * lo[u] : 0.0);
                if (p.halpern) {
                    const double y2 = fma(hr, (2.0 * yn - y[u]) - ys[u], ys[u]);
                    p.y[o] = y2;
                    p.ys[o] = yn;
                    m = fmax(m, fabs(y2));
                } else {
                    p.y[o] = yn;
                    p.ys[o] = ys[u] + yn;
                    m = fmax(m, fabs(yn));
                }
            } else if (r < p.Mp) {
                m = fmax(m, y_update_elem(p, r, b, (long long)r * p.Bp + b));
            }
        }
    }
    if (!p.mxy) return;
    sh[ty][tx] = m;
    __syncthreads();
    if (ty == 0) atomic_max_pos(p.mxy + b, fmax(fmax(sh[0][tx], sh[1][tx]), fmax(sh[2][tx], sh[3][tx])));
}

// The term  w * max_{i in stop rows} (K z)_i  of the objective (fir_ap_cvx.m:163-165: obj*ripple_stop with
// A_U(idx_stop,:) x <= ripple_stop) is handled through its conjugate: the multipliers of those rows live on the
// scaled simplex {y >= 0, sum y = w}, so the dual step is  y+ = Proj_simplex(y + sigma K zbar)  and ripple_stop
// never appears as a variable (it is the multiplier of the simplex constraint; its value is max_i (K z)_i).
// Rows of the block that do not belong to a design (hi = +inf) keep y = 0.  One warp per design; Michelot's
// fixed-point iteration  theta <- (sum_{v > theta} v - w) / #{v > theta}  (monotone, finite).
// Blocks of up to 32*R rows (the stop band of fir_ap_cvx has ~120): one warp per design keeps its rows in registers, so
// the Michelot passes are shuffles only instead of dependent global re-reads (the general kernel below is latency-bound).
template <int R>
__device__ __forceinline__ void simplex_update_reg(const Problem &p, int b, int lane)
{
    const double sig = p.ctl[b].sigma, w = p.sw[b], hr = hp_rho(p, b);
    double v[R], yo[R];
    double sum = 0.0;
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < R; ++u) {
        const int i = lane + 32 * u;
        v[u] = -INFINITY;
        yo[u] = 0.0;
        if (i < p.ns) {
            const size_t o = (size_t)(p.srow0 + i) * p.Bp + b;
            yo[u] = p.y[o];
            if (p.hi[o] == 0.0) v[u] = yo[u] + sig * p.S[o];
        }
    }
#pragma unroll
    for (int u = 0; u < R; ++u)
        if (v[u] > -INFINITY) { sum += v[u]; ++cnt; }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, k); cnt += __shfl_xor_sync(0xffffffffu, cnt, k); }
    double theta = cnt > 0 ? (sum - w) / cnt : 0.0;
    if (cnt > 0 && w > 0.0) {
        int prev = cnt;
        for (int pass = 0; pass < 64; ++pass) {
            double s2 = 0.0;
            int c2 = 0;
#pragma unroll
            for (int u = 0; u < R; ++u)
                if (v[u] > theta) { s2 += v[u]; ++c2; }
#pragma unroll
            for (int k = 16; k > 0; k >>= 1) { s2 += __shfl_xor_sync(0xffffffffu, s2, k); c2 += __shfl_xor_sync(0xffffffffu, c2, k); }
            if (c2 == 0) break;
            theta = (s2 - w) / c2;
            if (c2 == prev) break;
            prev = c2;
        }
    }
    double m = 0.0;
#pragma unroll
    for (int u = 0; u < R; ++u) {
        const int i = lane + 32 * u;
        if (i < p.ns) {
            const size_t o = (size_t)(p.srow0 + i) * p.Bp + b;
            const double yn = (w > 0.0 && v[u] > theta) ? v[u] - theta : 0.0;
            m = fmax(m, store_y(p, o, yo[u], yn, hr));
        }
    }
    if (p.mxy) {
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, k));
        if (lane == 0) atomic_max_pos(p.mxy + b, m);
    }
}
template <int R>
__global__ void __launch_bounds__(256) simplex_update_reg_kernel(Problem p)
{
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= p.Bp) return;
    simplex_update_reg<R>(p, b, threadIdx.x & 31);
}
// Thin batches: the interval rows and the simplex block in ONE launch (they touch disjoint rows): the first `yblocks` CTAs do
// the flat y update, the following ones the register-resident simplex projection (one warp per design).  One launch and one
// dependent latency less on the thin iteration's critical path.
template <int R>
__global__ void __launch_bounds__(256) y_simplex_thin_kernel(Problem p, int yblocks)
{
    if ((int)blockIdx.x < yblocks) {
        const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (idx < (long long)p.Mp * p.Bp) y_update_elem(p, (int)(idx / p.Bp), (int)(idx % p.Bp), idx);
        return;
    }
    const int b = (((int)blockIdx.x - yblocks) * blockDim.x + threadIdx.x) >> 5;
    if (b >= p.Bp) return;
    simplex_update_reg<R>(p, b, threadIdx.x & 31);
}
__global__ void simplex_update_kernel(Problem p)

{
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= p.Bp) return;
    const double sig = p.ctl[b].sigma, w = p.sw[b];
    double sum = 0.0;
    int cnt = 0;
    for (int i = lane; i < p.ns; i += 32) {
        const size_t o = (size_t)(p.srow0 + i) * p.Bp + b;
        if (p.hi[o] == 0.0) {
            const double v = p.y[o] + sig * p.S[o];
            p.y[o] = v;              // stage v in place
            sum += v;
            ++cnt;
        } else {
            p.y[o] = -INFINITY;      // not a member
        }
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, k); cnt += __shfl_xor_sync(0xffffffffu, cnt, k); }
    double theta = cnt > 0 ? (sum - w) / cnt : 0.0;
    if (cnt > 0 && w > 0.0) {
        int prev = cnt;
        for (int pass = 0; pass < 64; ++pass) {
            double s2 = 0.0;
            int c2 = 0;
            for (int i = lane; i < p.ns; i += 32) {
                const double v = p.y[(size_t)(p.srow0 + i) * p.Bp + b];
                if (v > theta) { s2 += v; ++c2; }
            }
#pragma unroll
            for (int k = 16; k > 0; k >>= 1) { s2 += __shfl_xor_sync(0xffffffffu, s2, k); c2 += __shfl_xor_sync(0xffffffffu, c2, k); }
            if (c2 == 0) break;      // cannot happen for w > 0 (the largest v always exceeds theta); guard anyway
            theta = (s2 - w) / c2;
            if (c2 == prev) break;
            prev = c2;
        }
    }
    double m = 0.0;
    const double hr = hp_rho(p, b);
    for (int i = lane; i < p.ns; i += 32) {
        const size_t o = (size_t)(p.srow0 + i) * p.Bp + b;
        const double v = p.y[o];
        const double yn = (w > 0.0 && v > theta) ? v - theta : 0.0;
        const double yo = v > -INFINITY ? v - sig * p.S[o] : 0.0;     // the staged v replaced y: recover it (non-members stay 0)
        m = fmax(m, store_y(p, o, yo, yn, hr));
    }
    if (p.mxy) {
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, k));
        if (lane == 0) atomic_max_pos(p.mxy + b, m);
    }
}

// Group block: the objective term  gw * max_i ||(K z)_pair_i||  (obj*Peak with ||(x_i, x_{n+i})|| <= Peak,
// fir_qp_cvx.m:147,158-160) has as conjugate the indicator of the l1,2 ball {sum_i ||u_i|| <= gw}: the dual step
// is the projection of v = y + sigma K zbar onto that ball (Michelot on the pair norms).  One warp per design.
// GroupSel picks the block: plain (obj*Peak: centres 0) or centred (delta: f(u) = max_i ||u_i - c_i|| has the conjugate
// <c, y> + indicator of the same ball, so the dual step projects  y + sigma (K zbar - c)  instead).
struct GroupSel { int row0, n, centred; const double *w; };
__device__ __forceinline__ GroupSel group_sel(const Problem &p, int which)
{
    return which == 0 ? GroupSel{p.grow0, p.ng, 0, p.gw} : GroupSel{p.grow2, p.ng2, 1, p.gw2};
}
__global__ void group_update_kernel(Problem pp, int which)
{
    const Problem &p0 = pp;
    const GroupSel g = group_sel(p0, which);
    struct { int grow0, ng, Bp; const double *gw; double *y, *S, *ys, *mxy; const Ctl *ctl; const double *lo; } p =
        {g.row0, g.n, p0.Bp, g.w, p0.y, p0.S, p0.ys, p0.mxy, p0.ctl, p0.lo};
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= p.Bp) return;
    const double sig = p.ctl[b].sigma, w = p.gw[b], hr = hp_rho(p0, b);
    double sum = 0.0;
    for (int i = lane; i < p.ng; i += 32) {
        const size_t o = (size_t)(p.grow0 + 2 * i) * p.Bp + b;
        const double c1 = g.centred ? p.lo[o] : 0.0, c2 = g.centred ? p.lo[o + p.Bp] : 0.0;
        const double v1 = p.y[o] + sig * (p.S[o] - c1), v2 = p.y[o + p.Bp] + sig * (p.S[o + p.Bp] - c2);
        p.y[o] = v1; p.y[o + p.Bp] = v2;      // stage v in place
        sum += hypot(v1, v2);
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, k);
    double theta = 0.0;
    if (sum > w) {                           // outside the ball: threshold the norms
        int prev = p.ng;
        theta = (sum - w) / p.ng;
        for (int pass = 0; pass < 64; ++pass) {
            double s2 = 0.0;
            int c2 = 0;
            for (int i = lane; i < p.ng; i += 32) {
                const size_t o = (size_t)(p.grow0 + 2 * i) * p.Bp + b;
                const double r = hypot(p.y[o], p.y[o + p.Bp]);
                if (r > theta) { s2 += r; ++c2; }
            }
#pragma unroll
            for (int k = 16; k > 0; k >>= 1) { s2 += __shfl_xor_sync(0xffffffffu, s2, k); c2 += __shfl_xor_sync(0xffffffffu, c2, k); }
            if (c2 == 0) break;
            theta = (s2 - w) / c2;
            if (c2 == prev) break;
            prev = c2;
        }
    }
    double m = 0.0;
    for (int i = lane; i < p.ng; i += 32) {
        const size_t o = (size_t)(p.grow0 + 2 * i) * p.Bp + b;
        double v1 = p.y[o], v2 = p.y[o + p.Bp];
        if (theta > 0.0) {
            const double r = hypot(v1, v2);
            const double f = r > theta ? (r - theta) / r : 0.0;
            v1 *= f; v2 *= f;
        }
        if (!(w > 0.0)) { v1 = 0.0; v2 = 0.0; }
        const double c1 = g.centred ? p.lo[o] : 0.0, c2 = g.centred ? p.lo[o + p.Bp] : 0.0;
        const double yo1 = p.y[o] - sig * (p.S[o] - c1), yo2 = p.y[o + p.Bp] - sig * (p.S[o + p.Bp] - c2);   // staged v -> old y
        m = fmax(m, fmax(store_y(p0, o, yo1, v1, hr), store_y(p0, o + p.Bp, yo2, v2, hr)));
    }
    if (p.mxy) {
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, k));
        if (lane == 0) atomic_max_pos(p.mxy + b, m);
    }
}

// Group blocks of up to 32*R pairs: the pairs of a design stay in its warp's registers across the Michelot passes
template <int R>
__global__ void __launch_bounds__(256) group_update_reg_kernel(Problem pp, int which)
{
    const Problem &p0 = pp;
    const GroupSel g = group_sel(p0, which);
    struct { int grow0, ng, Bp; const double *gw; double *y, *S, *ys, *mxy; const Ctl *ctl; const double *lo; } p =
        {g.row0, g.n, p0.Bp, g.w, p0.y, p0.S, p0.ys, p0.mxy, p0.ctl, p0.lo};
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= p.Bp) return;
    const double sig = p.ctl[b].sigma, w = p.gw[b], hr = hp_rho(p0, b);
    double v1[R], v2[R], nr[R];
    double sum = 0.0;
#pragma unroll
    for (int u = 0; u < R; ++u) {
        const int i = lane + 32 * u;
        v1[u] = v2[u] = nr[u] = 0.0;
        if (i < p.ng) {
            const size_t o = (size_t)(p.grow0 + 2 * i) * p.Bp + b;
            const double c1 = g.centred ? p.lo[o] : 0.0, c2 = g.centred ? p.lo[o + p.Bp] : 0.0;
            v1[u] = p.y[o] + sig * (p.S[o] - c1);
            v2[u] = p.y[o + p.Bp] + sig * (p.S[o + p.Bp] - c2);
            nr[u] = hypot(v1[u], v2[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < R; ++u) sum += nr[u];
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, k);
    double theta = 0.0;
    if (sum > w) {                           // outside the ball: threshold the norms
        int prev = p.ng;
        theta = (sum - w) / p.ng;
        for (int pass = 0; pass < 64; ++pass) {
            double s2 = 0.0;
            int c2 = 0;
#pragma unroll
            for (int u = 0; u < R; ++u)
                if (lane + 32 * u < p.ng && nr[u] > theta) { s2 += nr[u]; ++c2; }
#pragma unroll
            for (int k = 16; k > 0; k >>= 1) { s2 += __shfl_xor_sync(0xffffffffu, s2, k); c2 += __shfl_xor_sync(0xffffffffu, c2, k); }
            if (c2 == 0) break;
            theta = (s2 - w) / c2;
            if (c2 == prev) break;
            prev = c2;
        }
    }
    double m = 0.0;
#pragma unroll
    for (int u = 0; u < R; ++u) {
        const int i = lane + 32 * u;
        if (i < p.ng) {
            const size_t o = (size_t)(p.grow0 + 2 * i) * p.Bp + b;
            double a1 = v1[u], a2 = v2[u];
            if (theta > 0.0) {
                const double f = nr[u] > theta ? (nr[u] - theta) / nr[u] : 0.0;
                a1 *= f; a2 *= f;
            }
            if (!(w > 0.0)) { a1 = 0.0; a2 = 0.0; }
            const double c1 = g.centred ? p.lo[o] : 0.0, c2 = g.centred ? p.lo[o + p.Bp] : 0.0;
            const double yo1 = v1[u] - sig * (p.S[o] - c1), yo2 = v2[u] - sig * (p.S[o + p.Bp] - c2);   // old y from the staged v
            m = fmax(m, fmax(store_y(p0, o, yo1, a1, hr), store_y(p0, o + p.Bp, yo2, a2, hr)));
        }
    }
    if (p.mxy) {
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, k));
        if (lane == 0) atomic_max_pos(p.mxy + b, m);
    }
}

// reduce split-K slabs: G2 = sum_p G_p
__global__ void reduce_slabs_kernel(Problem p)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)p.Np * p.Bp;
    if (idx >= n) return;
    double g = 0.0;
    for (int s = 0; s < p.P; ++s) g += p.G[(size_t)s * n + idx];
    p.G2[idx] = g;
}

// Row-side metrics of a candidate (cand 0: running average = sums/cnt, cand 1: current iterate):
//   pr = max violation of K z, hs = h*(y), dy2 = ||y - y0||^2.   S holds K*(zs or z).
__global__ void __launch_bounds__(256) row_metrics_kernel(Problem p, int cand)
{
    // blockDim = (64 designs, 4 row lanes): rows strided over the row lanes and blockIdx.y, the row lanes are combined in
    // shared memory before the atomics (one thread per design and 64 row blocks was latency-bound: 130 us per call)
    __shared__ double red[6][4][64];
    const int b = blockIdx.x * 64 + threadIdx.x;
    const bool live = b < p.Bp;
    const int bb = live ? b : 0;
    {
    const int b = bb;
    const double inv = cand == 0 ? 1.0 / fmax(p.ctl[b].cnt, 1.0) : 1.0;
    const double *yv = cand == 0 ? p.ys : p.y;
    double pr = 0.0, hs = 0.0, dy2 = 0.0, tmax = 0.0, gmax = 0.0, gmax2 = 0.0;
    // disk / group pairs are handled by the thread that meets their first (even) row: keep pairs inside one row lane
    for (int i2 = blockIdx.y * 4 + threadIdx.y; i2 < (p.Mp >> 1); i2 += 4 * gridDim.y)
    for (int i = 2 * i2; i < 2 * i2 + 2 && live; ++i) {
        const size_t o = (size_t)i * p.Bp + b;
        const double kz = p.S[o] * inv, y = yv[o] * inv, lo = p.lo[o], hi = p.hi[o];
        if (i >= p.grow2 && i < p.grow2 + 2 * p.ng2) {   // centred group block: h* = <centre, y>, track max ||K z - centre||
            const double d = y - p.y0[o];
            dy2 = fma(d, d, dy2);
            if (((i - p.grow2) & 1) == 0) {
                const size_t o2 = o + p.Bp;
                const double kz2 = p.S[o2] * inv, y2 = yv[o2] * inv, c1 = lo, c2 = p.lo[o2];
                gmax2 = fmax(gmax2, hypot(kz - c1, kz2 - c2));
                hs += c1 * y + c2 * y2;
            }
            continue;
        }
        if (i >= p.drow0 && i < p.drow0 + 2 * p.nd) {  // disk pair (first row does both): violation, support function
            const double d = y - p.y0[o];
            dy2 = fma(d, d, dy2);
            if (((i - p.drow0) & 1) == 0) {
                const size_t o2 = o + p.Bp;
                const double kz2 = p.S[o2] * inv, y2 = yv[o2] * inv, c1 = lo, c2 = p.lo[o2], R = hi;
                pr = fmax(pr, hypot(kz - c1, kz2 - c2) - R);
                hs += c1 * y + c2 * y2 + R * hypot(y, y2);
            }
            continue;
        }
        if (i >= p.grow0 && i < p.grow0 + 2 * p.ng) {  // group block: h* = 0, track max pair norm of K z
            const double d = y - p.y0[o];
            dy2 = fma(d, d, dy2);
            if (((i - p.grow0) & 1) == 0) gmax = fmax(gmax, hypot(kz, p.S[o + p.Bp] * inv));
            continue;
        }
        if (i >= p.srow0 && i < p.srow0 + p.ns) {      // simplex block: no constraint of its own, h* = 0
            if (hi == 0.0) tmax = fmax(tmax, kz);
            const double d = y - p.y0[o];
            dy2 = fma(d, d, dy2);
            continue;
        }
        pr = fmax(pr, fmax(kz - hi, lo - kz));
        if (y > 0.0) hs += hi * y;
        else if (y < 0.0) hs += lo * y;
        const double d = y - p.y0[o];
        dy2 = fma(d, d, dy2);
    }
    const int tx = threadIdx.x, ty = threadIdx.y;
    red[0][ty][tx] = pr; red[1][ty][tx] = tmax; red[2][ty][tx] = gmax; red[3][ty][tx] = gmax2; red[4][ty][tx] = hs; red[5][ty][tx] = dy2;
    }
    __syncthreads();
    if (threadIdx.y != 0 || !live) return;
    const int tx = threadIdx.x;
    auto mx4 = [&](int k) { return fmax(fmax(red[k][0][tx], red[k][1][tx]), fmax(red[k][2][tx], red[k][3][tx])); };
    auto sm4 = [&](int k) { return (red[k][0][tx] + red[k][1][tx]) + (red[k][2][tx] + red[k][3][tx]); };
    double *acc = p.acc + (size_t)cand * NACC * p.Bp;
    atomic_max_pos(acc + A_PR * p.Bp + b, mx4(0));
    atomic_max_pos(acc + A_TMAX * p.Bp + b, mx4(1));
    atomic_max_pos(acc + A_GMAX * p.Bp + b, mx4(2));
    atomic_max_pos(acc + A_GMAX2 * p.Bp + b, mx4(3));
    atomicAdd(acc + A_HS * p.Bp + b, sm4(4));
    atomicAdd(acc + A_DY2 * p.Bp + b, sm4(5));
}

// Column-side metrics: pobj = c^T z, gz = g^T z, dr = ||z - P_X(z - g)||_inf, dz2 = ||z - z0||^2,
// rigx = min_{x in X} g^T x (for the rigorous dual bound).  G2 holds K^T (ys or y).
__global__ void col_metrics_kernel(Problem p, int cand)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.Bp) return;
    const double inv = cand == 0 ? 1.0 / fmax(p.ctl[b].cnt, 1.0) : 1.0;
    const double *zv = cand == 0 ? p.zs : p.z;
    double pobj = 0.0, gz = 0.0, dr = 0.0, dz2 = 0.0, rigx = 0.0, s_zz = 0.0, s_zv = 0.0, s_vv = 0.0;
    double rign = 0.0, s_gg = 0.0, s_rr = 0.0;   // norm-term coordinates: box bound, ||g||^2, radius^2 of a ball around the box
    for (int j = blockIdx.y; j < p.Np; j += gridDim.y) {
        const int pr = p.pair_of[j];
        if (pr >= 0 && p.pair_j[pr] == j) continue;
        const size_t o = (size_t)j * p.Bp + b;
        const double z = zv[o] * inv, c = p.c[o], g = c + p.G2[o] * inv;
        if (pr < 0 && j < p.nn) {   // norm-term coordinate: residual needs ||z - g|| first; collect the three inner products
            s_zz = fma(z, z, s_zz); s_zv = fma(z, z - g, s_zv); s_vv = fma(z - g, z - g, s_vv);
            pobj = fma(c, z, pobj);
            gz = fma(g, z, gz);
            const double d = z - p.z0[o];
            dz2 = fma(d, d, dz2);
            const double bl = p.bl[o], bu = p.bu[o];
            rign += g > 0.0 ? g * bl : (g < 0.0 ? g * bu : 0.0);
            s_gg = fma(g, g, s_gg);
            s_rr += fmax(bl * bl, bu * bu);
        } else if (pr < 0) {
            const double bl = p.bl[o], bu = p.bu[o];
            const double t = fmin(fmax(z - g, bl), bu);
            dr = fmax(dr, fabs(z - t));
            pobj = fma(c, z, pobj);
            gz = fma(g, z, gz);
            const double d = z - p.z0[o];
            dz2 = fma(d, d, dz2);
            rigx += g > 0.0 ? g * bl : (g < 0.0 ? g * bu : 0.0);
        } else {
            const size_t o2 = (size_t)p.pair_j[pr] * p.Bp + b;
            const double z2 = zv[o2] * inv, c2 = p.c[o2], g2 = c2 + p.G2[o2] * inv;
            double a = z - g, e = z2 - g2;
            const double r = hypot(a, e), rho = p.rho[(size_t)pr * p.Bp + b];
            if (r > rho) { const double s = rho / r; a *= s; e *= s; }
            dr = fmax(dr, fmax(fabs(z - a), fabs(z2 - e)));
            pobj = fma(c, z, fma(c2, z2, pobj));
            gz = fma(g, z, fma(g2, z2, gz));
            const double d1 = z - p.z0[o], d2 = z2 - p.z0[o2];
            dz2 = fma(d1, d1, fma(d2, d2, dz2));
            rigx -= rho * hypot(g, g2);
        }
    }
    double *acc = p.acc + (size_t)cand * NACC * p.Bp;
    atomic_max_pos(acc + A_DR * p.Bp + b, dr);
    atomicAdd(acc + A_POBJ * p.Bp + b, pobj);
    atomicAdd(acc + A_GZ * p.Bp + b, gz);
    atomicAdd(acc + A_DZ2 * p.Bp + b, dz2);
    atomicAdd(acc + A_RIGX * p.Bp + b, rigx);
    if (p.nn > 0) {
        atomicAdd(acc + A_ZZ * p.Bp + b, s_zz);
        atomicAdd(acc + A_ZV * p.Bp + b, s_zv);
        atomicAdd(acc + A_VV * p.Bp + b, s_vv);
        atomicAdd(acc + A_RIGN * p.Bp + b, rign);
        atomicAdd(acc + A_GG * p.Bp + b, s_gg);
        atomicAdd(acc + A_RR * p.Bp + b, s_rr);
    }
}

// Advance the per-design counters by one block of iterations (before the metrics use cnt).
__global__ void advance_kernel(Problem p)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) *p.iter_dev += p.check_every;
    if (b >= p.Bp) return;
    p.ctl[b].since += p.check_every;
    p.ctl[b].cnt = p.halpern ? 1.0 : p.ctl[b].cnt + p.check_every;     // Halpern mode: the "sums" hold the last PDHG output
}

// One thread per design: score both candidates, decide convergence / infeasibility / restart.
__global__ void control_kernel(Problem p, int max_iter)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.Bp) return;
    const int iter_now = *p.iter_dev;
    Ctl &c = p.ctl[b];
    c.restart = 0.0;
    if (b >= p.B) { c.status = 1.0; return; }     // padding designs
    double err[2], pr[2], dr[2], po[2], du[2], rig[2], dz[2], dy[2], tm[2];
    for (int k = 0; k < 2; ++k) {
        const double *a = p.acc + (size_t)k * NACC * p.Bp;
        pr[k] = a[A_PR * p.Bp + b];
        dr[k] = a[A_DR * p.Bp + b];
        tm[k] = a[A_TMAX * p.Bp + b];
        po[k] = a[A_POBJ * p.Bp + b] + (p.ns > 0 ? p.sw[b] * tm[k] : 0.0) + (p.ng > 0 ? p.gw[b] * a[A_GMAX * p.Bp + b] : 0.0) +
                (p.ng2 > 0 ? p.gw2[b] * a[A_GMAX2 * p.Bp + b] : 0.0);
        du[k] = -a[A_HS * p.Bp + b] + a[A_GZ * p.Bp + b];
        if (p.nn > 0) {            // norm term: lam*||z|| in both, residual ||z - shrink(z - g)||_2 from the inner products
            const double zz = a[A_ZZ * p.Bp + b], zv = a[A_ZV * p.Bp + b], vv = a[A_VV * p.Bp + b], lam = p.lam[b];
            const double nz = sqrt(fmax(zz, 0.0)), nv = sqrt(fmax(vv, 0.0));
            const double sh = nv > 0.0 ? fmax(0.0, 1.0 - lam / nv) : 1.0;
            po[k] += lam * nz;
            du[k] += lam * nz;
            dr[k] = fmax(a[A_DR * p.Bp + b], sqrt(fmax(zz - 2.0 * sh * zv + sh * sh * vv, 0.0)));
        }
        rig[k] = -a[A_HS * p.Bp + b] + a[A_RIGX * p.Bp + b];
        if (p.nn > 0) {
            // min over the box of lam*||z|| + g.z over the norm-term coordinates: >= the box minimum of g.z alone (lam*||z|| >= 0)
            // and >= -R*max(0, ||g|| - lam) with R the radius of a ball around the box ((lam - ||g||)*||z|| by Cauchy-Schwarz)
            const double gn = sqrt(fmax(a[A_GG * p.Bp + b], 0.0)), R = sqrt(fmax(a[A_RR * p.Bp + b], 0.0));
            const double ball = gn > p.lam[b] ? -R * (gn - p.lam[b]) : 0.0;
            rig[k] += fmax(a[A_RIGN * p.Bp + b], R < DBL_MAX ? ball : -DBL_MAX);
        }
        dz[k] = sqrt(a[A_DZ2 * p.Bp + b]);
        dy[k] = sqrt(a[A_DY2 * p.Bp + b]);
        err[k] = fmax(fmax(pr[k], dr[k]), fabs(po[k] - du[k]));
        if (!(err[k] == err[k])) err[k] = DBL_MAX;   // NaN never wins
    }
    const int k = p.halpern == 1 ? 0 : (err[0] < err[1] ? 0 : 1);   // Halpern mode 1: the PDHG output T(z, y) is the candidate
    c.use_avg = k == 0 ? 1.0 : 0.0;
    if (c.status == 0.0) {
        c.obj = po[k]; c.dual = du[k]; c.pr = pr[k]; c.dr = dr[k]; c.rigorous = p.halpern ? rig[0] : fmax(rig[0], rig[1]);   // Halpern: candidate 1 is the reflected
                                                              // iterate, which can leave the blocks' dual sets -- no valid bound
        c.tmax = p.ng > 0 ? p.acc[(size_t)k * NACC * p.Bp + A_GMAX * p.Bp + b] : tm[k];
        const bool solved = pr[k] <= p.eps_pr && dr[k] <= p.eps_dr &&
                            fabs(po[k] - du[k]) <= p.eps_gap * fmax(fabs(po[k]), 1e-12);
        const bool infeasible = !solved && p.obj_upper && c.rigorous > p.obj_upper[b];
        if (solved) { c.status = 1.0; c.iters = iter_now; }
        else if (infeasible) { c.status = 2.0; c.iters = iter_now; }
        else if (iter_now >= max_iter) { c.status = 3.0; c.iters = iter_now; }
        if (c.status != 0.0) atomicSub(p.active, 1);
        c.restart = 1.0;   // also snapshot the candidate into zbest / ybest (see apply kernel)
    }
    if (c.status != 0.0 && c.restart == 0.0) return;
    // restart rules (PDLP): sufficient decay, necessary decay + no progress, or long since last restart
    const double e = err[k];
    const bool doit = (e <= p.beta_suff * c.last_err || (e <= p.beta_nec * c.last_err && e > c.prev_err) ||
                       c.since >= p.beta_art * (double)iter_now) && c.since >= p.min_since;
    c.prev_err = e;
    if (doit && c.status == 0.0) {
        if (dz[k] > 1e-12 && dy[k] > 1e-12) c.omega = exp(p.omega_theta * log(dy[k] / dz[k]) + (1.0 - p.omega_theta) * log(c.omega));
        c.tau = p.eta / c.omega;
        c.sigma = p.eta * c.omega;
        c.last_err = e;
        c.since = 0.0;
        c.restart = 2.0;   // 2: real restart (iterate, anchor and sums are reset)
    }
}

// restart == 1: snapshot candidate into best.  restart == 2: additionally z,y <- candidate, anchors, sums = 0.
__global__ void apply_kernel(Problem p, int side)   // side 0: z arrays, 1: y arrays
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)(side == 0 ? p.Np : p.Mp) * p.Bp;
    if (idx >= n) return;
    const int b = (int)(idx % p.Bp);
    const Ctl &c = p.ctl[b];
    if (c.restart == 0.0) return;
    double *cur = side == 0 ? p.z : p.y, *sum = side == 0 ? p.zs : p.ys;
    double *anchor = side == 0 ? p.z0 : p.y0, *best = side == 0 ? p.zbest : p.ybest;
    // cnt was already advanced in the control kernel; on a real restart it is reset by reset_cnt_kernel
    const double cand = c.use_avg != 0.0 ? sum[idx] / fmax(c.cnt, 1.0) : cur[idx];
    best[idx] = cand;
    if (c.restart == 2.0) {
        cur[idx] = cand;
        anchor[idx] = cand;
        sum[idx] = 0.0;
    }
}

__global__ void reset_cnt_kernel(Problem p)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.Bp) return;
    if (p.ctl[b].restart == 2.0) p.ctl[b].cnt = 0.0;
}

// Batch compaction: finished designs leave the batch so that the GEMMs shrink with the work that is left.
// dst[r][nb] = src[r][slot[nb]] for nb < keep, pad otherwise  (dst != src)
__global__ void gather_cols_kernel(const double *__restrict__ src, int ld_old, double *__restrict__ dst, int ld_new,
                                   const int *__restrict__ slot, int keep, long long rows, double pad)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * ld_new) return;
    const long long r = idx / ld_new;
    const int nb = (int)(idx % ld_new);
    dst[idx] = nb < keep ? src[r * ld_old + slot[nb]] : pad;
}
// dst[r][orig[b]] = src[r][b] for b < live  (results back to the caller's design order)
__global__ void scatter_cols_kernel(const double *__restrict__ src, int ld_cur, double *__restrict__ dst, int ld0,
                                    const int *__restrict__ orig, int live, long long rows)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * live) return;
    const long long r = idx / live;
    const int b = (int)(idx % live);
    dst[r * ld0 + orig[b]] = src[r * ld_cur + b];
}

// power iteration helpers for ||K||_2
__global__ void scale_first_col_kernel(double *v, int n, int Bp, const double *nrm2)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    v[(size_t)j * Bp] = v[(size_t)j * Bp] / sqrt(nrm2[0]);
}
__global__ void norm2_first_col_kernel(const double *v, int n, int Bp, double *out)
{
    __shared__ double sh[256];
    double s = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) { const double x = v[(size_t)j * Bp]; s = fma(x, x, s); }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k]; __syncthreads(); }
    if (threadIdx.x == 0) out[0] = sh[0];
}

static inline int up(int v, int a) { return (v + a - 1) / a * a; }
// batch widths the products support: 1, 2, 4, 8 (matrix-vector pass) or a multiple of 64 (GEMM tiles)
static inline int batch_width(int b) { return b <= 1 ? 1 : b <= 2 ? 2 : b <= 4 ? 4 : b <= 8 ? 8 : up(b, 64); }

static double g_opt[6] = {0.9, 0.2, 0.8, 0.36, 0.5, 1.0};   // eta factor, beta_suff, beta_nec, beta_art, omega_theta, min restart interval
// product kernels of the iterations (mbrf_pdhg_set_gemm): 2 = tcgen05 int8 split-integer tiles (tc_gemm.cuh), 1 = FP64
// tensor path mma.sync m8n8k4, 0 = SIMT DFMA tiles.  The convergence checks always use an fp64 kernel (1 unless 0).
static int g_gemm_mode = 2;
static int g_halpern = 2;    // reflected Halpern iteration (mbrf_pdhg_set_halpern): 0 off (averaged restarts), 1 candidate = PDHG output,
                             // 2 (default) candidate = the better of PDHG output and Halpern iterate
static int g_tc_digits = 5;  // digit planes / level accumulators of the split-integer product (4..6)

// digit planes, scales and tensor maps of the tcgen05 path (device memory lives in the caller's workspace)
struct TcState {
    bool on = false;
    int nd = 0;
    int8_t *pK = nullptr, *pKT = nullptr, *pX = nullptr;   // [nd][Mp][Np], [nd][Np][Mp], [nd][Bp][max(Mp,Np)]
    double *saK = nullptr, *saKT = nullptr, *sx = nullptr, *mxy = nullptr, *mxz = nullptr;   // mx*: per-design max |iterate|
    CUtensorMap mK, mKT, mXn, mXm;                           // X maps: reduction over Np (K zbar) / over Mp (K^T y)
    int P = 1;                                               // split-K slabs of K^T y on this path
};

static size_t tc_bytes(int Mp, int Np, int Bp)
{
    if (Bp < 64) return 0;
    const size_t mxd = (size_t)(Mp > Np ? Mp : Np);
    return (size_t)tc::MAX_ND * (2 * (size_t)Mp * Np + (size_t)Bp * mxd) + ((size_t)Mp + Np + 3 * (size_t)Bp) * 8 + 4 * 256;
}
using ::mbrf::sm_count;       // per-device cache (common.cu)
static int split_k_tc(int Mp, int Np, int Bp)
{
    const int tiles = (Np / tc::TN) * ((Bp + tc::TM - 1) / tc::TM);
    int P = (2 * 148) / tiles;              // two full waves of one CTA per SM
    if (P > Mp / 256) P = Mp / 256;
    if (P > 32) P = 32;
    if (P < 1) P = 1;
    while (P < 32 && (double)up((Mp + P - 1) / P, tc::KB) * tc::MAX_ND * 16384.0 >= 2147483648.0) ++P;   // int32 level sums
    return P;
}

// mx: per-design max |X| ([Bp]); have_max: already accumulated by the kernel that produced X, otherwise computed here.
// zero_other: the other iterate's max-accumulator, cleared for its next producer.
template <int ND>
static int tc_product_nd(const TcState &t, const double *X, int kdim, int Bp, const CUtensorMap &mA, const CUtensorMap &mX,
                         const double *sa, int R, double *C, int P, long long slab, double *mx, bool have_max,
                         double *zero_other, cudaStream_t st, bool gemm_only = false)
{
    // per device (the attribute is), not per process: a host thread may move to another GPU (mbrf_set_device)
    MBRF_CUDA(cudaFuncSetAttribute(tc::tc_i8_gemm_kernel<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::smem_bytes(ND)));
    if (!have_max) {
        MBRF_CUDA(cudaMemsetAsync(mx, 0, (size_t)Bp * 8, st));
        int gy = kdim / 64;
        if (gy > 64) gy = 64;
        tc::col_absmax_kernel<<<dim3(Bp / 64, gy), 256, 0, st>>>(X, kdim, Bp, mx);
        MBRF_LAUNCH_CHECK();
    }
    if (!gemm_only) {
        tc::slice_cols_kernel<ND><<<dim3(Bp / 64, kdim / 64), 256, 0, st>>>(X, kdim, Bp, mx, t.pX, t.sx, zero_other);
        MBRF_LAUNCH_CHECK();
    }
    tc::Params q;
    q.C = C; q.slab = slab; q.ldc = Bp; q.R = R; q.kdim_total = kdim;
    q.kchunk = up((kdim + P - 1) / P, tc::KB); q.sa = sa; q.sx = t.sx; q.nslab = P;
    const int ntiles = (R / tc::TN) * ((Bp + tc::TM - 1) / tc::TM) * P;
    const int nsm = sm_count();
    tc::tc_i8_gemm_kernel<ND><<<ntiles < nsm ? ntiles : nsm, tc::THREADS, tc::smem_bytes(ND), st>>>(mA, mX, q);
    MBRF_LAUNCH_CHECK();
    return MBRF_OK;
}
template <typename... Args>
static int tc_product(const TcState &t, Args... args)
{
    switch (t.nd) {
    case 4: return tc_product_nd<4>(t, args...);
    case 5: return tc_product_nd<5>(t, args...);
    default: return tc_product_nd<6>(t, args...);
    }
}
template <int ND>
static void tc_slice_rows(const double *A, int ld, int R, int kdim, int8_t *out, double *sa, cudaStream_t st)
{
    tc::slice_rows_kernel<ND><<<(unsigned)(((size_t)R * 32 + 255) / 256), 256, 0, st>>>(A, ld, R, kdim, out, sa);
}

// C[Mp x Bp] = K X.  tcs != null: iteration path (may run on the tcgen05 tiles); null: fp64 kernels (checks, power iteration)
static int gemm_nn(const Problem &p, const double *X, double *C, cudaStream_t st, const TcState *tcs = nullptr, bool have_max = false)
{
    if (p.Bp <= 8) {   // thin batch (1..8 designs): one pass over K^T, bound by streaming the matrix
        launch_thin(p.Bp, st, p.K, p.ldk, p.KT, p.Mp, X, C, p.Mp, p.Np, p.Np, 1, 0LL);
        MBRF_LAUNCH_CHECK();
        return MBRF_OK;
    }
    if (tcs && tcs->on)
        return tc_product(*tcs, X, p.Np, p.Bp, tcs->mK, tcs->mXn, tcs->saK, p.Mp, C, 1, 0LL, tcs->mxz, have_max, tcs->mxy, st);
    dim3 grid(p.Bp / BN, p.Mp / BM, 1);
    if (g_gemm_mode) dgemm_mma_kernel<<<grid, 128, 0, st>>>(p.KT, p.Mp, X, p.Bp, C, p.Np, p.Np, 0);
    else dgemm_kernel<<<grid, GEMM_THREADS, 0, st>>>(p.KT, p.Mp, X, p.Bp, C, p.Np, p.Np, 0);
    MBRF_LAUNCH_CHECK();
    return MBRF_OK;
}
// G[p.P][Np x Bp] = K^T Y (p.P split-K slabs)
static int gemm_tn(const Problem &p, const double *Y, double *G, cudaStream_t st, const TcState *tcs = nullptr, bool have_max = false)
{
    if (p.Bp <= 8) {   // thin batch: one pass over K, the long reduction split into p.P slabs
        const int kc = up((p.Mp + p.P - 1) / p.P, 64);
        launch_thin(p.Bp, st, p.KT, p.Mp, p.K, p.ldk, Y, G, p.Np, p.Mp, kc, p.P, (long long)p.Np * p.Bp);
        MBRF_LAUNCH_CHECK();
        return MBRF_OK;
    }
    if (tcs && tcs->on)
        return tc_product(*tcs, Y, p.Mp, p.Bp, tcs->mKT, tcs->mXm, tcs->saKT, p.Np, G, p.P, (long long)p.Np * p.Bp, tcs->mxy,
                          have_max, tcs->mxz, st);
    const int kchunk = up((p.Mp + p.P - 1) / p.P, BK);
    dim3 grid(p.Bp / BN, p.Np / BM, p.P);
    if (g_gemm_mode) dgemm_mma_kernel<<<grid, 128, 0, st>>>(p.K, p.ldk, Y, p.Bp, G, p.Mp, kchunk, (long long)p.Np * p.Bp);
    else dgemm_kernel<<<grid, GEMM_THREADS, 0, st>>>(p.K, p.ldk, Y, p.Bp, G, p.Mp, kchunk, (long long)p.Np * p.Bp);
    MBRF_LAUNCH_CHECK();
    return MBRF_OK;
}

}  // namespace pdhg
}  // namespace mbrf

using namespace mbrf;
using namespace mbrf::pdhg;

// Warm start of this host thread's next mbrf_pdhg_solve_device call (consumed by it): device iterates [Np x Bp] / [Mp x Bp],
// host primal weights [Bp] in, host primal weights [Bp] out.  Any pointer may be null.
struct WarmHint { const double *z = nullptr, *y = nullptr, *omega = nullptr; double *omega_out = nullptr; };
static thread_local WarmHint t_warm;

extern "C" {

int mbrf_pdhg_warm_start_device(const double *z_init, const double *y_init, const double *omega_init, double *omega_out)
{
    t_warm.z = z_init; t_warm.y = y_init; t_warm.omega = omega_init; t_warm.omega_out = omega_out;
    return MBRF_OK;
}

// choose the product kernels of the iterations: 2 = tcgen05 int8 split-integer tiles (default), 1 = FP64 tensor path
// (mma.sync m8n8k4), 0 = SIMT DFMA
int mbrf_pdhg_set_gemm(int mode)
{
    if (mode < 0 || mode > 2) return MBRF_EINVAL;
    g_gemm_mode = mode;
    return MBRF_OK;
}
// 0: restarts to the better of running average and current iterate; 1: reflected Halpern PDHG, the PDHG output is the candidate;
// 2 (default): reflected Halpern PDHG, the better of PDHG output and Halpern iterate is the candidate
int mbrf_pdhg_set_halpern(int mode)
{
    if (mode < 0 || mode > 2) return MBRF_EINVAL;
    g_halpern = mode;
    return MBRF_OK;
}
// digit planes of the split-integer product: 4, 5 (default: ~1e-11 of |row|max * |column|max per term) or 6 (~1e-13)
int mbrf_pdhg_set_tc_digits(int nd)
{
    if (nd < 4 || nd > tc::MAX_ND) return MBRF_EINVAL;
    g_tc_digits = nd;
    return MBRF_OK;
}

// algorithm constants: which = 0 eta factor (0.9), 1 sufficient-decay beta (0.2), 2 necessary-decay beta (0.8),
// 3 artificial-restart fraction (0.36), 4 primal-weight smoothing (0.5)
int mbrf_pdhg_set_option(int which, double value)
{
    if (which < 0 || which > 5 || !(value > 0.0)) return MBRF_EINVAL;
    g_opt[which] = value;
    return MBRF_OK;
}

/*
 * The split-integer tcgen05 product on its own (tests: parity with fp64; bench.py: kernel duration for the roofline).
 * Device pointers.  C[nslab][R x Bp] = A[R x kdim] * X[kdim x Bp] split over nslab ranges of the reduction (the caller sums
 * the slabs); R, kdim, Bp multiples of 64, nd = 4..6 digit planes.  reps timed repetitions (CUDA events on `stream`):
 * ms_gemm = mean duration of the MMA kernel alone, ms_total = per-design maxima + digit planes of X + MMA kernel.
 */
int mbrf_tc_product_device(const double *A, int R, int kdim, const double *X, int Bp, int nd, int nslab, double *C, int reps,
                           float *ms_gemm, float *ms_total, void *stream)
{
    if (int rc = require_device()) return rc;
    if (!A || !X || !C || R < 64 || kdim < 64 || Bp < 64 || R % 64 || kdim % 64 || Bp % 64 || nd < 4 || nd > tc::MAX_ND || nslab < 1 ||
        reps < 1) {
        set_error("tc_product: bad arguments (R=%d kdim=%d Bp=%d nd=%d nslab=%d)", R, kdim, Bp, nd, nslab);
        return MBRF_EINVAL;
    }
    if ((double)up((kdim + nslab - 1) / nslab, tc::KB) * nd * 16384.0 >= 2147483648.0) {
        set_error("tc_product: reduction range too long for exact int32 level sums, use more slabs");
        return MBRF_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    TcState t;
    t.nd = nd;
    char *buf = nullptr;
    const size_t bA = (size_t)nd * R * kdim, bX = (size_t)nd * Bp * kdim;
    const size_t total = ((bA + 255) / 256 + (bX + 255) / 256) * 256 + ((size_t)R + 3 * (size_t)Bp) * 8 + 1024;
    MBRF_CUDA(cudaMalloc(&buf, total));
    struct Free { char *p; ~Free() { cudaFree(p); } } guard{buf};
    t.pK = (int8_t *)buf;
    t.pX = (int8_t *)(buf + (bA + 255) / 256 * 256);
    t.saK = (double *)((char *)t.pX + (bX + 255) / 256 * 256);
    t.sx = t.saK + R; t.mxy = t.sx + Bp; t.mxz = t.mxy + Bp;
    switch (nd) {
    case 4: tc_slice_rows<4>(A, kdim, R, kdim, t.pK, t.saK, st); break;
    case 5: tc_slice_rows<5>(A, kdim, R, kdim, t.pK, t.saK, st); break;
    default: tc_slice_rows<6>(A, kdim, R, kdim, t.pK, t.saK, st); break;
    }
    MBRF_LAUNCH_CHECK();
    if (!tc::make_map(&t.mK, t.pK, kdim, R, nd, tc::TN) || !tc::make_map(&t.mXn, t.pX, kdim, Bp, nd, tc::TM)) {
        set_error("tc_product: cuTensorMapEncodeTiled failed");
        return MBRF_ECUDA;
    }
    t.on = true;
    cudaEvent_t e0, e1;
    MBRF_CUDA(cudaEventCreate(&e0));
    MBRF_CUDA(cudaEventCreate(&e1));
    auto timed = [&](bool whole, float *out) -> int {
        int rc = tc_product(t, X, kdim, Bp, t.mK, t.mXn, t.saK, R, C, nslab, (long long)R * Bp, t.mxy, false, nullptr, st);   // warm
        if (rc) return rc;
        MBRF_CUDA(cudaEventRecord(e0, st));
        for (int i = 0; i < reps && !rc; ++i)
            rc = whole ? tc_product(t, X, kdim, Bp, t.mK, t.mXn, t.saK, R, C, nslab, (long long)R * Bp, t.mxy, false, nullptr, st)
                       : tc_product(t, X, kdim, Bp, t.mK, t.mXn, t.saK, R, C, nslab, (long long)R * Bp, t.mxy, true, nullptr, st, true);
        if (rc) return rc;
        MBRF_CUDA(cudaEventRecord(e1, st));
        MBRF_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        MBRF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (out) *out = ms / reps;
        return MBRF_OK;
    };
    int rc = timed(true, ms_total);
    if (!rc && ms_gemm) rc = timed(false, ms_gemm);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    MBRF_CUDA(cudaStreamSynchronize(st));
    return rc;
}

// sizes of the padded problem and of the workspace (bytes) for given live sizes
int mbrf_pdhg_padded_sizes(int M, int N, int B, int *Mp, int *Np, int *Bp)
{
    if (M <= 0 || N <= 0 || B <= 0 || !Mp || !Np || !Bp) return MBRF_EINVAL;
    *Mp = up(M, 64);
    *Np = up(N, 64);
    *Bp = batch_width(B);
    return MBRF_OK;
}

static int split_k(int Mp, int Np, int Bp)
{
    if (Bp <= 8) {   // thin path: slabs of >= 256 rows, at most 32 (every slab must own rows: ceil(Mp/P) rounded to 64)
        int P = Mp / 256 < 32 ? Mp / 256 : 32;
        if (P < 1) P = 1;
        while (P > 1 && up((Mp + P - 1) / P, 64) * (P - 1) >= Mp) --P;
        return P;
    }
    // K^T Y has only (Np/64)*(Bp/64) output tiles: split the long reduction so that ~6 CTAs land on each SM
    const int tiles = (Np / BM) * (Bp / BN);
    int P = (6 * 148 + tiles - 1) / tiles;
    if (P < 1) P = 1;
    if (P > Mp / 64) P = Mp / 64;
    if (P > 32) P = 32;
    return P;
}

// doubles needed for the split-K slabs at any batch width the compaction can reach
static size_t slab_doubles(int Mp, int Np, int Bp)
{
    size_t mx = (size_t)split_k(Mp, Np, 8) * Np * 8;
    for (int b = 64; b <= Bp; b += 64) {
        size_t v = (size_t)split_k(Mp, Np, b) * Np * b;
        const size_t vt = (size_t)split_k_tc(Mp, Np, b) * Np * b;
        if (vt > v) v = vt;
        if (v > mx) mx = v;
    }
    return mx;
}

unsigned long long mbrf_pdhg_workspace_bytes(int Mp, int Np, int Bp)
{
    const size_t zn = (size_t)Np * Bp, yn = (size_t)Mp * Bp;
    size_t d = 5 * zn + 4 * yn + yn + slab_doubles(Mp, Np, Bp) + zn + 2 * NACC * (size_t)Bp + 4 * (size_t)Bp;
    return d * 8 + (size_t)Bp * sizeof(Ctl) + 256 + (size_t)Np * 4 + 2 * (size_t)Bp * 4 + 128 + 512 + tc_bytes(Mp, Np, Bp);
}

/*
 * Solve a padded batch on the device.  All pointers are device pointers; per-design arrays are
 * [dim x Bp] with the design index fastest.  info_out: [Bp x 8] doubles
 * (status, iters, obj, dual, pr, dr, rigorous lower bound, omega).
 */
int mbrf_pdhg_solve_device(const double *K, const double *KT, int Mp, int Np, int ldk, double *c, double *lo,
                           double *hi, double *bl, double *bu, const int *pair_i,
                           const int *pair_j, int npairs, double *rho, int Bp, int B,
                           double *obj_upper, const mbrf_pdhg_blocks *blocks, int max_iter, int check_every,
                           double eps_pr, double eps_dr, double eps_gap, double *z_out, double *y_out,
                           double *info_out, void *workspace, void *stream)
{
    const WarmHint warm = t_warm;
    t_warm = WarmHint();
    if (int rc = require_device()) return rc;
    mbrf_pdhg_blocks bk;
    memset(&bk, 0, sizeof bk);
    if (blocks) bk = *blocks;
    const int srow0 = bk.simplex_row0, ns = bk.simplex_rows;
    double *simplex_w = bk.simplex_w;
    if (ns < 0 || (ns > 0 && (!simplex_w || srow0 < 0 || srow0 + ns > Mp)) || bk.disk_pairs < 0 ||
        (bk.disk_pairs > 0 && (bk.disk_row0 < 0 || bk.disk_row0 + 2 * bk.disk_pairs > Mp)) || bk.group_pairs < 0 ||
        (bk.group_pairs > 0 && (!bk.group_w || bk.group_row0 < 0 || bk.group_row0 + 2 * bk.group_pairs > Mp)) ||
        bk.group2_pairs < 0 ||
        (bk.group2_pairs > 0 && (!bk.group2_w || bk.group2_row0 < 0 || bk.group2_row0 + 2 * bk.group2_pairs > Mp)) ||
        bk.norm_coords < 0 || bk.norm_coords > Np || (bk.norm_coords > 0 && (!bk.norm_w || npairs > 0))) {
        set_error("pdhg: bad row/column blocks (simplex %d+%d, disks %d+2*%d, groups %d+2*%d, norm %d)", srow0, ns,
                  bk.disk_row0, bk.disk_pairs, bk.group_row0, bk.group_pairs, bk.norm_coords);
        return MBRF_EINVAL;
    }
    if (Mp % 64 || Np % 64 || Bp != batch_width(Bp) || B < 1 || B > Bp || ldk < Np || ldk % 2 ||
        max_iter < 1 || check_every < 1 || npairs < 0) {
        set_error("pdhg: bad padded sizes Mp=%d Np=%d Bp=%d B=%d ldk=%d", Mp, Np, Bp, B, ldk);
        return MBRF_EINVAL;
    }
    if (!K || !KT || !c || !lo || !hi || !bl || !bu || !z_out || !info_out || !workspace || (npairs && (!pair_i || !pair_j || !rho))) {
        set_error("pdhg: NULL required pointer");
        return MBRF_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Problem p;
    p.Mp = Mp; p.Np = Np; p.Bp = Bp; p.B = B; p.npairs = npairs; p.ldk = ldk; p.K = K; p.KT = KT;
    p.c = c; p.lo = lo; p.hi = hi; p.bl = bl; p.bu = bu; p.rho = rho; p.pair_i = pair_i; p.pair_j = pair_j;
    p.obj_upper = obj_upper; p.P = split_k(Mp, Np, Bp); p.mxy = p.mxz = nullptr;
    {
        const char *e = getenv("MBRF_HALPERN");     // developer override of mbrf_pdhg_set_halpern
        p.halpern = e ? atoi(e) : g_halpern;          // 2: Halpern with the better of {PDHG output, Halpern iterate} as candidate
    }
    p.kk = 0;
    p.srow0 = ns > 0 ? srow0 : 0; p.ns = ns; p.sw = ns > 0 ? simplex_w : nullptr;
    p.drow0 = bk.disk_pairs > 0 ? bk.disk_row0 : 0; p.nd = bk.disk_pairs;
    p.grow0 = bk.group_pairs > 0 ? bk.group_row0 : 0; p.ng = bk.group_pairs; p.gw = bk.group_pairs > 0 ? bk.group_w : nullptr;
    p.grow2 = bk.group2_pairs > 0 ? bk.group2_row0 : 0; p.ng2 = bk.group2_pairs; p.gw2 = bk.group2_pairs > 0 ? bk.group2_w : nullptr;
    p.nn = bk.norm_coords; p.lam = bk.norm_coords > 0 ? bk.norm_w : nullptr;
    p.sp_lo = Mp; p.sp_hi = 0;
    auto own = [&](int r0, int cnt) { if (cnt > 0) { if (r0 < p.sp_lo) p.sp_lo = r0; if (r0 + cnt > p.sp_hi) p.sp_hi = r0 + cnt; } };
    own(p.srow0, p.ns); own(p.drow0, 2 * p.nd); own(p.grow0, 2 * p.ng); own(p.grow2, 2 * p.ng2);
    p.eps_pr = eps_pr; p.eps_dr = eps_dr; p.eps_gap = eps_gap; p.check_every = check_every;
    const size_t zn = (size_t)Np * Bp, yn = (size_t)Mp * Bp;
    double *w = (double *)workspace;
    p.z = w; w += zn; p.zbar = w; w += zn; p.zs = w; w += zn; p.z0 = w; w += zn; p.zbest = w; w += zn;
    p.y = w; w += yn; p.ys = w; w += yn; p.y0 = w; w += yn; p.ybest = w; w += yn;
    p.S = w; w += yn;
    p.G = w; w += slab_doubles(Mp, Np, Bp);
    p.G2 = w; w += zn;
    p.acc = w; w += 2 * NACC * (size_t)Bp;
    p.nrm = w; w += 4 * (size_t)Bp;
    p.ctl = (Ctl *)w; w = (double *)((char *)w + (size_t)Bp * sizeof(Ctl));
    p.active = (int *)w; p.iter_dev = p.active + 1; w += 32;
    int *pair_of = (int *)w;
    p.pair_of = pair_of;
    int *d_slot = pair_of + Np, *d_orig = d_slot + Bp;   // compaction maps

    // ---- tcgen05 path: digit planes of K and K^T (once), buffers for the iterate's planes ----
    TcState tcs;
    auto tc_setup_width = [&]() -> int {   // (re)build what depends on the batch width p.Bp
        if (!tcs.on) return MBRF_OK;
        if (p.Bp < 64) { tcs.on = false; p.mxy = p.mxz = nullptr; p.P = split_k(p.Mp, p.Np, p.Bp); return MBRF_OK; }
        if (!tc::make_map(&tcs.mXn, tcs.pX, p.Np, p.Bp, tcs.nd, tc::TM) || !tc::make_map(&tcs.mXm, tcs.pX, p.Mp, p.Bp, tcs.nd, tc::TM)) {
            set_error("pdhg: cuTensorMapEncodeTiled failed for the iterate planes");
            return MBRF_ECUDA;
        }
        p.P = split_k_tc(p.Mp, p.Np, p.Bp);
        p.mxy = tcs.mxy; p.mxz = tcs.mxz;
        return MBRF_OK;
    };
    if (g_gemm_mode == 2 && Bp >= 64) {
        char *t = (char *)(((uintptr_t)(d_orig + Bp) + 255) & ~(uintptr_t)255);
        const size_t plane = (size_t)Mp * Np, mxd = (size_t)(Mp > Np ? Mp : Np);
        tcs.nd = g_tc_digits;
        tcs.pK = (int8_t *)t; t += tc::MAX_ND * plane;
        tcs.pKT = (int8_t *)t; t += tc::MAX_ND * plane;
        tcs.pX = (int8_t *)t; t += (size_t)tc::MAX_ND * Bp * mxd;
        t = (char *)(((uintptr_t)t + 255) & ~(uintptr_t)255);
        tcs.saK = (double *)t; t += (size_t)Mp * 8;
        tcs.saKT = (double *)t; t += (size_t)Np * 8;
        tcs.sx = (double *)t; t += (size_t)Bp * 8;
        tcs.mxy = (double *)t; t += (size_t)Bp * 8;
        tcs.mxz = (double *)t;
        switch (tcs.nd) {
        case 4: tc_slice_rows<4>(K, ldk, Mp, Np, tcs.pK, tcs.saK, st); tc_slice_rows<4>(KT, Mp, Np, Mp, tcs.pKT, tcs.saKT, st); break;
        case 5: tc_slice_rows<5>(K, ldk, Mp, Np, tcs.pK, tcs.saK, st); tc_slice_rows<5>(KT, Mp, Np, Mp, tcs.pKT, tcs.saKT, st); break;
        default: tc_slice_rows<6>(K, ldk, Mp, Np, tcs.pK, tcs.saK, st); tc_slice_rows<6>(KT, Mp, Np, Mp, tcs.pKT, tcs.saKT, st); break;
        }
        MBRF_LAUNCH_CHECK();
        if (!tc::make_map(&tcs.mK, tcs.pK, Np, Mp, tcs.nd, tc::TN) || !tc::make_map(&tcs.mKT, tcs.pKT, Mp, Np, tcs.nd, tc::TN)) {
            set_error("pdhg: cuTensorMapEncodeTiled failed for the matrix planes");
            return MBRF_ECUDA;
        }
        tcs.on = true;
    }

    // ---- init state ----
    MBRF_CUDA(cudaMemsetAsync(workspace, 0, (char *)pair_of - (char *)workspace, st));
    {
        std::vector<int> po((size_t)Np, -1), pi((size_t)npairs), pj((size_t)npairs);
        if (npairs) {
            MBRF_CUDA(cudaMemcpyAsync(pi.data(), pair_i, npairs * sizeof(int), cudaMemcpyDeviceToHost, st));
            MBRF_CUDA(cudaMemcpyAsync(pj.data(), pair_j, npairs * sizeof(int), cudaMemcpyDeviceToHost, st));
            MBRF_CUDA(cudaStreamSynchronize(st));
            for (int q = 0; q < npairs; ++q) {
                if (pi[q] < 0 || pi[q] >= Np || pj[q] < 0 || pj[q] >= Np || pi[q] == pj[q]) { set_error("pdhg: bad pair %d", q); return MBRF_EINVAL; }
                po[pi[q]] = q; po[pj[q]] = q;
            }
        }
        MBRF_CUDA(cudaMemcpyAsync(pair_of, po.data(), (size_t)Np * sizeof(int), cudaMemcpyHostToDevice, st));
        MBRF_CUDA(cudaStreamSynchronize(st));
    }
    // ---- ||K||_2 by power iteration on design column 0 of scratch vectors (zbar / S are free now) ----
    double knorm2 = 1.0;
    {
        std::vector<double> h((size_t)Np);
        unsigned s = 12345u;
        for (int j = 0; j < Np; ++j) { s = s * 1664525u + 1013904223u; h[j] = ((s >> 8) & 0xffff) / 65536.0 - 0.5; }
        MBRF_CUDA(cudaMemcpy2DAsync(p.zbar, (size_t)Bp * 8, h.data(), 8, 8, Np, cudaMemcpyHostToDevice, st));
        double *nrm = p.acc;  // scratch scalar
        Problem q = p;
        for (int itp = 0; itp < 40; ++itp) {
            norm2_first_col_kernel<<<1, 256, 0, st>>>(p.zbar, Np, Bp, nrm);
            MBRF_LAUNCH_CHECK();
            scale_first_col_kernel<<<(Np + 255) / 256, 256, 0, st>>>(p.zbar, Np, Bp, nrm);
            MBRF_LAUNCH_CHECK();
            if (int rc = gemm_nn(q, p.zbar, p.S, st)) return rc;
            if (int rc = gemm_tn(q, p.S, p.G, st)) return rc;
            reduce_slabs_kernel<<<(unsigned)((zn + 255) / 256), 256, 0, st>>>(p);   // G2 = K^T K v
            MBRF_LAUNCH_CHECK();
            MBRF_CUDA(cudaMemcpyAsync(p.zbar, p.G2, zn * 8, cudaMemcpyDeviceToDevice, st));
        }
        norm2_first_col_kernel<<<1, 256, 0, st>>>(p.zbar, Np, Bp, nrm);
        MBRF_LAUNCH_CHECK();
        double hn = 0.0;
        MBRF_CUDA(cudaMemcpyAsync(&hn, nrm, 8, cudaMemcpyDeviceToHost, st));
        MBRF_CUDA(cudaStreamSynchronize(st));
        knorm2 = sqrt(hn);       // ||K^T K v|| with ||v|| = 1  ->  largest eigenvalue of K^T K
        if (!(knorm2 > 0.0) || !std::isfinite(knorm2)) { set_error("pdhg: ||K|| estimate failed (%g)", knorm2); return MBRF_EINVAL; }
        MBRF_CUDA(cudaMemsetAsync(p.zbar, 0, zn * 8, st));
        MBRF_CUDA(cudaMemsetAsync(p.S, 0, yn * 8, st));
        MBRF_CUDA(cudaMemsetAsync(p.G, 0, slab_doubles(Mp, Np, Bp) * 8, st));
        MBRF_CUDA(cudaMemsetAsync(p.acc, 0, 2 * NACC * (size_t)Bp * 8, st));
    }
    if (int rc = tc_setup_width()) return rc;
    double opt[6];
    memcpy(opt, g_opt, sizeof opt);
    if (const char *e = getenv("MBRF_PDHG_OPTS")) {      // developer override of mbrf_pdhg_set_option: "which:value,which:value"
        while (*e) {
            char *end = nullptr;
            const long which = strtol(e, &end, 10);
            if (end == e || *end != ':') break;
            const double v = strtod(end + 1, &end);
            if (which >= 0 && which < 6 && v > 0.0) opt[which] = v;
            e = *end == ',' ? end + 1 : end;
            if (*end != ',') break;
        }
    }
    p.eta = opt[0] / sqrt(knorm2);
    p.beta_suff = opt[1]; p.beta_nec = opt[2]; p.beta_art = opt[3]; p.omega_theta = opt[4]; p.min_since = opt[5];
    {
        std::vector<Ctl> hc((size_t)Bp);
        for (int b = 0; b < Bp; ++b) {
            Ctl &c0 = hc[b];
            memset(&c0, 0, sizeof(Ctl));
            c0.omega = (warm.omega && b < B && warm.omega[b] > 0.0 && std::isfinite(warm.omega[b])) ? warm.omega[b] : 1.0;
            c0.tau = p.eta / c0.omega; c0.sigma = p.eta * c0.omega;
            c0.last_err = DBL_MAX; c0.prev_err = DBL_MAX;
            c0.status = b < B ? 0.0 : 1.0;
        }
        if (warm.z) {      // iterate and restart anchor; zbar = z makes the first dual step an ordinary one
            MBRF_CUDA(cudaMemcpyAsync(p.z, warm.z, zn * 8, cudaMemcpyDeviceToDevice, st));
            MBRF_CUDA(cudaMemcpyAsync(p.z0, warm.z, zn * 8, cudaMemcpyDeviceToDevice, st));
            MBRF_CUDA(cudaMemcpyAsync(p.zbar, warm.z, zn * 8, cudaMemcpyDeviceToDevice, st));
        }
        if (warm.y) {
            MBRF_CUDA(cudaMemcpyAsync(p.y, warm.y, yn * 8, cudaMemcpyDeviceToDevice, st));
            MBRF_CUDA(cudaMemcpyAsync(p.y0, warm.y, yn * 8, cudaMemcpyDeviceToDevice, st));
        }
        MBRF_CUDA(cudaMemcpyAsync(p.ctl, hc.data(), (size_t)Bp * sizeof(Ctl), cudaMemcpyHostToDevice, st));
        MBRF_CUDA(cudaMemcpyAsync(p.active, &B, sizeof(int), cudaMemcpyHostToDevice, st));
        MBRF_CUDA(cudaStreamSynchronize(st));
    }
    // z starts at P_X(0): one z-update with tau = 0 equivalent -> run it with G = 0, c contributes tau*c; instead
    // project zero directly: boxes containing 0 and disks are satisfied by 0 except boxes with bl > 0 or bu < 0.
    // (handled by the first iteration's projection; z = 0 is only a starting point.)

    const int TPB = 256;
    unsigned gz = (unsigned)((zn + TPB - 1) / TPB), gy = (unsigned)((yn + TPB - 1) / TPB);
    dim3 gm((Bp + 63) / 64, 64);

    // first: first iteration of a block -- y may have been replaced by a restart or a compaction since the last y-update,
    // so its per-design max is recomputed; afterwards the update kernels keep the maxima current.
    auto iteration = [&](int i_in_block) -> int {
        const bool first = i_in_block == 0;
        const bool wide = p.Bp >= 64;
        p.kk = i_in_block;
        if (int rc = gemm_tn(p, p.y, p.G, st, &tcs, !first)) return rc;
        if (p.nn > 0) {
            if (p.Bp <= 8) {
                z_hat_thin_kernel<<<p.Bp, 256, 0, st>>>(p);
            } else {
                MBRF_CUDA(cudaMemsetAsync(p.nrm, 0, (size_t)p.Bp * 8, st));
                z_hat_kernel<<<dim3((p.Bp + 63) / 64, 16), 64, 0, st>>>(p);
            }
            MBRF_LAUNCH_CHECK();
            z_shrink_kernel<<<gz, TPB, 0, st>>>(p);
            MBRF_LAUNCH_CHECK();
        } else if (wide) {
            z_update_wide_kernel<<<dim3(p.Bp / 64, p.Np / 4 < 128 ? p.Np / 4 : 128), 256, 0, st>>>(p);
            MBRF_LAUNCH_CHECK();
        } else if (p.Bp <= 8) {
            z_update_thin_kernel<<<(unsigned)(((size_t)p.Np * p.Bp * 32 + 255) / 256), 256, 0, st>>>(p);
            MBRF_LAUNCH_CHECK();
        } else {
            z_update_kernel<<<gz, TPB, 0, st>>>(p);
            MBRF_LAUNCH_CHECK();
        }
        if (int rc = gemm_nn(p, p.zbar, p.S, st, &tcs, p.nn == 0 && wide)) return rc;
        const bool fused_thin = !wide && p.ns > 0 && p.ns <= 256;
        if (wide) y_update_wide_kernel<<<dim3(p.Bp / 64, 128), 256, 0, st>>>(p);
        else if (fused_thin && p.ns <= 128) y_simplex_thin_kernel<4><<<gy + (p.Bp * 32 + 255) / 256, 256, 0, st>>>(p, (int)gy);
        else if (fused_thin) y_simplex_thin_kernel<8><<<gy + (p.Bp * 32 + 255) / 256, 256, 0, st>>>(p, (int)gy);
        else y_update_kernel<<<gy, TPB, 0, st>>>(p);
        MBRF_LAUNCH_CHECK();
        if (p.ns > 0 && !fused_thin) {
            if (p.ns <= 128) simplex_update_reg_kernel<4><<<(p.Bp * 32 + 255) / 256, 256, 0, st>>>(p);
            else if (p.ns <= 256) simplex_update_reg_kernel<8><<<(p.Bp * 32 + 255) / 256, 256, 0, st>>>(p);
            else simplex_update_kernel<<<(p.Bp * 32 + 255) / 256, 256, 0, st>>>(p);
            MBRF_LAUNCH_CHECK();
        }
        if (p.ng > 0) {
            if (p.ng <= 256) group_update_reg_kernel<8><<<(p.Bp * 32 + 255) / 256, 256, 0, st>>>(p, 0);
            else group_update_kernel<<<(p.Bp * 32 + 255) / 256, 256, 0, st>>>(p, 0);
            MBRF_LAUNCH_CHECK();
        }
        if (p.ng2 > 0) {
            if (p.ng2 <= 256) group_update_reg_kernel<8><<<(p.Bp * 32 + 255) / 256, 256, 0, st>>>(p, 1);
            else group_update_kernel<<<(p.Bp * 32 + 255) / 256, 256, 0, st>>>(p, 1);
            MBRF_LAUNCH_CHECK();
        }
        return MBRF_OK;
    };
    auto check = [&]() -> int {
        MBRF_CUDA(cudaMemsetAsync(p.acc, 0, 2 * NACC * (size_t)p.Bp * 8, st));
        advance_kernel<<<(p.Bp + 63) / 64, 64, 0, st>>>(p);
        MBRF_LAUNCH_CHECK();
        Problem pc = p;                      // the checks run on the fp64 kernels with their own split-K
        pc.P = split_k(p.Mp, p.Np, p.Bp);
        for (int cand = 0; cand < 2; ++cand) {
            if (int rc = gemm_nn(pc, cand == 0 ? p.zs : p.z, p.S, st)) return rc;
            row_metrics_kernel<<<gm, dim3(64, 4), 0, st>>>(p, cand);
            MBRF_LAUNCH_CHECK();
            if (int rc = gemm_tn(pc, cand == 0 ? p.ys : p.y, p.G, st)) return rc;
            reduce_slabs_kernel<<<gz, TPB, 0, st>>>(pc);
            MBRF_LAUNCH_CHECK();
            col_metrics_kernel<<<gm, 64, 0, st>>>(p, cand);
            MBRF_LAUNCH_CHECK();
        }
        control_kernel<<<(p.Bp + 63) / 64, 64, 0, st>>>(p, max_iter);
        MBRF_LAUNCH_CHECK();
        apply_kernel<<<gz, TPB, 0, st>>>(p, 0);
        MBRF_LAUNCH_CHECK();
        apply_kernel<<<gy, TPB, 0, st>>>(p, 1);
        MBRF_LAUNCH_CHECK();
        reset_cnt_kernel<<<(p.Bp + 63) / 64, 64, 0, st>>>(p);
        MBRF_LAUNCH_CHECK();
        return MBRF_OK;
    };

    // Capture one block of `check_every` iterations as a CUDA graph: the inner loop is launch-bound for
    // small batches (4 kernels of a few microseconds each per iteration).  Re-captured after a compaction.
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool use_graph = true;
    auto capture = [&]() {
        if (exec) { cudaGraphExecDestroy(exec); exec = nullptr; }
        if (graph) { cudaGraphDestroy(graph); graph = nullptr; }
        use_graph = true;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            int rc = MBRF_OK;
            for (int i = 0; i < check_every && rc == MBRF_OK; ++i) rc = iteration(i);
            if (rc == MBRF_OK) rc = check();   // the convergence check is part of the graph: three API calls per block on the host
            cudaError_t e = cudaStreamEndCapture(st, &graph);
            if (rc != MBRF_OK || e != cudaSuccess || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) use_graph = false;
        } else {
            use_graph = false;
        }
        if (!use_graph) cudaGetLastError();
    };
    capture();

    // results are kept in the caller's design order: z_out / y_out are [dim x Bp0], final control blocks on the host
    const int Bp0 = Bp;
    std::vector<int> orig((size_t)Bp0);
    for (int b = 0; b < Bp0; ++b) orig[b] = b;
    std::vector<Ctl> final_ctl((size_t)Bp0);
    for (auto &f : final_ctl) { memset(&f, 0, sizeof(Ctl)); f.status = 1.0; }
    MBRF_CUDA(cudaMemcpyAsync(d_orig, orig.data(), (size_t)Bp0 * 4, cudaMemcpyHostToDevice, st));
    MBRF_CUDA(cudaMemsetAsync(z_out, 0, zn * 8, st));
    if (y_out) MBRF_CUDA(cudaMemsetAsync(y_out, 0, yn * 8, st));

    std::vector<Ctl> hc;
    auto flush_results = [&]() -> int {   // snapshot every live slot into the caller-ordered outputs
        hc.resize((size_t)p.Bp);
        MBRF_CUDA(cudaMemcpyAsync(hc.data(), p.ctl, (size_t)p.Bp * sizeof(Ctl), cudaMemcpyDeviceToHost, st));
        const long long nz = (long long)p.Np * p.B, ny = (long long)p.Mp * p.B;
        scatter_cols_kernel<<<(unsigned)((nz + 255) / 256), 256, 0, st>>>(p.zbest, p.Bp, z_out, Bp0, d_orig, p.B, p.Np);
        MBRF_LAUNCH_CHECK();
        if (y_out) {
            scatter_cols_kernel<<<(unsigned)((ny + 255) / 256), 256, 0, st>>>(p.ybest, p.Bp, y_out, Bp0, d_orig, p.B, p.Mp);
            MBRF_LAUNCH_CHECK();
        }
        MBRF_CUDA(cudaStreamSynchronize(st));
        for (int b = 0; b < p.B; ++b) final_ctl[orig[b]] = hc[b];
        return MBRF_OK;
    };
    auto compact = [&]() -> int {
        if (int rc = flush_results()) return rc;
        std::vector<int> slot;
        for (int b = 0; b < p.B; ++b)
            if (hc[b].status == 0.0) slot.push_back(b);
        const int keep = (int)slot.size();
        const int nBp = batch_width(keep > 0 ? keep : 1);
        if (keep == 0 || nBp >= p.Bp) return MBRF_OK;
        MBRF_CUDA(cudaMemcpyAsync(d_slot, slot.data(), (size_t)keep * 4, cudaMemcpyHostToDevice, st));
        auto regather = [&](double *arr, long long rows, double *tmp, double pad) -> int {
            if (!arr || rows == 0) return MBRF_OK;
            const long long n = rows * nBp;
            gather_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(arr, p.Bp, tmp, nBp, d_slot, keep, rows, pad);
            MBRF_LAUNCH_CHECK();
            MBRF_CUDA(cudaMemcpyAsync(arr, tmp, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
            return MBRF_OK;
        };
        double *nside[] = {p.z, p.zs, p.z0, p.zbest, p.c, p.bl, p.bu};
        for (double *a : nside)
            if (int rc = regather(a, p.Np, p.G2, 0.0)) return rc;
        if (int rc = regather(p.rho, p.npairs, p.G2, 0.0)) return rc;
        if (int rc = regather(p.obj_upper, 1, p.G2, INFINITY)) return rc;
        if (int rc = regather(p.sw, 1, p.G2, 0.0)) return rc;
        if (int rc = regather(p.gw, 1, p.G2, 0.0)) return rc;
        if (int rc = regather(p.gw2, 1, p.G2, 0.0)) return rc;
        if (int rc = regather(p.lam, 1, p.G2, 0.0)) return rc;
        double *mside[] = {p.y, p.ys, p.y0, p.ybest};
        for (double *a : mside)
            if (int rc = regather(a, p.Mp, p.S, 0.0)) return rc;
        if (int rc = regather(p.lo, p.Mp, p.S, -INFINITY)) return rc;
        if (int rc = regather(p.hi, p.Mp, p.S, INFINITY)) return rc;
        std::vector<Ctl> nc((size_t)nBp);
        std::vector<int> norig((size_t)Bp0, 0);
        for (int b = 0; b < nBp; ++b) {
            if (b < keep) { nc[b] = hc[slot[b]]; norig[b] = orig[slot[b]]; }
            else { memset(&nc[b], 0, sizeof(Ctl)); nc[b].omega = 1.0; nc[b].tau = nc[b].sigma = p.eta; nc[b].status = 1.0; }
        }
        orig = norig;
        MBRF_CUDA(cudaMemcpyAsync(p.ctl, nc.data(), (size_t)nBp * sizeof(Ctl), cudaMemcpyHostToDevice, st));
        MBRF_CUDA(cudaMemcpyAsync(d_orig, orig.data(), (size_t)Bp0 * 4, cudaMemcpyHostToDevice, st));
        MBRF_CUDA(cudaStreamSynchronize(st));
        p.B = keep;
        p.Bp = nBp;
        p.P = split_k(p.Mp, p.Np, nBp);
        if (int rc = tc_setup_width()) return rc;
        gz = (unsigned)(((size_t)p.Np * nBp + TPB - 1) / TPB);
        gy = (unsigned)(((size_t)p.Mp * nBp + TPB - 1) / TPB);
        gm = dim3((nBp + 63) / 64, 64);
        capture();
        return MBRF_OK;
    };

    const bool trace = getenv("MBRF_PDHG_TRACE") != nullptr;
    int active = B, rcode = MBRF_OK;
    for (int it = 0; it < max_iter && active > 0 && rcode == MBRF_OK;) {
        if (use_graph) {
            if (cudaGraphLaunch(exec, st) != cudaSuccess) { set_error("pdhg: graph launch failed"); rcode = MBRF_ECUDA; break; }
            g_launches.fetch_add((4ull + (tcs.on ? 2 : 0) + (p.ns > 0 && !(p.Bp <= 8 && p.ns <= 256)) + (p.ng > 0) + (p.ng2 > 0) + (p.nn > 0)) * check_every, std::memory_order_relaxed);
        } else {
            for (int i = 0; i < check_every && rcode == MBRF_OK; ++i) rcode = iteration(i);
            if (rcode == MBRF_OK) rcode = check();
        }
        it += check_every;
        if (rcode != MBRF_OK) break;
        if (cudaMemcpyAsync(&active, p.active, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) { set_error("pdhg: status readback failed: %s", cudaGetErrorString(cudaGetLastError())); rcode = MBRF_ECUDA; break; }
        if (trace) {   // developer trace (MBRF_PDHG_TRACE=1): slot 0 after every check
            Ctl t0;
            if (cudaMemcpy(&t0, p.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost) == cudaSuccess)
                fprintf(stderr, "pdhg it %6d active %d width %d st %.0f obj %.8f dual %.8f pr %.2e dr %.2e omega %.3e restart %.0f avg %.0f since %.0f lasterr %.2e\n",
                        it, active, p.Bp, t0.status, t0.obj, t0.dual, t0.pr, t0.dr, t0.omega, t0.restart, t0.use_avg, t0.since, t0.last_err);
        }
        // finished designs leave the batch once they would free at least one 64-design column block
        if (active > 0 && batch_width(active) < p.Bp && it < max_iter) rcode = compact();
    }
    if (rcode == MBRF_OK) rcode = flush_results();
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    if (rcode != MBRF_OK) return rcode;

    {
        std::vector<double> info((size_t)Bp0 * 8);
        for (int b = 0; b < Bp0; ++b) {
            const Ctl &c0 = final_ctl[b];
            double *o = &info[(size_t)b * 8];
            o[0] = c0.status == 0.0 ? 3.0 : c0.status; o[1] = c0.iters; o[2] = c0.obj; o[3] = c0.dual;
            o[4] = c0.pr; o[5] = c0.dr; o[6] = c0.rigorous; o[7] = c0.tmax;
            if (warm.omega_out) warm.omega_out[b] = c0.omega;
        }
        MBRF_CUDA(cudaMemcpyAsync(info_out, info.data(), info.size() * 8, cudaMemcpyHostToDevice, st));
        MBRF_CUDA(cudaStreamSynchronize(st));
    }
    return MBRF_OK;
}

}  // extern "C"
