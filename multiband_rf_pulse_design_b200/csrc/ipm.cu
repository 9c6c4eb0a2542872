// Batched primal-dual interior-point solver for the convex FIR design step (fir_ap_cvx.m:160-169, ss/fir_linprog.m:246-252)
// on B200 -- the second-order companion of the first-order solver in pdhg.cu.
//
// Why it exists.  The reference hands these problems to CVX -> SeDuMi/SDPT3, i.e. to an interior-point method, and calls them
// with stop-band weights obj = 1e4 / 1e5 (dzrf_mb.m:167-170, fir_qp.m:47).  There the objective is lexicographic in all but
// name (3e-8 absolute accuracy of the ripple is needed for 1e-4 relative in the objective) and restarted PDHG does not get
// there in 60 000 iterations (round-1 finding).  A Newton-type method does, in 35-50 iterations, IF the Newton
// system is affordable -- and for frequency-sampled Fourier matrices it is:
//
//     K[i][j] = amp_j * {1, cos, sin}(w_i * q_j * u)        (q_j integer lags, u = 1 or 1/2)
//     (K' diag(d) K)[j][l] = amp_j amp_l / 2 * ( +-C[|q_j - q_l|] +- C[q_j + q_l]   or   S[..] )
//     C[p] = sum_i d_i cos(w_i p u),  S[p] = sum_i d_i sin(w_i p u)                   ("moments" of the row weights)
//
// so the normal matrix of a design costs ONE product of the [2(L+1) x M] trigonometric table with the weight vector
// (the same shape as K' y) plus O(N^2) index arithmetic, instead of the 2 M N^2 flops of a dense K' D K.  Batched over designs
// the moment step is a dense contraction table[2(L+1) x M] * D[M x B]; the factorisation is a batched Cholesky of
// (N+1) x (N+1) matrices, one CTA per design.
//
// Algorithm (per design, all designs of a batch in lockstep): homogeneous self-dual embedding, Nesterov-Todd scaling,
// Mehrotra predictor-corrector -- the method of CVXOPT's conelp / ECOS, restated for this structure:
//     minimise c'v  s.t.  G v + s = h,  s in K = R+^nl x Q3^npairs,      v = (z, t)
//         rows i:      a_i'z - e_i t <= hi_i,   -a_i'z <= -lo_i           (e_i = 1 on the rows of the stop block)
//         columns j:   z_j <= bu_j,  -z_j <= -bl_j                          (finite bounds only)
//         pairs k:     (rho_k, z_pi, z_pj) in Q3                            (||(x_i, x_{n+i-1})|| <= (n-i+1) Peak, fir_ap_cvx.m:166-168)
//         objective:   c'z + ct t                                            (x(1) + obj*ripple_stop, fir_ap_cvx.m:163)
// Every iteration: residuals (2 products with K), NT scaling, moments -> normal matrix -> Cholesky, three solves with one
// refinement step each on the un-reduced system (12 products with K), two step-length reductions.
//
// Precision.  State, residuals and products are fp64.  The normal matrix mixes weights z/s that span more than 1e16 in the
// last iterations; formed in fp64 the curvature of the inactive rows is rounded away and the iteration converges to a point
// 3e-4 off the optimum (measured, tools/ipm_proto.py, N = 256, obj = 1e4).  Moments, normal matrix, Cholesky factor and the
// triangular solves therefore run in double-double (dd.cuh) once the complementarity measure is small (precision = auto) --
// the trigonometric table is generated in double-double for that reason, and K is its rounding, so both describe one matrix.
#include "common.h"
#include "dd.cuh"
#include "dgemm.cuh"
#include "fir_assemble.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>
#include <vector>

namespace mbrf {
namespace ipm {

static inline int up(int v, int a) { return (v + a - 1) / a * a; }
static inline int batch_width(int b) { return b <= 1 ? 1 : b <= 2 ? 2 : b <= 4 ? 4 : b <= 8 ? 8 : up(b, 64); }

constexpr int NRHS = 3;      // 0: (-c, h) system, 1: affine (predictor), 2: combined (corrector)
constexpr int PANEL = 32;    // block-column width of the factorisation
constexpr int TILE = 64;     // trailing-update tile

// accumulator slots, [NACC][Bp] doubles, zeroed at the start of every iteration
enum {
    A_RZ2 = 0, A_RX2, A_SZ, A_HZ, A_CX, A_GTZ2, A_GXS2, A_T1,
    A_TQ0, A_TQ1, A_TQ2, A_TUZ0, A_TUZ1, A_TUZ2, A_HZ0, A_HZ1, A_HZ2, A_CX0, A_CX1, A_CX2,
    A_RATIO_A, A_RATIO_F, A_VIOL, A_NCON, A_TMAX, NACC
};

struct Ctl {                 // per-design scalars (device)
    double tau, kap, t;      // t: the stop-block epigraph variable (ripple_stop)
    double mu, sigma, alpha, rt, den;
    double dtau_a, dkap_a, dtau, dkap, dt_a, dt;
    double pcost, dcost, pres, dres, gap, relgap, hz, nrm_h, nrm_c;
    double status;           // 0 running, 1 optimal, 2 primal infeasible (certificate), 3 iteration limit / numerical failure, 4 unbounded
    double iters, chol_fail;
    double merit_best, tau_best, t_best, pcost_best, dcost_best, dres_best, improved, use_best;   // best iterate seen (see SC_PRE)
    double merit_ref, iter_ref;      // stall detection: the last iteration at which the merit had halved
};

struct P {                   // everything the kernels need, passed by value
    int M, Mp, N, Np, NV, NVp, B, Bp, npairs, srow0, ns, L, qmax, nsplit;
    const double *K, *KT;
    const dd *TC, *TS;                              // [L+1][Mp]
    const int *col_q, *col_type;
    const double *col_amp;
    const int *pair_i, *pair_j, *pair_of;           // pair_of[j] = 2*k + side or -1
    const double *c, *lo, *hi, *bl, *bu, *rho, *ct; // [Np x Bp], [Mp x Bp] x2, [Np x Bp] x2, [npairs x Bp], [Bp]
    double *x;                                      // [Np x Bp]
    double *su, *zu, *sl, *zl;                      // [Mp x Bp]
    double *sbu, *zbu, *sbl, *zbl;                  // [Np x Bp]
    double *sd, *zd;                                // [3 npairs x Bp]
    // row work arrays [Mp x Bp]
    double *AX, *YZ, *RZU, *RZL, *D, *DS, *CU, *CL, *DSU, *DZU, *DSL, *DZL, *YR;
    double *QU[NRHS], *QL[NRHS], *Y[NRHS], *GUX[NRHS];
    // column work arrays [Np x Bp]
    double *KTY, *RX, *RZBU, *RZBL, *DIAG, *CBU, *CBL, *DSBU, *DZBU, *DSBL, *DZBL, *DXV;
    double *QBU[NRHS], *QBL[NRHS], *RHS[NRHS], *UX[NRHS], *KTQ;
    double *UT[NRHS], *RHST[NRHS];                  // [Bp]: t components of solutions / right-hand sides
    // disks [dim x npairs x Bp]
    double *RZD, *WD /* eta, w0, w1, w2 */, *LAMD, *QD[NRHS], *CD, *DSD, *DZD, *HB /* 3: m11 m12 m22 */;
    double *XB;                                     // [Np x Bp] best iterate
    double *acc;                                    // [NACC][Bp]
    Ctl *ctl;
    int *active;                                    // designs still running
    double feastol, abstol, reltol;
    int max_iter;
};

__device__ __forceinline__ void atomic_max_pos(double *addr, double v)
{
    if (!(v > 0.0)) return;
    atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ bool fin(double v) { return fabs(v) < 1e300; }

// ------------------------------------------------------------------------------------------------------------------
// trigonometric table in double-double: TC[p][i] = cos(w_i p u), TS[p][i] = sin(w_i p u), p = 0..L, by rotation
// ------------------------------------------------------------------------------------------------------------------
__global__ void table_kernel(const double *__restrict__ w, int M, int Mp, double u, int L, dd *__restrict__ TC, dd *__restrict__ TS)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Mp) return;
    if (i >= M) {
        for (int p = 0; p <= L; ++p) { TC[(size_t)p * Mp + i] = dd{0.0, 0.0}; TS[(size_t)p * Mp + i] = dd{0.0, 0.0}; }
        return;
    }
    dd s1, c1;
    dd_sincos(dd{w[i] * u, 0.0}, &s1, &c1);          // u is 1 or 1/2: the product is exact
    dd c = dd{1.0, 0.0}, s = dd{0.0, 0.0};
    for (int p = 0; p <= L; ++p) {
        TC[(size_t)p * Mp + i] = c;
        TS[(size_t)p * Mp + i] = s;
        const dd cn = dd_sub(dd_mul(c, c1), dd_mul(s, s1));
        s = dd_add(dd_mul(s, c1), dd_mul(c, s1));
        c = cn;
    }
}

// K [Mp x Np] and KT [Np x Mp] as the rounding of the table: K[i][j] = amp_j * {1, cos, sin}
__global__ void matrix_kernel(P p, double *__restrict__ K, double *__restrict__ KT)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)p.Mp * p.Np) return;
    const int j = (int)(idx / p.Mp), i = (int)(idx % p.Mp);
    double v = 0.0;
    if (i < p.M && j < p.N) {
        const int t = p.col_type[j];
        if (t == 0) v = p.col_amp[j];
        else if (t == 1) v = p.col_amp[j] * p.TC[(size_t)p.col_q[j] * p.Mp + i].hi;
        else if (t == 2) v = p.col_amp[j] * p.TS[(size_t)p.col_q[j] * p.Mp + i].hi;
    }
    KT[(size_t)j * p.Mp + i] = v;
    K[(size_t)i * p.Np + j] = v;
}

__global__ void sum_slabs_kernel(const double *__restrict__ G, int nslab, long long slab, double *__restrict__ out, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int q = 0; q < nslab; ++q) s += G[(size_t)q * slab + i];
    out[i] = s;
}

// ------------------------------------------------------------------------------------------------------------------
// 3-dimensional second-order cone helpers (one disk ||(z_pi, z_pj)|| <= rho is the cone (rho, z_pi, z_pj) in Q3)
// ------------------------------------------------------------------------------------------------------------------
struct V3 { double a, b, c; };
__device__ __forceinline__ V3 soc_prod(V3 u, V3 v) { return V3{u.a * v.a + u.b * v.b + u.c * v.c, u.a * v.b + v.a * u.b, u.a * v.c + v.a * u.c}; }
__device__ __forceinline__ V3 soc_div(V3 u, V3 d)      // solve u o x = d
{
    const double det = u.a * u.a - u.b * u.b - u.c * u.c;
    const double x0 = (u.a * d.a - u.b * d.b - u.c * d.c) / det;
    return V3{x0, (d.b - x0 * u.b) / u.a, (d.c - x0 * u.c) / u.a};
}
// W v (inv = false) or W^-1 v for the NT scaling (eta, wbar)
__device__ __forceinline__ V3 soc_W(double eta, V3 w, V3 v, bool inv)
{
    const double w1 = inv ? -w.b : w.b, w2 = inv ? -w.c : w.c;
    const double dot = w1 * v.b + w2 * v.c;
    const double u0 = w.a * v.a + dot;
    const double f = v.a + dot / (1.0 + w.a);
    const double sc = inv ? 1.0 / eta : eta;
    return V3{sc * u0, sc * (v.b + w1 * f), sc * (v.c + w2 * f)};
}
__device__ __forceinline__ V3 soc_W2inv(double eta, V3 w, V3 v) { return soc_W(eta, w, soc_W(eta, w, v, true), true); }
// 1 / (largest alpha with u + alpha du in Q3), 0 if unbounded
__device__ __forceinline__ double soc_ratio(V3 u, V3 du)
{
    double best = INFINITY;
    if (du.a < 0.0) best = -u.a / du.a;
    const double a = du.a * du.a - du.b * du.b - du.c * du.c;
    const double b = 2.0 * (u.a * du.a - u.b * du.b - u.c * du.c);
    const double c = u.a * u.a - u.b * u.b - u.c * u.c;
    if (fabs(a) < 1e-300) {
        if (b < 0.0) best = fmin(best, -c / b);
    } else {
        const double disc = b * b - 4.0 * a * c;
        if (disc >= 0.0) {
            const double sq = sqrt(disc);
            const double q = -0.5 * (b + (b >= 0.0 ? sq : -sq));
            const double r1 = q / a;
            if (r1 > 0.0) best = fmin(best, r1);
            if (q != 0.0) { const double r2 = c / q; if (r2 > 0.0) best = fmin(best, r2); }
        }
    }
    return best == INFINITY ? 0.0 : 1.0 / best;
}

// ------------------------------------------------------------------------------------------------------------------
// phases of the row / column / disk kernels
// ------------------------------------------------------------------------------------------------------------------
enum { PH_YZ = 0, PH_PRE, PH_REFINE, PH_DOTS, PH_AFFINE, PH_COMB, PH_FINAL, PH_APPLY, PH_METRICS };

// rows: thread = design b (fastest), blockIdx.y = chunk of ROWS_PER_BLOCK rows
constexpr int ROWS_PER_BLOCK = 32;

__global__ void rows_kernel(P p, int phase, int r)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.Bp) return;
    const Ctl ct = p.ctl[b];
    const bool live = b < p.B && ct.status == 0.0;
    if (!live && phase != PH_METRICS) {
        if (phase == PH_YZ || phase == PH_PRE || phase == PH_REFINE || phase == PH_COMB) {
            // keep the arrays the products read finite
            const int i0 = blockIdx.y * ROWS_PER_BLOCK;
            for (int i = i0; i < i0 + ROWS_PER_BLOCK && i < p.Mp; ++i) {
                const size_t o = (size_t)i * p.Bp + b;
                if (phase == PH_YZ) p.YZ[o] = 0.0;
                else if (phase == PH_PRE) { p.Y[0][o] = 0.0; p.Y[1][o] = 0.0; p.D[o] = 0.0; p.DS[o] = 0.0; }
                else if (phase == PH_REFINE) p.YR[o] = 0.0;
                else p.Y[2][o] = 0.0;
            }
        }
        return;
    }
    if (b >= p.B) return;
    const double tau = ct.tau;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, amax = 0;
    const int i0 = blockIdx.y * ROWS_PER_BLOCK;
    const int i1 = min(p.M, i0 + ROWS_PER_BLOCK);
    for (int i = i0; i < i1; ++i) {
        const size_t o = (size_t)i * p.Bp + b;
        const double hi = p.hi[o], lo = p.lo[o];
        const bool hu = fin(hi), hl = fin(lo);
        const bool stop = i >= p.srow0 && i < p.srow0 + p.ns && hu && hi == 0.0;
        const double e = stop ? 1.0 : 0.0;
        const double su = p.su[o], zu = p.zu[o], sl = p.sl[o], zl = p.zl[o];
        const double du = hu ? zu / su : 0.0, dl = hl ? zl / sl : 0.0;
        switch (phase) {
        case PH_YZ:
            p.YZ[o] = (hu ? zu : 0.0) - (hl ? zl : 0.0);
            break;
        case PH_PRE: {
            const double a = p.AX[o] - e * ct.t;
            const double rzu = hu ? su + a - hi * tau : 0.0;
            const double rzl = hl ? sl - a + lo * tau : 0.0;
            p.RZU[o] = rzu; p.RZL[o] = rzl;
            p.D[o] = du + dl;
            p.DS[o] = e * du;
            // system 0: q = W^-2 h ;  system 1 (affine): q = W^-2 (s - rz)
            const double qu0 = hu ? du * hi : 0.0, ql0 = hl ? dl * (-lo) : 0.0;
            const double qu1 = hu ? du * (su - rzu) : 0.0, ql1 = hl ? dl * (sl - rzl) : 0.0;
            p.QU[0][o] = qu0; p.QL[0][o] = ql0; p.Y[0][o] = qu0 - ql0;
            p.QU[1][o] = qu1; p.QL[1][o] = ql1; p.Y[1][o] = qu1 - ql1;
            a0 += rzu * rzu + rzl * rzl;
            a1 += (hu ? su * zu : 0.0) + (hl ? sl * zl : 0.0);
            a2 += (hu ? hi * zu : 0.0) - (hl ? lo * zl : 0.0);
            a3 += e * zu;                                     // t component of G'z is -sum e zu
            a4 += e * qu1;
            // ||G x + s||^2 for the unboundedness certificate
            { const double g1 = hu ? a + su : 0.0, g2 = hl ? -a + sl : 0.0; amax += g1 * g1 + g2 * g2; }
            break;
        }
        case PH_REFINE: {                                     // YR = (uz_u - uz_l) of system r at the current ux
            const double g = p.GUX[r][o] - e * p.UT[r][b];
            const double uzu = du * g - p.QU[r][o];
            p.YR[o] = (du + dl) * g - p.Y[r][o];
            a0 += e * uzu;
            break;
        }
        case PH_DOTS: {                                       // h'z_r over the rows
            const double g = p.GUX[r][o] - e * p.UT[r][b];
            const double uzu = hu ? du * g - p.QU[r][o] : 0.0, uzl = hl ? -dl * g - p.QL[r][o] : 0.0;
            a0 += (hu ? hi * uzu : 0.0) - (hl ? lo * uzl : 0.0);
            break;
        }
        case PH_AFFINE:
        case PH_FINAL: {
            const int rr = phase == PH_AFFINE ? 1 : 2;
            const double dtau = phase == PH_AFFINE ? ct.dtau_a : ct.dtau;
            const double dtt = phase == PH_AFFINE ? ct.dt_a : ct.dt;
            const double g = (p.GUX[rr][o] + dtau * p.GUX[0][o]) - e * dtt;
            double dzu = 0, dsu = 0, dzl = 0, dsl = 0;
            if (hu) {
                dzu = du * g - (p.QU[rr][o] + dtau * p.QU[0][o]);
                const double dsrhs = phase == PH_AFFINE ? -su * zu : -su * zu - p.CU[o] + ct.sigma * ct.mu;
                dsu = (dsrhs - su * dzu) / zu;
                amax = fmax(amax, fmax(-dsu / su, -dzu / zu));
            }
            if (hl) {
                dzl = -dl * g - (p.QL[rr][o] + dtau * p.QL[0][o]);
                const double dsrhs = phase == PH_AFFINE ? -sl * zl : -sl * zl - p.CL[o] + ct.sigma * ct.mu;
                dsl = (dsrhs - sl * dzl) / zl;
                amax = fmax(amax, fmax(-dsl / sl, -dzl / zl));
            }
            if (phase == PH_AFFINE) { p.CU[o] = dsu * dzu; p.CL[o] = dsl * dzl; }
            else { p.DSU[o] = dsu; p.DZU[o] = dzu; p.DSL[o] = dsl; p.DZL[o] = dzl; }
            break;
        }
        case PH_COMB: {                                       // q of the combined system: W^-2 (dz - ds/z)
            const double sg = ct.sigma, smu = ct.sigma * ct.mu;
            const double qu = hu ? du * (-(1.0 - sg) * p.RZU[o] + su + (p.CU[o] - smu) / zu) : 0.0;
            const double ql = hl ? dl * (-(1.0 - sg) * p.RZL[o] + sl + (p.CL[o] - smu) / zl) : 0.0;
            p.QU[2][o] = qu; p.QL[2][o] = ql; p.Y[2][o] = qu - ql;
            a0 += e * qu;
            break;
        }
        case PH_APPLY: {
            const double al = ct.alpha;
            if (hu) { p.su[o] = su + al * p.DSU[o]; p.zu[o] = zu + al * p.DZU[o]; }
            if (hl) { p.sl[o] = sl + al * p.DSL[o]; p.zl[o] = zl + al * p.DZL[o]; }
            break;
        }
        case PH_METRICS: {                                    // max violation of the returned point, AX = K x (unscaled x)
            const double a = p.AX[o];
            if (stop) amax = fmax(amax, a);                   // here: max over the stop block = ripple_stop
            else { if (hu) a0 = fmax(a0, a - hi); }
            if (hl) a0 = fmax(a0, lo - a);
            break;
        }
        }
    }
    double *acc = p.acc;
    const int Bp = p.Bp;
    switch (phase) {
    case PH_PRE:
        atomicAdd(&acc[A_RZ2 * Bp + b], a0); atomicAdd(&acc[A_SZ * Bp + b], a1); atomicAdd(&acc[A_HZ * Bp + b], a2);
        atomicAdd(&acc[A_T1 * Bp + b], a3); atomicAdd(&acc[A_TQ1 * Bp + b], a4); atomicAdd(&acc[A_GXS2 * Bp + b], amax);
        break;
    case PH_REFINE: atomicAdd(&acc[(A_TUZ0 + r) * Bp + b], a0); break;
    case PH_DOTS: atomicAdd(&acc[(A_HZ0 + r) * Bp + b], a0); break;
    case PH_AFFINE: atomic_max_pos(&acc[A_RATIO_A * Bp + b], amax); break;
    case PH_FINAL: atomic_max_pos(&acc[A_RATIO_F * Bp + b], amax); break;
    case PH_COMB: atomicAdd(&acc[A_TQ2 * Bp + b], a0); break;
    case PH_METRICS: atomic_max_pos(&acc[A_VIOL * Bp + b], a0); atomic_max_pos(&acc[A_TMAX * Bp + b], amax); break;
    default: break;
    }
}

// disks: thread = design b, blockIdx.y = chunk of pairs.  Arrays [3][npairs][Bp] (component-major).
constexpr int PAIRS_PER_BLOCK = 16;
__device__ __forceinline__ V3 ld3(const double *a, int np, int Bp, int k, int b)
{
    return V3{a[((size_t)0 * np + k) * Bp + b], a[((size_t)1 * np + k) * Bp + b], a[((size_t)2 * np + k) * Bp + b]};
}
__device__ __forceinline__ void st3(double *a, int np, int Bp, int k, int b, V3 v)
{
    a[((size_t)0 * np + k) * Bp + b] = v.a; a[((size_t)1 * np + k) * Bp + b] = v.b; a[((size_t)2 * np + k) * Bp + b] = v.c;
}

__global__ void disks_kernel(P p, int phase, int r)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const Ctl ct = p.ctl[b];
    if (ct.status != 0.0 && phase != PH_METRICS) return;
    const int np = p.npairs, Bp = p.Bp;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, amax = 0;
    const int k0 = blockIdx.y * PAIRS_PER_BLOCK, k1 = min(np, k0 + PAIRS_PER_BLOCK);
    for (int k = k0; k < k1; ++k) {
        const int pi = p.pair_i[k], pj = p.pair_j[k];
        const double rho = p.rho[(size_t)k * Bp + b];
        if (phase == PH_METRICS) {
            const double xi = p.x[(size_t)pi * Bp + b], xj = p.x[(size_t)pj * Bp + b];
            a0 = fmax(a0, sqrt(xi * xi + xj * xj) - rho);
            continue;
        }
        const V3 s = ld3(p.sd, np, Bp, k, b), z = ld3(p.zd, np, Bp, k, b);
        switch (phase) {
        case PH_PRE: {
            const double xi = p.x[(size_t)pi * Bp + b], xj = p.x[(size_t)pj * Bp + b];
            const V3 rz = V3{s.a - rho * ct.tau, s.b - xi, s.c - xj};     // s + G v - h tau,  G v = (0, -x_pi, -x_pj)
            st3(p.RZD, np, Bp, k, b, rz);
            // NT scaling
            const double rs = hypot(s.b, s.c), rzn = hypot(z.b, z.c);
            const double sn = sqrt(fmax((s.a - rs) * (s.a + rs), 1e-300)), zn = sqrt(fmax((z.a - rzn) * (z.a + rzn), 1e-300));
            const V3 sb = V3{s.a / sn, s.b / sn, s.c / sn}, zb = V3{z.a / zn, z.b / zn, z.c / zn};
            const double gam = sqrt(0.5 * (1.0 + sb.a * zb.a + sb.b * zb.b + sb.c * zb.c));
            const V3 w = V3{(sb.a + zb.a) / (2.0 * gam), (sb.b - zb.b) / (2.0 * gam), (sb.c - zb.c) / (2.0 * gam)};
            const double eta = sqrt(sn / zn);
            p.WD[((size_t)0 * np + k) * Bp + b] = eta;
            p.WD[((size_t)1 * np + k) * Bp + b] = w.a;
            p.WD[((size_t)2 * np + k) * Bp + b] = w.b;
            p.WD[((size_t)3 * np + k) * Bp + b] = w.c;
            st3(p.LAMD, np, Bp, k, b, soc_W(eta, w, z, false));
            // normal-matrix block on (pi, pj): the lower-right 2x2 of W^-2 = eta^-2 (2 v v' - J), v = (w0, -w1, -w2)
            const double e2 = 1.0 / (eta * eta);
            p.HB[((size_t)0 * np + k) * Bp + b] = e2 * (2.0 * w.b * w.b + 1.0);
            p.HB[((size_t)1 * np + k) * Bp + b] = e2 * (2.0 * w.b * w.c);
            p.HB[((size_t)2 * np + k) * Bp + b] = e2 * (2.0 * w.c * w.c + 1.0);
            st3(p.QD[0], np, Bp, k, b, soc_W2inv(eta, w, V3{rho, 0.0, 0.0}));
            st3(p.QD[1], np, Bp, k, b, soc_W2inv(eta, w, V3{s.a - rz.a, s.b - rz.b, s.c - rz.c}));
            a0 += rz.a * rz.a + rz.b * rz.b + rz.c * rz.c;
            a1 += s.a * z.a + s.b * z.b + s.c * z.c;
            a2 += rho * z.a;
            { const double g1 = s.a, g2 = s.b - xi, g3 = s.c - xj; a3 += g1 * g1 + g2 * g2 + g3 * g3; }
            break;
        }
        case PH_DOTS: {
            const double eta = p.WD[((size_t)0 * np + k) * Bp + b];
            const V3 w = V3{p.WD[((size_t)1 * np + k) * Bp + b], p.WD[((size_t)2 * np + k) * Bp + b], p.WD[((size_t)3 * np + k) * Bp + b]};
            const double ui = p.UX[r][(size_t)pi * Bp + b], uj = p.UX[r][(size_t)pj * Bp + b];
            const V3 q = ld3(p.QD[r], np, Bp, k, b);
            a0 += rho * (2.0 * w.a * (w.b * ui + w.c * uj) / (eta * eta) - q.a);
            break;
        }
        case PH_AFFINE:
        case PH_FINAL: {
            const int rr = phase == PH_AFFINE ? 1 : 2;
            const double dtau = phase == PH_AFFINE ? ct.dtau_a : ct.dtau;
            const double eta = p.WD[((size_t)0 * np + k) * Bp + b];
            const V3 w = V3{p.WD[((size_t)1 * np + k) * Bp + b], p.WD[((size_t)2 * np + k) * Bp + b], p.WD[((size_t)3 * np + k) * Bp + b]};
            const double dxi = p.UX[rr][(size_t)pi * Bp + b] + dtau * p.UX[0][(size_t)pi * Bp + b];
            const double dxj = p.UX[rr][(size_t)pj * Bp + b] + dtau * p.UX[0][(size_t)pj * Bp + b];
            const V3 q2 = ld3(p.QD[rr], np, Bp, k, b), q0 = ld3(p.QD[0], np, Bp, k, b);
            // W^-2 (0, -dxi, -dxj): components 1, 2 through the SAME 2x2 block the normal matrix holds (HB), so that G' dz is
            // exactly what the solve assumed (two applications of W^-1 round differently by ~eps * |wbar|^2, which shows up as a
            // dual residual that grows once the cones are strongly active); component 0 from row 0 of eta^-2 (2 v v' - J)
            const double m11 = p.HB[((size_t)0 * np + k) * Bp + b], m12 = p.HB[((size_t)1 * np + k) * Bp + b],
                         m22 = p.HB[((size_t)2 * np + k) * Bp + b];
            const V3 wz = V3{2.0 * w.a * (w.b * dxi + w.c * dxj) / (eta * eta), -(m11 * dxi + m12 * dxj), -(m12 * dxi + m22 * dxj)};
            const V3 dz = V3{wz.a - (q2.a + dtau * q0.a), wz.b - (q2.b + dtau * q0.b), wz.c - (q2.c + dtau * q0.c)};
            const V3 lam = ld3(p.LAMD, np, Bp, k, b);
            V3 dsr;                                             // right-hand side of lam o (W dz + W^-1 ds) = dsr
            const V3 ll = soc_prod(lam, lam);
            if (phase == PH_AFFINE) dsr = V3{-ll.a, -ll.b, -ll.c};
            else { const V3 cc = ld3(p.CD, np, Bp, k, b); dsr = V3{-ll.a - cc.a + ct.sigma * ct.mu, -ll.b - cc.b, -ll.c - cc.c}; }
            const V3 lds = soc_div(lam, dsr);
            const V3 wdz = soc_W(eta, w, dz, false);
            const V3 ds = soc_W(eta, w, V3{lds.a - wdz.a, lds.b - wdz.b, lds.c - wdz.c}, false);
            amax = fmax(amax, fmax(soc_ratio(s, ds), soc_ratio(z, dz)));
            if (phase == PH_AFFINE) st3(p.CD, np, Bp, k, b, soc_prod(soc_W(eta, w, ds, true), wdz));
            else { st3(p.DSD, np, Bp, k, b, ds); st3(p.DZD, np, Bp, k, b, dz); }
            break;
        }
        case PH_COMB: {
            const double eta = p.WD[((size_t)0 * np + k) * Bp + b];
            const V3 w = V3{p.WD[((size_t)1 * np + k) * Bp + b], p.WD[((size_t)2 * np + k) * Bp + b], p.WD[((size_t)3 * np + k) * Bp + b]};
            const V3 lam = ld3(p.LAMD, np, Bp, k, b), cc = ld3(p.CD, np, Bp, k, b), rz = ld3(p.RZD, np, Bp, k, b);
            const V3 ll = soc_prod(lam, lam);
            const V3 dsr = V3{-ll.a - cc.a + ct.sigma * ct.mu, -ll.b - cc.b, -ll.c - cc.c};
            const V3 wl = soc_W(eta, w, soc_div(lam, dsr), false);
            const double f = -(1.0 - ct.sigma);
            st3(p.QD[2], np, Bp, k, b, soc_W2inv(eta, w, V3{f * rz.a - wl.a, f * rz.b - wl.b, f * rz.c - wl.c}));
            break;
        }
        case PH_APPLY: {
            const double al = ct.alpha;
            const V3 ds = ld3(p.DSD, np, Bp, k, b), dz = ld3(p.DZD, np, Bp, k, b);
            st3(p.sd, np, Bp, k, b, V3{s.a + al * ds.a, s.b + al * ds.b, s.c + al * ds.c});
            st3(p.zd, np, Bp, k, b, V3{z.a + al * dz.a, z.b + al * dz.b, z.c + al * dz.c});
            break;
        }
        default: break;
        }
    }
    double *acc = p.acc;
    switch (phase) {
    case PH_PRE:
        atomicAdd(&acc[A_RZ2 * Bp + b], a0); atomicAdd(&acc[A_SZ * Bp + b], a1); atomicAdd(&acc[A_HZ * Bp + b], a2);
        atomicAdd(&acc[A_GXS2 * Bp + b], a3);
        break;
    case PH_DOTS: atomicAdd(&acc[(A_HZ0 + r) * Bp + b], a0); break;
    case PH_AFFINE: atomic_max_pos(&acc[A_RATIO_A * Bp + b], amax); break;
    case PH_FINAL: atomic_max_pos(&acc[A_RATIO_F * Bp + b], amax); break;
    case PH_METRICS: atomic_max_pos(&acc[A_VIOL * Bp + b], a0); break;
    default: break;
    }
}

// contribution of the disk that owns variable j to a column-side quantity: component (side + 1) of a [3][np][Bp] array, negated
// (the rows of G for a disk are (0, -e_pi, -e_pj))
__device__ __forceinline__ double disk_comp(const double *a, int np, int Bp, int code, int b)
{
    return a[((size_t)((code & 1) + 1) * np + (code >> 1)) * Bp + b];
}

// columns: thread = design b, blockIdx.y = chunk of columns
constexpr int COLS_PER_BLOCK = 16;
enum { PC_PRE = 0, PC_RHS, PC_REFINE, PC_ADD, PC_DOTS, PC_AFFINE, PC_COMB, PC_FINAL, PC_APPLY, PC_METRICS };

__global__ void cols_kernel(P p, int phase, int r)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.Bp) return;
    const Ctl ct = p.ctl[b];
    const bool live = b < p.B && ct.status == 0.0;
    const int j0 = blockIdx.y * COLS_PER_BLOCK;
    if (!live && phase != PC_METRICS) {
        if (phase == PC_RHS || phase == PC_REFINE || phase == PC_ADD)
            for (int j = j0; j < j0 + COLS_PER_BLOCK && j < p.Np; ++j) {
                const size_t o = (size_t)j * p.Bp + b;
                if (phase == PC_RHS) p.RHS[r][o] = 0.0;
                else if (phase == PC_REFINE) p.DXV[o] = 0.0;
            }
        return;
    }
    if (b >= p.B) return;
    const int np = p.npairs, Bp = p.Bp;
    const double tau = ct.tau;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, amax = 0;
    const int j1 = min(p.N, j0 + COLS_PER_BLOCK);
    for (int j = j0; j < j1; ++j) {
        const size_t o = (size_t)j * Bp + b;
        const double bu = p.bu[o], bl = p.bl[o];
        const bool hu = fin(bu), hl = fin(bl);
        const int code = p.pair_of[j];
        const double sbu = p.sbu[o], zbu = p.zbu[o], sbl = p.sbl[o], zbl = p.zbl[o];
        const double du = hu ? zbu / sbu : 0.0, dl = hl ? zbl / sbl : 0.0;
        switch (phase) {
        case PC_PRE: {
            const double x = p.x[o], c = p.c[o];
            double gtz = p.KTY[o] + (hu ? zbu : 0.0) - (hl ? zbl : 0.0);
            if (code >= 0) gtz -= disk_comp(p.zd, np, Bp, code, b);
            const double rx = -gtz - c * tau;
            p.RX[o] = rx;
            const double rzu = hu ? sbu + x - bu * tau : 0.0, rzl = hl ? sbl - x + bl * tau : 0.0;
            p.RZBU[o] = rzu; p.RZBL[o] = rzl;
            p.DIAG[o] = du + dl;
            p.QBU[0][o] = hu ? du * bu : 0.0; p.QBL[0][o] = hl ? dl * (-bl) : 0.0;
            p.QBU[1][o] = hu ? du * (sbu - rzu) : 0.0; p.QBL[1][o] = hl ? dl * (sbl - rzl) : 0.0;
            a0 += rx * rx;
            a1 += c * x;
            a2 += gtz * gtz;
            a3 += rzu * rzu + rzl * rzl;
            a4 += (hu ? sbu * zbu : 0.0) + (hl ? sbl * zbl : 0.0);
            a5 += (hu ? bu * zbu : 0.0) - (hl ? bl * zbl : 0.0);
            { const double g1 = hu ? x + sbu : 0.0, g2 = hl ? -x + sbl : 0.0; amax += g1 * g1 + g2 * g2; }
            break;
        }
        case PC_RHS: {                                        // rhs_j = bx_j + (G' q)_j
            double bx;
            if (r == 0) bx = -p.c[o];
            else if (r == 1) bx = p.RX[o];                    // -dx with dx = -rx
            else bx = (1.0 - ct.sigma) * p.RX[o];
            double v = bx + p.KTQ[o] + p.QBU[r][o] - p.QBL[r][o];
            if (code >= 0) v -= disk_comp(p.QD[r], np, Bp, code, b);
            p.RHS[r][o] = v;
            break;
        }
        case PC_REFINE: {                                     // ex_j = bx_j - (G' uz)_j with uz = W^-2 G ux - q
            double bx;
            if (r == 0) bx = -p.c[o];
            else if (r == 1) bx = p.RX[o];
            else bx = (1.0 - ct.sigma) * p.RX[o];
            const double ux = p.UX[r][o];
            double gtuz = p.KTQ[o] + (du + dl) * ux - (p.QBU[r][o] - p.QBL[r][o]);
            if (code >= 0) {
                const int k = code >> 1;
                const double m11 = p.HB[((size_t)0 * np + k) * Bp + b], m12 = p.HB[((size_t)1 * np + k) * Bp + b],
                             m22 = p.HB[((size_t)2 * np + k) * Bp + b];
                const double ui = p.UX[r][(size_t)p.pair_i[k] * Bp + b], uj = p.UX[r][(size_t)p.pair_j[k] * Bp + b];
                // -(W^-2 (0, -ui, -uj))_{side+1} + q_{side+1}:  G' uz at this variable
                gtuz += ((code & 1) ? m12 * ui + m22 * uj : m11 * ui + m12 * uj) + disk_comp(p.QD[r], np, Bp, code, b);
            }
            p.DXV[o] = bx - gtuz;
            break;
        }
        case PC_ADD:                                          // ux += dx (dx solved into DXV)
            p.UX[r][o] += p.DXV[o];
            break;
        case PC_DOTS: {
            const double ux = p.UX[r][o];
            const double uzu = hu ? du * ux - p.QBU[r][o] : 0.0, uzl = hl ? -dl * ux - p.QBL[r][o] : 0.0;
            a0 += p.c[o] * ux;
            a1 += (hu ? bu * uzu : 0.0) - (hl ? bl * uzl : 0.0);
            break;
        }
        case PC_AFFINE:
        case PC_FINAL: {
            const int rr = phase == PC_AFFINE ? 1 : 2;
            const double dtau = phase == PC_AFFINE ? ct.dtau_a : ct.dtau;
            const double dx = p.UX[rr][o] + dtau * p.UX[0][o];
            double dzu = 0, dsu = 0, dzl = 0, dsl = 0;
            if (hu) {
                dzu = du * dx - (p.QBU[rr][o] + dtau * p.QBU[0][o]);
                const double rhs = phase == PC_AFFINE ? -sbu * zbu : -sbu * zbu - p.CBU[o] + ct.sigma * ct.mu;
                dsu = (rhs - sbu * dzu) / zbu;
                amax = fmax(amax, fmax(-dsu / sbu, -dzu / zbu));
            }
            if (hl) {
                dzl = -dl * dx - (p.QBL[rr][o] + dtau * p.QBL[0][o]);
                const double rhs = phase == PC_AFFINE ? -sbl * zbl : -sbl * zbl - p.CBL[o] + ct.sigma * ct.mu;
                dsl = (rhs - sbl * dzl) / zbl;
                amax = fmax(amax, fmax(-dsl / sbl, -dzl / zbl));
            }
            if (phase == PC_AFFINE) { p.CBU[o] = dsu * dzu; p.CBL[o] = dsl * dzl; }
            else { p.DSBU[o] = dsu; p.DZBU[o] = dzu; p.DSBL[o] = dsl; p.DZBL[o] = dzl; p.DXV[o] = dx; }
            break;
        }
        case PC_COMB: {
            const double sg = ct.sigma, smu = ct.sigma * ct.mu;
            p.QBU[2][o] = hu ? du * (-(1.0 - sg) * p.RZBU[o] + sbu + (p.CBU[o] - smu) / zbu) : 0.0;
            p.QBL[2][o] = hl ? dl * (-(1.0 - sg) * p.RZBL[o] + sbl + (p.CBL[o] - smu) / zbl) : 0.0;
            break;
        }
        case PC_APPLY: {
            const double al = ct.alpha;
            p.x[o] += al * p.DXV[o];
            if (hu) { p.sbu[o] = sbu + al * p.DSBU[o]; p.zbu[o] = zbu + al * p.DZBU[o]; }
            if (hl) { p.sbl[o] = sbl + al * p.DSBL[o]; p.zbl[o] = zbl + al * p.DZBL[o]; }
            break;
        }
        case PC_METRICS: {
            const double x = p.x[o];
            if (hu) a0 = fmax(a0, x - bu);
            if (hl) a0 = fmax(a0, bl - x);
            break;
        }
        }
    }
    double *acc = p.acc;
    switch (phase) {
    case PC_PRE:
        atomicAdd(&acc[A_RX2 * Bp + b], a0); atomicAdd(&acc[A_CX * Bp + b], a1); atomicAdd(&acc[A_GTZ2 * Bp + b], a2);
        atomicAdd(&acc[A_RZ2 * Bp + b], a3); atomicAdd(&acc[A_SZ * Bp + b], a4); atomicAdd(&acc[A_HZ * Bp + b], a5);
        atomicAdd(&acc[A_GXS2 * Bp + b], amax);
        break;
    case PC_DOTS: atomicAdd(&acc[(A_CX0 + r) * Bp + b], a0); atomicAdd(&acc[(A_HZ0 + r) * Bp + b], a1); break;
    case PC_AFFINE: atomic_max_pos(&acc[A_RATIO_A * Bp + b], amax); break;
    case PC_FINAL: atomic_max_pos(&acc[A_RATIO_F * Bp + b], amax); break;
    case PC_METRICS: atomic_max_pos(&acc[A_VIOL * Bp + b], a0); break;
    default: break;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// per-design scalar steps (one thread per design)
// ------------------------------------------------------------------------------------------------------------------
enum { SC_PRE = 0, SC_RHST, SC_REFT, SC_ADDT, SC_DIR_A, SC_SIGMA, SC_DIR_F, SC_STEP };

__global__ void scalars_kernel(P p, int phase, int r, int iter)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    Ctl &ct = p.ctl[b];
    if (ct.status != 0.0) return;
    const int Bp = p.Bp;
    double *acc = p.acc;
    const bool has_t = p.ns > 0;
    const double ctt = has_t ? p.ct[b] : 0.0;
    switch (phase) {
    case SC_PRE: {
        const double tau = ct.tau, kap = ct.kap;
        const double sz = acc[A_SZ * Bp + b], hz = acc[A_HZ * Bp + b];
        const double cx = acc[A_CX * Bp + b] + ctt * ct.t;
        const double rxt = has_t ? acc[A_T1 * Bp + b] - ctt * tau : 0.0;      // -(G'z)_t - ct tau,  (G'z)_t = -sum e zu
        const double gtzt = has_t ? acc[A_T1 * Bp + b] : 0.0;
        const double ncon = acc[A_NCON * Bp + b];
        ct.rt = kap + cx + hz;
        ct.mu = (sz + kap * tau) / (ncon + 1.0);
        ct.pcost = cx / tau;
        ct.dcost = -hz / tau;
        ct.hz = hz;
        ct.pres = sqrt(acc[A_RZ2 * Bp + b]) / tau / ct.nrm_h;
        ct.dres = sqrt(acc[A_RX2 * Bp + b] + rxt * rxt) / tau / ct.nrm_c;
        ct.gap = sz / (tau * tau);
        ct.relgap = ct.gap / fmax(fmax(fabs(ct.pcost), fabs(ct.dcost)), 1e-300);
        acc[A_T1 * Bp + b] = rxt;                              // from here on: the t component of rx
        int st = 0;
        // merit: how far the iterate is from the stopping rule (<= 1: stop).  The best iterate is kept: the last iterations of
        // a design with active peak cones can lose primal / dual feasibility again in fp64 (the cone scalings degrade at the
        // boundary); such a design ends with its best iterate if that was within 10x of the tolerances -- what CVX reports as
        // "Inaccurate/Solved", which fir_ap_cvx.m:176-182 accepts as 'Solved'.
        // the dual residual gets 100x the primal tolerance: K ux is an fp64 product (absolute error ~1e-17 |ux|), and once the
        // active slacks reach 1e-12 that noise, divided by the slack, is the floor of the dual residual (~1e-5 relative on
        // designs at the edge of feasibility).  The objective is insensitive to it: it stays put to 9 digits while dres wanders.
        double merit = fmax(fmax(ct.pres / p.feastol, ct.dres / (100.0 * p.feastol)), fmin(ct.gap / p.abstol, ct.relgap / p.reltol));
        if (!(merit == merit)) merit = INFINITY;
        ct.improved = 0.0;
        if (iter > 0 && merit < ct.merit_best) {
            ct.merit_best = merit; ct.tau_best = tau; ct.t_best = ct.t; ct.pcost_best = ct.pcost; ct.dcost_best = ct.dcost;
            ct.dres_best = ct.dres; ct.improved = 1.0;
        }
        // progress = the optimality merit OR the quality of the emerging infeasibility certificate halves
        const double infm = hz < 0.0 ? sqrt(acc[A_GTZ2 * Bp + b] + gtzt * gtzt) / (-hz) / p.feastol : INFINITY;
        const double prog = fmin(merit, infm);
        if (iter == 0 || prog < 0.5 * ct.merit_ref) { ct.merit_ref = prog; ct.iter_ref = iter; }
        const bool stalled = iter - ct.iter_ref > 30.0;       // no halving of the merit in 30 iterations: a design at the edge of
                                                              // feasibility that neither converges nor yields a certificate
        // Farkas certificate h'z < 0, ||G'z|| / (-h'z) <= feastol.  Once the embedding has collapsed (tau / kap -> 0: the
        // signature of infeasibility; tau then falls 100x per iteration) the iterate no longer improves and the quotient sits
        // at the accuracy of the fp64 products -- 1e-7 give or take a rounding, so a fixed 1e-7 is met in one run and missed in
        // the next (C-13 spec, n = 50: status 2 at iteration 39, at 47, or never).  There the test takes 100 x feastol, as the
        // dual residual does above; without the collapse the strict tolerance stands.
        const double cert = hz < 0.0 ? sqrt(acc[A_GTZ2 * Bp + b] + gtzt * gtzt) / (-hz) : INFINITY;
        const bool collapsed = tau <= 1e-8 * kap;
        if (merit <= 1.0) st = 1;
        else if (cert <= p.feastol || (collapsed && cert <= 100.0 * p.feastol)) st = 2;
        else if (cx < 0.0 && sqrt(acc[A_GXS2 * Bp + b]) / (-cx) <= p.feastol) st = 4;
        else if (iter >= p.max_iter || stalled || ct.chol_fail > 2.0 || merit == INFINITY || (ct.merit_best <= 10.0 && merit > 1e3 * ct.merit_best)) {
            st = 3;
            if (ct.merit_best <= 10.0) { st = 1; ct.use_best = 1.0; }
        }
        if (st) {
            ct.status = st;
            ct.iters = iter;
            atomicSub(p.active, 1);
        }
        break;
    }
    case SC_RHST: {                                            // t component of the right-hand side of system r
        double bx;
        if (r == 0) bx = -ctt;
        else if (r == 1) bx = acc[A_T1 * Bp + b];
        else bx = (1.0 - ct.sigma) * acc[A_T1 * Bp + b];
        // (G' q)_t = -sum e qu ; system 0 has hi = 0 on the stop rows: nothing
        const double tq = r == 0 ? 0.0 : acc[(A_TQ0 + r) * Bp + b];
        p.RHST[r][b] = has_t ? bx - tq : 0.0;
        break;
    }
    case SC_REFT: {                                            // t component of the refinement residual
        double bx;
        if (r == 0) bx = -ctt;
        else if (r == 1) bx = acc[A_T1 * Bp + b];
        else bx = (1.0 - ct.sigma) * acc[A_T1 * Bp + b];
        p.RHST[r][b] = has_t ? bx + acc[(A_TUZ0 + r) * Bp + b] : 0.0;
        acc[(A_TUZ0 + r) * Bp + b] = 0.0;
        break;
    }
    case SC_DIR_A:
    case SC_DIR_F: {
        // dtau = (d_tau - d_kap/tau - c'x2 - h'z2) / (c'x1 + h'z1 - kap/tau)
        const int rr = phase == SC_DIR_A ? 1 : 2;
        const double cx1 = acc[A_CX0 * Bp + b] + ctt * p.UT[0][b], hz1 = acc[A_HZ0 * Bp + b];
        const double cx2 = acc[(A_CX0 + rr) * Bp + b] + ctt * p.UT[rr][b], hz2 = acc[(A_HZ0 + rr) * Bp + b];
        const double den = cx1 + hz1 - ct.kap / ct.tau;
        double d_tau, d_kap;
        if (phase == SC_DIR_A) { d_tau = -ct.rt; d_kap = -ct.kap * ct.tau; }
        else { d_tau = -(1.0 - ct.sigma) * ct.rt; d_kap = -ct.kap * ct.tau - ct.dkap_a * ct.dtau_a + ct.sigma * ct.mu; }
        const double dtau = (d_tau - d_kap / ct.tau - cx2 - hz2) / den;
        const double dkap = (d_kap - ct.kap * dtau) / ct.tau;
        const double dtt = p.UT[rr][b] + dtau * p.UT[0][b];
        double ratio = 0.0;
        if (dtau < 0.0) ratio = fmax(ratio, -dtau / ct.tau);
        if (dkap < 0.0) ratio = fmax(ratio, -dkap / ct.kap);
        if (phase == SC_DIR_A) { ct.dtau_a = dtau; ct.dkap_a = dkap; ct.dt_a = dtt; atomic_max_pos(&acc[A_RATIO_A * Bp + b], ratio); }
        else { ct.dtau = dtau; ct.dkap = dkap; ct.dt = dtt; atomic_max_pos(&acc[A_RATIO_F * Bp + b], ratio); }
        ct.den = den;
        break;
    }
    case SC_SIGMA: {
        const double ratio = acc[A_RATIO_A * Bp + b];
        const double al = ratio > 1.0 ? 1.0 / ratio : 1.0;
        const double om = 1.0 - al;
        ct.sigma = om * om * om;
        break;
    }
    case SC_STEP: {
        const double ratio = acc[A_RATIO_F * Bp + b];
        double al = ratio > 0.0 ? 0.99 / ratio : 1.0;
        if (al > 1.0) al = 1.0;
        ct.alpha = al;
        ct.tau += al * ct.dtau;
        ct.kap += al * ct.dkap;
        ct.t += al * ct.dt;
        break;
    }
    default: break;
    }
}

// the t component of a solved system lives in row N of the factorised matrix: copy it out / in
__global__ void init_kernel(P p)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.Bp) return;
    // constraint count, ||h||, ||c||
    double ncon = 0, h2 = 0, c2 = 0;
    if (b < p.B) {
        for (int i = 0; i < p.M; ++i) {
            const double hi = p.hi[(size_t)i * p.Bp + b], lo = p.lo[(size_t)i * p.Bp + b];
            if (fin(hi)) { ncon += 1; h2 += hi * hi; }
            if (fin(lo)) { ncon += 1; h2 += lo * lo; }
        }
        for (int j = 0; j < p.N; ++j) {
            const double bu = p.bu[(size_t)j * p.Bp + b], bl = p.bl[(size_t)j * p.Bp + b], c = p.c[(size_t)j * p.Bp + b];
            if (fin(bu)) { ncon += 1; h2 += bu * bu; }
            if (fin(bl)) { ncon += 1; h2 += bl * bl; }
            c2 += c * c;
        }
        for (int k = 0; k < p.npairs; ++k) { const double rho = p.rho[(size_t)k * p.Bp + b]; ncon += 1; h2 += rho * rho; }
        if (p.ns > 0) c2 += p.ct[b] * p.ct[b];
    }
    Ctl ct;
    memset(&ct, 0, sizeof ct);
    ct.tau = 1.0; ct.kap = 1.0; ct.t = 0.0;
    ct.nrm_h = fmax(1.0, sqrt(h2)); ct.nrm_c = fmax(1.0, sqrt(c2));
    ct.status = b < p.B ? 0.0 : 1.0;
    ct.merit_best = INFINITY;
    p.ctl[b] = ct;
    p.acc[A_NCON * p.Bp + b] = ncon;     // re-written every iteration by the host-side memset + this value (see solve loop)
}

__global__ void keep_best_kernel(P p)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B || p.ctl[b].improved == 0.0) return;
    for (int j = blockIdx.y; j < p.N; j += gridDim.y) p.XB[(size_t)j * p.Bp + b] = p.x[(size_t)j * p.Bp + b];
}

// live[0 .. count) = the designs still running (any order); also commits the pivot failures the split factorisation of the
// previous iteration flagged (the one-CTA kernel counts them itself)
__global__ void live_list_kernel(P p, int *__restrict__ live, int *__restrict__ count, int *__restrict__ failflag)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    if (failflag[b]) { p.ctl[b].chol_fail += 1.0; failflag[b] = 0; }
    if (p.ctl[b].status == 0.0) live[atomicAdd(count, 1)] = b;
}

__global__ void fill_kernel(double *a, long long n, double v)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}
// x / tau in place for the finished designs' solutions, info rows
__global__ void finish_kernel(P p, double *__restrict__ z_out, double *__restrict__ info)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    Ctl ct = p.ctl[b];
    const bool best = ct.use_best != 0.0;
    if (best) { ct.tau = ct.tau_best; ct.t = ct.t_best; ct.pcost = ct.pcost_best; ct.dcost = ct.dcost_best; ct.dres = ct.dres_best; }
    const double it = 1.0 / ct.tau;
    const double *xs = best ? p.XB : p.x;
    for (int j = 0; j < p.Np; ++j) {
        const double v = j < p.N ? xs[(size_t)j * p.Bp + b] * it : 0.0;
        z_out[(size_t)j * p.Bp + b] = v;
        p.x[(size_t)j * p.Bp + b] = v;          // the metrics pass recomputes K x from this
    }
    double *o = info + (size_t)b * 8;
    o[0] = ct.status == 4.0 ? 3.0 : ct.status;
    o[1] = ct.iters;
    o[2] = ct.pcost;
    o[3] = ct.dcost;
    o[5] = ct.dres;
    o[6] = ct.dcost;
    o[7] = ct.t * it;
}
__global__ void finish2_kernel(P p, double *__restrict__ info)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    info[(size_t)b * 8 + 4] = p.acc[A_VIOL * p.Bp + b];
    if (p.ns > 0) info[(size_t)b * 8 + 7] = p.acc[A_TMAX * p.Bp + b];
}

// ------------------------------------------------------------------------------------------------------------------
// moments: out[p][b] = sum_i T[p][i] * d[i][b] for cos and sin tables, rows [i0, i1), split over blockIdx.z
// block = 32 designs x 8 lags
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
moments_kernel(P p, const double *__restrict__ d, int row0, int row1, int nlag, T *__restrict__ outC, T *__restrict__ outS)
{
    const int b = blockIdx.x * 32 + (threadIdx.x & 31);
    const int lag = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (lag >= nlag) return;
    const bool live = b < p.B && p.ctl[min(b, p.Bp - 1)].status == 0.0;
    const int nz = gridDim.z;
    const int chunk = ((row1 - row0 + nz - 1) / nz + 3) / 4 * 4;
    const int i0 = row0 + blockIdx.z * chunk, i1 = min(row1, i0 + chunk);
    T accC = Num<T>::zero(), accS = Num<T>::zero();
    if (live) {
        const dd *tc = p.TC + (size_t)lag * p.Mp, *ts = p.TS + (size_t)lag * p.Mp;
        for (int i = i0; i < i1; ++i) {
            const double w = d[(size_t)i * p.Bp + b];
            const dd c = tc[i], s = ts[i];
            if constexpr (sizeof(T) == sizeof(dd)) {
                // (c.hi + c.lo) * w accumulated with error-free transformations (sloppy double-double sum)
                dd pc = two_prod(c.hi, w); pc.lo = DD_FMA(c.lo, w, pc.lo);
                dd ps = two_prod(s.hi, w); ps.lo = DD_FMA(s.lo, w, ps.lo);
                dd t1 = two_sum(accC.hi, pc.hi); t1.lo = DD_ADD(t1.lo, DD_ADD(accC.lo, pc.lo)); accC = quick_two_sum(t1.hi, t1.lo);
                dd t2 = two_sum(accS.hi, ps.hi); t2.lo = DD_ADD(t2.lo, DD_ADD(accS.lo, ps.lo)); accS = quick_two_sum(t2.hi, t2.lo);
            } else {
                accC = fma(c.hi, w, accC);
                accS = fma(s.hi, w, accS);
            }
        }
    }
    if (b < p.Bp) {
        outC[((size_t)blockIdx.z * nlag + lag) * p.Bp + b] = accC;
        outS[((size_t)blockIdx.z * nlag + lag) * p.Bp + b] = accS;
    }
}

// normal matrix H [b][NVp][NVp] (lower triangle, row-major) from the moments
template <typename T>
__global__ void __launch_bounds__(256)
assemble_kernel(P p, const T *__restrict__ MC, const T *__restrict__ MS, const T *__restrict__ BC, const T *__restrict__ BS,
                int nsplit, int nlagM, int nlagB, T *__restrict__ Hall)
{
    const int b = blockIdx.x;
    T *H = Hall + (size_t)b * p.NVp * p.NVp;
    const bool live = b < p.B && p.ctl[b].status == 0.0;
    if (!live) return;                                           // the factorisation and the solves skip finished designs too
    const int NV = p.NV, NVp = p.NVp, N = p.N, Bp = p.Bp;
    auto mom = [&](const T *m, int nlag, int lag) {
        T s = m[(size_t)lag * Bp + b];
        const int ns_ = m == MC || m == MS ? nsplit : 1;          // the border moments are computed unsplit
        for (int q = 1; q < ns_; ++q) s = Num<T>::add(s, m[((size_t)q * nlag + lag) * Bp + b]);
        return s;
    };
    for (long long e = (long long)blockIdx.y * blockDim.x + threadIdx.x; e < (long long)NVp * NVp; e += (long long)gridDim.y * blockDim.x) {
        const int rr = (int)(e / NVp), cc = (int)(e % NVp);
        if (cc > rr) continue;
        T v = Num<T>::zero();
        if (!live || rr >= NV) {
            if (rr == cc) v = Num<T>::from(1.0);
        } else if (rr < N) {
            const int tr = p.col_type[rr], tc = p.col_type[cc];
            if (tr != 3 && tc != 3) {
                const int qr = tr == 0 ? 0 : p.col_q[rr], qc = tc == 0 ? 0 : p.col_q[cc];
                const bool sr = tr == 2, sc = tc == 2;
                const int dq = qr > qc ? qr - qc : qc - qr, sq = qr + qc;
                T t;
                if (!sr && !sc) t = Num<T>::add(mom(MC, nlagM, dq), mom(MC, nlagM, sq));
                else if (sr && sc) t = Num<T>::sub(mom(MC, nlagM, dq), mom(MC, nlagM, sq));
                else {
                    // cos(q_c w) sin(q_s w) = (S[q_s + q_c] + S[q_s - q_c]) / 2,  S[-p] = -S[p]
                    const int qs = sr ? qr : qc, qcs = sr ? qc : qr;
                    T d2 = mom(MS, nlagM, dq);
                    if (qs < qcs) d2 = Num<T>::neg(d2);
                    t = Num<T>::add(mom(MS, nlagM, sq), d2);
                }
                v = Num<T>::mul_d(t, 0.5 * p.col_amp[rr] * p.col_amp[cc]);
            }
            if (rr == cc) {
                v = Num<T>::add(v, Num<T>::from(p.DIAG[(size_t)rr * Bp + b]));
                if (tr == 3) v = Num<T>::add(v, Num<T>::from(1.0));
            }
            const int pr = p.pair_of[rr], pc = p.pair_of[cc];
            if (pr >= 0 && pc >= 0 && (pr >> 1) == (pc >> 1)) {
                const int k = pr >> 1;
                const int which = (pr & 1) + (pc & 1);               // 0: m11, 1: m12, 2: m22
                v = Num<T>::add(v, Num<T>::from(p.HB[((size_t)which * p.npairs + k) * Bp + b]));
            }
        } else {                                                     // rr == N: the t row
            if (p.ns == 0) { if (cc == N) v = Num<T>::from(1.0); }
            else if (cc == N) v = mom(BC, nlagB, 0);
            else {
                const int tc = p.col_type[cc];
                if (tc != 3) {
                    const T m = tc == 2 ? mom(BS, nlagB, p.col_q[cc]) : mom(BC, nlagB, tc == 0 ? 0 : p.col_q[cc]);
                    v = Num<T>::neg(Num<T>::mul_d(m, p.col_amp[cc]));
                }
            }
        }
        H[(size_t)rr * NVp + cc] = v;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// batched Cholesky, one CTA of 256 threads per design: right-looking, 32-wide panels, 64 x 64 trailing tiles
// ------------------------------------------------------------------------------------------------------------------
constexpr int CHOL_THREADS = 512;   // 16 warps: 4 per scheduler keep the FP64 pipe issuing through the long dd dependency chains

// The three stages of one panel step k0 of the right-looking factorisation, shared by the one-CTA-per-design kernel and by the
// split kernels (few designs left: one design's work spread over many CTAs).  NT = CHOL_THREADS threads.
// (1) diagonal block: load, factor in shared memory (Ld stays valid for the caller), store
template <typename T>
__device__ __forceinline__ void chol_diag_block(T *__restrict__ H, int n, int k0, T *Ld, int *fail, T *Xi, T *__restrict__ Xout)
{
    // Xi [PANEL][PANEL+1] (shared): inv(L11), built by the same elimination steps applied to the identity -- row j of the
    // inverse is final once column j of L11 is; the triangular solves then multiply by it instead of substituting serially.
    constexpr int LDS_ = PANEL + 1, NT = CHOL_THREADS;
    const int tid = threadIdx.x;
    for (int e = tid; e < PANEL * PANEL; e += NT) {
        const int r = e / PANEL, c = e % PANEL;
        Ld[r * LDS_ + c] = c <= r ? H[(size_t)(k0 + r) * n + k0 + c] : Num<T>::zero();
        Xi[r * LDS_ + c] = Num<T>::from(r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int j = 0; j < PANEL; ++j) {
        if (tid == 0) {
            T d = Ld[j * LDS_ + j];
            if (!Num<T>::positive(d)) { *fail = 1; d = Num<T>::from(1e-300); }
            Ld[j * LDS_ + j] = Num<T>::sqrt_(d);
        }
        __syncthreads();
        if (tid > j && tid < PANEL) Ld[tid * LDS_ + j] = Num<T>::div(Ld[tid * LDS_ + j], Ld[j * LDS_ + j]);
        else if (tid >= PANEL && tid - PANEL <= j) Xi[j * LDS_ + tid - PANEL] = Num<T>::div(Xi[j * LDS_ + tid - PANEL], Ld[j * LDS_ + j]);
        __syncthreads();
        // trailing part of the block: entries (r, c), j < c <= r < PANEL; inverse: rows r > j, columns c <= j
        for (int e = tid; e < PANEL * PANEL; e += NT) {
            const int r = e / PANEL, c = e % PANEL;
            if (c > j && c <= r) Ld[r * LDS_ + c] = Num<T>::fnma(Ld[r * LDS_ + j], Ld[c * LDS_ + j], Ld[r * LDS_ + c]);
            else if (r > j && c <= j) Xi[r * LDS_ + c] = Num<T>::fnma(Ld[r * LDS_ + j], Xi[j * LDS_ + c], Xi[r * LDS_ + c]);
        }
        __syncthreads();
    }
    for (int e = tid; e < PANEL * PANEL; e += NT) {
        const int r = e / PANEL, c = e % PANEL;
        if (c <= r) H[(size_t)(k0 + r) * n + k0 + c] = Ld[r * LDS_ + c];
        if (Xout) Xout[e] = Xi[r * LDS_ + c];
    }
}

// (2) panel rows [r0, r0 + nr) below the block, nr <= 256: X L11' = A21.  Two threads per row: the pair splits every inner sum
// in halves.  Ld: the factored diagonal block in shared memory.
template <typename T>
__device__ __forceinline__ void chol_panel_rows(T *__restrict__ H, int n, int k0, int r0, int nr, const T *Ld, T *Xs)
{
    constexpr int LDS_ = PANEL + 1, NT = CHOL_THREADS;
    const int tid = threadIdx.x;
    __syncthreads();
    for (int e = tid; e < nr * PANEL; e += NT) {
        const int r = e / PANEL, c = e % PANEL;
        Xs[r * LDS_ + c] = H[(size_t)(k0 + PANEL + r0 + r) * n + k0 + c];
    }
    __syncthreads();
    {
        const int row = tid >> 1, half = tid & 1;
        T *xr = Xs + (row < nr ? row : 0) * LDS_;
        for (int j = 0; j < PANEL; ++j) {
            // partial sums over l = half, half + 2, ... < j, combined through a shuffle of the pair
            T s = Num<T>::zero();
            if (row < nr)
                for (int l = half; l < j; l += 2) s = Num<T>::fnma_acc(xr[l], Ld[j * LDS_ + l], s);
            s = Num<T>::renorm(s);
            T o;
            if constexpr (sizeof(T) == sizeof(dd)) {
                o.hi = __shfl_xor_sync(0xffffffffu, s.hi, 1);
                o.lo = __shfl_xor_sync(0xffffffffu, s.lo, 1);
            } else {
                o = __shfl_xor_sync(0xffffffffu, s, 1);
            }
            if (row < nr && half == 0) xr[j] = Num<T>::div(Num<T>::add(xr[j], Num<T>::add(s, o)), Ld[j * LDS_ + j]);
            __syncwarp();
        }
    }
    __syncthreads();
    for (int e = tid; e < nr * PANEL; e += NT) {
        const int r = e / PANEL, c = e % PANEL;
        H[(size_t)(k0 + PANEL + r0 + r) * n + k0 + c] = Xs[r * LDS_ + c];
    }
}

// (3) one 64 x 64 tile (ti, tj), tj <= ti, of the trailing update A22[I][J] -= P[I] P[J]'; thread = 4 x 2 outputs
template <typename T>
__device__ __forceinline__ void chol_update_tile(T *__restrict__ H, int n, int k0, int ti, int tj, T *Xs)
{
    constexpr int NT = CHOL_THREADS;
    const int tid = threadIdx.x;
    T *Pi = Xs, *Pj = Xs + PANEL * TILE;                 // k-major: [PANEL][TILE]
    const int base = k0 + PANEL;
    const int ty = tid >> 5, tx = tid & 31;              // a warp shares its 4 rows (broadcast), lane tx owns columns tx and tx + 32 (conflict-free 16-byte loads)
    __syncthreads();
    for (int e = tid; e < TILE * PANEL; e += NT) {
        const int r = e / PANEL, c = e % PANEL;
        const int gi = base + ti * TILE + r, gj = base + tj * TILE + r;
        Pi[c * TILE + r] = gi < n ? H[(size_t)gi * n + k0 + c] : Num<T>::zero();
        Pj[c * TILE + r] = gj < n ? H[(size_t)gj * n + k0 + c] : Num<T>::zero();
    }
    __syncthreads();
    T acc[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int gi = base + ti * TILE + ty * 4 + i, gj = base + tj * TILE + tx + 32 * j;
            acc[i][j] = (gi < n && gj <= gi) ? H[(size_t)gi * n + gj] : Num<T>::zero();
        }
    // on a diagonal tile the warps whose rows lie in its upper half own no entry of the columns tx + 32 (above the diagonal)
    const bool upper_half_of_diag = ti == tj && ty < 8;
    if (upper_half_of_diag) {
#pragma unroll 8
        for (int k = 0; k < PANEL; ++k) {
            const T b0 = Pj[k * TILE + tx];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i][0] = Num<T>::fnma_acc(Pi[k * TILE + ty * 4 + i], b0, acc[i][0]);
        }
    } else {
#pragma unroll 8
        for (int k = 0; k < PANEL; ++k) {
            T a[4], bb[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = Pi[k * TILE + ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 2; ++j) bb[j] = Pj[k * TILE + tx + 32 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) acc[i][j] = Num<T>::fnma_acc(a[i], bb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int gi = base + ti * TILE + ty * 4 + i, gj = base + tj * TILE + tx + 32 * j;
            if (gi < n && gj <= gi) H[(size_t)gi * n + gj] = Num<T>::renorm(acc[i][j]);
        }
}

template <typename T>
__global__ void __launch_bounds__(CHOL_THREADS, 1)
cholesky_kernel(P p, T *__restrict__ Hall, T *__restrict__ Xall)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    if (b >= p.B || p.ctl[b].status != 0.0) return;
    T *H = Hall + (size_t)b * p.NVp * p.NVp;
    T *Xinv = Xall + (size_t)b * p.NVp * PANEL;                // [NVp / PANEL][PANEL][PANEL]: inverses of the diagonal blocks
    const int n = p.NVp, tid = threadIdx.x;
    constexpr int LDS_ = PANEL + 1;
    T *Ld = reinterpret_cast<T *>(smem_raw);                 // [PANEL][PANEL+1] diagonal block
    T *Xs = Ld + PANEL * LDS_;                               // [256][PANEL+1] panel rows / [2][PANEL][TILE] update tiles
    __shared__ int fail;
    if (tid == 0) fail = 0;
    for (int k0 = 0; k0 < n; k0 += PANEL) {
        chol_diag_block<T>(H, n, k0, Ld, &fail, Xs, Xinv + (size_t)k0 * PANEL);
        const int rows = n - k0 - PANEL;
        if (rows <= 0) break;
        for (int r0 = 0; r0 < rows; r0 += 256) chol_panel_rows<T>(H, n, k0, r0, min(256, rows - r0), Ld, Xs);
        __syncthreads();
        const int nt = (rows + TILE - 1) / TILE;
        for (int ti = 0; ti < nt; ++ti)
            for (int tj = 0; tj <= ti; ++tj) chol_update_tile<T>(H, n, k0, ti, tj, Xs);
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0 && fail) p.ctl[b].chol_fail += 1.0;
}

// Split factorisation for the late iterations of a batch and for single designs: with fewer live designs than SMs the
// one-CTA-per-design kernel leaves the GPU idle while every matrix takes its full single-SM latency (5.5 ms at order 512 in
// double-double).  Here one panel step is three launches whose grids spread ONE design's work over many CTAs:
// stage 0 the diagonal block (one CTA per design), stage 1 the panel rows in chunks of 256 (grid.y chunks), stage 2 the
// trailing tiles (grid.y = tile pairs).  `live` lists the designs still running.
template <typename T>
__global__ void __launch_bounds__(CHOL_THREADS, 1)
cholesky_split_kernel(P p, T *__restrict__ Hall, T *__restrict__ Xall, const int *__restrict__ live, int k0, int stage,
                      int *__restrict__ failflag)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = live[blockIdx.x];
    T *H = Hall + (size_t)b * p.NVp * p.NVp;
    const int n = p.NVp, tid = threadIdx.x;
    constexpr int LDS_ = PANEL + 1;
    T *Ld = reinterpret_cast<T *>(smem_raw);
    T *Xs = Ld + PANEL * LDS_;
    const int rows = n - k0 - PANEL;
    if (stage == 0) {
        __shared__ int fail;
        if (tid == 0) fail = 0;
        __syncthreads();
        chol_diag_block<T>(H, n, k0, Ld, &fail, Xs, Xall + ((size_t)b * p.NVp + k0) * PANEL);
        __syncthreads();
        if (tid == 0 && fail) failflag[b] = 1;                    // committed once per factorisation by live_list_kernel
    } else if (stage == 1) {
        const int r0 = blockIdx.y * 256;
        if (r0 >= rows) return;
        for (int e = tid; e < PANEL * PANEL; e += CHOL_THREADS) {
            const int r = e / PANEL, c = e % PANEL;
            Ld[r * LDS_ + c] = c <= r ? H[(size_t)(k0 + r) * n + k0 + c] : Num<T>::zero();
        }
        chol_panel_rows<T>(H, n, k0, r0, min(256, rows - r0), Ld, Xs);
    } else {
        // tile pair index -> (ti, tj), tj <= ti
        int t = blockIdx.y, ti = 0;
        while (t > ti) { t -= ti + 1; ++ti; }
        const int nt = (rows + TILE - 1) / TILE;
        if (ti >= nt) return;
        chol_update_tile<T>(H, n, k0, ti, t, Xs);
    }
}

// solve L L' u = rhs for system r of every live design: rhs = (RHS[r] | DXV , RHST[r]) -> (UX[r] | DXV, UT[r] | RHST)
// One CTA per design; 32-wide blocks.  The diagonal blocks are applied through their inverses (built by the factorisation):
// a 32 x 32 triangular matrix-vector product, 8 threads per output, instead of 32 serial substitution steps with a divide
// each -- the serial version was 60 % of a single design's solve time (profiles/r2_ipm_single_launches_summary.txt).
template <typename T>
__device__ __forceinline__ T shfl_xor_num(T v, int m)
{
    if constexpr (sizeof(T) == sizeof(dd)) {
        dd o;
        o.hi = __shfl_xor_sync(0xffffffffu, v.hi, m);
        o.lo = __shfl_xor_sync(0xffffffffu, v.lo, m);
        return o;
    } else {
        return __shfl_xor_sync(0xffffffffu, v, m);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
trsolve_kernel(P p, const T *__restrict__ Hall, const T *__restrict__ Xall, const double *rhs, const double *rhst, double *out,
               double *outt, int accumulate_t)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    if (b >= p.B || p.ctl[b].status != 0.0) return;
    const T *H = Hall + (size_t)b * p.NVp * p.NVp;
    const T *Xinv = Xall + (size_t)b * p.NVp * PANEL;
    const int n = p.NVp, tid = threadIdx.x, N = p.N, Bp = p.Bp;
    constexpr int LDS_ = PANEL + 1;
    T *v = reinterpret_cast<T *>(smem_raw);                  // [n]
    T *Ld = v + n;                                           // [PANEL][PANEL+1]: inverse of the current diagonal block
    T *yb = Ld + PANEL * LDS_;                               // [PANEL]
    for (int i = tid; i < n; i += 256)
        v[i] = Num<T>::from(i < N ? rhs[(size_t)i * Bp + b] : i == N ? rhst[b] : 0.0);
    __syncthreads();
    const int orow = tid >> 3, part = tid & 7;               // 8 threads per output of the block product
    // forward: L y = v
    for (int k0 = 0; k0 < n; k0 += PANEL) {
        for (int e = tid; e < PANEL * PANEL; e += 256) Ld[(e / PANEL) * LDS_ + e % PANEL] = Xinv[(size_t)k0 * PANEL + e];
        __syncthreads();
        {
            T s = Num<T>::zero();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = part * 4 + q;
                if (c <= orow) s = Num<T>::fnma_acc(Ld[orow * LDS_ + c], Num<T>::neg(v[k0 + c]), s);
            }
            s = Num<T>::renorm(s);
            s = Num<T>::add(s, shfl_xor_num<T>(s, 1));
            s = Num<T>::add(s, shfl_xor_num<T>(s, 2));
            s = Num<T>::add(s, shfl_xor_num<T>(s, 4));
            if (part == 0) yb[orow] = s;
        }
        __syncthreads();
        if (tid < PANEL) v[k0 + tid] = yb[tid];
        __syncthreads();
        for (int i = k0 + PANEL + tid; i < n; i += 256) {
            T s = v[i];
            const T *row = H + (size_t)i * n + k0;
#pragma unroll 4
            for (int j = 0; j < PANEL; ++j) s = Num<T>::fnma(row[j], v[k0 + j], s);
            v[i] = s;
        }
        __syncthreads();
    }
    // backward: L' u = y
    for (int k0 = n - PANEL; k0 >= 0; k0 -= PANEL) {
        for (int e = tid; e < PANEL * PANEL; e += 256) Ld[(e / PANEL) * LDS_ + e % PANEL] = Xinv[(size_t)k0 * PANEL + e];
        __syncthreads();
        {
            T s = Num<T>::zero();                            // u_c = sum_{r >= c} Xinv[r][c] v_r
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = part * 4 + q;
                if (r >= orow) s = Num<T>::fnma_acc(Ld[r * LDS_ + orow], Num<T>::neg(v[k0 + r]), s);
            }
            s = Num<T>::renorm(s);
            s = Num<T>::add(s, shfl_xor_num<T>(s, 1));
            s = Num<T>::add(s, shfl_xor_num<T>(s, 2));
            s = Num<T>::add(s, shfl_xor_num<T>(s, 4));
            if (part == 0) yb[orow] = s;
        }
        __syncthreads();
        if (tid < PANEL) v[k0 + tid] = yb[tid];
        __syncthreads();
        for (int c = tid; c < k0; c += 256) {
            T s = v[c];
#pragma unroll 4
            for (int j = 0; j < PANEL; ++j) s = Num<T>::fnma(H[(size_t)(k0 + j) * n + c], v[k0 + j], s);
            v[c] = s;
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += 256) {
        const double u = Num<T>::to_double(v[i]);
        if (i < N) out[(size_t)i * Bp + b] = u;
        else if (i == N) { if (accumulate_t) outt[b] += u; else outt[b] = u; }
    }
}

__global__ void add_rows_kernel(double *__restrict__ a, const double *__restrict__ d, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += d[i];
}

// SPD test matrices for mbrf_ipm_cholesky_bench: H = G G' / n + I with a fixed pseudo-random lower-triangular pattern
template <typename T>
__global__ void spd_fill_kernel(T *Hall, int n, int B)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * n * n) return;
    const int b = (int)(idx / ((long long)n * n)), r = (int)((idx / n) % n), c = (int)(idx % n);
    unsigned h = (unsigned)(r * 2654435761u) ^ (unsigned)(c * 40503u) ^ (unsigned)(b * 97u);
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    const double v = r == c ? (double)n : ((double)(h & 0xffff) / 65536.0 - 0.5);     // diagonally dominant: SPD
    Hall[idx] = Num<T>::from(c <= r ? v : 0.0);
}

static int g_precision = 2;          // 0 fp64, 1 double-double, 2 auto (double-double once mu is small)
static double g_dd_switch = 1e-3;    // auto: double-double when min over live designs of mu / mu0 falls below this
static int g_refine = 1, g_refine_fp64 = 0;
static int g_verbose = 0;
static int g_split_max = -1;         // split factorisation when at most this many designs are running (-1: 2/3 of the SMs, 0: never)

// Device-side input / output of a solve (mbrf_fir_ap_solve): `fill` writes c, bl, bu, lo, hi, rho, ct of the padded problem
// with kernels instead of an upload; `after` sees the solutions x [Np x Bp] (caller's units) and info [Bp x 8] on the device
// before anything is copied back.  `extra_bytes` of device scratch beyond the solver's own are handed to both.
struct Hooks {
    std::function<int(const P &, void *, cudaStream_t)> fill;
    std::function<int(const P &, const double *, const double *, void *, cudaStream_t)> after;
    size_t extra_bytes = 0;
};

}  // namespace ipm
}  // namespace mbrf

using namespace mbrf;
using namespace mbrf::ipm;

extern "C" {

int mbrf_ipm_set_option(int which, double value)
{
    switch (which) {
    case 0: if (value < 0 || value > 2) return MBRF_EINVAL; g_precision = (int)value; break;
    case 1: if (!(value > 0)) return MBRF_EINVAL; g_dd_switch = value; break;
    case 2: if (value < 0 || value > 4) return MBRF_EINVAL; g_refine = (int)value; break;
    case 3: g_verbose = (int)value; break;
    case 4: if (value < 0 || value > 4) return MBRF_EINVAL; g_refine_fp64 = (int)value; break;
    case 5: if (value < -1 || value > 1e6) return MBRF_EINVAL; g_split_max = (int)value; break;
    default: return MBRF_EINVAL;
    }
    return MBRF_OK;
}

/* Times the batched factorisation alone (CUDA events on its stream): `B` matrices of order nv (rounded up to the panel
 * width), double-double (use_dd != 0) or fp64, `reps` launches; *ms = mean per launch.  bench.py's solver roofline. */
int mbrf_ipm_cholesky_bench(int nv, int B, int use_dd, int reps, float *ms)
{
    if (int rc = require_device()) return rc;
    if (nv < 1 || B < 1 || reps < 1 || !ms) return MBRF_EINVAL;
    P p;
    memset(&p, 0, sizeof p);
    p.NV = nv; p.NVp = up(nv, PANEL); p.B = B; p.Bp = B;
    const size_t tsz = use_dd ? sizeof(dd) : sizeof(double), hb = (size_t)B * p.NVp * p.NVp * tsz;
    char *buf = nullptr;
    const size_t xb = (size_t)B * p.NVp * PANEL * tsz;           // inverses of the diagonal blocks (built by the same kernel)
    MBRF_CUDA(cudaMalloc(&buf, 2 * hb + xb + (size_t)B * sizeof(Ctl)));
    p.ctl = (Ctl *)(buf + 2 * hb + xb);
    MBRF_CUDA(cudaMemset(p.ctl, 0, (size_t)B * sizeof(Ctl)));
    const long long ne = (long long)B * p.NVp * p.NVp;
    if (use_dd) spd_fill_kernel<dd><<<(unsigned)((ne + 255) / 256), 256>>>((dd *)buf, p.NVp, B);
    else spd_fill_kernel<double><<<(unsigned)((ne + 255) / 256), 256>>>((double *)buf, p.NVp, B);
    MBRF_LAUNCH_CHECK();
    const size_t sm = (size_t)(PANEL * (PANEL + 1) + 256 * (PANEL + 1)) * tsz;
    if (use_dd) MBRF_CUDA(cudaFuncSetAttribute(cholesky_kernel<dd>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    else MBRF_CUDA(cudaFuncSetAttribute(cholesky_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    cudaEvent_t e0, e1;
    MBRF_CUDA(cudaEventCreate(&e0));
    MBRF_CUDA(cudaEventCreate(&e1));
    float total = 0.f;
    for (int r = 0; r <= reps; ++r) {                            // first pass: warm-up
        MBRF_CUDA(cudaMemcpyAsync(buf + hb, buf, hb, cudaMemcpyDeviceToDevice, 0));
        MBRF_CUDA(cudaEventRecord(e0, 0));
        if (use_dd) cholesky_kernel<dd><<<B, CHOL_THREADS, sm>>>(p, (dd *)(buf + hb), (dd *)(buf + 2 * hb));
        else cholesky_kernel<double><<<B, CHOL_THREADS, sm>>>(p, (double *)(buf + hb), (double *)(buf + 2 * hb));
        MBRF_LAUNCH_CHECK();
        MBRF_CUDA(cudaEventRecord(e1, 0));
        MBRF_CUDA(cudaEventSynchronize(e1));
        float t = 0.f;
        MBRF_CUDA(cudaEventElapsedTime(&t, e0, e1));
        if (r) total += t;
    }
    std::vector<Ctl> h((size_t)B);
    MBRF_CUDA(cudaMemcpy(h.data(), p.ctl, (size_t)B * sizeof(Ctl), cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
    for (int b = 0; b < B; ++b) if (h[(size_t)b].chol_fail != 0.0) { set_error("cholesky bench: pivot failure"); return MBRF_ECUDA; }
    *ms = total / reps;
    return MBRF_OK;
}

int mbrf_ipm_padded_sizes(int M, int N, int B, int *Mp, int *Np, int *Bp)
{
    if (Mp) *Mp = up(M, 64);
    if (Np) *Np = up(N, 64);
    if (Bp) *Bp = batch_width(B);
    return MBRF_OK;
}

}  // extern "C"

/*
 * Host-pointer entry point: same problem description as mbrf_fir_pdhg_solve (include/mbrf.h) minus the explicit column.
 */
static int ipm_solve_impl(const double *w_row, int M, const int *col_type, const double *col_kappa, const double *col_amp, int N,
                          const int *pair_i, const int *pair_j, int npairs, const double *c, const double *lo, const double *hi,
                          const double *bl, const double *bu, const double *rho, int B, int simplex_row0, int simplex_rows,
                          const double *simplex_w, int max_iter, double feastol, double reltol, double abstol, double *z_out,
                          double *info_out, const Hooks *hooks)
{
    if (int rc = require_device()) return rc;
    const bool dev_fill = hooks && hooks->fill;
    if (M <= 0 || N <= 0 || B <= 0 || !w_row || !col_type || !col_kappa || !col_amp || (!dev_fill && (!c || !lo || !hi)) ||
        (!z_out && !(hooks && hooks->after)) || !info_out ||
        npairs < 0 || (npairs && (!pair_i || !pair_j || (!rho && !dev_fill))) || simplex_rows < 0 ||
        (simplex_rows && ((!simplex_w && !dev_fill) || simplex_row0 < 0 || simplex_row0 + simplex_rows > M))) {
        set_error("fir_ipm_solve: bad arguments");
        return MBRF_EINVAL;
    }
    // lag unit: all column frequencies must be multiples of 1 or of 1/2
    double u = 1.0;
    for (int j = 0; j < N; ++j)
        if (col_type[j] == 1 || col_type[j] == 2)
            if (col_kappa[j] != floor(col_kappa[j])) u = 0.5;
    int qmax = 0;
    std::vector<int> hq((size_t)N, 0);
    for (int j = 0; j < N; ++j) {
        if (col_type[j] != 1 && col_type[j] != 2) continue;
        const double q = col_kappa[j] / u;
        if (q != floor(q) || q < 0 || q > 1e6) { set_error("fir_ipm_solve: column frequency %g is not a non-negative multiple of 1/2", col_kappa[j]); return MBRF_EINVAL; }
        hq[j] = (int)q;
        if (hq[j] > qmax) qmax = hq[j];
    }
    const int L = 2 * qmax;
    int Mp, Np, Bp;
    mbrf_ipm_padded_sizes(M, N, B, &Mp, &Np, &Bp);
    const int NV = N + 1, NVp = up(NV, PANEL);
    std::vector<int> hpair((size_t)Np, -1);
    for (int k = 0; k < npairs; ++k) {
        const int i = pair_i[k], j = pair_j[k];
        if (i < 0 || i >= N || j < 0 || j >= N || i == j || hpair[i] >= 0 || hpair[j] >= 0) { set_error("fir_ipm_solve: bad pair %d", k); return MBRF_EINVAL; }
        hpair[i] = 2 * k; hpair[j] = 2 * k + 1;
    }

    static thread_local DeviceScratch scratch;
    static thread_local cudaStream_t stream = nullptr;
    static thread_local int stream_dev = -1;
    int dev = 0;
    MBRF_CUDA(cudaGetDevice(&dev));
    if (!stream || stream_dev != dev) {
        if (stream) { cudaSetDevice(stream_dev); cudaStreamDestroy(stream); cudaSetDevice(dev); stream = nullptr; }
        MBRF_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        stream_dev = dev;
    }
    cudaStream_t st = stream;

    // ---- carve the workspace ----
    const size_t rowsz = (size_t)Mp * Bp * 8, colsz = (size_t)Np * Bp * 8, np1 = (size_t)(npairs > 0 ? npairs : 1);
    const size_t dsz = np1 * Bp * 8;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const int nsplit_hint = Bp >= 64 ? 1 : Bp >= 8 ? 4 : 16;
    const int nlagM = L + 1, nlagB = qmax + 1;
    const bool want_dd = g_precision != 0;
    const size_t tsz = want_dd ? sizeof(dd) : sizeof(double);
    size_t total = 0;
    auto need = [&](size_t b) { total += al(b); };
    need((size_t)Mp * 8);                                        // w
    need(2 * (size_t)Mp * Np * 8);                               // K, KT
    need(2 * (size_t)(L + 1) * Mp * sizeof(dd));                 // TC, TS
    need(3 * (size_t)Np * 8);                                    // col_q, col_type, col_amp (+ pair_of)
    need((size_t)Np * 4); need(2 * np1 * 4);
    need(3 * colsz + 2 * rowsz + dsz + (size_t)Bp * 8);          // c bl bu lo hi rho ct
    need(2 * colsz + 4 * rowsz + 4 * colsz + 6 * dsz);           // state (+ best iterate)
    need((13 + 4 * NRHS) * rowsz);                               // row work
    need((12 + 4 * NRHS + 1) * colsz + 2 * NRHS * (size_t)Bp * 8);
    need((3 + 4 + 3 + 3 * NRHS + 3 + 3 + 3 + 3) * dsz);
    need((size_t)NACC * Bp * 8 + (size_t)Bp * sizeof(Ctl) + 256);
    need(2 * (size_t)nsplit_hint * nlagM * Bp * tsz + 2 * (size_t)nsplit_hint * nlagB * Bp * tsz);   // moments
    need((size_t)B * NVp * NVp * tsz);                           // normal matrices
    need((size_t)B * NVp * PANEL * tsz);                         // inverses of the diagonal blocks of the factors
    need(16 * (size_t)Np * Bp * 8);                              // split-K slabs of K' y
    need(colsz + (size_t)Bp * 64);                               // z_out, info
    need(hooks ? hooks->extra_bytes : 0);
    need(2 * (size_t)Bp * 4 + 256);                              // live list, pivot-failure flags, counter
    if (int rc = scratch.reserve(total + (1 << 20))) return rc;
    char *dptr = (char *)scratch.ptr;
    auto take = [&](size_t b) { char *q = dptr; dptr += al(b); return q; };

    P p;
    memset(&p, 0, sizeof p);
    p.M = M; p.Mp = Mp; p.N = N; p.Np = Np; p.NV = NV; p.NVp = NVp; p.B = B; p.Bp = Bp; p.npairs = npairs;
    p.srow0 = simplex_rows ? simplex_row0 : 0; p.ns = simplex_rows; p.L = L; p.qmax = qmax; p.nsplit = nsplit_hint;
    p.feastol = feastol > 0 ? feastol : 1e-7; p.reltol = reltol > 0 ? reltol : 2e-6; p.abstol = abstol > 0 ? abstol : 1e-12;
    p.max_iter = max_iter > 0 ? max_iter : 100;
    double *dw = (double *)take((size_t)Mp * 8);
    double *dK = (double *)take((size_t)Mp * Np * 8), *dKT = (double *)take((size_t)Mp * Np * 8);
    dd *dTC = (dd *)take((size_t)(L + 1) * Mp * sizeof(dd)), *dTS = (dd *)take((size_t)(L + 1) * Mp * sizeof(dd));
    int *dq = (int *)take((size_t)Np * 4), *dtype = (int *)take((size_t)Np * 4), *dpof = (int *)take((size_t)Np * 4);
    double *damp = (double *)take((size_t)Np * 8);
    int *dpi = (int *)take(np1 * 4), *dpj = (int *)take(np1 * 4);
    double *dc = (double *)take(colsz), *dbl = (double *)take(colsz), *dbu = (double *)take(colsz);
    double *dlo = (double *)take(rowsz), *dhi = (double *)take(rowsz), *drho = (double *)take(dsz), *dct = (double *)take((size_t)Bp * 8);
    p.K = dK; p.KT = dKT; p.TC = dTC; p.TS = dTS; p.col_q = dq; p.col_type = dtype; p.col_amp = damp; p.pair_of = dpof;
    p.pair_i = dpi; p.pair_j = dpj; p.c = dc; p.lo = dlo; p.hi = dhi; p.bl = dbl; p.bu = dbu; p.rho = drho; p.ct = dct;
    p.x = (double *)take(colsz);
    p.XB = (double *)take(colsz);
    p.su = (double *)take(rowsz); p.zu = (double *)take(rowsz); p.sl = (double *)take(rowsz); p.zl = (double *)take(rowsz);
    p.sbu = (double *)take(colsz); p.zbu = (double *)take(colsz); p.sbl = (double *)take(colsz); p.zbl = (double *)take(colsz);
    p.sd = (double *)take(3 * dsz); p.zd = (double *)take(3 * dsz);
    double **rowarrs[] = {&p.AX, &p.YZ, &p.RZU, &p.RZL, &p.D, &p.DS, &p.CU, &p.CL, &p.DSU, &p.DZU, &p.DSL, &p.DZL, &p.YR};
    for (auto a : rowarrs) *a = (double *)take(rowsz);
    for (int r = 0; r < NRHS; ++r) { p.QU[r] = (double *)take(rowsz); p.QL[r] = (double *)take(rowsz); p.Y[r] = (double *)take(rowsz); p.GUX[r] = (double *)take(rowsz); }
    double **colarrs[] = {&p.KTY, &p.RX, &p.RZBU, &p.RZBL, &p.DIAG, &p.CBU, &p.CBL, &p.DSBU, &p.DZBU, &p.DSBL, &p.DZBL, &p.DXV, &p.KTQ};
    for (auto a : colarrs) *a = (double *)take(colsz);
    for (int r = 0; r < NRHS; ++r) {
        p.QBU[r] = (double *)take(colsz); p.QBL[r] = (double *)take(colsz); p.RHS[r] = (double *)take(colsz); p.UX[r] = (double *)take(colsz);
        p.UT[r] = (double *)take((size_t)Bp * 8); p.RHST[r] = (double *)take((size_t)Bp * 8);
    }
    p.RZD = (double *)take(3 * dsz); p.WD = (double *)take(4 * dsz); p.LAMD = (double *)take(3 * dsz);
    for (int r = 0; r < NRHS; ++r) p.QD[r] = (double *)take(3 * dsz);
    p.CD = (double *)take(3 * dsz); p.DSD = (double *)take(3 * dsz); p.DZD = (double *)take(3 * dsz); p.HB = (double *)take(3 * dsz);
    p.acc = (double *)take((size_t)NACC * Bp * 8);
    p.ctl = (Ctl *)take((size_t)Bp * sizeof(Ctl));
    p.active = (int *)take(256);
    void *dMC = take((size_t)nsplit_hint * nlagM * Bp * tsz), *dMS = take((size_t)nsplit_hint * nlagM * Bp * tsz);
    void *dBC = take((size_t)nsplit_hint * nlagB * Bp * tsz), *dBS = take((size_t)nsplit_hint * nlagB * Bp * tsz);
    void *dH = take((size_t)B * NVp * NVp * tsz);
    void *dXinv = take((size_t)B * NVp * PANEL * tsz);
    double *dslab = (double *)take(16 * (size_t)Np * Bp * 8);
    double *dzout = (double *)take(colsz), *dinfo = (double *)take((size_t)Bp * 64);
    void *dextra = hooks && hooks->extra_bytes ? take(hooks->extra_bytes) : nullptr;
    int *dlive = (int *)take(2 * (size_t)Bp * 4 + 256), *dfailflag = dlive + Bp, *dlivecount = dfailflag + Bp;
    MBRF_CUDA(cudaMemsetAsync(dlive, 0, 2 * (size_t)Bp * 4 + 256, st));

    // ---- upload ----
    std::vector<double> h;
    MBRF_CUDA(cudaMemcpyAsync(dw, w_row, (size_t)M * 8, cudaMemcpyHostToDevice, st));
    std::vector<int> htype((size_t)Np, 3);
    std::vector<double> hamp((size_t)Np, 0.0);
    hq.resize((size_t)Np, 0);
    for (int j = 0; j < N; ++j) { htype[j] = col_type[j]; hamp[j] = col_amp[j]; }
    MBRF_CUDA(cudaMemcpyAsync(dq, hq.data(), (size_t)Np * 4, cudaMemcpyHostToDevice, st));
    MBRF_CUDA(cudaMemcpyAsync(dtype, htype.data(), (size_t)Np * 4, cudaMemcpyHostToDevice, st));
    MBRF_CUDA(cudaMemcpyAsync(damp, hamp.data(), (size_t)Np * 8, cudaMemcpyHostToDevice, st));
    MBRF_CUDA(cudaMemcpyAsync(dpof, hpair.data(), (size_t)Np * 4, cudaMemcpyHostToDevice, st));
    if (npairs) {
        MBRF_CUDA(cudaMemcpyAsync(dpi, pair_i, (size_t)npairs * 4, cudaMemcpyHostToDevice, st));
        MBRF_CUDA(cudaMemcpyAsync(dpj, pair_j, (size_t)npairs * 4, cudaMemcpyHostToDevice, st));
    }
    auto upload = [&](const double *src, double *dst, int dim, int dimp, double pad, double absent) -> int {
        h.assign((size_t)dimp * Bp, pad);
        if (src)
            for (int i = 0; i < dim; ++i)
                for (int b = 0; b < B; ++b) h[(size_t)i * Bp + b] = src[(size_t)i * B + b];
        else
            for (int i = 0; i < dim; ++i)
                for (int b = 0; b < B; ++b) h[(size_t)i * Bp + b] = absent;
        MBRF_CUDA(cudaMemcpyAsync(dst, h.data(), (size_t)dimp * Bp * 8, cudaMemcpyHostToDevice, st));
        MBRF_CUDA(cudaStreamSynchronize(st));
        return MBRF_OK;
    };
    const double INF = INFINITY;
    if (dev_fill) {                                              // assembled on the device (mbrf_fir_ap_solve)
        if (int rc = hooks->fill(p, dextra, st)) return rc;
    } else {
        if (int rc = upload(c, dc, N, Np, 0.0, 0.0)) return rc;
        if (int rc = upload(bl, dbl, N, Np, -INF, -INF)) return rc;
        if (int rc = upload(bu, dbu, N, Np, INF, INF)) return rc;
        if (int rc = upload(lo, dlo, M, Mp, -INF, -INF)) return rc;
        if (int rc = upload(hi, dhi, M, Mp, INF, INF)) return rc;
        if (npairs) if (int rc = upload(rho, drho, npairs, npairs, 1.0, 1.0)) return rc;
        h.assign((size_t)Bp, 0.0);
        if (simplex_rows) for (int b = 0; b < B; ++b) h[b] = simplex_w[b];
        MBRF_CUDA(cudaMemcpyAsync(dct, h.data(), (size_t)Bp * 8, cudaMemcpyHostToDevice, st));
    }
    MBRF_CUDA(cudaStreamSynchronize(st));

    // ---- matrix ----
    table_kernel<<<(Mp + 127) / 128, 128, 0, st>>>(dw, M, Mp, u, L, dTC, dTS);
    MBRF_LAUNCH_CHECK();
    { const long long nK = (long long)Mp * Np; matrix_kernel<<<(unsigned)((nK + 255) / 256), 256, 0, st>>>(p, dK, dKT); MBRF_LAUNCH_CHECK(); }

    // ---- initial point: x = 0, s = z = e, tau = kap = 1 ----
    auto fill = [&](double *a, size_t n, double v) { fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, (long long)n, v); };
    fill(p.x, (size_t)Np * Bp, 0.0);
    for (double *a : {p.su, p.zu, p.sl, p.zl}) fill(a, (size_t)Mp * Bp, 1.0);
    for (double *a : {p.sbu, p.zbu, p.sbl, p.zbl}) fill(a, (size_t)Np * Bp, 1.0);
    fill(p.sd, 3 * np1 * Bp, 0.0); fill(p.zd, 3 * np1 * Bp, 0.0);
    fill(p.sd, np1 * Bp, 1.0); fill(p.zd, np1 * Bp, 1.0);
    for (int r = 0; r < NRHS; ++r) { fill(p.UT[r], Bp, 0.0); fill(p.RHST[r], Bp, 0.0); fill(p.UX[r], (size_t)Np * Bp, 0.0); }
    fill(p.DXV, (size_t)Np * Bp, 0.0);
    fill(p.CU, (size_t)Mp * Bp, 0.0); fill(p.CL, (size_t)Mp * Bp, 0.0);
    // inputs of the products: the padding rows [M, Mp) are never written again and must be finite (0 * NaN = NaN)
    for (double *a : {p.YZ, p.YR, p.D, p.DS, p.Y[0], p.Y[1], p.Y[2], p.AX, p.GUX[0], p.GUX[1], p.GUX[2]}) fill(a, (size_t)Mp * Bp, 0.0);
    for (double *a : {p.KTY, p.KTQ, p.RHS[0], p.RHS[1], p.RHS[2], p.XB}) fill(a, (size_t)Np * Bp, 0.0);
    MBRF_CUDA(cudaMemsetAsync(p.acc, 0, (size_t)NACC * Bp * 8, st));
    init_kernel<<<(Bp + 127) / 128, 128, 0, st>>>(p);
    MBRF_LAUNCH_CHECK();
    std::vector<double> hncon((size_t)Bp);
    MBRF_CUDA(cudaMemcpyAsync(hncon.data(), p.acc + (size_t)A_NCON * Bp, (size_t)Bp * 8, cudaMemcpyDeviceToHost, st));
    int hactive = B;
    MBRF_CUDA(cudaMemcpyAsync(p.active, &hactive, 4, cudaMemcpyHostToDevice, st));
    MBRF_CUDA(cudaStreamSynchronize(st));

    // ---- product helpers ----
    const int nsm = sm_count();
    int PK = 1;                                                  // split-K slabs of K' y
    if (Bp >= 64) { const int tiles = (Np / 64) * (Bp / 64); PK = (2 * nsm + tiles - 1) / tiles; }
    else PK = 16;
    if (PK > 16) PK = 16;
    if (PK > Mp / 64) PK = Mp / 64;
    if (PK < 1) PK = 1;
    auto gemm_K = [&](const double *X, double *C) -> int {      // C [Mp x Bp] = K X
        if (Bp <= 8) pdhg::launch_thin(Bp, st, dK, Np, dKT, Mp, X, C, Mp, Np, Np, 1, 0LL);
        else pdhg::dgemm_mma_kernel<<<dim3(Bp / 64, Mp / 64, 1), 128, 0, st>>>(dKT, Mp, X, Bp, C, Np, Np, 0);
        MBRF_LAUNCH_CHECK();
        return MBRF_OK;
    };
    auto gemm_KT = [&](const double *Y, double *G) -> int {     // G [Np x Bp] = K' Y
        const long long slab = (long long)Np * Bp;
        if (Bp <= 8) {
            const int kc = up((Mp + PK - 1) / PK, 64);
            pdhg::launch_thin(Bp, st, dKT, Mp, dK, Np, Y, dslab, Np, Mp, kc, PK, slab);
        } else {
            const int kc = up((Mp + PK - 1) / PK, 16);
            pdhg::dgemm_mma_kernel<<<dim3(Bp / 64, Np / 64, PK), 128, 0, st>>>(dK, Np, Y, Bp, dslab, Mp, kc, slab);
        }
        MBRF_LAUNCH_CHECK();
        sum_slabs_kernel<<<(unsigned)((slab + 255) / 256), 256, 0, st>>>(dslab, PK, slab, G, slab);
        MBRF_LAUNCH_CHECK();
        return MBRF_OK;
    };
    const int tb = Bp < 128 ? (Bp < 32 ? 32 : Bp) : 128;
    const dim3 grows((Bp + tb - 1) / tb, (Mp + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK), gcols((Bp + tb - 1) / tb, (Np + COLS_PER_BLOCK - 1) / COLS_PER_BLOCK),
        gdisk((Bp + tb - 1) / tb, (npairs + PAIRS_PER_BLOCK - 1) / PAIRS_PER_BLOCK);
    auto rows = [&](int ph, int r) { rows_kernel<<<grows, tb, 0, st>>>(p, ph, r); };
    auto cols = [&](int ph, int r) { cols_kernel<<<gcols, tb, 0, st>>>(p, ph, r); };
    auto disks = [&](int ph, int r) { if (npairs) disks_kernel<<<gdisk, tb, 0, st>>>(p, ph, r); };
    auto scal = [&](int ph, int r, int it) { scalars_kernel<<<(B + 127) / 128, 128, 0, st>>>(p, ph, r, it); };

    // shared memory of the factorisation kernels
    const size_t sm_chol_dd = (size_t)(PANEL * (PANEL + 1) + 256 * (PANEL + 1)) * sizeof(dd);
    const size_t sm_chol_d = (size_t)(PANEL * (PANEL + 1) + 256 * (PANEL + 1)) * sizeof(double);
    const size_t sm_trs_dd = (size_t)(NVp + PANEL * (PANEL + 1) + PANEL) * sizeof(dd), sm_trs_d = (size_t)(NVp + PANEL * (PANEL + 1) + PANEL) * sizeof(double);
    MBRF_CUDA(cudaFuncSetAttribute(cholesky_kernel<dd>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_chol_dd));
    MBRF_CUDA(cudaFuncSetAttribute(cholesky_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_chol_d));
    MBRF_CUDA(cudaFuncSetAttribute(cholesky_split_kernel<dd>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_chol_dd));
    MBRF_CUDA(cudaFuncSetAttribute(cholesky_split_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_chol_d));
    // a constant limit: concurrent host threads (order searches) solve different sizes, and the attribute is per function
    constexpr int SM_TRS_MAX = 200 * 1024;
    if (sm_trs_dd > (size_t)SM_TRS_MAX) { set_error("fir_ipm_solve: %d variables exceed the triangular-solve kernel", N); return MBRF_EINVAL; }
    MBRF_CUDA(cudaFuncSetAttribute(trsolve_kernel<dd>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TRS_MAX));
    MBRF_CUDA(cudaFuncSetAttribute(trsolve_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TRS_MAX));

    bool use_dd = g_precision == 1;
    // Split factorisation (cholesky_split_kernel) once fewer designs are running than about 2/3 of the SMs: three launches per
    // panel step spread each matrix over many CTAs instead of leaving it to the latency of one SM.
    const int split_max = g_split_max >= 0 ? g_split_max : (2 * nsm) / 3;
    auto chol_split = [&](auto *Hptr, size_t smem) -> int {
        using T = std::remove_pointer_t<decltype(Hptr)>;
        const int L = hactive;
        for (int k0 = 0; k0 < NVp; k0 += PANEL) {
            cholesky_split_kernel<T><<<dim3(L, 1), CHOL_THREADS, smem, st>>>(p, Hptr, (T *)dXinv, dlive, k0, 0, dfailflag);
            MBRF_LAUNCH_CHECK();
            const int rows = NVp - k0 - PANEL;
            if (rows <= 0) break;
            cholesky_split_kernel<T><<<dim3(L, (rows + 255) / 256), CHOL_THREADS, smem, st>>>(p, Hptr, (T *)dXinv, dlive, k0, 1, dfailflag);
            MBRF_LAUNCH_CHECK();
            const int nt = (rows + TILE - 1) / TILE;
            cholesky_split_kernel<T><<<dim3(L, nt * (nt + 1) / 2), CHOL_THREADS, smem, st>>>(p, Hptr, (T *)dXinv, dlive, k0, 2, dfailflag);
            MBRF_LAUNCH_CHECK();
        }
        return MBRF_OK;
    };
    auto chol_split_dd = [&]() -> int { return chol_split((dd *)dH, sm_chol_dd); };
    auto chol_split_d = [&]() -> int { return chol_split((double *)dH, sm_chol_d); };
    bool split = false;
    auto factor = [&]() -> int {
        split = hactive > 0 && hactive <= split_max;
        if (split || g_split_max != 0) {                         // the live list also commits pivot failures flagged by the split kernels
            MBRF_CUDA(cudaMemsetAsync(dlivecount, 0, 4, st));
            live_list_kernel<<<(B + 255) / 256, 256, 0, st>>>(p, dlive, dlivecount, dfailflag);
            MBRF_LAUNCH_CHECK();
        }
        const dim3 gm((Bp + 31) / 32, (nlagM + 7) / 8, nsplit_hint), gb((Bp + 31) / 32, (nlagB + 7) / 8, 1);
        int ay = (int)(((long long)NVp * NVp + 255) / 256);
        if (ay > 64) ay = 64;
        if (use_dd) {
            moments_kernel<dd><<<gm, 256, 0, st>>>(p, p.D, 0, M, nlagM, (dd *)dMC, (dd *)dMS);
            MBRF_LAUNCH_CHECK();
            if (p.ns) { moments_kernel<dd><<<gb, 256, 0, st>>>(p, p.DS, p.srow0, p.srow0 + p.ns, nlagB, (dd *)dBC, (dd *)dBS); MBRF_LAUNCH_CHECK(); }
            assemble_kernel<dd><<<dim3(B, ay), 256, 0, st>>>(p, (dd *)dMC, (dd *)dMS, (dd *)dBC, (dd *)dBS, nsplit_hint, nlagM, nlagB, (dd *)dH);
            MBRF_LAUNCH_CHECK();
            if (split) { if (int rc = chol_split_dd()) return rc; }
            else { cholesky_kernel<dd><<<B, CHOL_THREADS, sm_chol_dd, st>>>(p, (dd *)dH, (dd *)dXinv); MBRF_LAUNCH_CHECK(); }
        } else {
            moments_kernel<double><<<gm, 256, 0, st>>>(p, p.D, 0, M, nlagM, (double *)dMC, (double *)dMS);
            MBRF_LAUNCH_CHECK();
            if (p.ns) { moments_kernel<double><<<gb, 256, 0, st>>>(p, p.DS, p.srow0, p.srow0 + p.ns, nlagB, (double *)dBC, (double *)dBS); MBRF_LAUNCH_CHECK(); }
            assemble_kernel<double><<<dim3(B, ay), 256, 0, st>>>(p, (double *)dMC, (double *)dMS, (double *)dBC, (double *)dBS, nsplit_hint, nlagM, nlagB, (double *)dH);
            MBRF_LAUNCH_CHECK();
            if (split) { if (int rc = chol_split_d()) return rc; }
            else { cholesky_kernel<double><<<B, CHOL_THREADS, sm_chol_d, st>>>(p, (double *)dH, (double *)dXinv); MBRF_LAUNCH_CHECK(); }
        }
        return MBRF_OK;
    };
    auto trsolve = [&](const double *rhs, const double *rhst, double *out, double *outt, int acc_t) -> int {
        if (use_dd) trsolve_kernel<dd><<<B, 256, sm_trs_dd, st>>>(p, (const dd *)dH, (const dd *)dXinv, rhs, rhst, out, outt, acc_t);
        else trsolve_kernel<double><<<B, 256, sm_trs_d, st>>>(p, (const double *)dH, (const double *)dXinv, rhs, rhst, out, outt, acc_t);
        MBRF_LAUNCH_CHECK();
        return MBRF_OK;
    };
    // solve system r (its q arrays, Y[r] and the t sums are in place): UX[r], UT[r], GUX[r]
    auto ksolve = [&](int r) -> int {
        if (int rc = gemm_KT(p.Y[r], p.KTQ)) return rc;
        cols(PC_RHS, r); MBRF_LAUNCH_CHECK();
        scal(SC_RHST, r, 0); MBRF_LAUNCH_CHECK();
        if (int rc = trsolve(p.RHS[r], p.RHST[r], p.UX[r], p.UT[r], 0)) return rc;
        if (int rc = gemm_K(p.UX[r], p.GUX[r])) return rc;
        // the refinement step matters once the weights spread (the double-double phase); the early fp64 iterations only need a
        // direction that makes progress
        for (int k = 0; k < (use_dd ? g_refine : g_refine_fp64); ++k) {
            rows(PH_REFINE, r); MBRF_LAUNCH_CHECK();
            if (int rc = gemm_KT(p.YR, p.KTQ)) return rc;
            cols(PC_REFINE, r); MBRF_LAUNCH_CHECK();
            scal(SC_REFT, r, 0); MBRF_LAUNCH_CHECK();
            if (int rc = trsolve(p.DXV, p.RHST[r], p.DXV, p.UT[r], 1)) return rc;
            cols(PC_ADD, r); MBRF_LAUNCH_CHECK();
            if (int rc = gemm_K(p.DXV, p.YR)) return rc;
            add_rows_kernel<<<(unsigned)(((size_t)Mp * Bp + 255) / 256), 256, 0, st>>>(p.GUX[r], p.YR, (long long)Mp * Bp);
            MBRF_LAUNCH_CHECK();
        }
        rows(PH_DOTS, r); MBRF_LAUNCH_CHECK();
        disks(PH_DOTS, r); MBRF_LAUNCH_CHECK();
        cols(PC_DOTS, r); MBRF_LAUNCH_CHECK();
        return MBRF_OK;
    };

    std::vector<Ctl> hctl((size_t)Bp);
    double mu0 = -1.0;
    int it = 0;
    for (;; ++it) {
        MBRF_CUDA(cudaMemsetAsync(p.acc, 0, (size_t)NACC * Bp * 8, st));
        MBRF_CUDA(cudaMemcpyAsync(p.acc + (size_t)A_NCON * Bp, hncon.data(), (size_t)Bp * 8, cudaMemcpyHostToDevice, st));
        rows(PH_YZ, 0); MBRF_LAUNCH_CHECK();
        if (int rc = gemm_K(p.x, p.AX)) return rc;
        if (int rc = gemm_KT(p.YZ, p.KTY)) return rc;
        disks(PH_PRE, 0); MBRF_LAUNCH_CHECK();
        rows(PH_PRE, 0); MBRF_LAUNCH_CHECK();
        cols(PC_PRE, 0); MBRF_LAUNCH_CHECK();
        scal(SC_PRE, 0, it); MBRF_LAUNCH_CHECK();
        keep_best_kernel<<<dim3((B + 127) / 128, 16), 128, 0, st>>>(p); MBRF_LAUNCH_CHECK();
        MBRF_CUDA(cudaMemcpyAsync(&hactive, p.active, 4, cudaMemcpyDeviceToHost, st));
        if (g_precision == 2 || g_verbose) MBRF_CUDA(cudaMemcpyAsync(hctl.data(), p.ctl, (size_t)Bp * sizeof(Ctl), cudaMemcpyDeviceToHost, st));
        MBRF_CUDA(cudaStreamSynchronize(st));
        if (g_verbose)
            for (int b = 0; b < B && b < g_verbose; ++b)
                fprintf(stderr, "ipm it %3d b %d st %g pcost %+.10e dcost %+.10e gap %.2e pres %.1e dres %.1e tau %.2e kap %.2e mu %.2e%s\n", it, b,
                        hctl[b].status, hctl[b].pcost, hctl[b].dcost, hctl[b].gap, hctl[b].pres, hctl[b].dres, hctl[b].tau, hctl[b].kap, hctl[b].mu,
                        use_dd ? " dd" : "");
        if (hactive <= 0) break;
        if (g_precision == 2) {
            double mumin = INFINITY;
            for (int b = 0; b < B; ++b) if (hctl[b].status == 0.0) mumin = fmin(mumin, hctl[b].mu);
            if (mu0 < 0) mu0 = mumin;
            use_dd = mumin < g_dd_switch * mu0;
        }
        if (int rc = factor()) return rc;
        if (int rc = ksolve(0)) return rc;
        if (int rc = ksolve(1)) return rc;
        scal(SC_DIR_A, 0, 0); MBRF_LAUNCH_CHECK();
        rows(PH_AFFINE, 0); MBRF_LAUNCH_CHECK();
        cols(PC_AFFINE, 0); MBRF_LAUNCH_CHECK();
        disks(PH_AFFINE, 0); MBRF_LAUNCH_CHECK();
        scal(SC_SIGMA, 0, 0); MBRF_LAUNCH_CHECK();
        rows(PH_COMB, 0); MBRF_LAUNCH_CHECK();
        cols(PC_COMB, 0); MBRF_LAUNCH_CHECK();
        disks(PH_COMB, 0); MBRF_LAUNCH_CHECK();
        if (int rc = ksolve(2)) return rc;
        scal(SC_DIR_F, 0, 0); MBRF_LAUNCH_CHECK();
        rows(PH_FINAL, 0); MBRF_LAUNCH_CHECK();
        cols(PC_FINAL, 0); MBRF_LAUNCH_CHECK();
        disks(PH_FINAL, 0); MBRF_LAUNCH_CHECK();
        scal(SC_STEP, 0, 0); MBRF_LAUNCH_CHECK();
        rows(PH_APPLY, 0); MBRF_LAUNCH_CHECK();
        cols(PC_APPLY, 0); MBRF_LAUNCH_CHECK();
        disks(PH_APPLY, 0); MBRF_LAUNCH_CHECK();
    }
    // ---- solutions and metrics in the caller's units ----
    finish_kernel<<<(B + 127) / 128, 128, 0, st>>>(p, dzout, dinfo);
    MBRF_LAUNCH_CHECK();
    if (int rc = gemm_K(p.x, p.AX)) return rc;
    MBRF_CUDA(cudaMemsetAsync(p.acc, 0, (size_t)NACC * Bp * 8, st));
    rows(PH_METRICS, 0); MBRF_LAUNCH_CHECK();
    cols(PC_METRICS, 0); MBRF_LAUNCH_CHECK();
    disks(PH_METRICS, 0); MBRF_LAUNCH_CHECK();
    finish2_kernel<<<(B + 127) / 128, 128, 0, st>>>(p, dinfo);
    MBRF_LAUNCH_CHECK();
    if (hooks && hooks->after)
        if (int rc = hooks->after(p, dzout, dinfo, dextra, st)) return rc;
    std::vector<double> info((size_t)Bp * 8);
    if (z_out) {
        h.assign((size_t)Np * Bp, 0.0);
        MBRF_CUDA(cudaMemcpyAsync(h.data(), dzout, (size_t)Np * Bp * 8, cudaMemcpyDeviceToHost, st));
    }
    MBRF_CUDA(cudaMemcpyAsync(info.data(), dinfo, (size_t)Bp * 64, cudaMemcpyDeviceToHost, st));
    MBRF_CUDA(cudaStreamSynchronize(st));
    if (z_out)
        for (int j = 0; j < N; ++j)
            for (int b = 0; b < B; ++b) z_out[(size_t)j * B + b] = h[(size_t)j * Bp + b];
    memcpy(info_out, info.data(), (size_t)B * 64);
    return MBRF_OK;
}

namespace mbrf {
namespace ipm {

// Host part of the device-side assembly: the union grid of a batch, the specification on the device, pass 1 (per-design
// reductions) and the rows of the stop block.
struct ApPrep {
    std::vector<double> w;               // union grid, sorted, unique
    std::vector<unsigned char> is_base;
    std::vector<int> srows;              // union rows that are a stop row of at least one design
    int M1 = 0, ns = 0;
    assemble::ApSpec sp;
    double *dw = nullptr;
    unsigned char *dbase = nullptr;
    int *dsrows = nullptr;
    assemble::ApRed *dred = nullptr;

    int run(int n, int nband, const double *f, const double *a, const double *d, const double *obj, const double *peak, int B,
            int oversamp)
    {
        using namespace mbrf::assemble;
        if (n < 2 || nband < 1 || nband > MAX_BANDS || B < 1 || !f || !a || !d || !obj || !peak || oversamp < 1) {
            set_error("fir_ap: bad arguments (n=%d nband=%d B=%d)", n, nband, B);
            return MBRF_EINVAL;
        }
        // ---- union grid: w = sort([linspace(-pi, pi, 2 n oversamp), f*pi]) of every design (fir_ap_cvx.m:44-48), unique ----
        const int m = 2 * n * oversamp, ne = 2 * nband;
        std::vector<double> fw((size_t)B * ne);
        for (size_t k = 0; k < fw.size(); ++k) fw[k] = f[k] * M_PI;              // :44
        std::vector<std::pair<double, unsigned char>> grid;
        grid.reserve((size_t)m + fw.size());
        {
            const double start = -M_PI, stop = M_PI;
            volatile double step = (stop - start) / (double)(m - 1);             // numpy / MATLAB linspace: i*step + start, last = stop
            for (int i = 0; i < m; ++i) {
                volatile double t = (double)i * step;
                t = t + start;
                grid.push_back({i == m - 1 ? stop : (double)t, (unsigned char)1});
            }
        }
        for (double v : fw) grid.push_back({v, (unsigned char)0});
        std::sort(grid.begin(), grid.end(),
                  [](const auto &x, const auto &y) { return x.first < y.first || (x.first == y.first && x.second > y.second); });
        for (const auto &g : grid) {
            if (!w.empty() && w.back() == g.first) continue;                     // base samples sort first among equals
            w.push_back(g.first);
            is_base.push_back(g.second);
        }
        M1 = (int)w.size();

        // ---- pass 1 on the device: per-design reductions and the rows of the stop block ----
        static thread_local DeviceScratch pre;
        auto al = [](size_t v) { return (v + 255) / 256 * 256; };
        const size_t b_spec = al((size_t)B * ne * 8), b_d = al((size_t)B * nband * 8), b_B = al((size_t)B * 8);
        const size_t pre_bytes = al((size_t)M1 * 8) + 2 * al((size_t)M1) + 2 * b_spec + b_d + 2 * b_B + al((size_t)B * sizeof(ApRed)) +
                                 al((size_t)M1 * 4);
        if (int rc = pre.reserve(pre_bytes)) return rc;
        char *q = (char *)pre.ptr;
        auto take = [&](size_t bytes) { char *r = q; q += al(bytes); return r; };
        dw = (double *)take((size_t)M1 * 8);
        dbase = (unsigned char *)take((size_t)M1);
        unsigned char *dany = (unsigned char *)take((size_t)M1);
        double *df = (double *)take((size_t)B * ne * 8), *da = (double *)take((size_t)B * ne * 8), *dd_ = (double *)take((size_t)B * nband * 8);
        double *dobj = (double *)take((size_t)B * 8), *dpeak = (double *)take((size_t)B * 8);
        dred = (ApRed *)take((size_t)B * sizeof(ApRed));
        dsrows = (int *)take((size_t)M1 * 4);
        cudaStream_t s0 = 0;
        MBRF_CUDA(cudaMemcpyAsync(dw, w.data(), (size_t)M1 * 8, cudaMemcpyHostToDevice, s0));
        MBRF_CUDA(cudaMemcpyAsync(dbase, is_base.data(), (size_t)M1, cudaMemcpyHostToDevice, s0));
        MBRF_CUDA(cudaMemcpyAsync(df, fw.data(), (size_t)B * ne * 8, cudaMemcpyHostToDevice, s0));
        MBRF_CUDA(cudaMemcpyAsync(da, a, (size_t)B * ne * 8, cudaMemcpyHostToDevice, s0));
        MBRF_CUDA(cudaMemcpyAsync(dd_, d, (size_t)B * nband * 8, cudaMemcpyHostToDevice, s0));
        MBRF_CUDA(cudaMemcpyAsync(dobj, obj, (size_t)B * 8, cudaMemcpyHostToDevice, s0));
        MBRF_CUDA(cudaMemcpyAsync(dpeak, peak, (size_t)B * 8, cudaMemcpyHostToDevice, s0));
        sp.f = df; sp.a = da; sp.d = dd_; sp.obj = dobj; sp.peak = dpeak; sp.nband = nband; sp.n = n; sp.B = B;
        ap_reduce_kernel<<<B, 256, 0, s0>>>(sp, dw, dbase, M1, dred);
        MBRF_LAUNCH_CHECK();
        ap_stop_any_kernel<<<(M1 + 7) / 8, 256, 0, s0>>>(sp, dw, dbase, M1, dred, dany);
        MBRF_LAUNCH_CHECK();
        std::vector<unsigned char> any((size_t)M1);
        MBRF_CUDA(cudaMemcpyAsync(any.data(), dany, (size_t)M1, cudaMemcpyDeviceToHost, s0));
        MBRF_CUDA(cudaStreamSynchronize(s0));
        for (int i = 0; i < M1; ++i) if (any[i]) srows.push_back(i);
        ns = (int)srows.size();
        if (ns) MBRF_CUDA(cudaMemcpyAsync(dsrows, srows.data(), (size_t)ns * 4, cudaMemcpyHostToDevice, s0));
        MBRF_CUDA(cudaStreamSynchronize(s0));
        return MBRF_OK;
    }
};

}  // namespace ipm
}  // namespace mbrf

extern "C" {

int mbrf_fir_ipm_solve(const double *w_row, int M, const int *col_type, const double *col_kappa, const double *col_amp, int N,
                       const int *pair_i, const int *pair_j, int npairs, const double *c, const double *lo, const double *hi,
                       const double *bl, const double *bu, const double *rho, int B, int simplex_row0, int simplex_rows,
                       const double *simplex_w, int max_iter, double feastol, double reltol, double abstol, double *z_out,
                       double *info_out)
{
    return ipm_solve_impl(w_row, M, col_type, col_kappa, col_amp, N, pair_i, pair_j, npairs, c, lo, hi, bl, bu, rho, B, simplex_row0,
                          simplex_rows, simplex_w, max_iter, feastol, reltol, abstol, z_out, info_out, nullptr);
}

/*
 * fir_ap_cvx (fir_ap_cvx.m:1-245) for a batch of designs of one order n, specification in, taps out: the union grid is built
 * here on the host (a few thousand doubles), everything per design -- band masks, bounds, stop rows, radii (:51-142) -- is
 * assembled on the device (fir_assemble.cuh), the batch is solved by the interior-point method, and the minimum-phase factor
 * h = fmp2(r) (:185-202) of every solved design is taken on the device from the solutions before anything is copied back.
 */
int mbrf_fir_ap_solve(int n, int nband, const double *f, const double *a, const double *d, const double *obj, const double *peak,
                      int B, int oversamp, int max_iter, double feastol, double reltol, double abstol, double *x_out, double *h_re,
                      double *h_im, double *info_out, int *rows_out)
{
    using namespace mbrf::assemble;
    if (int rc = require_device()) return rc;
    if (!info_out) { set_error("fir_ap_solve: info_out is required"); return MBRF_EINVAL; }
    if ((h_re || h_im) && (!h_re || !h_im || n > mbrf_fmp2_max_taps())) {
        set_error("fir_ap_solve: taps need both planes and n <= %d", mbrf_fmp2_max_taps());
        return MBRF_EINVAL;
    }
    ApPrep pr;
    if (int rc = pr.run(n, nband, f, a, d, obj, peak, B, oversamp)) return rc;
    const int M1 = pr.M1, ns = pr.ns, M = M1 + ns;
    const std::vector<double> &w = pr.w;
    const std::vector<int> &srows = pr.srows;
    const ApSpec sp = pr.sp;
    double *dw = pr.dw;
    unsigned char *dbase = pr.dbase;
    int *dsrows = pr.dsrows;
    ApRed *dred = pr.dred;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    if (rows_out) { rows_out[0] = M1; rows_out[1] = ns; }

    // ---- matrix description: A = [1, 2cos(w k), 2sin(w k)], k = 1..n-1 (:100); pairs (x_i, x_{n+i-1}), i = 2..n (:133-139) ----
    const int N = 2 * n - 1, np = n - 1;
    std::vector<double> w_row((size_t)M), kappa((size_t)N), amp((size_t)N);
    std::vector<int> type((size_t)N), pi((size_t)np), pj((size_t)np);
    for (int i = 0; i < M1; ++i) w_row[i] = w[i];
    for (int k = 0; k < ns; ++k) w_row[M1 + k] = w[srows[k]];
    type[0] = 0; kappa[0] = 0.0; amp[0] = 1.0;
    for (int k = 1; k < n; ++k) {
        type[k] = 1; type[n - 1 + k] = 2;
        kappa[k] = kappa[n - 1 + k] = (double)k;
        amp[k] = amp[n - 1 + k] = 2.0;
        pi[k - 1] = k; pj[k - 1] = n - 1 + k;
    }
    const int len = 2 * n - 1;
    Hooks hk;
    const size_t b_r = al((size_t)B * len * 8), b_h = al((size_t)B * n * 8), b_tw = al((size_t)mbrf_fmp2_workspace_bytes(n));
    hk.extra_bytes = 3 * b_r + 2 * b_h + b_tw;
    hk.fill = [&](const P &p, void *, cudaStream_t st) -> int {
        if (p.Mp > 65535) { set_error("fir_ap_solve: %d grid rows exceed the assembly kernel's grid (n too large)", p.Mp); return MBRF_EINVAL; }
        ap_fill_rows_kernel<<<dim3((p.Bp + 127) / 128, p.Mp), 128, 0, st>>>(sp, dw, dbase, M1, dsrows, ns, dred, p.Mp, p.Bp,
                                                                             (double *)p.lo, (double *)p.hi);
        MBRF_LAUNCH_CHECK();
        const int rows = p.Np > p.npairs ? p.Np : p.npairs;
        ap_fill_cols_kernel<<<dim3((p.Bp + 127) / 128, rows), 128, 0, st>>>(sp, p.Np, p.Bp, p.npairs, (double *)p.c, (double *)p.bl,
                                                                             (double *)p.bu, (double *)p.rho, (double *)p.ct);
        MBRF_LAUNCH_CHECK();
        return MBRF_OK;
    };
    hk.after = [&](const P &p, const double *x, const double *, void *extra, cudaStream_t st) -> int {
        char *e = (char *)extra;
        double *rr = (double *)e, *ri = (double *)(e + b_r), *xr = (double *)(e + 2 * b_r);
        double *hr = (double *)(e + 3 * b_r), *hi_ = (double *)(e + 3 * b_r + b_h);
        void *tw = e + 3 * b_r + 2 * b_h;
        ap_x_to_r_kernel<<<dim3((len + 127) / 128, B), 128, 0, st>>>(x, n, B, p.Bp, rr, ri, xr);
        MBRF_LAUNCH_CHECK();
        if (x_out) MBRF_CUDA(cudaMemcpyAsync(x_out, xr, (size_t)B * len * 8, cudaMemcpyDeviceToHost, st));
        if (h_re) {
            if (int rc = mbrf_fmp2_batch_device(rr, ri, n, B, hr, hi_, tw, (void *)st)) return rc;
            MBRF_CUDA(cudaMemcpyAsync(h_re, hr, (size_t)B * n * 8, cudaMemcpyDeviceToHost, st));
            MBRF_CUDA(cudaMemcpyAsync(h_im, hi_, (size_t)B * n * 8, cudaMemcpyDeviceToHost, st));
        }
        return MBRF_OK;
    };
    return ipm_solve_impl(w_row.data(), M, type.data(), kappa.data(), amp.data(), N, pi.data(), pj.data(), np, nullptr, nullptr, nullptr,
                          nullptr, nullptr, nullptr, B, M1, ns, nullptr, max_iter, feastol, reltol, abstol, nullptr, info_out, &hk);
}

/* Assembly alone (the arrays mbrf_fir_ap_solve hands to the solver), for callers that pose the problem to another solver and
 * for the parity tests.  Call once with lo_out == NULL to learn rows_out = {grid rows M1, stop-block rows ns}. */
int mbrf_fir_ap_assemble(int n, int nband, const double *f, const double *a, const double *d, const double *obj, const double *peak,
                         int B, int oversamp, int *rows_out, double *w_row_out, double *lo_out, double *hi_out, double *c_out,
                         double *bl_out, double *bu_out, double *rho_out, double *ct_out)
{
    using namespace mbrf::assemble;
    if (int rc = require_device()) return rc;
    if (!rows_out) { set_error("fir_ap_assemble: rows_out is required"); return MBRF_EINVAL; }
    ApPrep pr;
    if (int rc = pr.run(n, nband, f, a, d, obj, peak, B, oversamp)) return rc;
    rows_out[0] = pr.M1; rows_out[1] = pr.ns;
    if (!lo_out) return MBRF_OK;
    if (!w_row_out || !hi_out || !c_out || !bl_out || !bu_out || !rho_out || !ct_out) { set_error("fir_ap_assemble: output arrays missing"); return MBRF_EINVAL; }
    const int M = pr.M1 + pr.ns, N = 2 * n - 1, np = n - 1;
    if (M > 65535) { set_error("fir_ap_assemble: %d grid rows exceed the assembly kernel's grid (n too large)", M); return MBRF_EINVAL; }
    static thread_local DeviceScratch out;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t rb = al((size_t)M * B * 8), cb = al((size_t)N * B * 8), pb = al((size_t)np * B * 8), bb = al((size_t)B * 8);
    if (int rc = out.reserve(2 * rb + 3 * cb + pb + bb)) return rc;
    char *q = (char *)out.ptr;
    double *lo = (double *)q, *hi = (double *)(q + rb), *c = (double *)(q + 2 * rb), *bl = (double *)(q + 2 * rb + cb),
           *bu = (double *)(q + 2 * rb + 2 * cb), *rho = (double *)(q + 2 * rb + 3 * cb), *ct = (double *)(q + 2 * rb + 3 * cb + pb);
    ap_fill_rows_kernel<<<dim3((B + 127) / 128, M), 128>>>(pr.sp, pr.dw, pr.dbase, pr.M1, pr.dsrows, pr.ns, pr.dred, M, B, lo, hi);
    MBRF_LAUNCH_CHECK();
    ap_fill_cols_kernel<<<dim3((B + 127) / 128, N > np ? N : np), 128>>>(pr.sp, N, B, np, c, bl, bu, rho, ct);
    MBRF_LAUNCH_CHECK();
    MBRF_CUDA(cudaMemcpyAsync(lo_out, lo, (size_t)M * B * 8, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaMemcpyAsync(hi_out, hi, (size_t)M * B * 8, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaMemcpyAsync(c_out, c, (size_t)N * B * 8, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaMemcpyAsync(bl_out, bl, (size_t)N * B * 8, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaMemcpyAsync(bu_out, bu, (size_t)N * B * 8, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaMemcpyAsync(rho_out, rho, (size_t)np * B * 8, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaMemcpyAsync(ct_out, ct, (size_t)B * 8, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaStreamSynchronize(0));
    for (int i = 0; i < pr.M1; ++i) w_row_out[i] = pr.w[i];
    for (int k = 0; k < pr.ns; ++k) w_row_out[pr.M1 + k] = pr.w[pr.srows[k]];
    return MBRF_OK;
}

}  // extern "C"
