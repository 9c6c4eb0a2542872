// Host-pointer front end of the FIR design solver: builds the frequency-sampled Fourier matrix on the
// device from a column/row description (so the 30 MB matrix never crosses PCIe), scales its columns,
// uploads the per-design vectors, runs the batched PDHG of pdhg.cu and returns the solutions.
//
// The matrices of the reference all have entries amp_j * {1, cos, sin}(w_i * kappa_j):
//   fir_ap_cvx.m:100      A = [1, 2cos(w k), 2sin(w k)],  k = 1..n-1
//   ss/fir_linprog.m:195-217  Acos = [1, 2cos(w k)] or 2cos(w (k+1/2)),  Asin likewise
//   fir_qp_cvx.m:96-109   [cos(w k), sin(w k); -sin(w k), cos(w k)],  k = 0..n-1
// plus one optional extra column with explicit per-row coefficients (ripple_stop of fir_ap_cvx.m:165).
#include "common.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" {
int mbrf_pdhg_padded_sizes(int M, int N, int B, int *Mp, int *Np, int *Bp);
unsigned long long mbrf_pdhg_workspace_bytes(int Mp, int Np, int Bp);
int mbrf_pdhg_warm_start_device(const double *z_init, const double *y_init, const double *omega_init, double *omega_out);
int mbrf_pdhg_solve_device(const double *K, const double *KT, int Mp, int Np, int ldk, double *c, double *lo,
                           double *hi, double *bl, double *bu, const int *pair_i,
                           const int *pair_j, int npairs, double *rho, int Bp, int B,
                           double *obj_upper, const mbrf_pdhg_blocks *blocks, int max_iter, int check_every,
                           double eps_pr, double eps_dr, double eps_gap, double *z_out, double *y_out,
                           double *info_out, void *workspace, void *stream);
}

namespace mbrf {
namespace fir {

// K[i][j] = scale_i * amp_j * trig_j(w_i * kappa_j + phase_i)   (type 0: 1, 1: cos, 2: sin, 3: 0);
// K[i][tcol] = tcoef[i]
__global__ void build_matrix_kernel(const double *__restrict__ w, const double *__restrict__ phase,
                                    const double *__restrict__ scale, const double *__restrict__ tcoef, int M,
                                    const int *__restrict__ type, const double *__restrict__ kappa,
                                    const double *__restrict__ amp, int N, int tcol, double *__restrict__ K,
                                    int Mp, int ldk)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)Mp * ldk) return;
    const int i = (int)(idx / ldk), j = (int)(idx % ldk);
    double v = 0.0;
    if (i < M && j < N) {
        if (j == tcol) v = tcoef ? tcoef[i] : 0.0;
        else {
            const int t = type[j];
            const double ph = phase ? phase[i] : 0.0, sc = scale ? scale[i] : 1.0;
            if (t == 0) v = sc * amp[j];
            else if (t == 1) v = sc * amp[j] * cos(w[i] * kappa[j] + ph);
            else if (t == 2) v = sc * amp[j] * sin(w[i] * kappa[j] + ph);
        }
    }
    K[idx] = v;
}

__global__ void add_entries_kernel(double *__restrict__ K, int ldk, int nnz, const int *__restrict__ ti,
                                   const int *__restrict__ tj, const double *__restrict__ tv)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz) atomicAdd(&K[(size_t)ti[k] * ldk + tj[k]], tv[k]);
}

// cs2[j] = sum_i K[i][j]^2 : one block per 32 columns, threads stride the rows
__global__ void col_norm2_kernel(const double *__restrict__ K, int Mp, int ldk, double *__restrict__ cs2)
{
    __shared__ double sh[8][33];
    const int j = blockIdx.x * 32 + threadIdx.x;
    double s = 0.0;
    for (int i = threadIdx.y; i < Mp; i += 8) { const double v = K[(size_t)i * ldk + j]; s = fma(v, v, s); }
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0) {
        for (int r = 1; r < 8; ++r) s += sh[r][threadIdx.x];
        cs2[j] = s;
    }
}

// scale the columns of K in place and write the transposed copy KT [ldk x Mp] (32 x 32 smem transpose)
__global__ void scale_transpose_kernel(double *__restrict__ K, double *__restrict__ KT, int Mp, int ldk,
                                       const double *__restrict__ inv_cs)
{
    __shared__ double t[32][33];
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const size_t o = (size_t)(i0 + r) * ldk + j0 + threadIdx.x;
        const double v = K[o] * inv_cs[j0 + threadIdx.x];
        K[o] = v;
        t[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) KT[(size_t)(j0 + r) * Mp + i0 + threadIdx.x] = t[threadIdx.x][r];
}

struct Ctx {
    DeviceScratch dev;
    cudaStream_t stream = nullptr;
    int stream_device = -1;
};
static thread_local Ctx t_ctx;
// warm start / multiplier output of this host thread's next solve (host pointers, caller's units); consumed by that call
struct WarmHost { const double *z = nullptr, *y = nullptr, *omega = nullptr; double *y_out = nullptr, *omega_out = nullptr; };
static thread_local WarmHost t_warm_host;

}  // namespace fir
}  // namespace mbrf

using namespace mbrf;
using namespace mbrf::fir;

extern "C" {

/*
 * Solve B convex FIR design problems that share one frequency-sampled matrix.  HOST pointers.
 *   minimise c^T z  s.t.  lo <= K z <= hi,  bl <= z <= bu,  ||(z_pi, z_pj)|| <= rho
 * per-design arrays are [dim x B] row-major (design index fastest).  The solver scales the columns of K
 * to unit norm internally (pair members share a scale) and returns z in the caller's units.
 * Rows [simplex_row0, simplex_row0 + simplex_rows) add  simplex_w[b] * max_i (K z)_i  to design b's objective
 * (over the rows whose hi is 0; hi = +inf excludes a row from a design).
 * info: [B x 8] = status (1 solved, 2 infeasible, 3 iteration limit), iterations, objective, dual objective,
 *       max row violation, natural residual, rigorous lower bound on the optimum, max_i (K z)_i of the block.
 */
static int solve_impl(const double *w_row, const double *row_phase, const double *row_scale, const double *tcoef,
                      int M, const int *col_type, const double *col_kappa, const double *col_amp, int N, int tcol,
                      int nnz, const int *ti, const int *tj, const double *tv, const int *pair_i, const int *pair_j,
                      int npairs, const double *c, const double *lo, const double *hi, const double *bl,
                      const double *bu, const double *rho, int B, const double *obj_upper,
                      const mbrf_pdhg_blocks *blocks, int max_iter, int check_every, double eps_pr, double eps_dr,
                      double eps_gap, double *z_out, double *info_out, double *colscale_out)
{
    const WarmHost warm = t_warm_host;
    t_warm_host = WarmHost();
    if (int rc = require_device()) return rc;
    const bool timing = getenv("MBRF_TIMING") != nullptr;     // developer trace of the host-side phases (stderr)
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count(); };
    if (M <= 0 || N <= 0 || B <= 0 || !w_row || !col_type || !col_kappa || !col_amp || !c || !lo || !hi || !bl ||
        !bu || !z_out || !info_out || npairs < 0 || (npairs && (!pair_i || !pair_j || !rho)) || tcol >= N || nnz < 0 ||
        (nnz && (!ti || !tj || !tv))) {
        set_error("fir_pdhg_solve: bad arguments");
        return MBRF_EINVAL;
    }
    int Mp, Np, Bp;
    mbrf_pdhg_padded_sizes(M, N, B, &Mp, &Np, &Bp);
    const int ldk = Np;
    Ctx &cx = t_ctx;
    int dev = 0;
    MBRF_CUDA(cudaGetDevice(&dev));
    if (!cx.stream || cx.stream_device != dev) {
        if (cx.stream) { cudaSetDevice(cx.stream_device); cudaStreamDestroy(cx.stream); cudaSetDevice(dev); cx.stream = nullptr; }
        MBRF_CUDA(cudaStreamCreateWithFlags(&cx.stream, cudaStreamNonBlocking));
        cx.stream_device = dev;
    }
    cudaStream_t st = cx.stream;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t zn = (size_t)Np * Bp, yn = (size_t)Mp * Bp;
    const size_t bK = al((size_t)Mp * ldk * 8), bw = al((size_t)Mp * 8), bcol = al((size_t)Np * 8),
                 bz = al(zn * 8), by = al(yn * 8), bpair = al((size_t)(npairs > 0 ? npairs : 1) * 4),
                 brho = al((size_t)(npairs > 0 ? npairs : 1) * Bp * 8), binfo = al((size_t)Bp * 8 * 8),
                 bws = al(mbrf_pdhg_workspace_bytes(Mp, Np, Bp));
    const size_t bnz = al((size_t)(nnz > 0 ? nnz : 1) * 8);
    const size_t total = 3 * bnz + 2 * bw + 2 * al((size_t)Bp * 8) + 2 * bK + 2 * bw + 4 * bcol + 4 * bz /*c,bl,bu,zout*/ + 2 * by /*lo,hi*/ + 2 * bpair + brho +
                         binfo + 2 * al((size_t)Bp * 8) + bws + bz /*z init*/ + 2 * by /*y init, y out*/ + al((size_t)Bp * 8) /*group2 w*/;
    if (int rc = cx.dev.reserve(total)) return rc;
    char *d = (char *)cx.dev.ptr;
    auto take = [&](size_t b) { char *p = d; d += b; return p; };
    double *dK = (double *)take(bK), *dKT = (double *)take(bK), *dw = (double *)take(bw), *dt = (double *)take(bw);
    double *dph = (double *)take(bw), *dsc = (double *)take(bw);
    int *dti = (int *)take(bnz), *dtj = (int *)take(bnz);
    double *dtv = (double *)take(bnz), *dgw = (double *)take(al((size_t)Bp * 8)), *dlam = (double *)take(al((size_t)Bp * 8));
    int *dtype = (int *)take(bcol);
    double *dkap = (double *)take(bcol), *damp = (double *)take(bcol), *dcs = (double *)take(bcol);
    double *dc = (double *)take(bz), *dbl = (double *)take(bz), *dbu = (double *)take(bz), *dz = (double *)take(bz);
    double *dlo = (double *)take(by), *dhi = (double *)take(by);
    int *dpi = (int *)take(bpair), *dpj = (int *)take(bpair);
    double *drho = (double *)take(brho), *dinfo = (double *)take(binfo), *dupper = (double *)take(al((size_t)Bp * 8)),
           *dsw = (double *)take(al((size_t)Bp * 8));
    void *dws = take(bws);
    double *dzi = (double *)take(bz), *dyi = (double *)take(by), *dyo = (double *)take(by);
    double *dgw2 = (double *)take(al((size_t)Bp * 8));

    // ---- matrix: build, column norms, scale ----
    MBRF_CUDA(cudaMemcpyAsync(dw, w_row, (size_t)M * 8, cudaMemcpyHostToDevice, st));
    if (tcoef) MBRF_CUDA(cudaMemcpyAsync(dt, tcoef, (size_t)M * 8, cudaMemcpyHostToDevice, st));
    MBRF_CUDA(cudaMemcpyAsync(dtype, col_type, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    MBRF_CUDA(cudaMemcpyAsync(dkap, col_kappa, (size_t)N * 8, cudaMemcpyHostToDevice, st));
    MBRF_CUDA(cudaMemcpyAsync(damp, col_amp, (size_t)N * 8, cudaMemcpyHostToDevice, st));
    const long long nK = (long long)Mp * ldk;
    if (row_phase) MBRF_CUDA(cudaMemcpyAsync(dph, row_phase, (size_t)M * 8, cudaMemcpyHostToDevice, st));
    if (row_scale) MBRF_CUDA(cudaMemcpyAsync(dsc, row_scale, (size_t)M * 8, cudaMemcpyHostToDevice, st));
    build_matrix_kernel<<<(unsigned)((nK + 255) / 256), 256, 0, st>>>(dw, row_phase ? dph : nullptr, row_scale ? dsc : nullptr,
                                                                     tcoef ? dt : nullptr, M, dtype, dkap, damp, N, tcol,
                                                                     dK, Mp, ldk);
    MBRF_LAUNCH_CHECK();
    if (nnz) {
        for (int k = 0; k < nnz; ++k)
            if (ti[k] < 0 || ti[k] >= M || tj[k] < 0 || tj[k] >= N) { set_error("fir_pdhg_solve: entry %d out of range", k); return MBRF_EINVAL; }
        MBRF_CUDA(cudaMemcpyAsync(dti, ti, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
        MBRF_CUDA(cudaMemcpyAsync(dtj, tj, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
        MBRF_CUDA(cudaMemcpyAsync(dtv, tv, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
        add_entries_kernel<<<(nnz + 255) / 256, 256, 0, st>>>(dK, ldk, nnz, dti, dtj, dtv);
        MBRF_LAUNCH_CHECK();
    }
    col_norm2_kernel<<<ldk / 32, dim3(32, 8), 0, st>>>(dK, Mp, ldk, dcs);
    MBRF_LAUNCH_CHECK();
    std::vector<double> cs((size_t)Np), inv((size_t)Np);
    MBRF_CUDA(cudaMemcpyAsync(cs.data(), dcs, (size_t)Np * 8, cudaMemcpyDeviceToHost, st));
    MBRF_CUDA(cudaStreamSynchronize(st));
    for (int j = 0; j < Np; ++j) cs[j] = cs[j] > 0.0 ? sqrt(cs[j]) : 1.0;
    for (int q = 0; q < npairs; ++q) {  // a disk must stay a disk: both members share one scale
        const int i = pair_i[q], j = pair_j[q];
        if (i < 0 || i >= N || j < 0 || j >= N) { set_error("fir_pdhg_solve: pair %d out of range", q); return MBRF_EINVAL; }
        const double pm = sqrt(0.5 * (cs[i] * cs[i] + cs[j] * cs[j]));
        cs[i] = cs[j] = pm;
    }
    mbrf_pdhg_blocks bk;
    memset(&bk, 0, sizeof bk);
    if (blocks) bk = *blocks;
    if (bk.norm_coords > 0) {           // a ball must stay a ball: the norm-term coordinates share one scale
        if (bk.norm_coords > N || npairs > 0) { set_error("fir_pdhg_solve: norm term excludes pairs and needs norm_coords <= N"); return MBRF_EINVAL; }
        double m2 = 0.0;
        for (int j = 0; j < bk.norm_coords; ++j) m2 += cs[j] * cs[j];
        const double sm = sqrt(m2 / bk.norm_coords);
        for (int j = 0; j < bk.norm_coords; ++j) cs[j] = sm;
    }
    for (int j = 0; j < Np; ++j) inv[j] = 1.0 / cs[j];
    MBRF_CUDA(cudaMemcpyAsync(dcs, inv.data(), (size_t)Np * 8, cudaMemcpyHostToDevice, st));
    scale_transpose_kernel<<<dim3(ldk / 32, Mp / 32), dim3(32, 8), 0, st>>>(dK, dKT, Mp, ldk, dcs);
    MBRF_LAUNCH_CHECK();
    if (colscale_out) memcpy(colscale_out, cs.data(), (size_t)N * 8);

    // ---- per-design vectors: pad to [dim_p x Bp], apply the column scaling on the way ----
    const double INF = INFINITY;
    std::vector<double> h;
    auto upload_cols = [&](const double *src, double *dst, int mode) -> int {   // N-side arrays
        // mode 0: c / cs ; 1: bound * cs (padding: 0 / [0,0])
        h.assign(zn, 0.0);
        for (int j = 0; j < N; ++j)
            for (int b = 0; b < B; ++b) {
                const double v = src[(size_t)j * B + b];
                h[(size_t)j * Bp + b] = mode == 0 ? v / cs[j] : v * cs[j];
            }
        MBRF_CUDA(cudaMemcpyAsync(dst, h.data(), zn * 8, cudaMemcpyHostToDevice, st));
        MBRF_CUDA(cudaStreamSynchronize(st));
        return MBRF_OK;
    };
    if (int rc = upload_cols(c, dc, 0)) return rc;
    if (int rc = upload_cols(bl, dbl, 1)) return rc;
    if (int rc = upload_cols(bu, dbu, 1)) return rc;
    auto upload_rows = [&](const double *src, double *dst, double pad) -> int {
        h.assign(yn, pad);
        for (int i = 0; i < M; ++i)
            for (int b = 0; b < B; ++b) h[(size_t)i * Bp + b] = src[(size_t)i * B + b];
        MBRF_CUDA(cudaMemcpyAsync(dst, h.data(), yn * 8, cudaMemcpyHostToDevice, st));
        MBRF_CUDA(cudaStreamSynchronize(st));
        return MBRF_OK;
    };
    if (int rc = upload_rows(lo, dlo, -INF)) return rc;
    if (int rc = upload_rows(hi, dhi, INF)) return rc;
    if (npairs) {
        MBRF_CUDA(cudaMemcpyAsync(dpi, pair_i, (size_t)npairs * 4, cudaMemcpyHostToDevice, st));
        MBRF_CUDA(cudaMemcpyAsync(dpj, pair_j, (size_t)npairs * 4, cudaMemcpyHostToDevice, st));
        h.assign((size_t)npairs * Bp, 0.0);
        for (int q = 0; q < npairs; ++q)
            for (int b = 0; b < B; ++b) h[(size_t)q * Bp + b] = rho[(size_t)q * B + b] * cs[pair_i[q]];
        MBRF_CUDA(cudaMemcpyAsync(drho, h.data(), (size_t)npairs * Bp * 8, cudaMemcpyHostToDevice, st));
        MBRF_CUDA(cudaStreamSynchronize(st));
    }
    if (obj_upper) {
        h.assign((size_t)Bp, INF);
        for (int b = 0; b < B; ++b) h[b] = obj_upper[b];
        MBRF_CUDA(cudaMemcpyAsync(dupper, h.data(), (size_t)Bp * 8, cudaMemcpyHostToDevice, st));
        MBRF_CUDA(cudaStreamSynchronize(st));
    }

    auto upload_w = [&](const double *src, double *dst, double scale) -> int {
        h.assign((size_t)Bp, 0.0);
        for (int b = 0; b < B; ++b) h[b] = src[b] * scale;
        MBRF_CUDA(cudaMemcpyAsync(dst, h.data(), (size_t)Bp * 8, cudaMemcpyHostToDevice, st));
        MBRF_CUDA(cudaStreamSynchronize(st));
        return MBRF_OK;
    };
    mbrf_pdhg_blocks dbk = bk;
    if (bk.simplex_rows > 0) {
        if (!bk.simplex_w || bk.simplex_row0 < 0 || bk.simplex_row0 + bk.simplex_rows > M) { set_error("fir_pdhg_solve: bad simplex block"); return MBRF_EINVAL; }
        if (int rc = upload_w(bk.simplex_w, dsw, 1.0)) return rc;
        dbk.simplex_w = dsw;
    }
    if (bk.group_pairs > 0) {
        if (!bk.group_w || bk.group_row0 < 0 || bk.group_row0 + 2 * bk.group_pairs > M) { set_error("fir_pdhg_solve: bad group block"); return MBRF_EINVAL; }
        if (int rc = upload_w(bk.group_w, dgw, 1.0)) return rc;
        dbk.group_w = dgw;
    }
    if (bk.group2_pairs > 0) {
        if (!bk.group2_w || bk.group2_row0 < 0 || bk.group2_row0 + 2 * bk.group2_pairs > M) { set_error("fir_pdhg_solve: bad centred group block"); return MBRF_EINVAL; }
        if (int rc = upload_w(bk.group2_w, dgw2, 1.0)) return rc;
        dbk.group2_w = dgw2;
    }
    if (bk.disk_pairs > 0 && (bk.disk_row0 < 0 || bk.disk_row0 + 2 * bk.disk_pairs > M)) { set_error("fir_pdhg_solve: bad disk block"); return MBRF_EINVAL; }
    if (bk.norm_coords > 0) {
        if (!bk.norm_w) { set_error("fir_pdhg_solve: norm_w missing"); return MBRF_EINVAL; }
        if (int rc = upload_w(bk.norm_w, dlam, 1.0 / cs[0])) return rc;     // ||x|| = ||z|| / scale
        dbk.norm_w = dlam;
    }

    std::vector<double> om_in, om_out;
    if (warm.z || warm.y || warm.omega || warm.omega_out) {
        if (warm.z) {
            h.assign(zn, 0.0);
            for (int j = 0; j < N; ++j)
                for (int b = 0; b < B; ++b) h[(size_t)j * Bp + b] = warm.z[(size_t)j * B + b] * cs[j];
            MBRF_CUDA(cudaMemcpyAsync(dzi, h.data(), zn * 8, cudaMemcpyHostToDevice, st));
            MBRF_CUDA(cudaStreamSynchronize(st));
        }
        if (warm.y) {
            h.assign(yn, 0.0);
            for (int i = 0; i < M; ++i)
                for (int b = 0; b < B; ++b) h[(size_t)i * Bp + b] = warm.y[(size_t)i * B + b];
            MBRF_CUDA(cudaMemcpyAsync(dyi, h.data(), yn * 8, cudaMemcpyHostToDevice, st));
            MBRF_CUDA(cudaStreamSynchronize(st));
        }
        if (warm.omega) { om_in.assign((size_t)Bp, 1.0); for (int b = 0; b < B; ++b) om_in[b] = warm.omega[b]; }
        if (warm.omega_out) om_out.assign((size_t)Bp, 1.0);
        mbrf_pdhg_warm_start_device(warm.z ? dzi : nullptr, warm.y ? dyi : nullptr, warm.omega ? om_in.data() : nullptr,
                                    warm.omega_out ? om_out.data() : nullptr);
    }
    const double t_setup = since();
    int rc = mbrf_pdhg_solve_device(dK, dKT, Mp, Np, ldk, dc, dlo, dhi, dbl, dbu, npairs ? dpi : nullptr,
                                    npairs ? dpj : nullptr, npairs, npairs ? drho : nullptr, Bp, B,
                                    obj_upper ? dupper : nullptr, &dbk, max_iter, check_every, eps_pr, eps_dr, eps_gap,
                                    dz, warm.y_out ? dyo : nullptr, dinfo, dws, st);
    if (rc) return rc;
    const double t_solved = since();
    if (warm.y_out) {
        h.assign(yn, 0.0);
        MBRF_CUDA(cudaMemcpyAsync(h.data(), dyo, yn * 8, cudaMemcpyDeviceToHost, st));
        MBRF_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < M; ++i)
            for (int b = 0; b < B; ++b) warm.y_out[(size_t)i * B + b] = h[(size_t)i * Bp + b];
    }
    if (warm.omega_out)
        for (int b = 0; b < B; ++b) warm.omega_out[b] = om_out[b];
    h.assign(zn, 0.0);
    std::vector<double> info((size_t)Bp * 8);
    MBRF_CUDA(cudaMemcpyAsync(h.data(), dz, zn * 8, cudaMemcpyDeviceToHost, st));
    MBRF_CUDA(cudaMemcpyAsync(info.data(), dinfo, (size_t)Bp * 64, cudaMemcpyDeviceToHost, st));
    MBRF_CUDA(cudaStreamSynchronize(st));
    for (int j = 0; j < N; ++j)
        for (int b = 0; b < B; ++b) z_out[(size_t)j * B + b] = h[(size_t)j * Bp + b] / cs[j];
    memcpy(info_out, info.data(), (size_t)B * 64);
    if (timing)
        fprintf(stderr, "fir_pdhg_solve M=%d N=%d B=%d: setup+upload %.3f s, solve %.3f s, download %.3f s\n", M, N, B, t_setup,
                t_solved - t_setup, since() - t_solved);
    return MBRF_OK;
}

/* Warm start of the next mbrf_fir_pdhg_solve / _solve2 call made by this host thread (consumed by that call).  Host pointers
 * in the caller's units that must stay valid until that call returns; any may be NULL:
 *   z_init [N x B], y_init [M x B]: starting iterate and multipliers;  omega_init [B]: primal weights (default 1);
 *   y_out [M x B], omega_out [B]: receive the final multipliers / primal weights (to warm-start a neighbouring design). */
extern "C" int mbrf_fir_pdhg_warm_start(const double *z_init, const double *y_init, const double *omega_init, double *y_out,
                                        double *omega_out)
{
    t_warm_host.z = z_init; t_warm_host.y = y_init; t_warm_host.omega = omega_init;
    t_warm_host.y_out = y_out; t_warm_host.omega_out = omega_out;
    return MBRF_OK;
}

extern "C" int mbrf_fir_pdhg_solve2(const double *w_row, const double *row_phase, const double *row_scale, int M,
                                    const int *col_type, const double *col_kappa, const double *col_amp, int N, int nnz,
                                    const int *ti, const int *tj, const double *tv, const int *pair_i,
                                    const int *pair_j, int npairs, const double *c, const double *lo, const double *hi,
                                    const double *bl, const double *bu, const double *rho, int B,
                                    const double *obj_upper, const mbrf_pdhg_blocks *blocks, int max_iter,
                                    int check_every, double eps_pr, double eps_dr, double eps_gap, double *z_out,
                                    double *info_out, double *colscale_out)
{
    return solve_impl(w_row, row_phase, row_scale, nullptr, M, col_type, col_kappa, col_amp, N, -1, nnz, ti, tj, tv, pair_i,
                      pair_j, npairs, c, lo, hi, bl, bu, rho, B, obj_upper, blocks, max_iter, check_every, eps_pr, eps_dr,
                      eps_gap, z_out, info_out, colscale_out);
}

extern "C" int mbrf_fir_pdhg_solve(const double *w_row, const double *tcoef, int M, const int *col_type,
                                   const double *col_kappa, const double *col_amp, int N, int tcol, const int *pair_i,
                                   const int *pair_j, int npairs, const double *c, const double *lo, const double *hi,
                                   const double *bl, const double *bu, const double *rho, int B,
                                   const double *obj_upper, int simplex_row0, int simplex_rows,
                                   const double *simplex_w, int max_iter, int check_every, double eps_pr, double eps_dr,
                                   double eps_gap, double *z_out, double *info_out, double *colscale_out)
{
    mbrf_pdhg_blocks bk;
    memset(&bk, 0, sizeof bk);
    bk.simplex_row0 = simplex_row0;
    bk.simplex_rows = simplex_rows;
    bk.simplex_w = const_cast<double *>(simplex_w);
    return solve_impl(w_row, nullptr, nullptr, tcoef, M, col_type, col_kappa, col_amp, N, tcol, 0, nullptr, nullptr, nullptr,
                      pair_i, pair_j, npairs, c, lo, hi, bl, bu, rho, B, obj_upper, &bk, max_iter, check_every, eps_pr, eps_dr,
                      eps_gap, z_out, info_out, colscale_out);
}

}  // extern "C"
