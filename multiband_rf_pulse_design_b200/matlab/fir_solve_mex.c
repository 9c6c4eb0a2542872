/*
 * MEX gateway of the convex FIR design solvers, general form (every problem fir_ap_cvx.m, fir_qp_cvx.m and ss/fir_linprog.m pose):
 *
 *   [z, info] = fir_solve_mex(method, w_row, row_phase, row_scale, col_type, col_kappa, col_amp, entries, ...
 *                             pair_i, pair_j, c, lo, hi, bl, bu, rho, obj_upper, opts, blocks, block_w)
 *
 *   method    0 or 'pdhg': first-order solver, mbrf_fir_pdhg_solve2 (all block types: fir_qp_cvx.m needs them)
 *             1 or 'ipm' : interior-point solver, mbrf_fir_ipm_solve (interval rows, stop block, peak cones: fir_ap_cvx.m,
 *                          ss/fir_linprog.m) -- row_phase, row_scale, entries must be [] and only the stop block may be given
 *   w_row     M row frequencies;  row_phase, row_scale: M-vectors or [] (phase 0, scale 1)
 *   col_*     N-vectors: type 0 constant / 1 cos / 2 sin, frequency, amplitude           (K of include/mbrf.h)
 *   entries   nnz-by-3 [row col value], 1-based, added to K; or []
 *   pair_i/j  1-based variable pairs with ||(z_i, z_j)|| <= rho (npairs-by-B); or []
 *   c         N-by-B;  lo, hi: M-by-B;  bl, bu: N-by-B or [] (no bounds);  obj_upper: B-vector or []
 *   opts      pdhg: [max_iter check_every eps_pr eps_dr eps_gap];   ipm: [max_iter feastol reltol abstol] (0 = default)
 *   blocks    [] or [simplex_row0 simplex_rows disk_row0 disk_pairs group_row0 group_pairs norm_coords group2_row0 group2_pairs],
 *             rows 1-based (mbrf_pdhg_blocks of include/mbrf.h); block_w: 4-by-B [simplex_w; group_w; norm_w; group2_w]
 *   z         N-by-B;  info: 8-by-B (status 1 solved / 2 infeasible (certificate) / 3 iteration limit, iterations, objective,
 *             dual objective, max violation, residual, lower bound, ripple_stop)
 *
 * Per-design arrays are dim-by-B MATLAB matrices (column-major) and are transposed to the C ABI's [dim x B] row-major here.
 */
#include "mex.h"
#include "mbrf.h"
#include <stdlib.h>
#include <string.h>

static int numel(const mxArray *a) { return (int)(mxGetM(a) * mxGetN(a)); }

static double *to_rows(const mxArray *a, int dim, int B)
{
    const double *s = mxGetPr(a);
    double *d = (double *)malloc(sizeof(double) * (size_t)(dim > 0 ? dim : 1) * (size_t)B);
    int i, b;
    for (b = 0; b < B; b++)
        for (i = 0; i < dim; i++) d[(size_t)i * B + b] = s[(size_t)b * dim + i];
    return d;
}

static int *to_int(const mxArray *a, int n, int offset)
{
    const double *s = mxGetPr(a);
    int *d = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1)), i;
    for (i = 0; i < n; i++) d[i] = (int)s[i] - offset;
    return d;
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    int method, M, N, B, npairs, nnz, rc, i, b;
    int *ctype = NULL, *pi = NULL, *pj = NULL, *ti = NULL, *tj = NULL;
    double *c, *lo, *hi, *bl = NULL, *bu = NULL, *rho = NULL, *z, *info, *tv = NULL, *bw[4] = {NULL, NULL, NULL, NULL};
    const double *opts, *upper = NULL, *phase = NULL, *scale = NULL;
    mbrf_pdhg_blocks blk;

    if (nrhs != 20 || nlhs > 2)
        mexErrMsgTxt("Usage: [z, info] = fir_solve_mex(method,w_row,row_phase,row_scale,col_type,col_kappa,col_amp,entries,pair_i,pair_j,c,lo,hi,bl,bu,rho,obj_upper,opts,blocks,block_w)");
    /* method: number, or a char array whose first letter is 'i' (ipm) / 'p' (pdhg) */
    method = (int)mxGetScalar(prhs[0]);
    if (method == 'i') method = 1;
    else if (method == 'p') method = 0;
    if (method != 0 && method != 1) mexErrMsgTxt("fir_solve_mex: method must be 0 / 'pdhg' or 1 / 'ipm'");
    M = numel(prhs[1]);
    N = numel(prhs[4]);
    B = (int)mxGetN(prhs[10]);
    npairs = numel(prhs[8]);
    nnz = (int)mxGetM(prhs[7]);
    if ((int)mxGetM(prhs[10]) != N || (int)mxGetM(prhs[11]) != M || (int)mxGetM(prhs[12]) != M || (int)mxGetN(prhs[11]) != B ||
        (int)mxGetN(prhs[12]) != B || numel(prhs[5]) != N || numel(prhs[6]) != N || numel(prhs[9]) != npairs ||
        (nnz && (int)mxGetN(prhs[7]) != 3) || numel(prhs[17]) < (method ? 4 : 5) ||
        (numel(prhs[13]) && ((int)mxGetM(prhs[13]) != N || (int)mxGetN(prhs[13]) != B)) ||
        (numel(prhs[14]) && ((int)mxGetM(prhs[14]) != N || (int)mxGetN(prhs[14]) != B)) ||
        (npairs && ((int)mxGetM(prhs[15]) != npairs || (int)mxGetN(prhs[15]) != B)))
        mexErrMsgTxt("fir_solve_mex: inconsistent sizes");
    if (numel(prhs[2]) == M) phase = mxGetPr(prhs[2]);
    if (numel(prhs[3]) == M) scale = mxGetPr(prhs[3]);
    memset(&blk, 0, sizeof blk);
    if (numel(prhs[18]) == 9) {
        const double *q = mxGetPr(prhs[18]);
        if ((int)mxGetM(prhs[19]) != 4 || (int)mxGetN(prhs[19]) != B) mexErrMsgTxt("fir_solve_mex: block_w must be 4-by-B");
        for (i = 0; i < 4; i++) {
            bw[i] = (double *)malloc(sizeof(double) * (size_t)B);
            for (b = 0; b < B; b++) bw[i][b] = mxGetPr(prhs[19])[(size_t)b * 4 + i];
        }
        blk.simplex_row0 = (int)q[0] - 1; blk.simplex_rows = (int)q[1]; blk.simplex_w = bw[0];
        blk.disk_row0 = (int)q[2] - 1; blk.disk_pairs = (int)q[3];
        blk.group_row0 = (int)q[4] - 1; blk.group_pairs = (int)q[5]; blk.group_w = bw[1];
        blk.norm_coords = (int)q[6]; blk.norm_w = bw[2];
        blk.group2_row0 = (int)q[7] - 1; blk.group2_pairs = (int)q[8]; blk.group2_w = bw[3];
    } else if (numel(prhs[18]) != 0)
        mexErrMsgTxt("fir_solve_mex: blocks must be [] or a 9-vector");
    if (method == 1 && (phase || scale || nnz || blk.disk_pairs || blk.group_pairs || blk.norm_coords || blk.group2_pairs))
        mexErrMsgTxt("fir_solve_mex: the interior-point solver takes interval rows, the stop block and peak cones only");

    ctype = to_int(prhs[4], N, 0);
    pi = to_int(prhs[8], npairs, 1);
    pj = to_int(prhs[9], npairs, 1);
    if (nnz) {
        const double *e = mxGetPr(prhs[7]);
        ti = (int *)malloc(sizeof(int) * (size_t)nnz); tj = (int *)malloc(sizeof(int) * (size_t)nnz);
        tv = (double *)malloc(sizeof(double) * (size_t)nnz);
        for (i = 0; i < nnz; i++) { ti[i] = (int)e[i] - 1; tj[i] = (int)e[nnz + i] - 1; tv[i] = e[2 * (size_t)nnz + i]; }
    }
    c = to_rows(prhs[10], N, B); lo = to_rows(prhs[11], M, B); hi = to_rows(prhs[12], M, B);
    if (numel(prhs[13])) bl = to_rows(prhs[13], N, B);
    if (numel(prhs[14])) bu = to_rows(prhs[14], N, B);
    if (method == 0 && (!bl || !bu)) {          /* the first-order solver wants explicit (possibly infinite) bounds */
        if (!bl) { bl = (double *)malloc(sizeof(double) * (size_t)N * B); for (i = 0; i < N * B; i++) bl[i] = -1.0 / 0.0; }
        if (!bu) { bu = (double *)malloc(sizeof(double) * (size_t)N * B); for (i = 0; i < N * B; i++) bu[i] = 1.0 / 0.0; }
    }
    if (npairs) rho = to_rows(prhs[15], npairs, B);
    if (numel(prhs[16]) == B) upper = mxGetPr(prhs[16]);
    opts = mxGetPr(prhs[17]);
    z = (double *)malloc(sizeof(double) * (size_t)N * B);
    info = (double *)malloc(sizeof(double) * 8 * (size_t)B);

    if (method == 1)
        rc = mbrf_fir_ipm_solve(mxGetPr(prhs[1]), M, ctype, mxGetPr(prhs[5]), mxGetPr(prhs[6]), N, npairs ? pi : NULL,
                                npairs ? pj : NULL, npairs, c, lo, hi, bl, bu, rho, B, blk.simplex_rows ? blk.simplex_row0 : 0,
                                blk.simplex_rows, blk.simplex_w, (int)opts[0], opts[1], opts[2], opts[3], z, info);
    else
        rc = mbrf_fir_pdhg_solve2(mxGetPr(prhs[1]), phase, scale, M, ctype, mxGetPr(prhs[5]), mxGetPr(prhs[6]), N, nnz, ti, tj, tv,
                                  npairs ? pi : NULL, npairs ? pj : NULL, npairs, c, lo, hi, bl, bu, rho, B, upper, &blk,
                                  (int)opts[0], (int)opts[1], opts[2], opts[3], opts[4], z, info, NULL);
    if (rc == MBRF_OK) {
        plhs[0] = mxCreateDoubleMatrix((size_t)N, (size_t)B, mxREAL);
        for (b = 0; b < B; b++)
            for (i = 0; i < N; i++) mxGetPr(plhs[0])[(size_t)b * N + i] = z[(size_t)i * B + b];
        if (nlhs > 1) {
            plhs[1] = mxCreateDoubleMatrix(8, (size_t)B, mxREAL);
            for (i = 0; i < 8 * B; i++) mxGetPr(plhs[1])[i] = info[i];
        }
    }
    for (i = 0; i < 4; i++) free(bw[i]);
    free(ctype); free(pi); free(pj); free(ti); free(tj); free(tv);
    free(c); free(lo); free(hi); free(bl); free(bu); free(rho); free(z); free(info);
    if (rc != MBRF_OK) mexErrMsgTxt(mbrf_last_error());
}
