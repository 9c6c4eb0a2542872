/*
 * MEX gateway of the convex FIR design solver:
 *
 *   [z, info] = fir_pdhg_mex(w_row, tcoef, col_type, col_kappa, col_amp, tcol, pair_i, pair_j, ...
 *                            c, lo, hi, bl, bu, rho, obj_upper, opts, simplex)
 *
 * One call solves B problems  min c'z  s.t. lo <= K z <= hi, bl <= z <= bu, ||(z_pi,z_pj)|| <= rho  that
 * share the frequency-sampled matrix K described by (w_row, tcoef, col_*, tcol) — see mbrf_fir_pdhg_solve in
 * include/mbrf.h.  It is what the replacement fir_ap_cvx.m / fir_linprog.m below call instead of
 * `cvx_begin ... cvx_end` (fir_ap_cvx.m:160-169) and `linprog` (ss/fir_linprog.m:246-252).
 *   per-design arrays are dim-by-B MATLAB matrices (column-major = the C ABI's "design index slowest"; they are
 *   transposed here), index vectors are 1-based doubles, tcol = 0 means "no extra column",
 *   opts = [max_iter check_every eps_pr eps_dr eps_gap];  simplex = {} or [first_row n_rows w_1 .. w_B]: rows
 *   first_row .. first_row+n_rows-1 (1-based) add w_b*max_i (K z)_i to design b's objective (see mbrf.h).
 *   z is N-by-B, info is 8-by-B (status 1 solved / 2 infeasible / 3 iteration limit, iterations, objective,
 *   dual objective, max violation, residual, lower bound, primal weight).
 */
#include "mex.h"
#include "mbrf.h"
#include <stdlib.h>

static int numel(const mxArray *a) { return (int)(mxGetM(a) * mxGetN(a)); }

/* MATLAB dim-by-B (column-major) -> C ABI [dim x B] row-major */
static double *to_rows(const mxArray *a, int dim, int B)
{
    const double *s = mxGetPr(a);
    double *d = (double *)malloc(sizeof(double) * (size_t)(dim > 0 ? dim : 1) * B);
    int i, b;
    for (b = 0; b < B; b++)
        for (i = 0; i < dim; i++) d[(size_t)i * B + b] = s[(size_t)b * dim + i];
    return d;
}

static int *to_int0(const mxArray *a, int n, int offset)
{
    const double *s = mxGetPr(a);
    int *d = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1)), i;
    for (i = 0; i < n; i++) d[i] = (int)s[i] - offset;
    return d;
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    int M, N, B, npairs, tcol, rc, i, b;
    int *ctype, *pi, *pj;
    double *c, *lo, *hi, *bl, *bu, *rho = NULL, *z, *info;
    const double *opts, *upper = NULL;

    int srow0 = 0, srows = 0;
    const double *sw = NULL;
    double *zi = NULL, *yi = NULL, *yo = NULL, *oo = NULL;
    const double *oi = NULL;
    if (nrhs < 16 || nrhs > 20 || nlhs > 4)
        mexErrMsgTxt("Usage: [z, info, y, omega] = fir_pdhg_mex(w_row,tcoef,col_type,col_kappa,col_amp,tcol,pair_i,pair_j,c,lo,hi,bl,bu,rho,obj_upper,opts,simplex,z_init,y_init,omega_init)");
    M = numel(prhs[0]);
    N = numel(prhs[2]);
    B = (int)mxGetN(prhs[8]);
    npairs = numel(prhs[6]);
    tcol = (int)(*mxGetPr(prhs[5])) - 1;
    if ((int)mxGetM(prhs[8]) != N || (int)mxGetM(prhs[9]) != M || (int)mxGetM(prhs[10]) != M || numel(prhs[15]) < 5)
        mexErrMsgTxt("fir_pdhg_mex: inconsistent sizes");
    ctype = to_int0(prhs[2], N, 0);
    pi = to_int0(prhs[6], npairs, 1);
    pj = to_int0(prhs[7], npairs, 1);
    c = to_rows(prhs[8], N, B); lo = to_rows(prhs[9], M, B); hi = to_rows(prhs[10], M, B);
    bl = to_rows(prhs[11], N, B); bu = to_rows(prhs[12], N, B);
    if (npairs) rho = to_rows(prhs[13], npairs, B);
    if (numel(prhs[14]) == B) upper = mxGetPr(prhs[14]);
    opts = mxGetPr(prhs[15]);
    if (nrhs == 17 && numel(prhs[16]) == 2 + B) {
        srow0 = (int)mxGetPr(prhs[16])[0] - 1;
        srows = (int)mxGetPr(prhs[16])[1];
        sw = mxGetPr(prhs[16]) + 2;
    }
    z = (double *)malloc(sizeof(double) * (size_t)N * B);
    info = (double *)malloc(sizeof(double) * 8 * (size_t)B);
    /* warm start from a neighbouring design (optional inputs 18-20) and multipliers / primal weights out (outputs 3-4) */
    if (nrhs > 17 && (int)mxGetM(prhs[17]) == N && (int)mxGetN(prhs[17]) == B) zi = to_rows(prhs[17], N, B);
    if (nrhs > 18 && (int)mxGetM(prhs[18]) == M && (int)mxGetN(prhs[18]) == B) yi = to_rows(prhs[18], M, B);
    if (nrhs > 19 && numel(prhs[19]) == B) oi = mxGetPr(prhs[19]);
    if (nlhs > 2) yo = (double *)malloc(sizeof(double) * (size_t)M * B);
    if (nlhs > 3) oo = (double *)malloc(sizeof(double) * (size_t)B);
    if (zi || yi || oi || yo || oo) mbrf_fir_pdhg_warm_start(zi, yi, oi, yo, oo);

    rc = mbrf_fir_pdhg_solve(mxGetPr(prhs[0]), numel(prhs[1]) == M ? mxGetPr(prhs[1]) : NULL, M, ctype,
                             mxGetPr(prhs[3]), mxGetPr(prhs[4]), N, tcol, npairs ? pi : NULL, npairs ? pj : NULL, npairs,
                             c, lo, hi, bl, bu, rho, B, upper, srow0, srows, sw, (int)opts[0], (int)opts[1], opts[2], opts[3],
                             opts[4],
                             z, info, NULL);
    if (rc == MBRF_OK) {
        plhs[0] = mxCreateDoubleMatrix((size_t)N, (size_t)B, mxREAL);
        for (b = 0; b < B; b++)
            for (i = 0; i < N; i++) mxGetPr(plhs[0])[(size_t)b * N + i] = z[(size_t)i * B + b];
        if (nlhs > 1) {
            plhs[1] = mxCreateDoubleMatrix(8, (size_t)B, mxREAL);
            for (i = 0; i < 8 * B; i++) mxGetPr(plhs[1])[i] = info[i];
        }
        if (nlhs > 2) {
            plhs[2] = mxCreateDoubleMatrix((size_t)M, (size_t)B, mxREAL);
            for (b = 0; b < B; b++)
                for (i = 0; i < M; i++) mxGetPr(plhs[2])[(size_t)b * M + i] = yo[(size_t)i * B + b];
        }
        if (nlhs > 3) {
            plhs[3] = mxCreateDoubleMatrix((size_t)B, 1, mxREAL);
            for (b = 0; b < B; b++) mxGetPr(plhs[3])[b] = oo[b];
        }
    }
    free(zi); free(yi); free(yo); free(oo);
    free(ctype); free(pi); free(pj); free(c); free(lo); free(hi); free(bl); free(bu); free(rho); free(z); free(info);
    if (rc != MBRF_OK) mexErrMsgTxt(mbrf_last_error());
}
