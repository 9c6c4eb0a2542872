/*
 * MEX gateway of mbrf_fir_ap_solve: fir_ap_cvx (fir_ap_cvx.m:44-202) for a batch of designs of one order as ONE call --
 * specification in, minimum-phase taps out; problem assembly, interior-point solve and fmp2 all happen on the GPU.
 *
 *   [H, info, X, rows] = fir_ap_mex(n, F, A, D, obj, Peak)
 *
 *   F     2*nband-by-B band edges (fractions of pi, one design per column);  A 2*nband-by-B (or one column, shared);
 *   D     nband-by-B (or one column);  obj, Peak: B-vectors (or scalars)
 *   H     n-by-B complex taps (columns of designs with info(1,b) ~= 1 are meaningless);  info 8-by-B (status 1 solved /
 *         2 infeasible (certificate) / 3 iteration limit, iterations, objective, dual objective, max violation, residual,
 *         bound, ripple_stop);  X (2n-1)-by-B solutions;  rows = [grid rows of the union, rows of the stop block]
 *
 * MATLAB's column-major dim-by-B matrices are the C ABI's row-major [B x dim] arrays: nothing is transposed.
 * Build:  mex -output fir_ap_mex fir_ap_mex.c -I<repo>/include -L<pkg> -lmbrf
 */
#include "mex.h"
#include "mbrf.h"
#include <stdlib.h>

static double *expand(const mxArray *a, int dim, int B, const char *what)
{
    const int m = (int)mxGetM(a), k = (int)mxGetN(a), total = m * k;
    const double *s = mxGetPr(a);
    double *d;
    int b, i;
    if (!(total == dim * B || total == dim)) mexErrMsgTxt(what);
    d = (double *)malloc(sizeof(double) * (size_t)dim * (size_t)B);
    for (b = 0; b < B; b++)
        for (i = 0; i < dim; i++) d[(size_t)b * dim + i] = total == dim ? s[i] : s[(size_t)b * dim + i];
    return d;
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    int n, B, ne, nband, rc, rows[2] = {0, 0};
    double *a, *d, *obj, *peak;

    if (nrhs != 6 || nlhs > 4) mexErrMsgTxt("Usage: [H, info, X, rows] = fir_ap_mex(n, F, A, D, obj, Peak)");
    n = (int)mxGetScalar(prhs[0]);
    ne = (int)mxGetM(prhs[1]);
    B = (int)mxGetN(prhs[1]);
    if (ne == 1) { ne = B; B = 1; }                       /* a row vector is one design */
    if (n < 2 || ne < 2 || ne % 2 || B < 1) mexErrMsgTxt("fir_ap_mex: F must be 2*nband-by-B");
    nband = ne / 2;
    a = expand(prhs[2], ne, B, "fir_ap_mex: A must be 2*nband-by-B or one column");
    d = expand(prhs[3], nband, B, "fir_ap_mex: D must be nband-by-B or one column");
    obj = expand(prhs[4], 1, B, "fir_ap_mex: obj must be a scalar or a B-vector");
    peak = expand(prhs[5], 1, B, "fir_ap_mex: Peak must be a scalar or a B-vector");
    plhs[0] = mxCreateDoubleMatrix((size_t)n, (size_t)B, mxCOMPLEX);
    plhs[1] = mxCreateDoubleMatrix(8, (size_t)B, mxREAL);
    plhs[2] = mxCreateDoubleMatrix((size_t)(2 * n - 1), (size_t)B, mxREAL);
    rc = mbrf_fir_ap_solve(n, nband, mxGetPr(prhs[1]), a, d, obj, peak, B, 15, 0, 0.0, 0.0, 0.0, mxGetPr(plhs[2]), mxGetPr(plhs[0]),
                           mxGetPi(plhs[0]), mxGetPr(plhs[1]), rows);
    free(a); free(d); free(obj); free(peak);
    if (rc != MBRF_OK) mexErrMsgTxt(mbrf_last_error());
    if (nlhs > 3) { plhs[3] = mxCreateDoubleMatrix(1, 2, mxREAL); mxGetPr(plhs[3])[0] = rows[0]; mxGetPr(plhs[3])[1] = rows[1]; }
}
