/*
 * MEX gateways of the batched post-processing stages (one file, the stage chosen at compile time):
 *
 *   -DMBRF_STAGE=1   aca = b2a(bc)         drop-in for rf_tools/b2a.m:13-28      (mbrf_b2a_batch)
 *   -DMBRF_STAGE=2   rf  = ab2rf(ac, bc)   drop-in for rf_tools/ab2rf.m:12-26    (mbrf_ab2rf_batch)
 *   -DMBRF_STAGE=3   hmp = fmp2(h)         the local function fmp2 of fir_ap_cvx.m:262-283 (mbrf_fmp2_batch)
 *
 * A vector argument is one polynomial, as in the reference; an n-by-B matrix is a batch of B column polynomials
 * (MATLAB's column-major n x B is the C ABI's row-major [B x n]), and the result has the same shape (fmp2: (n+1)/2 rows).
 * Complex data uses the split real / imaginary planes of the pre-R2018a API (mxGetPr / mxGetPi), which is also Octave's.
 *
 * Build:  mex -DMBRF_STAGE=1 -output b2a islr_mex.c -I<repo>/include -L<pkg> -lmbrf      (likewise ab2rf, fmp2)
 */
#include "mex.h"
#include "mbrf.h"

#ifndef MBRF_STAGE
#define MBRF_STAGE 1
#endif

static void shape(const mxArray *a, int *n, int *B, int *row)
{
    const int M = (int)mxGetM(a), N = (int)mxGetN(a);
    *row = (M == 1);                      /* a row vector is one polynomial, like a column vector */
    if (M == 1 || N == 1) { *n = M * N; *B = 1; }
    else { *n = M; *B = N; }
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    int n, B, row, rc;
#if MBRF_STAGE == 2
    int n2, B2, row2;
    if (nrhs != 2 || nlhs > 1) mexErrMsgTxt("Usage: rf = ab2rf(ac, bc)");
    shape(prhs[0], &n, &B, &row);
    shape(prhs[1], &n2, &B2, &row2);
    if (n != n2 || B != B2) mexErrMsgTxt("ab2rf: ac and bc must have the same size");
    plhs[0] = row ? mxCreateDoubleMatrix(1, (size_t)n, mxCOMPLEX) : mxCreateDoubleMatrix((size_t)n, (size_t)B, mxCOMPLEX);
    rc = mbrf_ab2rf_batch(mxGetPr(prhs[0]), mxGetPi(prhs[0]), mxGetPr(prhs[1]), mxGetPi(prhs[1]), n, B, mxGetPr(plhs[0]),
                          mxGetPi(plhs[0]));
#elif MBRF_STAGE == 3
    if (nrhs != 1 || nlhs > 1) mexErrMsgTxt("Usage: hmp = fmp2(h)");
    shape(prhs[0], &n, &B, &row);
    if (n % 2 == 0) mexErrMsgTxt("filter length must be odd");           /* fir_ap_cvx.m:265-268 */
    plhs[0] = row || B == 1 ? mxCreateDoubleMatrix(1, (size_t)((n + 1) / 2), mxCOMPLEX)   /* fmp2 returns a row (:263) */
                            : mxCreateDoubleMatrix((size_t)((n + 1) / 2), (size_t)B, mxCOMPLEX);
    rc = mbrf_fmp2_batch(mxGetPr(prhs[0]), mxGetPi(prhs[0]), (n + 1) / 2, B, mxGetPr(plhs[0]), mxGetPi(plhs[0]));
#else
    if (nrhs != 1 || nlhs > 1) mexErrMsgTxt("Usage: aca = b2a(bc)");
    shape(prhs[0], &n, &B, &row);
    plhs[0] = row ? mxCreateDoubleMatrix(1, (size_t)n, mxCOMPLEX) : mxCreateDoubleMatrix((size_t)n, (size_t)B, mxCOMPLEX);
    rc = mbrf_b2a_batch(mxGetPr(prhs[0]), mxGetPi(prhs[0]), n, B, mxGetPr(plhs[0]), mxGetPi(plhs[0]));
#endif
    if (rc != MBRF_OK) mexErrMsgTxt(mbrf_last_error());
}
