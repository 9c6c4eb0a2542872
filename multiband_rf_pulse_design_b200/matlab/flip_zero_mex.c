/*
 * MEX gateway of the batched zero flipping (mbrf_flip_zero_batch, csrc/flipzero.cu): the loop of fir_flip_zero.m:66-99.
 *
 *   [h_new, best, peak, power] = flip_zero_mex(Z, idx_pb, mask, hsum)
 *
 *   Z       zeros of the filter, roots(h) (fir_flip_zero.m:25), real or complex vector
 *   idx_pb  1-based indices of the passband zeros (:28)
 *   mask    N_z-by-Num, column = one flip pattern, 1 flip / 0 keep (:45-64) -- MATLAB's column-major N_z x Num is the
 *           C ABI's row-major [Num x N_z]
 *   hsum    sum(h) (:71)
 *   h_new   N-by-1 complex, the candidate with the smallest peak (:96-99); best its 1-based column; peak, power 1-by-Num
 *
 * Build:  mex -output flip_zero_mex flip_zero_mex.c -I<repo>/include -L<pkg> -lmbrf
 */
#include "mex.h"
#include "mbrf.h"
#include <stdlib.h>

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    int nroots, n_pb, num, i, rc, best = 0;
    int *idx;
    unsigned char *mask;
    const double *m, *ip;
    double hs_re, hs_im = 0.0, *peak, *power;

    if (nrhs != 4 || nlhs > 4) mexErrMsgTxt("Usage: [h_new, best, peak, power] = flip_zero_mex(Z, idx_pb, mask, hsum)");
    nroots = (int)(mxGetM(prhs[0]) * mxGetN(prhs[0]));
    n_pb = (int)(mxGetM(prhs[1]) * mxGetN(prhs[1]));
    num = n_pb ? (int)mxGetN(prhs[2]) : 1;
    if (n_pb && (int)mxGetM(prhs[2]) != n_pb) mexErrMsgTxt("flip_zero_mex: mask must have one row per passband zero");
    if (nroots < 1) mexErrMsgTxt("flip_zero_mex: no zeros");
    idx = (int *)malloc(sizeof(int) * (size_t)(n_pb + 1));
    mask = (unsigned char *)malloc((size_t)num * (size_t)(n_pb + 1));
    ip = mxGetPr(prhs[1]);
    for (i = 0; i < n_pb; i++) idx[i] = (int)ip[i] - 1;
    m = mxGetPr(prhs[2]);
    for (i = 0; i < num * n_pb; i++) mask[i] = m[i] != 0.0;
    hs_re = mxGetPr(prhs[3])[0];
    if (mxIsComplex(prhs[3])) hs_im = mxGetPi(prhs[3])[0];
    plhs[0] = mxCreateDoubleMatrix((size_t)nroots + 1, 1, mxCOMPLEX);
    peak = (double *)malloc(sizeof(double) * (size_t)num);
    power = (double *)malloc(sizeof(double) * (size_t)num);
    rc = mbrf_flip_zero_batch(mxGetPr(prhs[0]), mxIsComplex(prhs[0]) ? mxGetPi(prhs[0]) : NULL, nroots, idx, n_pb, mask, num, hs_re,
                              hs_im, &best, mxGetPr(plhs[0]), mxGetPi(plhs[0]), peak, power, NULL, NULL);
    if (rc == MBRF_OK) {
        if (nlhs > 1) { plhs[1] = mxCreateDoubleMatrix(1, 1, mxREAL); mxGetPr(plhs[1])[0] = best + 1; }
        if (nlhs > 2) { plhs[2] = mxCreateDoubleMatrix(1, (size_t)num, mxREAL); for (i = 0; i < num; i++) mxGetPr(plhs[2])[i] = peak[i]; }
        if (nlhs > 3) { plhs[3] = mxCreateDoubleMatrix(1, (size_t)num, mxREAL); for (i = 0; i < num; i++) mxGetPr(plhs[3])[i] = power[i]; }
    }
    free(idx); free(mask); free(peak); free(power);
    if (rc != MBRF_OK) mexErrMsgTxt(mbrf_last_error());
}
