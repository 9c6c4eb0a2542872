/*
 * MEX gateway:  [mx,my,mz] = blochC(b1,gr,tp,t1,t2,df,dp,mode,mx,my,mz)      (also blochH)
 *
 * Drop-in for bloch_simulation/blochC.c and blochH.c of the reference (mexFunction at
 * blochC.c:514): same MATLAB signature, same output shapes.  It only unpacks mxArrays — all
 * argument handling and the simulation itself are in libmbrf.so (mbrf_bloch, include/mbrf.h).
 *
 * Build (MATLAB):   mex -DMBRF_GAMMA=6726.1 -output blochC bloch_mex.c -I<repo>/include -L<pkg> -lmbrf
 *                   mex -DMBRF_GAMMA=26754  -output blochH bloch_mex.c -I<repo>/include -L<pkg> -lmbrf
 * Build (Octave):   mkoctfile --mex -DMBRF_GAMMA=6726.1 -o blochC.mex bloch_mex.c -I... -L... -lmbrf
 * Uses the separate real/imag API (mxGetPr/mxGetPi) like the reference; with MATLAB >= R2018a
 * compile without -R2018a (the default), Octave has only this API.
 */
#include "mex.h"
#include "mbrf.h"

#ifndef MBRF_GAMMA
#define MBRF_GAMMA MBRF_GAMMA_C13
#endif

static int numel(const mxArray *a) { return (int)(mxGetM(a) * mxGetN(a)); }

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    int ntime, ngr, ntp, nf, npos_m, npos_n, npos, mode, ntout, n_m0 = 0, dims[4], i;
    const double *mx0 = NULL, *my0 = NULL, *mz0 = NULL;
    long total;
    (void)nlhs;

    if (nrhs < 7)
        mexErrMsgTxt("Usage: [mx,my,mz] = bloch(b1,gr,tp,t1,t2,df,dp,mode,mx,my,mz)");

    ntime = numel(prhs[0]);                        /* blochC.c:571 */
    ngr = numel(prhs[1]);                          /* blochC.c:595 */
    ntp = numel(prhs[2]);                          /* blochC.c:662 */
    nf = numel(prhs[5]);                           /* blochC.c:692 */
    npos_m = (int)mxGetM(prhs[6]);                 /* blochC.c:697-698 */
    npos_n = (int)mxGetN(prhs[6]);
    npos = (npos_n == 3 || npos_n == 2) ? npos_m : npos_m * npos_n;
    mode = nrhs > 7 ? (int)(*mxGetPr(prhs[7])) : 0; /* blochC.c:772-775 */
    ntout = (mode & 2) ? ntime : 1;                /* blochC.c:778-781 */
    total = (long)ntout * npos * nf;

    if (nrhs > 10) {                               /* blochC.c:820-823 */
        mx0 = mxGetPr(prhs[8]); my0 = mxGetPr(prhs[9]); mz0 = mxGetPr(prhs[10]);
        n_m0 = numel(prhs[8]);
        if (numel(prhs[9]) != n_m0 || numel(prhs[10]) != n_m0) n_m0 = -1;  /* -> (0,0,1), blochC.c:851-865 */
    }

    for (i = 0; i < 3; i++) plhs[i] = mxCreateDoubleMatrix((size_t)total, 1, mxREAL);   /* blochC.c:806-808 */

    if (mbrf_bloch(mxGetPr(prhs[0]), mxIsComplex(prhs[0]) ? mxGetPi(prhs[0]) : NULL, ntime,
                   mxGetPr(prhs[1]), ngr, mxGetPr(prhs[2]), ntp,
                   *mxGetPr(prhs[3]), *mxGetPr(prhs[4]), mxGetPr(prhs[5]), nf,
                   mxGetPr(prhs[6]), npos_m, npos_n, mode, mx0, my0, mz0, n_m0,
                   mxGetPr(plhs[0]), mxGetPr(plhs[1]), mxGetPr(plhs[2]), dims, MBRF_GAMMA) != MBRF_OK)
        mexErrMsgTxt(mbrf_last_error());

    {   /* blochC.c:880-904; mxSetDimensions takes mwSize (size_t on 64-bit MATLAB / Octave), not int */
        mwSize d[3];
        d[0] = (mwSize)dims[0]; d[1] = (mwSize)dims[1]; d[2] = (mwSize)dims[2];
        for (i = 0; i < 3; i++) mxSetDimensions(plhs[i], d, (mwSize)dims[3]);
    }
}
