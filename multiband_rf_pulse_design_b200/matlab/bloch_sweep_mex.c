/*
 * MEX gateway of the B1-scale sweep (mbrf_bloch_scale_sweep): the loop of sim_rf_scale.m:82-89 -- one blochC / blochH call per
 * scaling of the pulse -- as ONE call.
 *
 *   [mx, my, mz] = bloch_sweep_mex(b1, tp, t1, t2, df, scale, gamma)
 *
 *   b1     pulse in Gauss (real or complex vector);  tp: time step in s;  t1, t2 in s;  df: off-resonances in Hz
 *   scale  B1 scalings;  gamma: 6726.1 (blochC.c:6, C-13) or 26754 (blochH.c:6, H-1)
 *   mx, my, mz: numel(df)-by-numel(scale), column k = the result of blochC(b1*scale(k), 0*b1, tp, t1, t2, df, 0, 0)
 *
 * Build:  mex -output bloch_sweep_mex bloch_sweep_mex.c -I<repo>/include -L<pkg> -lmbrf
 */
#include "mex.h"
#include "mbrf.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    int ntime, nf, ns, i, rc;
    if (nrhs != 7 || nlhs > 3) mexErrMsgTxt("Usage: [mx, my, mz] = bloch_sweep_mex(b1, tp, t1, t2, df, scale, gamma)");
    ntime = (int)(mxGetM(prhs[0]) * mxGetN(prhs[0]));
    nf = (int)(mxGetM(prhs[4]) * mxGetN(prhs[4]));
    ns = (int)(mxGetM(prhs[5]) * mxGetN(prhs[5]));
    if (ntime < 1 || nf < 1 || ns < 1) mexErrMsgTxt("bloch_sweep_mex: b1, df and scale must be non-empty");
    for (i = 0; i < 3; i++) plhs[i] = mxCreateDoubleMatrix((size_t)nf, (size_t)ns, mxREAL);
    rc = mbrf_bloch_scale_sweep(mxGetPr(prhs[0]), mxIsComplex(prhs[0]) ? mxGetPi(prhs[0]) : NULL, ntime, mxGetScalar(prhs[1]),
                                mxGetScalar(prhs[2]), mxGetScalar(prhs[3]), mxGetPr(prhs[4]), nf, mxGetPr(prhs[5]), ns,
                                mxGetPr(plhs[0]), mxGetPr(plhs[1]), mxGetPr(plhs[2]), mxGetScalar(prhs[6]));
    if (rc != MBRF_OK) mexErrMsgTxt(mbrf_last_error());
}
