/*
 * MEX gateway:  [alpha, beta] = abrx(rf, g, x {, y})
 *
 * Drop-in for rf_tools/mex5/abrx.c of the reference (mexFunction at abrx.c:35): same usage
 * and length checks with the same messages, outputs nx-by-ny complex.  rf_tools/abr.m keeps
 * working unchanged on top of it.  An optional 5th argument selects the convention
 * (0 abrx, 1 abrm, 2 abr) so that abrm.m can be served by the same file.
 *
 * Build:  mex -output abrx abrx_mex.c -I<repo>/include -L<pkg> -lmbrf
 */
#include "mex.h"
#include "mbrf.h"

#define MAXD(a, b) ((a) > (b) ? (a) : (b))

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    int ns, nx, ny = 1, conv = MBRF_SLR_ABRX;
    const double *y = NULL, *gy = NULL;

    if ((nrhs < 3) || (nrhs > 5) || (nlhs != 2))                       /* abrx.c:40-41 */
        mexErrMsgTxt("Usage: [alpha, beta] = abrx(rf, g, x {, y})");
    ns = (int)MAXD(mxGetN(prhs[0]), mxGetM(prhs[0]));                  /* abrx.c:43 */
    if (ns != (int)MAXD(mxGetN(prhs[1]), mxGetM(prhs[1])))            /* abrx.c:44-45 */
        mexErrMsgTxt("rf and gradient vectors are of different lengths");
    nx = (int)MAXD(mxGetN(prhs[2]), mxGetM(prhs[2]));                  /* abrx.c:53 */
    if (nrhs >= 4 && mxGetM(prhs[3]) * mxGetN(prhs[3]) > 0) {           /* abrx.c:50,54-57,64 */
        y = mxGetPr(prhs[3]);
        ny = (int)MAXD(mxGetN(prhs[3]), mxGetM(prhs[3]));
        gy = mxGetPi(prhs[1]);                                         /* NULL for a real gradient */
    }
    if (nrhs == 5) conv = (int)(*mxGetPr(prhs[4]));

    plhs[0] = mxCreateDoubleMatrix((size_t)nx, (size_t)ny, mxCOMPLEX); /* abrx.c:59-62 */
    plhs[1] = mxCreateDoubleMatrix((size_t)nx, (size_t)ny, mxCOMPLEX);

    if (mbrf_abr(mxGetPr(prhs[0]), mxGetPi(prhs[0]), mxGetPr(prhs[1]), gy, ns, mxGetPr(prhs[2]), nx, y, ny, conv,
                 mxGetPr(plhs[0]), mxGetPi(plhs[0]), mxGetPr(plhs[1]), mxGetPi(plhs[1])) != MBRF_OK)
        mexErrMsgTxt(mbrf_last_error());
}
