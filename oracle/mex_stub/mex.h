/*
 * TEST INFRASTRUCTURE ONLY — stand-in for MATLAB's mex.h.
 *
 * Neither MATLAB nor Octave exists in the build container, so the reference MEX
 * sources (/root/reference/bloch_simulation/blochC.c, blochH.c,
 * rf_tools/mex5/abrx.c, b2rf.c) are compiled for the oracle against this header.
 * It provides exactly the symbols those files use (pre-R2018a separate
 * real/imag API): mxArray, mxGetM/N/Pr/Pi, mxIsComplex, mxCreateDoubleMatrix,
 * mxSetDimensions, mexErrMsgTxt, mexPrintf.
 *
 * `mex_stub_call` runs the translation unit's mexFunction under setjmp so that
 * mexErrMsgTxt (which MATLAB implements as a longjmp out of the MEX file) comes
 * back as a return code + message instead of killing the test process.
 *
 * The same header is used to syntax/ABI-check our own MEX gateways
 * (multiband_rf_pulse_design_b200/matlab, the *_mex.c files).
 */
#ifndef MBRF_MEX_STUB_H
#define MBRF_MEX_STUB_H

#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <stdarg.h>
#include <setjmp.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mxArray_stub {
    size_t m, n;        /* rows, columns (n = product of trailing dims) */
    double *pr, *pi;    /* separate real / imaginary planes; pi == NULL if real */
    int ndim;
    int dims[3];
    int is_char;        /* 1 when the array holds a string (status outputs) */
} mxArray;

typedef int mxComplexity;
#define mxREAL 0
#define mxCOMPLEX 1
typedef size_t mwSize;

static inline size_t mxGetM(const mxArray *a) { return a->m; }
static inline size_t mxGetN(const mxArray *a) { return a->n; }
static inline double *mxGetPr(const mxArray *a) { return a->pr; }
static inline double *mxGetPi(const mxArray *a) { return a->pi; }
static inline int mxIsComplex(const mxArray *a) { return a->pi != NULL; }
static inline size_t mxGetNumberOfElements(const mxArray *a) { return a->m * a->n; }
static inline double mxGetScalar(const mxArray *a) { return a->pr[0]; }
static inline int mxIsEmpty(const mxArray *a) { return a->m * a->n == 0; }

static inline mxArray *mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity c)
{
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
    size_t cnt = (m * n) > 0 ? m * n : 1;
    a->m = m; a->n = n; a->ndim = 2;
    a->dims[0] = (int)m; a->dims[1] = (int)n; a->dims[2] = 1;
    a->pr = (double *)calloc(cnt, sizeof(double));
    a->pi = c == mxCOMPLEX ? (double *)calloc(cnt, sizeof(double)) : NULL;
    return a;
}

static inline mxArray *mxCreateString(const char *s)
{
    size_t n = strlen(s), i;
    mxArray *a = mxCreateDoubleMatrix(1, n, mxREAL);
    for (i = 0; i < n; i++) a->pr[i] = (double)(unsigned char)s[i];
    a->is_char = 1;
    return a;
}

/* The real API is  int mxSetDimensions(mxArray *, const mwSize *dims, mwSize ndim)  with mwSize = size_t on 64-bit MATLAB
 * and 64-bit-index Octave.  The reference's 1990s sources pass an `int` array (blochC.c:885-903) -- they are compiled with
 * -DMEX_STUB_INT_DIMS (oracle/Makefile) so that the stub reads what they write; everything else (our gateways) compiles
 * against the real signature, so a gateway that passes an int array does not compile cleanly here either. */
#ifdef MEX_STUB_INT_DIMS
typedef int mex_stub_dim_t;
#else
typedef mwSize mex_stub_dim_t;
#endif
static inline int mxSetDimensions(mxArray *a, const mex_stub_dim_t *dims, mex_stub_dim_t ndim)
{
    int i; size_t tail = 1;
    a->ndim = (int)ndim;
    for (i = 0; i < 3; i++) a->dims[i] = i < (int)ndim ? (int)dims[i] : 1;
    for (i = 1; i < (int)ndim; i++) tail *= (size_t)dims[i];
    a->m = (size_t)dims[0]; a->n = tail;
    return 0;
}

static inline void mxDestroyArray(mxArray *a)
{
    if (!a) return;
    free(a->pr); free(a->pi); free(a);
}

/* ---- error path: mexErrMsgTxt longjmps back into mex_stub_call ---- */
static jmp_buf mex_stub_jmp;
static int mex_stub_jmp_armed = 0;
static char mex_stub_errmsg[512];

static inline void mexErrMsgTxt(const char *msg)
{
    strncpy(mex_stub_errmsg, msg ? msg : "", sizeof(mex_stub_errmsg) - 1);
    mex_stub_errmsg[sizeof(mex_stub_errmsg) - 1] = 0;
    if (mex_stub_jmp_armed) longjmp(mex_stub_jmp, 1);
    fprintf(stderr, "mexErrMsgTxt: %s\n", mex_stub_errmsg);
    abort();
}

static inline void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...)
{
    char buf[400]; va_list ap;
    va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    (void)id;
    mexErrMsgTxt(buf);
}

static inline int mexPrintf(const char *fmt, ...)
{
    va_list ap; int r;
    va_start(ap, fmt); r = vprintf(fmt, ap); va_end(ap);
    return r;
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);

/* returns 0 on normal return, 1 if mexErrMsgTxt fired (message via mex_stub_last_error) */
int mex_stub_call(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    int rc;
    mex_stub_errmsg[0] = 0;
    mex_stub_jmp_armed = 1;
    if (setjmp(mex_stub_jmp) == 0) { mexFunction(nlhs, plhs, nrhs, prhs); rc = 0; }
    else rc = 1;
    mex_stub_jmp_armed = 0;
    return rc;
}
const char *mex_stub_last_error(void) { return mex_stub_errmsg; }
mxArray *mex_stub_create(size_t m, size_t n, int cplx) { return mxCreateDoubleMatrix(m, n, cplx); }
void mex_stub_destroy(mxArray *a) { mxDestroyArray(a); }

#ifdef __cplusplus
}
#endif
#endif
