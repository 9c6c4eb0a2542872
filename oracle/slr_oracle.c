/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement of the reference forward-SLR
 * (Cayley-Klein) recursions.
 *
 *   oracle_abrx : /root/reference/rf_tools/mex5/abrx.c:35-115 (mexFunction loop + abrot)
 *   oracle_abrm : /root/reference/rf_tools/abrm.m:26-64
 *
 * Parity status: oracle_abrx is PINNED against the unmodified reference C
 * (oracle/_ref/libabrx.so, driven through its own mexFunction) in
 * tests/test_oracle.py.  abrm.m cannot run here (no MATLAB/Octave); oracle_abrm
 * is pinned through the identity abrm(rf,g,x) == (a_abrx(-x), conj(b_abrx(-x)))
 *  and |a|^2+|b|^2 == 1.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this.
 */
#include <math.h>
#include <stddef.h>

/* abrx.c:81-115 abrot for one (x, y).  rfq / gy may be NULL (real rf, 1-D g). */
static void abrx_one(const double *rfi, const double *rfq, const double *gx, const double *gy,
                     int ns, double x, double y, double a[2], double b[2])
{
    int k;
    for (k = 0; k < ns; k++) {
        double cg = x * gx[k];                  /* :88 */
        double cpr, cpi, phi, nx, ny, nz, csp, snp, al[2], be[2], ap[2], bp[2];
        if (gy != NULL) cg += y * gy[k];        /* :89 */
        cpr = rfi[k];
        cpi = rfq != NULL ? rfq[k] : 0.0;       /* :90-92 */
        phi = sqrt(cg * cg + cpr * cpr + cpi * cpi); /* :93 */
        if (phi > 0.0) { nx = cpr / phi; ny = cpi / phi; nz = cg / phi; } /* :94-95 */
        else { nx = 0.0; ny = 0.0; nz = 1.0; }  /* :97 */
        csp = cos(phi / 2); snp = sin(phi / 2); /* :99 */
        al[0] = csp; al[1] = nz * snp;          /* :100 */
        be[0] = ny * snp; be[1] = nx * snp;     /* :101 */
        /* :103-109, same association order as written there */
        bp[0] = al[0] * b[0] - al[1] * b[1] + be[0] * a[0] - be[1] * (-a[1]);
        bp[1] = al[0] * b[1] + al[1] * b[0] + be[1] * a[0] + be[0] * (-a[1]);
        ap[0] = -(be[0] * b[0] - (-be[1]) * b[1]) + al[0] * a[0] - (-al[1]) * (-a[1]);
        ap[1] = -(-(-(be[1]) * b[0] + be[0] * b[1]) + (-al[1]) * a[0] + al[0] * (-a[1]));
        a[0] = ap[0]; a[1] = ap[1]; b[0] = bp[0]; b[1] = bp[1]; /* :111 */
    }
}

/* abrx.c:67-78: outputs indexed ix + iy*nx, split real/imag planes.
 * have_y == 0 mirrors the 3-argument call (ny = 1, y = 0, gy ignored). */
void oracle_abrx(const double *rfi, const double *rfq, const double *gx, const double *gy, int ns,
                 const double *xp, int nx, const double *yp, int ny, int have_y, double *alpr,
                 double *alpi, double *btpr, double *btpi)
{
    int ix, iy;
    if (!have_y) ny = 1;
    for (iy = 0; iy < ny; iy++) {
        double y = have_y ? yp[iy] : 0.0;
        for (ix = 0; ix < nx; ix++) {
            double a[2] = {1.0, 0.0}, b[2] = {0.0, 0.0};
            abrx_one(rfi, rfq, gx, have_y ? gy : NULL, ns, xp[ix], y, a, b);
            alpr[ix + iy * nx] = a[0]; alpi[ix + iy * nx] = a[1];
            btpr[ix + iy * nx] = b[0]; btpi[ix + iy * nx] = b[1];
        }
    }
}

/* abrm.m:46-60.  g is complex (gr = Re g, gi = Im g; gi may be NULL = 0),
 * outputs are lx-by-ly column-major (index kk + jj*lx).  As in the .m file
 * there is no phi == 0 guard: 0/0 gives NaN there and here. */
void oracle_abrm(const double *rfr, const double *rfi, const double *gr, const double *gi, int ns,
                 const double *x, int lx, const double *y, int ly, double *ar, double *ai,
                 double *br, double *bi)
{
    int jj, kk, m;
    for (jj = 0; jj < ly; jj++)
        for (kk = 0; kk < lx; kk++) {
            double a_re = 1.0, a_im = 0.0, b_re = 0.0, b_im = 0.0;
            for (m = 0; m < ns; m++) {
                double fr = rfr[m], fi = rfi ? rfi[m] : 0.0;
                double om = x[kk] * gr[m] + y[jj] * (gi ? gi[m] : 0.0);   /* :48 */
                double phi = sqrt((fr * fr + fi * fi) + om * om);          /* :49 */
                double n1 = fr / phi, n2 = fi / phi, n3 = om / phi;        /* :50 */
                double s = sin(phi / 2), c = cos(phi / 2);
                double av_re = c, av_im = -n3 * s;                         /* :51 */
                double bv_re = n2 * s, bv_im = -n1 * s;                    /* :52: -i*(n1+i*n2)*s */
                /* :55: [a;b] <- [av -conj(bv); bv conj(av)] [a;b] */
                double na_re = (av_re * a_re - av_im * a_im) - (bv_re * b_re + bv_im * b_im);
                double na_im = (av_re * a_im + av_im * a_re) - (bv_re * b_im - bv_im * b_re);
                double nb_re = (bv_re * a_re - bv_im * a_im) + (av_re * b_re + av_im * b_im);
                double nb_im = (bv_re * a_im + bv_im * a_re) + (av_re * b_im - av_im * b_re);
                a_re = na_re; a_im = na_im; b_re = nb_re; b_im = nb_im;
            }
            ar[kk + jj * lx] = a_re; ai[kk + jj * lx] = a_im;
            br[kk + jj * lx] = b_re; bi[kk + jj * lx] = b_im;
        }
}
