/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement of the reference Bloch simulator.
 *
 * Follows /root/reference/bloch_simulation/blochC.c (blochH.c is identical but
 * for GAMMA, blochC.c:6 vs blochH.c:6).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may use this file; the product path
 * (multiband_rf_pulse_design_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED for modes 0 and 2 — checked bit-for-bit against the
 * unmodified reference C compiled into oracle/_ref/libblochC.so / libblochH.so
 * (tests/test_oracle.py).  Modes 1 and 3 of the reference are unusable as an
 * oracle: blochC.c:131-132 (`*imat = *imat++ / det`) is undefined behaviour and
 * gcc 13 builds abort or return garbage.  They are restated here with the
 * intended semantics inv = adj/det, and pinned only through properties
 * (mode-3 last sample == mode-1 result, fixed point of one period).
 *
 * Arithmetic order mirrors the reference expression by expression so that,
 * compiled with -ffp-contract=off, modes 0/2 reproduce it exactly.
 */
#include <math.h>
#include <stdlib.h>

#define ORACLE_TWOPI 6.283185 /* blochC.c:7 — the reference's truncated 2*pi */

/* 3x3 matrices are column-major like the reference (blochC.c:13-20). */

/* blochC.c:171-234 calcrotmat: rotation of |n| rad about n, via Cayley-Klein. */
static void rot_about(double nx, double ny, double nz, double R[9])
{
    double phi = sqrt(nx * nx + ny * ny + nz * nz);
    if (phi == 0.0) { /* blochC.c:182-193 */
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        return;
    }
    {
        double hp = phi / 2, cp = cos(hp), sp = sin(hp) / phi; /* :200-202 */
        double ar = cp, ai = -nz * sp, br = ny * sp, bi = -nx * sp; /* :203-206 */
        double arar = ar * ar, aiai = ai * ai, brbr = br * br, bibi = bi * bi; /* :210-219 */
        double arai2 = 2 * ar * ai, brbi2 = 2 * br * bi, arbi2 = 2 * ar * bi;
        double aibr2 = 2 * ai * br, arbr2 = 2 * ar * br, aibi2 = 2 * ai * bi;
        R[0] = arar - aiai - brbr + bibi; /* :224-232 */
        R[1] = -arai2 - brbi2;
        R[2] = -arbr2 + aibi2;
        R[3] = arai2 - brbi2;
        R[4] = arar - aiai + brbr - bibi;
        R[5] = -aibr2 - arbi2;
        R[6] = arbr2 + aibi2;
        R[7] = arbi2 - aibr2;
        R[8] = arar + aiai - brbr - bibi;
    }
}

static void mat_vec(const double M[9], const double v[3], double out[3]) /* blochC.c:13-20 */
{
    out[0] = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
    out[1] = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
    out[2] = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
}

static void mat_mat(const double A[9], const double B[9], double out[9]) /* blochC.c:153-167 */
{
    int c;
    for (c = 0; c < 3; c++) mat_vec(A, B + 3 * c, out + 3 * c);
}

/* blochC.c:38-52 (adjugate), :84-99 (determinant), :121-133 with the UB at :132
 * replaced by the intended inv[i] = adj[i] / det. */
static void mat_inv(const double m[9], double inv[9])
{
    double det, adj[9];
    int i;
    det = m[0] * m[4] * m[8];
    det += m[3] * m[7] * m[2];
    det += m[6] * m[1] * m[5];
    det -= m[0] * m[7] * m[5];
    det -= m[3] * m[1] * m[8];
    det -= m[6] * m[4] * m[2];
    adj[0] = (m[4] * m[8] - m[7] * m[5]);
    adj[1] = -(m[1] * m[8] - m[7] * m[2]);
    adj[2] = (m[1] * m[5] - m[4] * m[2]);
    adj[3] = -(m[3] * m[8] - m[6] * m[5]);
    adj[4] = (m[0] * m[8] - m[6] * m[2]);
    adj[5] = -(m[0] * m[5] - m[3] * m[2]);
    adj[6] = (m[3] * m[7] - m[6] * m[4]);
    adj[7] = -(m[0] * m[7] - m[6] * m[1]);
    adj[8] = (m[0] * m[4] - m[3] * m[1]);
    for (i = 0; i < 9; i++) inv[i] = adj[i] / det;
}

/* blochC.c:283-418 blochsim: one spin through all time samples.
 * mode 0: end point; 1: steady state; 2: every sample.  m* are in/out. */
static void spin_through_time(const double *b1r, const double *b1i, const double *gx,
                              const double *gy, const double *gz, const double *dt, int ntime,
                              const double *e1, const double *e2, double df, double dx, double dy,
                              double dz, double *mx, double *my, double *mz, int mode,
                              double gamma)
{
    double A[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, b[3] = {0, 0, 0};
    double gdx = dx * gamma, gdy = dy * gamma, gdz = dz * gamma; /* :313-315 */
    double m[3], R[9], D[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int t;
    m[0] = *mx; m[1] = *my; m[2] = *mz; /* :318-320 */

    for (t = 0; t < ntime; t++) {
        double rotz = -(gx[t] * gdx + gy[t] * gdy + gz[t] * gdz + df * ORACLE_TWOPI) * dt[t]; /* :330-331 */
        double rotx = (-b1r[t] * gamma * dt[t]); /* :332 */
        double roty = (+b1i[t] * gamma * dt[t]); /* :333 */
        double rec = 1 - e1[t];                  /* :350 */
        rot_about(rotx, roty, rotz, R);          /* :334 */
        D[0] = e2[t]; D[4] = e2[t]; D[8] = e1[t]; /* :351-353 */
        if (mode == 1) { /* :336-340, :355-359 — A <- D R A ; b <- D R b + (0,0,1-e1) */
            double RA[9], Rb[3];
            mat_mat(R, A, RA);
            mat_vec(R, b, Rb);
            mat_mat(D, RA, A);
            mat_vec(D, Rb, b);
            b[0] = b[0] + 0.0; b[1] = b[1] + 0.0; b[2] = b[2] + rec;
        } else { /* :342, :362-365 */
            double r[3];
            mat_vec(R, m, r);
            mat_vec(D, r, m);
            m[0] = m[0] + 0.0; m[1] = m[1] + 0.0; m[2] = m[2] + rec;
        }
        if (mode == 2) { *mx++ = m[0]; *my++ = m[1]; *mz++ = m[2]; } /* :381-391 */
    }
    if (mode == 0) { *mx = m[0]; *my = m[1]; *mz = m[2]; } /* :399-404 */
    else if (mode == 1) { /* :406-415: M = inv(I - A) b */
        double ImA[9], inv[9], ss[3];
        int i;
        for (i = 0; i < 9; i++) ImA[i] = -1.0 * A[i] + ((i % 4 == 0) ? 1.0 : 0.0);
        mat_inv(ImA, inv);
        mat_vec(inv, b, ss);
        *mx = ss[0]; *my = ss[1]; *mz = ss[2];
    }
}

/* blochC.c:422-511 blochsimfz.  Same argument order as the reference plus the
 * trailing gamma (6726.1 for blochC, 26754 for blochH).  mx/my/mz hold the
 * initial magnetisation at stride ntout and receive the result in place. */
int oracle_blochsimfz(const double *b1real, const double *b1imag, const double *xgrad,
                      const double *ygrad, const double *zgrad, const double *tsteps, int ntime,
                      double t1, double t2, const double *dfreq, int nfreq, const double *dxpos,
                      const double *dypos, const double *dzpos, int npos, double *mx, double *my,
                      double *mz, int mode, double gamma)
{
    int ntout = (mode & 2) ? ntime : 1; /* :444-447 */
    double *e1 = (double *)malloc(sizeof(double) * (ntime > 0 ? ntime : 1));
    double *e2 = (double *)malloc(sizeof(double) * (ntime > 0 ? ntime : 1));
    int t, f, p;
    if (!e1 || !e2) { free(e1); free(e2); return -1; }
    for (t = 0; t < ntime; t++) { /* :460-464 */
        e1[t] = exp(-tsteps[t] / t1);
        e2[t] = exp(-tsteps[t] / t2);
    }
    for (f = 0; f < nfreq; f++)       /* :468 — frequency outer */
        for (p = 0; p < npos; p++) {  /* :473 — position inner */
            if (mode == 3) {          /* :477-489 steady state, then transient from it */
                spin_through_time(b1real, b1imag, xgrad, ygrad, zgrad, tsteps, ntime, e1, e2,
                                  dfreq[f], dxpos[p], dypos[p], dzpos[p], mx, my, mz, 1, gamma);
                spin_through_time(b1real, b1imag, xgrad, ygrad, zgrad, tsteps, ntime, e1, e2,
                                  dfreq[f], dxpos[p], dypos[p], dzpos[p], mx, my, mz, 2, gamma);
            } else
                spin_through_time(b1real, b1imag, xgrad, ygrad, zgrad, tsteps, ntime, e1, e2,
                                  dfreq[f], dxpos[p], dypos[p], dzpos[p], mx, my, mz, mode, gamma);
            mx += ntout; my += ntout; mz += ntout; /* :497-499 */
        }
    free(e1); free(e2);
    return 0;
}

/* blochC.c:249-276 times2intervals: end times -> intervals, 1 iff all > 0. */
int oracle_times2intervals(const double *endtimes, double *intervals, long n)
{
    int allpos = 1;
    long i;
    double last = 0.0;
    for (i = 0; i < n; i++) {
        intervals[i] = endtimes[i] - last;
        last = endtimes[i];
        if (intervals[i] <= 0) allpos = 0;
    }
    return allpos;
}
