/*
 * mbrf.h — C ABI of libmbrf.so, the B200 (sm_100a) engine behind the MATLAB/MEX
 * signatures of shanghong/Multiband-RF-pulse-Design's data-parallel hot paths.
 *
 * Plain C: pointers, sizes and scalars only — no torch, no C++ types.  Every entry
 * point names the reference interface it replaces (file:line under /root/reference).
 * All functions return MBRF_OK (0) or a negative MBRF_E* code; the message of the
 * last failure on the calling thread is available from mbrf_last_error().
 *
 * There is NO CPU fallback: without a CUDA device every compute entry point returns
 * MBRF_ENODEVICE.
 *
 * Pointer conventions
 *   *_host entry points (no suffix)  : host memory, pageable or pinned; the call is
 *                                      synchronous and performs its own H2D / D2H.
 *   *_device entry points            : device memory of the current device; work is
 *                                      enqueued on `stream` (a cudaStream_t passed as
 *                                      void*, NULL = legacy default stream) and NOT
 *                                      synchronised.
 * Arrays are column-major doubles; complex data is split into real / imaginary
 * planes exactly as the pre-R2018a MEX API hands them over (mxGetPr / mxGetPi).
 */
#ifndef MBRF_H
#define MBRF_H

#ifdef __cplusplus
extern "C" {
#endif

#define MBRF_OK          0
#define MBRF_EINVAL     -1   /* bad argument (sizes, NULL where data is required, bad mode) */
#define MBRF_ENODEVICE  -2   /* no usable CUDA device: the engine has no CPU path */
#define MBRF_ECUDA      -3   /* CUDA runtime error, text in mbrf_last_error() */
#define MBRF_ENOMEM     -4

#define MBRF_GAMMA_C13  6726.1   /* blochC.c:6 */
#define MBRF_GAMMA_H1   26754.0  /* blochH.c:6 */
#define MBRF_TWOPI      6.283185 /* blochC.c:7 — the reference's truncated 2*pi, kept for parity */

/* ---- library / device ---------------------------------------------------- */
const char *mbrf_version(void);
const char *mbrf_last_error(void);
int  mbrf_device_count(void);             /* >= 0, never fails */
int  mbrf_set_device(int device);         /* device used by this host thread's later calls */
int  mbrf_get_device(int *device);        /* the calling host thread's current device (worker threads inherit nothing) */
int  mbrf_device_sm_count(int *sm_count); /* multiprocessors of the current device */
/* Host-pointer entry points (mbrf_bloch, mbrf_blochsimfz, mbrf_abr) spread one call over `ndevices` GPUs of the box: contiguous
 * spin / position ranges, one per device, results written into the caller's arrays (no collective).  0 = all devices; default 1,
 * or the environment variable MBRF_FANOUT.  A MATLAB host is one process: this is how it uses more than one GPU. */
int  mbrf_set_fanout(int ndevices);
int  mbrf_get_fanout(void);
/* Peer-mapped result buffer of a sharded run (one process per GPU): the destination rank allocates [planes x S] doubles and
 * exports a 64-byte CUDA IPC handle; the other ranks open it and hand pointers into it to mbrf_bloch_device / mbrf_abr_device as
 * output arrays, so every kernel stores its slice straight into the destination GPU over NVLink -- the "one gather at the end"
 * of SURVEY.md 8(e) without a transfer step.  The caller orders completion (a barrier / 1-element all-reduce on the stream). */
int  mbrf_peer_alloc(unsigned long long bytes, void **dptr, unsigned char handle[64]);
int  mbrf_peer_open(const unsigned char handle[64], void **dptr);
int  mbrf_peer_close(void *dptr);   /* a pointer from mbrf_peer_open */
int  mbrf_peer_free(void *dptr);    /* a pointer from mbrf_peer_alloc */
/* counts kernel launches made by this library on the calling process (for bench `gpu_launches`) */
unsigned long long mbrf_launch_count(void);

/* Diagnostic: FP64 FMA throughput of the current device (TFLOP/s, FMA = 2 flops), measured with a
 * DFMA-only kernel.  The Bloch / SLR kernels are bound by this pipe; bench.py reports against it. */
int  mbrf_measure_fp64_peak(double *tflops, double *kernel_ms);

/* ---- Bloch simulation ---------------------------------------------------- */

/*
 * Replaces  int blochsimfz(...)  — blochC.c:422-511 / blochH.c:422-511.
 * Same argument order and meaning, plus the trailing gyromagnetic ratio that
 * distinguishes blochC (MBRF_GAMMA_C13) from blochH (MBRF_GAMMA_H1).
 *
 *   b1real,b1imag,xgrad,ygrad,zgrad,tsteps : ntime samples each (G, G/cm, s);
 *        b1imag / ygrad / zgrad may be NULL meaning all-zero.
 *   dfreq  : nfreq off-resonances (Hz);  dxpos,dypos,dzpos : npos positions (cm),
 *        dypos / dzpos may be NULL meaning all-zero.
 *   mode   : bit0 steady state, bit1 record every sample (blochC.c:772-781).
 *   mx,my,mz : IN/OUT, ntout*npos*nfreq doubles, ntout = (mode&2) ? ntime : 1.
 *        On entry element [ntout*(p + npos*f)] holds the initial magnetisation of
 *        spin (p,f) (blochC.c:838-865); on return element [t + ntout*(p + npos*f)]
 *        holds the result (blochC.c:497-499).
 * Spins are independent; they are spread over one thread each on the current device.
 */
int mbrf_blochsimfz(const double *b1real, const double *b1imag,
                    const double *xgrad, const double *ygrad, const double *zgrad,
                    const double *tsteps, int ntime, double t1, double t2,
                    const double *dfreq, int nfreq,
                    const double *dxpos, const double *dypos, const double *dzpos, int npos,
                    double *mx, double *my, double *mz, int mode, double gamma);

/*
 * Replaces the argument handling of  mexFunction  — blochC.c:514-927 — i.e. the MATLAB call
 *   [mx,my,mz] = blochC(b1, gr, tp, t1, t2, df, dp, mode, mx0, my0, mz0)
 * so that the MEX gateway (matlab/bloch_mex.c) only unpacks mxArrays.
 *
 *   b1r,b1i : ntime samples, b1i NULL for a real pulse            (blochC.c:571-587)
 *   gr      : ngr = numel(gr) doubles, column-major ntime x {1,2,3}; missing axes are
 *             zero; ngr not in {1,2,3}*ntime only warns in the reference (:595-638)
 *   tp, ntp : ntp == 1 constant step; ntp == ntime: end times if strictly increasing
 *             from 0, else intervals (:660-681); any other ntp -> MBRF_EINVAL
 *             (the reference prints a warning and then reads out of bounds)
 *   df, nf  : off-resonances                                       (:691-692)
 *   dp      : column-major npos_m x npos_n; npos_n == 3 / 2 -> xyz / xy columns,
 *             anything else -> npos_m*npos_n one-dimensional x positions (:701-758)
 *   mode    : 0..3                                                 (:772-781)
 *   mx0,my0,mz0,n_m0 : optional initial magnetisation, used only if all three are
 *             non-NULL and n_m0 == npos*nf, else (0,0,1)           (:820-865)
 *   mx,my,mz: OUT, ntout*npos*nf doubles each, layout [t + ntout*(p + npos*f)]
 *   out_dims: OUT, the MATLAB shape of the outputs; returns ndim (2 or 3) in
 *             out_dims[3]                                          (:880-904)
 */
int mbrf_bloch(const double *b1r, const double *b1i, int ntime,
               const double *gr, int ngr,
               const double *tp, int ntp, double t1, double t2,
               const double *df, int nf,
               const double *dp, int npos_m, int npos_n, int mode,
               const double *mx0, const double *my0, const double *mz0, int n_m0,
               double *mx, double *my, double *mz, int out_dims[4], double gamma);

/*
 * Device-resident form of mbrf_blochsimfz for callers that keep data in HBM
 * (bench `value`, multi-GPU shards).  All pointers are device pointers.
 *   spin0, nspins : this call simulates flattened spins s in [spin0, spin0+nspins),
 *                   s = p + npos*f (blochC.c:468-473), which is how shards are cut.
 *   m0x,m0y,m0z   : initial magnetisation indexed by LOCAL spin (s - spin0) with stride
 *                   m0_stride doubles, or all NULL for (0,0,1).
 *   mx,my,mz      : outputs indexed [t + ntout*(s - spin0)].
 *   workspace     : device scratch of mbrf_bloch_workspace_bytes(ntime) bytes, 16-byte aligned (the kernels stream it with
 *                   cp.async.bulk); a misaligned pointer is MBRF_EINVAL.  The same holds for mbrf_abr_device.
 */
unsigned long long mbrf_bloch_workspace_bytes(int ntime);
int mbrf_bloch_device(const double *b1real, const double *b1imag,
                      const double *xgrad, const double *ygrad, const double *zgrad,
                      const double *tsteps, int ntime, double t1, double t2,
                      const double *dfreq, int nfreq,
                      const double *dxpos, const double *dypos, const double *dzpos, int npos,
                      long long spin0, long long nspins,
                      const double *m0x, const double *m0y, const double *m0z, int m0_stride,
                      double *mx, double *my, double *mz, int mode, double gamma,
                      void *workspace, void *stream);

/*
 * Sweep extension (no reference twin: the reference runs it as one blochC call per
 * scale, sim_rf_scale.m:82-89).  Spin s = i_f + nfreq*i_s sees b1 * b1scale[i_s] and
 * off-resonance dfreq[i_f], at position (0,0,0)-free gradient-less conditions unless
 * gradients/positions are given (position index 0 is used for every spin).
 */
int mbrf_bloch_scale_sweep_device(const double *b1real, const double *b1imag,
                                  const double *tsteps, int ntime, double t1, double t2,
                                  const double *dfreq, int nfreq,
                                  const double *b1scale, int nscale,
                                  long long spin0, long long nspins,
                                  double *mx, double *my, double *mz, double gamma,
                                  void *workspace, void *stream);

/* host-pointer form of the sweep: constant time step tp (s), b1imag may be NULL, M0 = (0, 0, 1), mode 0;
 * mx, my, mz: nfreq*nscale doubles, element [i_f + nfreq*i_s] */
int mbrf_bloch_scale_sweep(const double *b1real, const double *b1imag, int ntime, double tp, double t1, double t2,
                           const double *dfreq, int nfreq, const double *b1scale, int nscale,
                           double *mx, double *my, double *mz, double gamma);

/* tuning knobs of the Bloch kernel (0 = automatic): resident CTAs per SM, spins per thread (1|2) */
int mbrf_bloch_set_tuning(int ctas_per_sm, int spins_per_thread);

/* ---- forward SLR (Cayley-Klein) ------------------------------------------ */

#define MBRF_SLR_ABRX 0   /* rf_tools/mex5/abrx.c convention                         */
#define MBRF_SLR_ABRM 1   /* rf_tools/abrm.m convention: (a_abrx(-x), conj b_abrx(-x)), NaN at phi==0 */
#define MBRF_SLR_ABR  2   /* rf_tools/abr.m:34: abrx followed by b = -conj(b)         */

/*
 * Replaces  mexFunction + abrot  — rf_tools/mex5/abrx.c:35-115 — and, through
 * `convention`, rf_tools/abrm.m:46-60 and rf_tools/abr.m:19-37.
 *   rfr,rfi : ns samples (radians); rfi NULL for a real pulse     (abrx.c:51,92)
 *   gx,gy   : ns samples; gy NULL = no y gradient                  (abrx.c:49-50,89)
 *   x, nx   : positions; y, ny: second axis, y NULL => ny = 1, y = 0 (abrx.c:53-57,68)
 *   outputs : nx*ny doubles each, index ix + iy*nx                 (abrx.c:73-76)
 */
int mbrf_abr(const double *rfr, const double *rfi, const double *gx, const double *gy, int ns,
             const double *x, int nx, const double *y, int ny, int convention,
             double *alpha_r, double *alpha_i, double *beta_r, double *beta_i);

/* device-resident form; positions [pos0, pos0+npos) of the flattened ix + iy*nx index */
int mbrf_abr_device(const double *rfr, const double *rfi, const double *gx, const double *gy, int ns,
                    const double *x, int nx, const double *y, int ny, int convention,
                    long long pos0, long long npos,
                    double *alpha_r, double *alpha_i, double *beta_r, double *beta_i,
                    void *workspace, void *stream);
unsigned long long mbrf_abr_workspace_bytes(int ns);

/* ---- convex FIR design step: batched restarted PDHG --------------------------- */

/*
 * Replaces the SOLVE inside  fir_ap_cvx.m:160-169 (CVX -> SeDuMi/SDPT3),  ss/fir_linprog.m:246-252
 * (MATLAB linprog) and the probes of fir_ap.m:51-173 / ss/fir_min_order_linprog.m:98-211: B problems
 *
 *      minimise c^T z   s.t.   lo <= K z <= hi,   bl <= z <= bu,   ||(z_pi, z_pj)|| <= rho
 *
 * that share one frequency-sampled Fourier matrix K (M x N), generated ON THE DEVICE from
 *      K[i][j] = col_amp[j] * {1, cos, sin}[col_type[j]] (w_row[i] * col_kappa[j]),  col_type 3 = zero,
 *      K[i][tcol] = tcoef[i]   (tcol < 0: no such column)
 * which covers A = [1, 2cos(w k), 2sin(w k)] (fir_ap_cvx.m:100), [Acos Asin] (fir_linprog.m:195-217)
 * and an optional explicit column.  Rows [simplex_row0, simplex_row0+simplex_rows) add the term
 * simplex_w[b] * max_i (K z)_i (over the rows of the block with hi == 0) to design b's objective: this is
 * `obj*ripple_stop` with `A_U(idx_stop,:)*x <= ripple_stop` (fir_ap_cvx.m:163-165) with ripple_stop eliminated
 * (its multipliers live on a simplex, handled exactly in the dual step).  HOST pointers; per-design arrays are [dim x B]
 * row-major (design index fastest).  obj_upper[b] (optional): an upper bound on the objective of any
 * feasible point; a dual bound above it certifies infeasibility (status 2).
 *   z_out [N x B]: solutions in the caller's units;  info_out [B x 8]: status (1 solved, 2 infeasible,
 *   3 iteration limit), iterations, objective, dual objective, max row violation, natural residual,
 *   rigorous lower bound on the optimum, max_i (K z)_i of the simplex block (= ripple_stop);  colscale_out [N] optional.
 * The status maps onto the reference's strings: 1 -> 'Solved', 2/3 -> 'Failed' (fir_ap_cvx.m:176-182).
 */
int mbrf_fir_pdhg_solve(const double *w_row, const double *tcoef, int M,
                        const int *col_type, const double *col_kappa, const double *col_amp, int N, int tcol,
                        const int *pair_i, const int *pair_j, int npairs,
                        const double *c, const double *lo, const double *hi,
                        const double *bl, const double *bu, const double *rho, int B,
                        const double *obj_upper,
                        int simplex_row0, int simplex_rows, const double *simplex_w,
                        int max_iter, int check_every,
                        double eps_pr, double eps_dr, double eps_gap,
                        double *z_out, double *info_out, double *colscale_out);

/* Optional structure of the rows / columns beyond plain intervals (all fields 0 / NULL = absent).
 * Weight arrays are per design: [B] host doubles for mbrf_fir_pdhg_solve2, [Bp] device doubles for
 * mbrf_pdhg_solve_device. */
typedef struct mbrf_pdhg_blocks {
    int simplex_row0, simplex_rows;   /* rows adding  simplex_w[b] * max_i (K z)_i  over rows with hi == 0   */
    double *simplex_w;                /*   (obj*ripple_stop, fir_ap_cvx.m:163-165)                              */
    int disk_row0, disk_pairs;        /* row pairs (r, r+1): ||(K z)_pair - (lo[r], lo[r+1])|| <= hi[r]        */
                                      /*   (norm(A_i*x - Hd_i) <= D_i, fir_qp_cvx.m:148-157)                    */
    int group_row0, group_pairs;      /* row pairs adding  group_w[b] * max_i ||(K z)_pair_i||                  */
    double *group_w;                  /*   (obj*Peak with norm(F_i*x) <= Peak, fir_qp_cvx.m:147,158-160)        */
    int norm_coords;                  /* adds  norm_w[b] * ||z[0 .. norm_coords)||_2                            */
    double *norm_w;                   /*   (E_total with norm(x,2) <= E_total, fir_qp_cvx.m:147,161)            */
    int group2_row0, group2_pairs;    /* row pairs adding  group2_w[b] * max_i ||(K z)_pair_i - (lo[r], lo[r+1])||  */
    double *group2_w;                 /*   (delta of the minimax form, fir_qp_cvx.m:170-177, rows pre-scaled 1/D_i) */
} mbrf_pdhg_blocks;

/* Matrix description with the extras fir_qp_cvx needs: K[i][j] = row_scale[i] * col_amp[j] *
 * trig_j(w_row[i]*col_kappa[j] + row_phase[i]) (+ explicit entries K[ti[k]][tj[k]] += tv[k]); NULL arrays mean
 * phase 0 / scale 1 / no entries.  [cos sin; -sin cos] (fir_qp_cvx.m:96-109) is phase 0 and pi/2. */
int mbrf_fir_pdhg_solve2(const double *w_row, const double *row_phase, const double *row_scale, int M,
                         const int *col_type, const double *col_kappa, const double *col_amp, int N,
                         int nnz, const int *ti, const int *tj, const double *tv,
                         const int *pair_i, const int *pair_j, int npairs,
                         const double *c, const double *lo, const double *hi,
                         const double *bl, const double *bu, const double *rho, int B,
                         const double *obj_upper, const mbrf_pdhg_blocks *blocks,
                         int max_iter, int check_every, double eps_pr, double eps_dr, double eps_gap,
                         double *z_out, double *info_out, double *colscale_out);

/* Warm start of the next mbrf_fir_pdhg_solve / _solve2 call made by this host thread (consumed by that call).  Host pointers in
 * the caller's units, valid until that call returns; any may be NULL: z_init [N x B], y_init [M x B] starting iterate and
 * multipliers, omega_init [B] primal weights; y_out [M x B], omega_out [B] receive the final multipliers / primal weights
 * (for warm-starting a neighbouring design of a sweep, which fir_ap.m-style searches solve one after the other). */
int mbrf_fir_pdhg_warm_start(const double *z_init, const double *y_init, const double *omega_init, double *y_out,
                             double *omega_out);
/* the same for mbrf_pdhg_solve_device: device iterates [Np x Bp] / [Mp x Bp], host primal weights [Bp] in / out */
int mbrf_pdhg_warm_start_device(const double *z_init, const double *y_init, const double *omega_init, double *omega_out);

/* Device-resident core of the above on a padded batch (Mp, Np, Bp multiples of 64; arrays [dim x Bp]);
 * K row-major [Mp x ldk] with columns already scaled, KT its transpose [Np x Mp]. */
int mbrf_pdhg_padded_sizes(int M, int N, int B, int *Mp, int *Np, int *Bp);
int mbrf_pdhg_set_option(int which, double value);   /* 0 eta factor, 1-3 restart betas, 4 primal-weight smoothing,
                                                        5 minimum iterations between restarts (default 1: none) */
int mbrf_pdhg_set_gemm(int mode);       /* product kernels of the iterations: 2 (default) tcgen05 int8 split-integer tiles,
                                           1 FP64 tensor tiles mma.sync.m8n8k4, 0 SIMT DFMA tiles; checks always run in fp64 */
int mbrf_pdhg_set_halpern(int mode);    /* reflected Halpern PDHG iteration (r2HPDHG): 2 (default) on, candidate = better of PDHG output and
                                           Halpern iterate; 1 on, candidate = PDHG output; 0 off (restarts to running averages) */
int mbrf_pdhg_set_tc_digits(int nd);    /* base-256 digit planes of the split-integer product: 4, 5 (default) or 6 */
unsigned long long mbrf_pdhg_workspace_bytes(int Mp, int Np, int Bp);
/* The tcgen05 split-integer product alone, device pointers: C[nslab][R x Bp] = A[R x kdim] * X[kdim x Bp] over nslab ranges of
 * the reduction (sum the slabs); R, kdim, Bp multiples of 64; nd = 4..6.  Timed with CUDA events on `stream` over `reps`
 * repetitions: ms_gemm = MMA kernel alone, ms_total = digit planes of X + MMA kernel (either may be NULL). */
int mbrf_tc_product_device(const double *A, int R, int kdim, const double *X, int Bp, int nd, int nslab, double *C, int reps,
                           float *ms_gemm, float *ms_total, void *stream);
int mbrf_pdhg_solve_device(const double *K, const double *KT, int Mp, int Np, int ldk,
                           double *c, double *lo, double *hi, double *bl, double *bu,   /* destroyed: compacted */
                           const int *pair_i, const int *pair_j, int npairs, double *rho,
                           int Bp, int B, double *obj_upper, const mbrf_pdhg_blocks *blocks,
                           int max_iter, int check_every, double eps_pr, double eps_dr, double eps_gap,
                           double *z_out, double *y_out, double *info_out,
                           void *workspace, void *stream);

/* ---- convex FIR design step: batched interior-point solver (csrc/ipm.cu) --------------------------------
 *
 * Second-order companion of mbrf_fir_pdhg_solve for the SAME problem description (same arrays, same meaning; the
 * explicit column tcoef/tcol is not supported).  It replaces the solve inside fir_ap_cvx.m:160-169 and ss/fir_linprog.m:246-252
 * where the reference's own callers need interior-point accuracy: stop-band weights obj = 1e4 / 1e5 (dzrf_mb.m:167-170,
 * fir_qp.m:47) make the objective lexicographic in all but name, which a first-order method does not resolve to 1e-4.
 * Homogeneous self-dual embedding, Nesterov-Todd scaling, Mehrotra predictor-corrector; the normal matrix of every design is
 * assembled from the Toeplitz/Hankel moments of its row weights (one contraction of the shared trigonometric table with the
 * weight vectors of the batch) and factorised by a batched Cholesky, in double-double once the complementarity gap is small.
 *   Column frequencies col_kappa must be multiples of 1/2 (all the reference's matrices: fir_ap_cvx.m:100 k = 1..n-1,
 *   ss/fir_linprog.m:195-217 k or k + 1/2).  bl / bu may be NULL (no bounds); infinite entries mean "no bound".
 *   feastol (relative primal residual, default 1e-7; the dual residual gets 100 x feastol), reltol (relative gap, default 2e-6), abstol (default 1e-12).
 *   A design whose last iterations lose feasibility again (fp64 cone scalings at the boundary) ends with its best iterate if that was
 *   within 10x of the tolerances (CVX's "Inaccurate/Solved", accepted as 'Solved' by fir_ap_cvx.m:176-182).
 *   info_out [B x 8] as mbrf_fir_pdhg_solve: status 1 optimal, 2 primal infeasible (a Farkas certificate was found),
 *   3 iteration limit or numerical failure; iterations; objective; dual objective; max violation of the returned point
 *   (recomputed from it); relative dual residual; dual objective; max_i (K z)_i over the stop block (= ripple_stop).
 */
int mbrf_fir_ipm_solve(const double *w_row, int M, const int *col_type, const double *col_kappa, const double *col_amp, int N,
                       const int *pair_i, const int *pair_j, int npairs, const double *c, const double *lo, const double *hi,
                       const double *bl, const double *bu, const double *rho, int B, int simplex_row0, int simplex_rows,
                       const double *simplex_w, int max_iter, double feastol, double reltol, double abstol, double *z_out,
                       double *info_out);
/*
 * fir_ap_cvx as ONE call, specification in -> taps out (device-side problem assembly).  Replaces
 * fir_ap_cvx.m:44-202 for B designs of one order n: grid (:44-48), band masks and interpolated bounds (:51-82), squared /
 * floored power bounds (:103-120), stop rows (:125), peak radii (:166-168), the solve (:160-169) and h = fmp2(r) (:185-202).
 * The union of the designs' grids is built on the host (one sorted vector); every per-design quantity is computed by kernels
 * straight into the solver's arrays (bit-identical to the host assembly of the Python mirror), the batch is solved by the
 * interior-point method of mbrf_fir_ipm_solve, and the spectral factors are taken on the device from the solutions.
 *   f [B x 2 nband] band edges as the reference takes them (fractions of pi), a [B x 2 nband], d [B x nband], obj [B], peak [B];
 *   oversamp: 15 (fir_ap_cvx.m:45);  max_iter / feastol / reltol / abstol as mbrf_fir_ipm_solve (0 = defaults).
 *   x_out [B x (2n-1)] solutions (may be NULL), h_re / h_im [B x n] minimum-phase taps (both NULL = skip; rows of designs that
 *   did not end with status 1 are meaningless), info_out [B x 8] as mbrf_fir_ipm_solve, rows_out[2] (may be NULL): grid rows
 *   of the union and rows of the stop block.  Host pointers, row-major, design index slowest.
 */
int mbrf_fir_ap_solve(int n, int nband, const double *f, const double *a, const double *d, const double *obj, const double *peak,
                      int B, int oversamp, int max_iter, double feastol, double reltol, double abstol, double *x_out, double *h_re,
                      double *h_im, double *info_out, int *rows_out);
/* The assembly alone: the arrays mbrf_fir_ap_solve hands to its solver, in the layout of mbrf_fir_ipm_solve / mbrf_fir_pdhg_solve
 * ([dim x B], design index fastest): w_row [M], lo / hi [M x B], c / bl / bu [(2n-1) x B], rho [(n-1) x B], ct [B] (the stop-band
 * weights), M = rows_out[0] + rows_out[1] (grid rows of the union + rows of the stop block, which starts at rows_out[0]).
 * Call with lo_out == NULL first to learn rows_out and size the arrays. */
int mbrf_fir_ap_assemble(int n, int nband, const double *f, const double *a, const double *d, const double *obj, const double *peak,
                         int B, int oversamp, int *rows_out, double *w_row_out, double *lo_out, double *hi_out, double *c_out,
                         double *bl_out, double *bu_out, double *rho_out, double *ct_out);
int mbrf_ipm_padded_sizes(int M, int N, int B, int *Mp, int *Np, int *Bp);
/* 0: precision of the Newton systems (0 fp64, 1 double-double, 2 auto = double-double once mu < switch * mu0; default 2),
 * 1: that switch (default 1e-3), 2: refinement steps per solve in the double-double phase (default 1), 3: trace the first `value`
 * designs on stderr, 4: refinement steps per solve in the fp64 phase (default 0), 5: the factorisation is split over many CTAs
 * per matrix when at most `value` designs of a batch are still running (default -1 = two thirds of the SMs, 0 = never) */
int mbrf_ipm_set_option(int which, double value);
/* Diagnostic: the batched Cholesky of the interior-point solver alone -- B matrices of order nv, double-double (use_dd) or
 * fp64, timed with CUDA events over `reps` launches (mean ms per launch): the kernel bench.py's solver roofline is quoted on. */
int mbrf_ipm_cholesky_bench(int nv, int B, int use_dd, int reps, float *ms);

/* ------------------------------------------------------------------------------------------------
 * Batched minimum-phase spectral factorisation: hmp = fmp2(h) of fir_ap_cvx.m:262-283 (with fftc :253-255 and
 * mag2mp :292-303), the step after the solve in fir_ap_cvx.m:185-202.  B sequences of odd length 2n-1 in, the n taps of
 * the minimum-phase factor out; complex data as split re / im planes like the MEX API (r_im NULL = real input);
 * row-major [B x (2n-1)] and [B x n].  n <= mbrf_fmp2_max_taps() (512: the padded transform lives in shared memory).
 * ------------------------------------------------------------------------------------------------ */
int mbrf_fmp2_max_taps(void);
int mbrf_fmp2_batch(const double *r_re, const double *r_im, int n, int B, double *h_re, double *h_im);   /* host pointers */
unsigned long long mbrf_fmp2_workspace_bytes(int n);
int mbrf_fmp2_batch_device(const double *r_re, const double *r_im, int n, int B, double *h_re, double *h_im,
                           void *workspace, void *stream);                                                 /* device pointers */

/* ------------------------------------------------------------------------------------------------
 * Batched inverse SLR transform (the step after the FIR design, dzrf_mb.m:239-240): aca = b2a(bc) of rf_tools/b2a.m:13-28
 * (n = 2^k <= 1024: radix-2 transform of length 8n; any other n <= 512: Bluestein) and rf = ab2rf(ac, bc) of rf_tools/ab2rf.m:12-26 (n <= 2048).
 * Host pointers, complex data as split re / im planes (imaginary inputs may be NULL), row-major [B x n].
 * ------------------------------------------------------------------------------------------------ */
int mbrf_b2a_batch(const double *b_re, const double *b_im, int n, int B, double *a_re, double *a_im);
int mbrf_ab2rf_batch(const double *a_re, const double *a_im, const double *b_re, const double *b_im, int n, int B,
                     double *rf_re, double *rf_im);

/* ------------------------------------------------------------------------------------------------
 * Batched zero flipping: the loop of fir_flip_zero.m:66-99 (called by fir_ap.m:199-208, fir_qp.m:136-146, dzrf_mb.m:225).
 * Z = roots(h) (nroots = N-1 zeros, split planes, z_im NULL = all real) and the passband zeros idx_pb (0-based indices into
 * Z, fir_flip_zero.m:28) stay the caller's work, as does the choice of flip patterns mask [nmask x n_pb] (row-major,
 * 1 = reflect that passband zero about the unit circle, :112-117; MATLAB's n_pb-by-Num `mask` is exactly this layout).
 * Per pattern the polynomial is expanded in the order of Z like poly() (:70), scaled by hsum / sum (hsum = sum(h), :71),
 * and its power sum|h|^2 and peak max|h| are recorded (:74-75); the pattern with the smallest peak (first one on ties, as
 * min(), :96) is returned: best (0-based), h [N].  Optional outputs (may be NULL): peak [nmask], power [nmask],
 * h_all [nmask x N] (the reference's h_array).  Host pointers.  N <= mbrf_flip_zero_max_taps().
 * ------------------------------------------------------------------------------------------------ */
int mbrf_flip_zero_max_taps(void);
int mbrf_flip_zero_batch(const double *z_re, const double *z_im, int nroots, const int *idx_pb, int n_pb,
                         const unsigned char *mask, int nmask, double hsum_re, double hsum_im, int *best, double *h_re,
                         double *h_im, double *peak, double *power, double *h_all_re, double *h_all_im);

#ifdef __cplusplus
}
#endif
#endif /* MBRF_H */
