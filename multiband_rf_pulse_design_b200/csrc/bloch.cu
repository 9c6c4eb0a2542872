// Bloch simulation on B200: one spin (or two) per thread, fp64 registers, the per-sample
// waveform table streamed through shared memory with TMA 1-D bulk copies.
//
// Replaces blochsim / blochsimfz / calcrotmat of the reference
// (/root/reference/bloch_simulation/blochC.c:171-234, :283-418, :422-511; blochH.c is the
// same code with another GAMMA).  The arithmetic is re-derived, not transcribed:
//
//  * Everything that depends only on the time sample is hoisted into a 10-double table
//    entry built once per call by `bloch_prep_kernel` (blochC.c:330-333, :350-353, :460-464):
//        rx = -b1r*gamma*dt   ry = +b1i*gamma*dt   c = rx^2+ry^2   dtn = -TWOPI*dt
//        e1 = exp(-dt/T1)     e2 = exp(-dt/T2)     rec = 1-e1
//        gxd = -gx*dt  gyd = -gy*dt  gzd = -gz*dt
//    so per spin-step   rz = dtn*df + gxd*(gamma*dx) + gyd*(gamma*dy) + gzd*(gamma*dz)
//    and                u  = phi^2 = rz^2 + c.
//  * calcrotmat's sqrt, divide, sin and cos (blochC.c:180,201,202) collapse into two
//    polynomials in u: C(u)=cos(phi/2) and S2(u)=2 sin(phi/2)/phi are entire functions of
//    u (rot_poly.h).  A per-spin bound on u picks the polynomial tier outside the time
//    loop; only spins that can exceed phi = 2*pi per sample take the sqrt/sincos path.
//  * The Cayley-Klein matrix of blochC.c:224-232 is the rotation by the unit quaternion
//    q = (C, n*S2/2); it is applied as  M += C*t + (v x t)/2,  t = v x M,  v = n*S2
//    (18 flops) instead of forming nine matrix entries and a 3x3 product (55 flops).
//  * Decay is the diagonal it is (blochC.c:347-361 multiplies by a matrix with six zeros).
//
// Work per spin-step on the common path (1 gradient axis, |phi| <= pi/2): 42 FP64
// instructions, all on the FP64 pipe, no SFU, no divide.  The kernel is FP64-pipe bound;
// HBM traffic is 56 B per spin per call (df/dp in, M out).
#include "common.h"
#include "rot_coeffs.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace mbrf {
namespace bloch {

constexpr int TT = 128;    // time samples per shared-memory tile
constexpr int SD = 10;     // doubles per table entry
constexpr int NBUF = 2;    // tiles in flight
constexpr int BLOCK = 128; // threads per CTA
constexpr int WS_HEADER = 8;  // doubles of bounds in front of the table

enum { E_RX = 0, E_RY, E_C, E_REC, E_E1, E_E2, E_DTN, E_GXD, E_GYD, E_GZD };  // 16-byte pairs: (rx,ry)(c,rec)(e1,e2)(dtn,gxd)(gyd,gzd)
enum { B_C = 0, B_DTN, B_GXD, B_GYD, B_GZD, B_NONUNIFORM };

// ---------------------------------------------------------------------------
// prep: waveform -> table + bounds
// ---------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v)
{
    // non-negative doubles (and NaN, which sorts above +inf) order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long *>(addr),
              static_cast<unsigned long long>(__double_as_longlong(fabs(v))));
}

__global__ void bloch_prep_kernel(const double *__restrict__ b1r, const double *__restrict__ b1i,
                                  const double *__restrict__ gx, const double *__restrict__ gy,
                                  const double *__restrict__ gz, const double *__restrict__ dt, int ntime,
                                  double t1, double t2, double gamma, double *__restrict__ ws)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntime) return;
    double *e = ws + WS_HEADER + (size_t)t * SD;
    const double d = dt[t];
    const double rx = (-b1r[t] * gamma * d);              // blochC.c:332
    const double ry = b1i ? (b1i[t] * gamma * d) : 0.0;   // blochC.c:333
    const double e1 = exp(-d / t1), e2 = exp(-d / t2);    // blochC.c:462-463
    e[E_RX] = rx;
    e[E_RY] = ry;
    e[E_C] = rx * rx + ry * ry;
    e[E_DTN] = -MBRF_TWOPI * d;
    e[E_E1] = e1;
    e[E_E2] = e2;
    e[E_REC] = 1 - e1;                                   // blochC.c:350
    e[E_GXD] = gx ? -gx[t] * d : 0.0;
    e[E_GYD] = gy ? -gy[t] * d : 0.0;
    e[E_GZD] = gz ? -gz[t] * d : 0.0;
    atomic_max_nonneg(ws + B_C, e[E_C]);
    atomic_max_nonneg(ws + B_DTN, e[E_DTN]);
    atomic_max_nonneg(ws + B_GXD, e[E_GXD]);
    atomic_max_nonneg(ws + B_GYD, e[E_GYD]);
    atomic_max_nonneg(ws + B_GZD, e[E_GZD]);
    // Constant time step and constant gradients (every call site of the reference: scalar tp, G = 0
    // or a constant slice-select gradient) make rz a per-spin constant; flag anything else.
    const double d0 = dt[0];
    const bool same = e[E_DTN] == -MBRF_TWOPI * d0 && e[E_GXD] == (gx ? -gx[0] * d0 : 0.0) &&
                      e[E_GYD] == (gy ? -gy[0] * d0 : 0.0) && e[E_GZD] == (gz ? -gz[0] * d0 : 0.0);
    if (!same) ws[B_NONUNIFORM] = 1.0;
}

// ---------------------------------------------------------------------------
// rotation coefficients  w = cos(phi/2),  s2 = 2 sin(phi/2)/phi  from u = phi^2
// ---------------------------------------------------------------------------
// M <- R(q) M  with q = (w, v/2), v = (vx,vy,vz) = n*s2.
__device__ __forceinline__ void rotate(double w, double vx, double vy, double vz, double &mx, double &my,
                                       double &mz)
{
    const double tx = fma(vy, mz, -(vz * my));
    const double ty = fma(vz, mx, -(vx * mz));
    const double tz = fma(vx, my, -(vy * mx));
    const double cx = fma(vy, tz, -(vz * ty));
    const double cy = fma(vz, tx, -(vx * tz));
    const double cz = fma(vx, ty, -(vy * tx));
    mx = fma(0.5, cx, fma(w, tx, mx));
    my = fma(0.5, cy, fma(w, ty, my));
    mz = fma(0.5, cz, fma(w, tz, mz));
}

struct Params {
    const double *ws;  // bounds + table
    int ntime;
    const double *df;
    const double *dx, *dy, *dz;  // dy/dz may be null
    int npos, nfreq;
    long long spin0, nspins;
    const double *m0x, *m0y, *m0z;  // null => (0,0,1)
    int m0_stride;
    double *mx, *my, *mz;
    double gamma;
    const double *b1scale;  // sweep only
};

struct Spin {
    double px, py, pz, fz;  // gamma*position, off-resonance
    double sc, sc2;         // b1 scale (sweep)
    double rz0, rz2;        // rz and rz^2 when they do not depend on the sample (CRZ)
    double mx, my, mz;
    double A[9], b[3];      // steady-state propagators (modes 1, 3); column-major like the reference
};

template <int NG, int TIER, bool SWEEP, bool STEADY, bool CRZ>
__device__ __forceinline__ void spin_step(const double *__restrict__ e, Spin &s)
{
    const double2 rxy = *reinterpret_cast<const double2 *>(e + E_RX);
    const double2 cr = *reinterpret_cast<const double2 *>(e + E_C);    // c, rec
    const double2 e12 = *reinterpret_cast<const double2 *>(e + E_E1);
    const double rec = cr.y;
    double rz, u, w, s2;
    if (CRZ) {
        rz = s.rz0;
        if (SWEEP) u = fma(cr.x, s.sc2, s.rz2);
        else u = s.rz2 + cr.x;
    } else {
        const double2 dg = *reinterpret_cast<const double2 *>(e + E_DTN);  // dtn, gxd
        rz = dg.x * s.fz;
        if (NG >= 1) rz = fma(dg.y, s.px, rz);
        if (NG >= 2) {
            const double2 gyz = *reinterpret_cast<const double2 *>(e + E_GYD);
            rz = fma(gyz.x, s.py, rz);
            rz = fma(gyz.y, s.pz, rz);
        }
        if (SWEEP) u = fma(rz, rz, cr.x * s.sc2);
        else u = fma(rz, rz, cr.x);
    }
    rot_coeffs<TIER>(u, w, s2);
    const double sxy = SWEEP ? s2 * s.sc : s2;
    const double vx = rxy.x * sxy, vy = rxy.y * sxy, vz = rz * s2;
    if (!STEADY) {
        rotate(w, vx, vy, vz, s.mx, s.my, s.mz);
        s.mx *= e12.y;                       // blochC.c:347-365, the diagonal it is
        s.my *= e12.y;
        s.mz = fma(s.mz, e12.x, rec);
    } else {                                 // blochC.c:336-340, :355-359
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            rotate(w, vx, vy, vz, s.A[3 * c], s.A[3 * c + 1], s.A[3 * c + 2]);
            s.A[3 * c] *= e12.y;
            s.A[3 * c + 1] *= e12.y;
            s.A[3 * c + 2] *= e12.x;
        }
        rotate(w, vx, vy, vz, s.b[0], s.b[1], s.b[2]);
        s.b[0] *= e12.y;
        s.b[1] *= e12.y;
        s.b[2] = fma(s.b[2], e12.x, rec);
    }
}

// M = inv(I - A) b by adjugate / determinant, as blochC.c:406-415 intends (its :132 is UB).
__device__ void steady_state(Spin &s)
{
    double m[9], adj[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = ((i % 4 == 0) ? 1.0 : 0.0) - s.A[i];
    double det = m[0] * m[4] * m[8];
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
#pragma unroll
    for (int i = 0; i < 9; ++i) adj[i] = adj[i] / det;
    s.mx = adj[0] * s.b[0] + adj[3] * s.b[1] + adj[6] * s.b[2];
    s.my = adj[1] * s.b[0] + adj[4] * s.b[1] + adj[7] * s.b[2];
    s.mz = adj[2] * s.b[0] + adj[5] * s.b[1] + adj[8] * s.b[2];
}

// Mode bit 1 (record every sample, blochC.c:381-391) writes element [t + ntime*spin]: consecutive threads are
// ntime doubles apart, so direct stores touch one 8-byte word per 32-byte sector (measured 0.6 TB/s).  Each warp
// therefore stages RT = 8 samples of its 32 spins per component in shared memory and writes them transposed:
// 8 consecutive samples of one spin = 64 contiguous bytes, i.e. whole sectors.
constexpr int RT = 8;              // samples staged per flush
constexpr int RLD = RT + 1;        // padded row: conflict-free column access

template <int SPT>
__device__ __forceinline__ void record_flush(double (*stage)[32 * RLD], int count, int t_first, long long warp_ls0,
                                             long long nspins, int ntime, double *mx, double *my, double *mz, int lane)
{
    __syncwarp();
    double *const outs[3] = {mx, my, mz};
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
        const long long base = warp_ls0 + (long long)j * BLOCK;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double *st = stage[j * 3 + c];
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                const int idx = r * 32 + lane, sl = idx / RT, tt = idx % RT;
                if (tt < count && base + sl < nspins)
                    outs[c][(size_t)(base + sl) * ntime + t_first + tt] = st[sl * RLD + tt];
            }
        }
    }
    __syncwarp();
}

// One tile of n samples for this thread's SPT spins.  RECORD: write every sample (mode bit 1).
template <int NG, int SPT, int TIER, bool SWEEP, bool STEADY, bool RECORD, bool CRZ>
__device__ __forceinline__ void run_tile(const double *__restrict__ tile, int n, Spin (&sp)[SPT],
                                         const bool (&active)[SPT], double *const (&ox)[SPT],
                                         double *const (&oy)[SPT], double *const (&oz)[SPT], int t0,
                                         double (*stage)[32 * RLD], long long warp_ls0, long long nspins, int ntime,
                                         double *mx, double *my, double *mz)
{
    const int lane = threadIdx.x & 31;
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
        const double *e = tile + i * SD;
#pragma unroll
        for (int j = 0; j < SPT; ++j) spin_step<NG, TIER, SWEEP, STEADY, CRZ>(e, sp[j]);
        if (RECORD) {
#pragma unroll
            for (int j = 0; j < SPT; ++j) {   // blochC.c:381-391
                stage[j * 3 + 0][lane * RLD + (i % RT)] = sp[j].mx;
                stage[j * 3 + 1][lane * RLD + (i % RT)] = sp[j].my;
                stage[j * 3 + 2][lane * RLD + (i % RT)] = sp[j].mz;
            }
            if ((i % RT) == RT - 1 || i == n - 1)
                record_flush<SPT>(stage, (i % RT) + 1, t0 + i - (i % RT), warp_ls0, nspins, ntime, mx, my, mz, lane);
        }
    }
}

// MODE: 0 end point, 1 steady state, 2 every sample, 3 steady state then every sample.
template <int MODE, int NG, int SPT, bool SWEEP>
__global__ void __launch_bounds__(BLOCK) bloch_kernel(const Params p)
{
    __shared__ __align__(128) double tiles[NBUF][TT * SD];
    __shared__ __align__(8) uint64_t full[NBUF];
    // per-warp transpose staging for the record modes (one slot otherwise so the array is never empty)
    constexpr int NSTAGE = (MODE & 2) ? (BLOCK / 32) * SPT * 3 : 1;
    __shared__ double stage_all[NSTAGE][(MODE & 2) ? 32 * RLD : 1];
    double (*stage)[32 * RLD] = (MODE & 2) ? reinterpret_cast<double (*)[32 * RLD]>(&stage_all[(threadIdx.x >> 5) * SPT * 3][0]) : nullptr;

    const int tid = threadIdx.x;
    const int ntiles = (p.ntime + TT - 1) / TT;
    const long long group_spins = (long long)BLOCK * SPT;
    const long long ngroups = (p.nspins + group_spins - 1) / group_spins;
    if ((long long)blockIdx.x >= ngroups) return;
    const unsigned my_groups = (unsigned)((ngroups - blockIdx.x + gridDim.x - 1) / gridDim.x);
    constexpr int PASSES = (MODE == 3) ? 2 : 1;
    const unsigned total = my_groups * PASSES * (unsigned)ntiles;
    const double *__restrict__ tab = p.ws + WS_HEADER;
    const int ntout = (MODE & 2) ? p.ntime : 1;

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NBUF; ++b) mbar_init(&full[b], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // producer (thread 0): tile k of this CTA's stream -> buffer k % NBUF
    auto issue = [&](unsigned k) {
        const int ti = (int)(k % (unsigned)ntiles);
        const int n = min(TT, p.ntime - ti * TT);
        const unsigned bytes = (unsigned)n * SD * 8u;
        uint64_t *bar = &full[k % NBUF];
        mbar_expect_tx(bar, bytes);
        tma_bulk_g2s(tiles[k % NBUF], tab + (size_t)ti * TT * SD, bytes, bar);
    };
    if (tid == 0) {
#pragma unroll
        for (unsigned k = 0; k < NBUF; ++k)
            if (k < total) issue(k);
    }

    // bounds for tier selection: u <= cmax + (|fz| dtn + |px| gxd + |py| gyd + |pz| gzd)^2
    const double bc = p.ws[B_C], bdtn = p.ws[B_DTN], bgx = p.ws[B_GXD], bgy = p.ws[B_GYD], bgz = p.ws[B_GZD];
    // sample-independent rz: entry 0 of the table holds the (constant) dtn, gxd, gyd, gzd
    const bool crz = (MODE == 0 || MODE == 2) && p.ws[B_NONUNIFORM] == 0.0;
    const double dtn0 = tab[E_DTN], gxd0 = tab[E_GXD], gyd0 = tab[E_GYD], gzd0 = tab[E_GZD];

    unsigned k = 0;
    for (long long g = blockIdx.x; g < ngroups; g += gridDim.x) {
        Spin sp[SPT];
        bool active[SPT];
        double *ox[SPT], *oy[SPT], *oz[SPT];
        int tier = TIER_TINY;
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
            const long long ls = g * group_spins + (long long)j * BLOCK + tid;  // local spin
            active[j] = ls < p.nspins;
            const long long s = p.spin0 + (active[j] ? ls : 0);
            Spin &q = sp[j];
            q.sc = 1.0;
            q.sc2 = 1.0;
            q.px = q.py = q.pz = 0.0;
            if (SWEEP) {  // s = i_f + nfreq * i_s
                const long long is = s / p.nfreq;
                q.fz = p.df[s - is * p.nfreq];
                q.sc = p.b1scale[is];
                q.sc2 = q.sc * q.sc;
            } else {      // s = p + npos * f  (blochC.c:468-473)
                const long long f = s / p.npos;
                const long long pi = s - f * p.npos;
                q.fz = p.df[f];
                if (NG >= 1) q.px = p.dx[pi] * p.gamma;              // blochC.c:313
                if (NG >= 2) {
                    q.py = p.dy ? p.dy[pi] * p.gamma : 0.0;
                    q.pz = p.dz ? p.dz[pi] * p.gamma : 0.0;
                }
            }
            const size_t o = (size_t)(active[j] ? ls : 0);
            if (p.m0x) {
                q.mx = p.m0x[o * p.m0_stride];
                q.my = p.m0y[o * p.m0_stride];
                q.mz = p.m0z[o * p.m0_stride];
            } else {
                q.mx = 0.0;
                q.my = 0.0;
                q.mz = 1.0;
            }
            ox[j] = p.mx + o * ntout;
            oy[j] = p.my + o * ntout;
            oz[j] = p.mz + o * ntout;
            const double rzb = fabs(q.fz) * bdtn + fabs(q.px) * bgx + fabs(q.py) * bgy + fabs(q.pz) * bgz;
            const double ub = fma(rzb, rzb, bc * q.sc2);
            const int tj = rot_tier(ub);
            tier = max(tier, active[j] ? tj : TIER_TINY);
            q.rz0 = fma(gzd0, q.pz, fma(gyd0, q.py, fma(gxd0, q.px, dtn0 * q.fz)));
            q.rz2 = q.rz0 * q.rz0;
        }
        tier = __reduce_max_sync(0xffffffffu, tier);  // one code path per warp
        const long long warp_ls0 = g * group_spins + (tid & ~31);   // local spin of this warp's lane 0 (j = 0)

#pragma unroll
        for (int pass = 0; pass < PASSES; ++pass) {
            const bool steady = (MODE == 1) || (MODE == 3 && pass == 0);
            if (steady) {
#pragma unroll
                for (int j = 0; j < SPT; ++j) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) sp[j].A[i] = (i % 4 == 0) ? 1.0 : 0.0;
                    sp[j].b[0] = sp[j].b[1] = sp[j].b[2] = 0.0;
                }
            }
            for (int ti = 0; ti < ntiles; ++ti, ++k) {
                const int n = min(TT, p.ntime - ti * TT);
                const double *tile = tiles[k % NBUF];
                mbar_wait(&full[k % NBUF], (k / NBUF) & 1u);
#define MBRF_RUN(TIER_, CRZ_)                                                                                   \
    if (steady) run_tile<NG, SPT, TIER_, SWEEP, true, false, false>(tile, n, sp, active, ox, oy, oz, ti * TT, stage,    \
                                                                    warp_ls0, p.nspins, p.ntime, p.mx, p.my, p.mz);       \
    else run_tile<NG, SPT, TIER_, SWEEP, false, (MODE & 2) != 0, CRZ_>(tile, n, sp, active, ox, oy, oz, ti * TT, stage, \
                                                                       warp_ls0, p.nspins, p.ntime, p.mx, p.my, p.mz);
                if (MODE == 1 || MODE == 3) {
                    // steady-state modes are rare (no caller in the reference): one general path
                    MBRF_RUN(TIER_ANY, false)
                } else if (crz && tier <= TIER_MED) {
                    switch (tier) {
                    case TIER_TINY: MBRF_RUN(TIER_TINY, true) break;
                    case TIER_SMALL: MBRF_RUN(TIER_SMALL, true) break;
                    default: MBRF_RUN(TIER_MED, true) break;
                    }
                } else {
                    switch (tier) {
                    case TIER_TINY: MBRF_RUN(TIER_TINY, false) break;
                    case TIER_SMALL: MBRF_RUN(TIER_SMALL, false) break;
                    case TIER_MED: MBRF_RUN(TIER_MED, false) break;
                    case TIER_BIG: MBRF_RUN(TIER_BIG, false) break;
                    default: MBRF_RUN(TIER_ANY, false) break;
                    }
                }
#undef MBRF_RUN
                __syncthreads();  // every warp is done with this buffer
                if (tid == 0 && k + NBUF < total) issue(k + NBUF);
            }
            if (steady) {
#pragma unroll
                for (int j = 0; j < SPT; ++j) steady_state(sp[j]);
            }
        }
        if (!(MODE & 2)) {
#pragma unroll
            for (int j = 0; j < SPT; ++j)
                if (active[j]) {  // blochC.c:399-404 / :412-414
                    ox[j][0] = sp[j].mx;
                    oy[j][0] = sp[j].my;
                    oz[j][0] = sp[j].mz;
                }
        }
    }
}

// ---------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------
static int g_tune_ctas_per_sm = 0;
static int g_tune_spt = 0;

template <typename K>
static int launch_balanced(K kernel, const Params &p, int spt, cudaStream_t stream)
{
    const long long group_spins = (long long)BLOCK * spt;
    const long long ngroups = (p.nspins + group_spins - 1) / group_spins;
    if (ngroups == 0) return MBRF_OK;
    int occ = 0;
    MBRF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, BLOCK, 0));
    if (occ < 1) occ = 1;
    // Default: one group of BLOCK*SPT spins per CTA and let the hardware block scheduler balance the
    // SMs (a finished CTA is replaced at once; the tail is at most one group).  A positive tuning value
    // selects the persistent form instead: sms*ctas_per_sm CTAs striding over the groups.
    long long grid = ngroups;
    if (g_tune_ctas_per_sm > 0) {
        const int bps = g_tune_ctas_per_sm < occ ? g_tune_ctas_per_sm : occ;
        grid = (long long)sm_count() * bps;
        if (grid > ngroups) grid = ngroups;
    }
    kernel<<<(unsigned)grid, BLOCK, 0, stream>>>(p);
    MBRF_LAUNCH_CHECK();
    return MBRF_OK;
}

template <int MODE, bool SWEEP>
static int dispatch_ng_spt(const Params &p, int ng, int spt, cudaStream_t stream)
{
#define MBRF_CASE(NG_, SPT_) return launch_balanced(bloch_kernel<MODE, NG_, SPT_, SWEEP>, p, SPT_, stream)
    if constexpr (SWEEP) {
        MBRF_CASE(0, 2);
    } else if constexpr (MODE == 1 || MODE == 3) {
        MBRF_CASE(2, 1);
    } else {
        if constexpr (MODE == 0) {
            if (spt == 2) {
                if (ng == 0) MBRF_CASE(0, 2);
                if (ng == 1) MBRF_CASE(1, 2);
                MBRF_CASE(2, 2);
            }
        }
        if (ng == 0) MBRF_CASE(0, 1);
        if (ng == 1) MBRF_CASE(1, 1);
        MBRF_CASE(2, 1);
    }
#undef MBRF_CASE
}

static int run_prep(const double *b1r, const double *b1i, const double *gx, const double *gy, const double *gz,
                    const double *dt, int ntime, double t1, double t2, double gamma, double *ws,
                    cudaStream_t stream)
{
    MBRF_CUDA(cudaMemsetAsync(ws, 0, WS_HEADER * sizeof(double), stream));
    bloch_prep_kernel<<<(ntime + 255) / 256, 256, 0, stream>>>(b1r, b1i, gx, gy, gz, dt, ntime, t1, t2, gamma, ws);
    MBRF_LAUNCH_CHECK();
    return MBRF_OK;
}

}  // namespace bloch
}  // namespace mbrf

using namespace mbrf;
using namespace mbrf::bloch;

extern "C" {

int mbrf_bloch_set_tuning(int ctas_per_sm, int spins_per_thread)
{
    if (ctas_per_sm < 0 || (spins_per_thread != 0 && spins_per_thread != 1 && spins_per_thread != 2)) {
        set_error("mbrf_bloch_set_tuning: ctas_per_sm >= 0 and spins_per_thread in {0,1,2}");
        return MBRF_EINVAL;
    }
    g_tune_ctas_per_sm = ctas_per_sm;
    g_tune_spt = spins_per_thread;
    return MBRF_OK;
}

unsigned long long mbrf_bloch_workspace_bytes(int ntime)
{
    if (ntime < 0) ntime = 0;
    return (unsigned long long)(WS_HEADER + (size_t)ntime * SD) * sizeof(double);
}

int mbrf_bloch_device(const double *b1real, const double *b1imag, const double *xgrad, const double *ygrad,
                      const double *zgrad, const double *tsteps, int ntime, double t1, double t2,
                      const double *dfreq, int nfreq, const double *dxpos, const double *dypos,
                      const double *dzpos, int npos, long long spin0, long long nspins, const double *m0x,
                      const double *m0y, const double *m0z, int m0_stride, double *mx, double *my, double *mz,
                      int mode, double gamma, void *workspace, void *stream)
{
    if (int rc = require_device()) return rc;
    if (mode < 0 || mode > 3) { set_error("bloch: mode must be 0..3, got %d", mode); return MBRF_EINVAL; }
    if (ntime < 0 || nfreq < 0 || npos < 0 || nspins < 0 || spin0 < 0 ||
        spin0 + nspins > (long long)nfreq * npos) {
        set_error("bloch: bad sizes ntime=%d nfreq=%d npos=%d spin0=%lld nspins=%lld", ntime, nfreq, npos, spin0, nspins);
        return MBRF_EINVAL;
    }
    if (nspins == 0) return MBRF_OK;
    if (!b1real || !tsteps || !dfreq || !mx || !my || !mz || !workspace || (ntime > 0 && npos > 0 && !dxpos)) {
        set_error("bloch: NULL required pointer");
        return MBRF_EINVAL;
    }
    if ((m0x || m0y || m0z) && !(m0x && m0y && m0z)) { set_error("bloch: m0x,m0y,m0z must be all set or all NULL"); return MBRF_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    if (((uintptr_t)workspace & 15) != 0) { set_error("bloch: workspace must be 16-byte aligned (TMA bulk copies)"); return MBRF_EINVAL; }
    double *ws = (double *)workspace;
    if (ntime == 0) {
        // no samples: modes 0/2 leave M untouched (mode 2 has zero outputs); steady state of nothing is 0/0
        set_error("bloch: ntime == 0");
        return MBRF_EINVAL;
    }
    if (int rc = run_prep(b1real, b1imag, xgrad, ygrad, zgrad, tsteps, ntime, t1, t2, gamma, ws, st)) return rc;

    Params p;
    p.ws = ws; p.ntime = ntime; p.df = dfreq; p.dx = dxpos; p.dy = dypos; p.dz = dzpos;
    p.npos = npos; p.nfreq = nfreq; p.spin0 = spin0; p.nspins = nspins;
    p.m0x = m0x; p.m0y = m0y; p.m0z = m0z; p.m0_stride = m0_stride > 0 ? m0_stride : 1;
    p.mx = mx; p.my = my; p.mz = mz; p.gamma = gamma; p.b1scale = nullptr;
    const int ng = (ygrad || zgrad) ? 2 : (xgrad ? 1 : 0);
    int spt = g_tune_spt ? g_tune_spt : 1;
    switch (mode) {
    case 0: return dispatch_ng_spt<0, false>(p, ng, spt, st);
    case 1: return dispatch_ng_spt<1, false>(p, ng, 1, st);
    case 2: return dispatch_ng_spt<2, false>(p, ng, spt, st);
    default: return dispatch_ng_spt<3, false>(p, ng, 1, st);
    }
}

int mbrf_bloch_scale_sweep_device(const double *b1real, const double *b1imag, const double *tsteps, int ntime,
                                  double t1, double t2, const double *dfreq, int nfreq, const double *b1scale,
                                  int nscale, long long spin0, long long nspins, double *mx, double *my,
                                  double *mz, double gamma, void *workspace, void *stream)
{
    if (int rc = require_device()) return rc;
    if (ntime <= 0 || nfreq <= 0 || nscale <= 0 || nspins < 0 || spin0 < 0 ||
        spin0 + nspins > (long long)nfreq * nscale) {
        set_error("bloch sweep: bad sizes");
        return MBRF_EINVAL;
    }
    if (nspins == 0) return MBRF_OK;
    if (!b1real || !tsteps || !dfreq || !b1scale || !mx || !my || !mz || !workspace) {
        set_error("bloch sweep: NULL required pointer");
        return MBRF_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (((uintptr_t)workspace & 15) != 0) { set_error("bloch: workspace must be 16-byte aligned (TMA bulk copies)"); return MBRF_EINVAL; }
    double *ws = (double *)workspace;
    if (int rc = run_prep(b1real, b1imag, nullptr, nullptr, nullptr, tsteps, ntime, t1, t2, gamma, ws, st)) return rc;
    Params p;
    p.ws = ws; p.ntime = ntime; p.df = dfreq; p.dx = nullptr; p.dy = nullptr; p.dz = nullptr;
    p.npos = 1; p.nfreq = nfreq; p.spin0 = spin0; p.nspins = nspins;
    p.m0x = p.m0y = p.m0z = nullptr; p.m0_stride = 1;
    p.mx = mx; p.my = my; p.mz = mz; p.gamma = gamma; p.b1scale = b1scale;
    return dispatch_ng_spt<0, true>(p, 0, 2, st);
}

}  // extern "C"
