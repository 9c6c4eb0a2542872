// Forward Shinnar-Le Roux (Cayley-Klein) recursion on B200: one position per thread,
// fp64 spinor (alpha, beta) in registers, rf/gradient table streamed through shared
// memory with TMA bulk copies.
//
// Replaces abrot + the (x,y) loops of rf_tools/mex5/abrx.c:67-115, and through the
// `convention` argument rf_tools/abrm.m:46-60 and rf_tools/abr.m:34.
//
// Per sample the reference forms phi = sqrt(cg^2+|rf|^2), n = (rf, cg)/phi (3 divides),
// cos(phi/2), sin(phi/2).  Only  al = (C, cg*S)  and  be = (rfq*S, rfi*S)  are needed, with
// C = cos(phi/2) and S = sin(phi/2)/phi entire functions of u = phi^2 (rot_poly.h), so the
// sqrt, the divides and the sincos disappear; the phi == 0 branch of abrx.c:94-98 is the
// u = 0 value of the same polynomials.  Update (abrx.c:103-111):
//        a <- al*a - be*conj(b)        b <- al*b + be*conj(a)
#include "common.h"
#include "hostpipe.h"
#include "rot_coeffs.cuh"
#include <cstring>

namespace mbrf {
namespace slr {

constexpr int TT = 128;
constexpr int SD = 6;  // rfr, rfi, c=|rf|^2, gx, gy, pad   (48 B, keeps 16-B alignment)
constexpr int NBUF = 2;
constexpr int BLOCK = 128;
constexpr int WS_HEADER = 8;
enum { E_RFR = 0, E_RFI, E_C, E_GX, E_GY };
enum { B_C = 0, B_GX, B_GY, B_VARG };   // header: bounds of |rf|^2, |gx|, |gy|; B_VARG != 0 when the gradient varies over the pulse

__global__ void slr_prep_kernel(const double *__restrict__ rfr, const double *__restrict__ rfi,
                                const double *__restrict__ gx, const double *__restrict__ gy, int ns,
                                double *__restrict__ ws)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ns) return;
    double *e = ws + WS_HEADER + (size_t)k * SD;
    const double r = rfr[k], q = rfi ? rfi[k] : 0.0;
    e[E_RFR] = r;
    e[E_RFI] = q;
    e[E_C] = r * r + q * q;
    e[E_GX] = gx[k];
    e[E_GY] = gy ? gy[k] : 0.0;
    e[5] = 0.0;
    auto amax = [](double *addr, double v) {
        atomicMax(reinterpret_cast<unsigned long long *>(addr),
                  static_cast<unsigned long long>(__double_as_longlong(fabs(v))));
    };
    amax(ws + B_C, e[E_C]);
    amax(ws + B_GX, e[E_GX]);
    amax(ws + B_GY, e[E_GY]);
    if (gx[k] != gx[0] || (gy && gy[k] != gy[0])) amax(ws + B_VARG, 1.0);
}

struct Params {
    const double *ws;
    int ns;
    const double *x, *y;  // y may be null
    int nx, ny;
    long long pos0, npos;
    double *ar, *ai, *br, *bi;
    int convention;
};

struct State {
    double x, y;
    double cg, cg2;   // constant-gradient case (the default g = 2 pi / N of abr.m:25 / abrm.m:28): x*gx + y*gy and its square
    double a0, a1, b0, b1;
    bool degenerate;  // some sample had phi == 0 (abrm.m has no guard there: 0/0)
};

template <bool HAVE_Y, int TIER, bool TRACK_ZERO, bool CONSTG>
__device__ __forceinline__ void slr_step(const double *__restrict__ e, State &s)
{
    const double2 rf = *reinterpret_cast<const double2 *>(e + E_RFR);
    const double2 cg2 = *reinterpret_cast<const double2 *>(e + E_C);  // c, gx
    double cg, u;
    if (CONSTG) {
        cg = s.cg;
        u = s.cg2 + cg2.x;
    } else {
        cg = s.x * cg2.y;                                             // abrx.c:88
        if (HAVE_Y) cg = fma(s.y, e[E_GY], cg);                       // abrx.c:89
        u = fma(cg, cg, cg2.x);                                       // abrx.c:93, squared
    }
    if (TRACK_ZERO) s.degenerate |= (u == 0.0);
    double w, s2;
    rot_coeffs<TIER>(u, w, s2);
    const double sp = 0.5 * s2;
    const double al1 = cg * sp, be0 = rf.y * sp, be1 = rf.x * sp;     // abrx.c:100-101
    const double a0 = s.a0, a1 = s.a1, b0 = s.b0, b1 = s.b1;
    // b' = al*b + be*conj(a)      (abrx.c:103-104)
    s.b0 = fma(be1, a1, fma(be0, a0, fma(-al1, b1, w * b0)));
    s.b1 = fma(-be0, a1, fma(be1, a0, fma(al1, b0, w * b1)));
    // a' = al*a - be*conj(b)      (abrx.c:106-109)
    s.a0 = fma(-be1, b1, fma(-be0, b0, fma(-al1, a1, w * a0)));
    s.a1 = fma(be0, b1, fma(-be1, b0, fma(al1, a0, w * a1)));
}

template <bool HAVE_Y, int SPT, int TIER, bool TRACK_ZERO>
__device__ __forceinline__ void run_tile(const double *__restrict__ tile, int n, State (&st)[SPT], bool constg)
{
    if (constg) {
#pragma unroll 4
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j < SPT; ++j) slr_step<HAVE_Y, TIER, TRACK_ZERO, true>(tile + i * SD, st[j]);
        }
    } else {
#pragma unroll 2
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j < SPT; ++j) slr_step<HAVE_Y, TIER, TRACK_ZERO, false>(tile + i * SD, st[j]);
        }
    }
}

template <bool HAVE_Y, int SPT, bool TRACK_ZERO>
__global__ void __launch_bounds__(BLOCK) slr_kernel(const Params p)
{
    __shared__ __align__(128) double tiles[NBUF][TT * SD];
    __shared__ __align__(8) uint64_t full[NBUF];
    const int tid = threadIdx.x;
    const int ntiles = (p.ns + TT - 1) / TT;
    const long long group_pos = (long long)BLOCK * SPT;
    const long long ngroups = (p.npos + group_pos - 1) / group_pos;
    if ((long long)blockIdx.x >= ngroups) return;
    const unsigned my_groups = (unsigned)((ngroups - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const unsigned total = my_groups * (unsigned)ntiles;
    const double *__restrict__ tab = p.ws + WS_HEADER;

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NBUF; ++b) mbar_init(&full[b], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](unsigned k) {
        const int ti = (int)(k % (unsigned)ntiles);
        const int n = min(TT, p.ns - ti * TT);
        const unsigned bytes = (unsigned)n * SD * 8u;
        mbar_expect_tx(&full[k % NBUF], bytes);
        tma_bulk_g2s(tiles[k % NBUF], tab + (size_t)ti * TT * SD, bytes, &full[k % NBUF]);
    };
    if (tid == 0) {
#pragma unroll
        for (unsigned k = 0; k < NBUF; ++k)
            if (k < total) issue(k);
    }
    const double bc = p.ws[B_C], bgx = p.ws[B_GX], bgy = p.ws[B_GY];
    const bool constg = p.ns > 0 && p.ws[B_VARG] == 0.0;
    const double g0x = p.ns > 0 ? tab[E_GX] : 0.0, g0y = p.ns > 0 ? tab[E_GY] : 0.0;
    // abrm(rf,g,x,y) == (a_abrx(-x,-y), conj(b_abrx(-x,-y)))  (abrm.m:51-55 vs abrx.c:100-109)
    const double sign = p.convention == MBRF_SLR_ABRM ? -1.0 : 1.0;

    unsigned k = 0;
    for (long long g = blockIdx.x; g < ngroups; g += gridDim.x) {
        State st[SPT];
        bool active[SPT];
        long long lp[SPT];
        int tier = TIER_TINY;
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
            lp[j] = g * group_pos + (long long)j * BLOCK + tid;
            active[j] = lp[j] < p.npos;
            const long long q = p.pos0 + (active[j] ? lp[j] : 0);  // q = ix + iy*nx  (abrx.c:73)
            const long long iy = q / p.nx;
            st[j].x = sign * p.x[q - iy * p.nx];
            st[j].y = (HAVE_Y && p.y) ? sign * p.y[iy] : 0.0;
            st[j].a0 = 1.0; st[j].a1 = 0.0; st[j].b0 = 0.0; st[j].b1 = 0.0;  // abrx.c:71
            st[j].degenerate = false;
            st[j].cg = HAVE_Y ? fma(st[j].y, g0y, st[j].x * g0x) : st[j].x * g0x;   // same operations as the per-step form
            st[j].cg2 = st[j].cg * st[j].cg;
            const double cgb = fabs(st[j].x) * bgx + fabs(st[j].y) * bgy;
            const double ub = fma(cgb, cgb, bc);
            const int tj = rot_tier(ub);
            tier = max(tier, active[j] ? tj : TIER_TINY);
        }
        tier = __reduce_max_sync(0xffffffffu, tier);
        for (int ti = 0; ti < ntiles; ++ti, ++k) {
            const int n = min(TT, p.ns - ti * TT);
            const double *tile = tiles[k % NBUF];
            mbar_wait(&full[k % NBUF], (k / NBUF) & 1u);
            switch (tier) {
            case TIER_TINY: run_tile<HAVE_Y, SPT, TIER_TINY, TRACK_ZERO>(tile, n, st, constg); break;
            case TIER_SMALL: run_tile<HAVE_Y, SPT, TIER_SMALL, TRACK_ZERO>(tile, n, st, constg); break;
            case TIER_MED: run_tile<HAVE_Y, SPT, TIER_MED, TRACK_ZERO>(tile, n, st, constg); break;
            case TIER_BIG: run_tile<HAVE_Y, SPT, TIER_BIG, TRACK_ZERO>(tile, n, st, constg); break;
            default: run_tile<HAVE_Y, SPT, TIER_ANY, TRACK_ZERO>(tile, n, st, constg); break;
            }
            __syncthreads();
            if (tid == 0 && k + NBUF < total) issue(k + NBUF);
        }
#pragma unroll
        for (int j = 0; j < SPT; ++j)
            if (active[j]) {
                double a0 = st[j].a0, a1 = st[j].a1, b0 = st[j].b0, b1 = st[j].b1;
                if (p.convention == MBRF_SLR_ABRM) {
                    b1 = -b1;
                    if (TRACK_ZERO && st[j].degenerate) a0 = a1 = b0 = b1 = __longlong_as_double(0x7ff8000000000000LL);
                } else if (p.convention == MBRF_SLR_ABR) {  // abr.m:34  b = -conj(b)
                    b0 = -b0;
                }
                p.ar[lp[j]] = a0; p.ai[lp[j]] = a1; p.br[lp[j]] = b0; p.bi[lp[j]] = b1;
            }
    }
}

template <typename K>
static int launch(K kernel, const Params &p, int spt, cudaStream_t stream)
{
    const long long group = (long long)BLOCK * spt;
    const long long ngroups = (p.npos + group - 1) / group;
    if (ngroups == 0) return MBRF_OK;
    // one group of BLOCK*SPT positions per CTA: the hardware block scheduler balances the SMs (see bloch.cu)
    const long long grid = ngroups;
    kernel<<<(unsigned)grid, BLOCK, 0, stream>>>(p);
    MBRF_LAUNCH_CHECK();
    return MBRF_OK;
}


}  // namespace slr
}  // namespace mbrf

using namespace mbrf;
using namespace mbrf::slr;

extern "C" {

unsigned long long mbrf_abr_workspace_bytes(int ns)
{
    if (ns < 0) ns = 0;
    return (unsigned long long)(WS_HEADER + (size_t)ns * SD) * sizeof(double);
}

int mbrf_abr_device(const double *rfr, const double *rfi, const double *gx, const double *gy, int ns,
                    const double *x, int nx, const double *y, int ny, int convention, long long pos0,
                    long long npos, double *alpha_r, double *alpha_i, double *beta_r, double *beta_i,
                    void *workspace, void *stream)
{
    if (int rc = require_device()) return rc;
    if (convention < 0 || convention > 2) { set_error("abr: unknown convention %d", convention); return MBRF_EINVAL; }
    if (!y) ny = 1;
    if (ns < 0 || nx < 0 || ny < 0 || pos0 < 0 || npos < 0 || pos0 + npos > (long long)nx * ny) {
        set_error("abr: bad sizes ns=%d nx=%d ny=%d pos0=%lld npos=%lld", ns, nx, ny, pos0, npos);
        return MBRF_EINVAL;
    }
    if (npos == 0) return MBRF_OK;
    if (!rfr || !gx || !x || !alpha_r || !alpha_i || !beta_r || !beta_i || !workspace) {
        set_error("abr: NULL required pointer");
        return MBRF_EINVAL;
    }
    if (((uintptr_t)workspace & 15) != 0) { set_error("abr: workspace must be 16-byte aligned (TMA bulk copies)"); return MBRF_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    double *ws = (double *)workspace;
    MBRF_CUDA(cudaMemsetAsync(ws, 0, WS_HEADER * sizeof(double), st));
    if (ns > 0) {
        slr_prep_kernel<<<(ns + 255) / 256, 256, 0, st>>>(rfr, rfi, gx, gy, ns, ws);
        MBRF_LAUNCH_CHECK();
    }
    Params p;
    p.ws = ws; p.ns = ns; p.x = x; p.y = y; p.nx = nx; p.ny = ny; p.pos0 = pos0; p.npos = npos;
    p.ar = alpha_r; p.ai = alpha_i; p.br = beta_r; p.bi = beta_i; p.convention = convention;
    const bool have_y = gy && y;
    const bool track = convention == MBRF_SLR_ABRM;
    if (have_y) {
        if (track) return launch(slr_kernel<true, 2, true>, p, 2, st);
        return launch(slr_kernel<true, 2, false>, p, 2, st);
    }
    if (track) return launch(slr_kernel<false, 2, true>, p, 2, st);
    return launch(slr_kernel<false, 2, false>, p, 2, st);
}

int mbrf_abr(const double *rfr, const double *rfi, const double *gx, const double *gy, int ns, const double *x,
             int nx, const double *y, int ny, int convention, double *alpha_r, double *alpha_i, double *beta_r,
             double *beta_i)
{
    if (int rc = require_device()) return rc;
    if (!y) ny = 1;
    if (ns < 0 || nx < 0 || ny < 0) { set_error("abr: negative size"); return MBRF_EINVAL; }
    const long long npos = (long long)nx * ny;
    if (npos == 0) return MBRF_OK;
    if (!rfr || !gx || !x || !alpha_r || !alpha_i || !beta_r || !beta_i) { set_error("abr: NULL required pointer"); return MBRF_EINVAL; }
    // positions spread over mbrf_set_fanout() devices, chunks pipelined through a pinned ring (hostpipe.h)
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t nsb = up((size_t)ns * 8), nxb = up((size_t)nx * 8), nyb = up((size_t)ny * 8);
    double *outs[4] = {alpha_r, alpha_i, beta_r, beta_i};
    hostpipe::Desc d;
    d.in_bytes = nsb * 4 + nxb + nyb;
    d.pack = [&](char *hp) {
        memcpy(hp, rfr, (size_t)ns * 8);
        if (rfi) memcpy(hp + nsb, rfi, (size_t)ns * 8);
        memcpy(hp + 2 * nsb, gx, (size_t)ns * 8);
        if (gy) memcpy(hp + 3 * nsb, gy, (size_t)ns * 8);
        memcpy(hp + 4 * nsb, x, (size_t)nx * 8);
        if (y) memcpy(hp + 4 * nsb + nxb, y, (size_t)ny * 8);
    };
    d.ws_bytes = mbrf_abr_workspace_bytes(ns);
    d.ncomp = 4;
    d.host_out = outs;
    d.item_doubles = 1;
    d.items = npos;
    d.launch = [&](cudaStream_t st, const char *dp, char *d_ws, long long p0, long long n, const double *const *,
                   double *const *dout) -> int {
        return mbrf_abr_device((const double *)dp, rfi ? (const double *)(dp + nsb) : nullptr, (const double *)(dp + 2 * nsb),
                               gy ? (const double *)(dp + 3 * nsb) : nullptr, ns, (const double *)(dp + 4 * nsb), nx,
                               y ? (const double *)(dp + 4 * nsb + nxb) : nullptr, ny, convention, p0, n, dout[0], dout[1],
                               dout[2], dout[3], d_ws, st);
    };
    return hostpipe::run(d);
}

}  // extern "C"
