// Host-pointer entry points of the Bloch path: mbrf_blochsimfz (the reference's inner C
// ABI, blochC.c:422-426) and mbrf_bloch (the argument handling of its mexFunction,
// blochC.c:514-927).  They stage inputs through one pinned buffer, run the device path
// of bloch.cu and bring the result back; there is no CPU computation of the physics.
#include "common.h"
#include "hostpipe.h"

#include <cstdio>
#include <cstring>
#include <vector>

namespace mbrf {
namespace bloch {

static bool all_zero(const double *v, long long n)
{
    if (!v) return true;
    for (long long i = 0; i < n; ++i)
        if (v[i] != 0.0) return false;  // NaN counts as non-zero, as it must
    return true;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// m0*: host, element i at [i*m0_stride], or all NULL for (0,0,1).  Outputs: host, [t + ntout*s].
// The spins are spread over mbrf_set_fanout() devices and pipelined in chunks (hostpipe.h): pageable result arrays -- what
// mxCreateDoubleMatrix hands a MEX gateway -- are filled from a pinned ring by host threads while the GPUs simulate the
// next chunks; page-locked ones are written by the GPU directly.
static int run_host(const double *b1r, const double *b1i, const double *gx, const double *gy, const double *gz,
                    const double *dt, int ntime, double t1, double t2, const double *df, int nf,
                    const double *dx, const double *dy, const double *dz, int npos, const double *m0x,
                    const double *m0y, const double *m0z, long long m0_stride, double *mx, double *my,
                    double *mz, int mode, double gamma)
{
    if (int rc = require_device()) return rc;
    if (mode < 0 || mode > 3) { set_error("bloch: mode must be 0..3, got %d", mode); return MBRF_EINVAL; }
    if (ntime <= 0 || nf < 0 || npos < 0) { set_error("bloch: bad sizes ntime=%d nf=%d npos=%d", ntime, nf, npos); return MBRF_EINVAL; }
    const long long nspins = (long long)nf * npos;
    if (nspins == 0) return MBRF_OK;
    if (!b1r || !dt || !df || !dx || !mx || !my || !mz) { set_error("bloch: NULL required pointer"); return MBRF_EINVAL; }
    const long long ntout = (mode & 2) ? ntime : 1;

    // gradient axes that cannot contribute are dropped so the kernel skips their FMAs
    const bool x_live = !all_zero(gx, ntime) && !all_zero(dx, npos);
    const bool y_live = !all_zero(gy, ntime) && !all_zero(dy, npos);
    const bool z_live = !all_zero(gz, ntime) && !all_zero(dz, npos);
    const bool any_yz = y_live || z_live;
    const bool use_x = x_live || any_yz;

    // ---- the small inputs, packed into one block -> one H2D per device ----------------------
    const size_t nt = (size_t)ntime;
    size_t off = 0;
    auto take = [&](size_t n_doubles) { size_t o = off; off += align_up(n_doubles * sizeof(double), 64); return o; };
    const size_t o_b1r = take(nt), o_b1i = b1i ? take(nt) : 0, o_dt = take(nt);
    const size_t o_gx = (use_x && gx) ? take(nt) : 0, o_gy = (y_live) ? take(nt) : 0, o_gz = (z_live) ? take(nt) : 0;
    const size_t o_df = take((size_t)nf);
    const size_t o_dx = take((size_t)npos), o_dy = (any_yz && dy) ? take((size_t)npos) : 0,
                 o_dz = (any_yz && dz) ? take((size_t)npos) : 0;

    const bool have_m0 = m0x && m0y && m0z;
    const double *m0[3] = {m0x, m0y, m0z};
    double *outs[3] = {mx, my, mz};
    hostpipe::Desc d;
    d.in_bytes = off;
    d.pack = [&](char *hp) {
        memcpy(hp + o_b1r, b1r, nt * 8);
        if (b1i) memcpy(hp + o_b1i, b1i, nt * 8);
        memcpy(hp + o_dt, dt, nt * 8);
        if (use_x && gx) memcpy(hp + o_gx, gx, nt * 8);
        if (y_live) memcpy(hp + o_gy, gy, nt * 8);
        if (z_live) memcpy(hp + o_gz, gz, nt * 8);
        memcpy(hp + o_df, df, (size_t)nf * 8);
        memcpy(hp + o_dx, dx, (size_t)npos * 8);
        if (any_yz && dy) memcpy(hp + o_dy, dy, (size_t)npos * 8);
        if (any_yz && dz) memcpy(hp + o_dz, dz, (size_t)npos * 8);
    };
    d.ws_bytes = mbrf_bloch_workspace_bytes(ntime);
    d.ncomp = 3;
    d.host_out = outs;
    d.item_doubles = (size_t)ntout;
    d.items = nspins;
    d.ncomp_in = have_m0 ? 3 : 0;
    d.host_in = m0;
    d.in_stride = m0_stride;
    d.launch = [&](cudaStream_t st, const char *dp, char *d_ws, long long s0, long long n, const double *const *dm0,
                   double *const *dout) -> int {
        auto dptr = [&](size_t o) { return (const double *)(dp + o); };
        return mbrf_bloch_device(dptr(o_b1r), b1i ? dptr(o_b1i) : nullptr, (use_x && gx) ? dptr(o_gx) : nullptr,
                                 y_live ? dptr(o_gy) : nullptr, z_live ? dptr(o_gz) : nullptr, dptr(o_dt), ntime, t1, t2,
                                 dptr(o_df), nf, dptr(o_dx), (any_yz && dy) ? dptr(o_dy) : nullptr,
                                 (any_yz && dz) ? dptr(o_dz) : nullptr, npos, s0, n, dm0 ? dm0[0] : nullptr,
                                 dm0 ? dm0[1] : nullptr, dm0 ? dm0[2] : nullptr, 1, dout[0], dout[1], dout[2], mode, gamma,
                                 d_ws, st);
    };
    return hostpipe::run(d);
}

}  // namespace bloch
}  // namespace mbrf

using namespace mbrf;
using namespace mbrf::bloch;

extern "C" {

int mbrf_blochsimfz(const double *b1real, const double *b1imag, const double *xgrad, const double *ygrad,
                    const double *zgrad, const double *tsteps, int ntime, double t1, double t2,
                    const double *dfreq, int nfreq, const double *dxpos, const double *dypos,
                    const double *dzpos, int npos, double *mx, double *my, double *mz, int mode, double gamma)
{
    const long long ntout = (mode & 2) ? ntime : 1;
    // mx/my/mz carry the initial magnetisation at stride ntout (blochC.c:838-865) and receive the result
    return run_host(b1real, b1imag, xgrad, ygrad, zgrad, tsteps, ntime, t1, t2, dfreq, nfreq, dxpos, dypos, dzpos,
                    npos, mx, my, mz, ntout, mx, my, mz, mode, gamma);
}

int mbrf_bloch(const double *b1r, const double *b1i, int ntime, const double *gr, int ngr, const double *tp,
               int ntp, double t1, double t2, const double *df, int nf, const double *dp, int npos_m, int npos_n,
               int mode, const double *mx0, const double *my0, const double *mz0, int n_m0, double *mx,
               double *my, double *mz, int out_dims[4], double gamma)
{
    if (ntime <= 0 || !b1r) { set_error("bloch: b1 is empty"); return MBRF_EINVAL; }
    if (ngr < 0 || (ngr > 0 && !gr) || !tp || !df || nf < 0 || npos_m < 0 || npos_n < 0 ||
        ((long long)npos_m * npos_n > 0 && !dp) || !out_dims) {
        set_error("bloch: NULL / negative argument");
        return MBRF_EINVAL;
    }
    // ---- gradients (blochC.c:595-638): first ntime values are x; y and z only if present ----
    if (ngr < ntime) {
        // the reference reads ntime x-gradient samples regardless; fewer is an out-of-bounds read there
        set_error("bloch: gradient has %d samples, b1 has %d", ngr, ntime);
        return MBRF_EINVAL;
    }
    const double *gx = gr;
    const double *gy = (ngr < 2 * ntime) ? nullptr : gr + ntime;
    const double *gz = (ngr < 3 * ntime) ? nullptr : gr + 2 * (size_t)ntime;
    if (ngr != ntime && ngr != 2 * ntime && ngr != 3 * ntime)
        printf("Gradient length differs from B1 length\n");  // blochC.c:637-638, same text
    // ---- time (blochC.c:660-681) ----
    std::vector<double> dt((size_t)ntime);
    if (ntp == 1) {
        for (int i = 0; i < ntime; ++i) dt[i] = tp[0];
    } else if (ntp != ntime) {
        // reference: prints "Time-point length differs from B1 length" and then indexes tp[0..ntime)
        set_error("bloch: time vector has %d entries, b1 has %d", ntp, ntime);
        return MBRF_EINVAL;
    } else {
        bool allpos = true;  // times2intervals, blochC.c:249-276
        double last = 0.0;
        for (int i = 0; i < ntime; ++i) {
            dt[i] = tp[i] - last;
            last = tp[i];
            if (dt[i] <= 0) allpos = false;
        }
        if (!allpos) memcpy(dt.data(), tp, sizeof(double) * (size_t)ntime);  // they were intervals
    }
    // ---- positions (blochC.c:701-758) ----
    int npos;
    const double *dx = dp, *dy = nullptr, *dz = nullptr;
    if (npos_n == 3) { npos = npos_m; dy = dx + npos; dz = dy + npos; }
    else if (npos_n == 2) { npos = npos_m; dy = dx + npos; }
    else npos = npos_m * npos_n;
    const long long nfnpos = (long long)nf * npos;
    const long long ntout = (mode & 2) ? ntime : 1;
    // ---- output shape (blochC.c:880-904) ----
    if (ntout > 1 && nf > 1 && npos > 1) { out_dims[0] = (int)ntout; out_dims[1] = npos; out_dims[2] = nf; out_dims[3] = 3; }
    else if (ntout > 1) { out_dims[0] = (int)ntout; out_dims[1] = npos * nf; out_dims[2] = 1; out_dims[3] = 2; }
    else { out_dims[0] = npos; out_dims[1] = nf; out_dims[2] = 1; out_dims[3] = 2; }
    // ---- initial magnetisation (blochC.c:820-865): used only when all three have npos*nf elements ----
    const bool have_m0 = mx0 && my0 && mz0 && (long long)n_m0 == nfnpos;
    if (nfnpos == 0) return MBRF_OK;
    if (!mx || !my || !mz) { set_error("bloch: NULL output"); return MBRF_EINVAL; }
    return run_host(b1r, b1i, gx, gy, gz, dt.data(), ntime, t1, t2, df, nf, dx, dy, dz, npos,
                    have_m0 ? mx0 : nullptr, have_m0 ? my0 : nullptr, have_m0 ? mz0 : nullptr, 1, mx, my, mz,
                    mode, gamma);
}

}  // extern "C"

/*
 * Host-pointer form of the sweep extension: one call simulates nfreq off-resonances x nscale B1 scalings of ONE pulse
 * (sim_rf_scale.m:82-89 runs one blochC / blochH call per scale).  b1imag may be NULL; tp is the constant time step (s).
 * mx, my, mz: nfreq*nscale doubles each, element [i_f + nfreq*i_s].  Magnetisation starts at (0, 0, 1), mode 0 (end point).
 */
extern "C" int mbrf_bloch_scale_sweep(const double *b1real, const double *b1imag, int ntime, double tp, double t1, double t2,
                                      const double *dfreq, int nfreq, const double *b1scale, int nscale, double *mx, double *my,
                                      double *mz, double gamma)
{
    using namespace mbrf;
    if (int rc = require_device()) return rc;
    if (ntime <= 0 || nfreq <= 0 || nscale <= 0 || !b1real || !dfreq || !b1scale || !mx || !my || !mz || !(tp > 0.0)) {
        set_error("bloch sweep: bad arguments (ntime=%d nfreq=%d nscale=%d)", ntime, nfreq, nscale);
        return MBRF_EINVAL;
    }
    static thread_local DeviceScratch scratch;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t nt = (size_t)ntime, ns = (size_t)nfreq * nscale;
    const size_t wsb = al(mbrf_bloch_workspace_bytes(ntime));
    if (int rc = scratch.reserve(wsb + 3 * al(nt * 8) + al((size_t)nfreq * 8) + al((size_t)nscale * 8) + 3 * al(ns * 8))) return rc;
    char *d = (char *)scratch.ptr;
    void *ws = d; d += wsb;
    double *dbr = (double *)d; d += al(nt * 8);
    double *dbi = (double *)d; d += al(nt * 8);
    double *ddt = (double *)d; d += al(nt * 8);
    double *ddf = (double *)d; d += al((size_t)nfreq * 8);
    double *dsc = (double *)d; d += al((size_t)nscale * 8);
    double *dmx = (double *)d; d += al(ns * 8);
    double *dmy = (double *)d; d += al(ns * 8);
    double *dmz = (double *)d;
    std::vector<double> steps(nt, tp);
    MBRF_CUDA(cudaMemcpyAsync(dbr, b1real, nt * 8, cudaMemcpyHostToDevice, 0));
    if (b1imag) MBRF_CUDA(cudaMemcpyAsync(dbi, b1imag, nt * 8, cudaMemcpyHostToDevice, 0));
    MBRF_CUDA(cudaMemcpyAsync(ddt, steps.data(), nt * 8, cudaMemcpyHostToDevice, 0));
    MBRF_CUDA(cudaMemcpyAsync(ddf, dfreq, (size_t)nfreq * 8, cudaMemcpyHostToDevice, 0));
    MBRF_CUDA(cudaMemcpyAsync(dsc, b1scale, (size_t)nscale * 8, cudaMemcpyHostToDevice, 0));
    MBRF_CUDA(cudaStreamSynchronize(0));                        // `steps` is pageable and goes out of scope with this call
    if (int rc = mbrf_bloch_scale_sweep_device(dbr, b1imag ? dbi : nullptr, ddt, ntime, t1, t2, ddf, nfreq, dsc, nscale, 0, (long long)ns,
                                               dmx, dmy, dmz, gamma, ws, nullptr)) return rc;
    MBRF_CUDA(cudaMemcpyAsync(mx, dmx, ns * 8, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaMemcpyAsync(my, dmy, ns * 8, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaMemcpyAsync(mz, dmz, ns * 8, cudaMemcpyDeviceToHost, 0));
    MBRF_CUDA(cudaStreamSynchronize(0));
    return MBRF_OK;
}
