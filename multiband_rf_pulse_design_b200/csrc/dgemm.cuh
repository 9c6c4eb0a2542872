// fp64 product kernels shared by the first-order solver (pdhg.cu) and the interior-point solver (ipm.cu):
// SIMT DFMA tiles, FP64 tensor tiles (mma.sync.m8n8k4.f64, DMMA) and the thin-batch matrix-vector kernels.
#pragma once
#include "common.h"

#include <cstdlib>
#include <cstring>

namespace mbrf {
namespace pdhg {

// ---------------------------------------------------------------------------
// fp64 SIMT GEMM:  C[R x Bp] = AT^T * X  with AT stored [kdim x R] row-major (ld = ldat), X [kdim x Bp].
//   K * Zbar  : AT = K^T (kept as a second copy, [Np x Mp]),  kdim = Np, R = Mp
//   K^T * Y   : AT = K   ([Mp x Np]),                          kdim = Mp, R = Np, split over kdim (blockIdx.z)
// Tile 64 x 64 per CTA of 64 threads (2 warps); each thread owns an 8 x 8 micro-tile: per k it issues
// 8 LDS.128 for 64 DFMA, so the FP64 pipe (2 issue cycles per DFMA) is the only busy unit.  Operand tiles
// arrive with 16-byte cp.async (LDGSTS) into a 2-stage shared-memory ring: no staging registers.
// Many small CTAs (976 for a 7872 x 512 output) keep the tail of the last wave short on 148 SMs.
// ---------------------------------------------------------------------------
constexpr int BM = 64, BN = 64, BK = 16, GEMM_THREADS = 128;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 128 threads, each an 8 (rows) x 4 (columns) micro-tile: rows q*16 + ty*2 + {0,1} (q = 0..3, ty = 0..7),
// columns q*32 + tx*2 + {0,1} (q = 0..1, tx = 0..15).  The 8 threads of an LDS.128 phase read 128 contiguous
// bytes: no bank conflicts.  ~110 registers -> 4 CTAs = 16 warps per SM, which the FP64 pipe needs: one warp
// alone cannot issue a DFMA every 2 cycles (measured: 71 % pipe-busy at 2 warps per scheduler, ncu).
static __global__ void __launch_bounds__(GEMM_THREADS, 4)
dgemm_kernel(const double *__restrict__ AT, int ldat,         // [kdim x R] row-major
             const double *__restrict__ X, int Bp,            // [kdim x Bp]
             double *__restrict__ C,                          // [R x Bp], one slab per blockIdx.z
             int kdim_total, int kchunk, long long slab)
{
    __shared__ __align__(16) double As[2][BK][BM];
    __shared__ __align__(16) double Bs[2][BK][BN];
    const int tid = threadIdx.x;
    const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * kchunk;
    const int k_end = min(kdim_total, k_begin + kchunk);
    const int ty = tid / 16, tx = tid % 16;

    // each cp.async moves 2 doubles; a 16 x 64 tile is 512 such pieces = 4 per thread, for A and for B
    auto load_slice = [&](int buf, int k0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int piece = e * GEMM_THREADS + tid;        // kk = piece / 32, c = (piece % 32) * 2
            const int kk = piece >> 5, c = (piece & 31) * 2;
            cp_async16(&As[buf][kk][c], AT + (size_t)(k0 + kk) * ldat + row0 + c);
            cp_async16(&Bs[buf][kk][c], X + (size_t)(k0 + kk) * Bp + col0 + c);
        }
        cp_async_commit();
    };

    double acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    int buf = 0;
    if (k_begin < k_end) load_slice(0, k_begin);
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        const bool more = k0 + BK < k_end;
        if (more) {
            load_slice(buf ^ 1, k0 + BK);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            double a[8], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double2 av = *reinterpret_cast<const double2 *>(&As[buf][kk][q * 16 + ty * 2]);
                a[2 * q] = av.x; a[2 * q + 1] = av.y;
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const double2 bv = *reinterpret_cast<const double2 *>(&Bs[buf][kk][q * 32 + tx * 2]);
                b[2 * q] = bv.x; b[2 * q + 1] = bv.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();   // everyone is done with `buf` before the next iteration's cp.async overwrites it
        buf ^= 1;
    }
    double *Cz = C + (size_t)blockIdx.z * slab;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = row0 + (i >> 1) * 16 + ty * 2 + (i & 1);
#pragma unroll
        for (int q = 0; q < 2; ++q)
            *reinterpret_cast<double2 *>(Cz + (size_t)r * Bp + col0 + q * 32 + tx * 2) =
                make_double2(acc[i][2 * q], acc[i][2 * q + 1]);
    }
}

// ---------------------------------------------------------------------------
// Same product on the FP64 tensor path: mma.sync.m8n8k4.f64 (DMMA).  tcgen05 has no fp64 kind, so this is
// the only tensor-core route that keeps the solver's fp64 tolerance.  One warp instruction performs
// 8x8x4 = 256 FMAs, i.e. 8 per thread for 2 operand doubles, which takes the issue-slot pressure of the
// SIMT kernel away (there: 32 DFMA + 6 LDS per thread and k).
// CTA 64 x 64, 4 warps in a 2 x 2 arrangement, warp tile 32 x 32 = 4 x 4 mma tiles, BK = 16.
// Fragment layout (PTX ISA, m8n8k4 .f64): A(row = lane/4, k = lane%4), B(k = lane%4, col = lane/4),
// C(row = lane/4, cols 2*(lane%4) + {0,1}).  Shared rows are padded to 72 doubles so that the four k-rows a
// fragment load touches fall into the two halves of the banks: 2 wavefronts per LDS.64, the minimum for
// 256 bytes.
// ---------------------------------------------------------------------------
constexpr int MMA_LD = 72;

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

static __global__ void __launch_bounds__(128)
dgemm_mma_kernel(const double *__restrict__ AT, int ldat, const double *__restrict__ X, int Bp,
                 double *__restrict__ C, int kdim_total, int kchunk, long long slab)
{
    __shared__ __align__(16) double As[2][BK][MMA_LD];
    __shared__ __align__(16) double Bs[2][BK][MMA_LD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * kchunk;
    const int k_end = min(kdim_total, k_begin + kchunk);
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;   // warp tile origin inside the CTA tile
    const int lr = lane >> 2, lk = lane & 3;

    auto load_slice = [&](int buf, int k0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int piece = e * 128 + tid;
            const int kk = piece >> 5, c = (piece & 31) * 2;
            cp_async16(&As[buf][kk][c], AT + (size_t)(k0 + kk) * ldat + row0 + c);
            cp_async16(&Bs[buf][kk][c], X + (size_t)(k0 + kk) * Bp + col0 + c);
        }
        cp_async_commit();
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    int buf = 0;
    if (k_begin < k_end) load_slice(0, k_begin);
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        const bool more = k0 + BK < k_end;
        if (more) {
            load_slice(buf ^ 1, k0 + BK);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[buf][k4 + lk][wm + i * 8 + lr];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[buf][k4 + lk][wn + j * 8 + lr];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
        buf ^= 1;
    }
    double *Cz = C + (size_t)blockIdx.z * slab;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<double2 *>(Cz + (size_t)(row0 + wm + i * 8 + lr) * Bp + col0 + wn + j * 8 + 2 * lk) =
                make_double2(acc[i][j][0], acc[i][j][1]);
}

// ---------------------------------------------------------------------------
// Single design (Bp == 1): the two products are matrix-vector products, bound by streaming the 32 MB matrix
// from L2/HBM.  One warp per output element's row of the k-major operand (K for K z, K^T for K^T y), 16-byte
// coalesced loads, the vector kept in shared memory, warp-shuffle reduction.
//   out[r] = sum_k A[r][k] * x[k]      A row-major [R x ld], kdim multiple of 64
// ---------------------------------------------------------------------------
// out[r][b] = sum_k A[r][k] * x[k][b]  for NB = 1, 2, 4 or 8 designs; A row-major [R x ld] (k contiguous: K for K z, K^T for
// K^T y).  One warp per output row: 16-byte coalesced loads (512 contiguous bytes per warp instruction), NB accumulators
// per lane, warp-shuffle reduction.  blockIdx.y splits a long reduction into slabs like the GEMM's split-K.  The matrix
// (tens of MB) stays in the 126 MB L2, so the pass is L2-bandwidth-bound; the column-major variant this replaces had 89
// CTAs for K z and ran at 0.4 TB/s.
template <int NB>
__global__ void __launch_bounds__(256)
gemv_rows_kernel(const double *__restrict__ A, int ld, const double *__restrict__ X, double *__restrict__ C, int R, int kdim,
                 int kchunk, long long slab)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + warp;
    if (r >= R) return;
    const int k_begin = blockIdx.y * kchunk, k_end = min(kdim, k_begin + kchunk);
    const double *a = A + (size_t)r * ld;
    double acc[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = 0.0;
#pragma unroll 4
    for (int k = k_begin + 2 * lane; k < k_end; k += 64) {          // k ranges are multiples of 64
        const double2 av = *reinterpret_cast<const double2 *>(a + k);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            acc[b] = fma(av.x, __ldg(X + (size_t)k * NB + b), acc[b]);
            acc[b] = fma(av.y, __ldg(X + (size_t)(k + 1) * NB + b), acc[b]);
        }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
    if (lane == 0) {
        double *out = C + (size_t)blockIdx.y * slab + (size_t)r * NB;
#pragma unroll
        for (int b = 0; b < NB; ++b) out[b] = acc[b];
    }
}

// The same product for NB = 2, 4, 8 designs with the x tile staged in shared memory, design-major ([b][k]): a CTA owns 32 rows
// (8 warps x 4 rows), per 64-k block every lane loads its k-pair of each of its 4 rows (LDG.128) and, per design, the matching
// x pair (conflict-free LDS.128), i.e. 8 NB FMAs for 4 + NB loads.  One pass over the L2-resident matrix for the whole batch.
template <int NB>
__global__ void __launch_bounds__(256)
gemv_rows4_kernel(const double *__restrict__ A, int ld, const double *__restrict__ X, double *__restrict__ C, int R, int kdim,
                  int kchunk, long long slab)
{
    __shared__ __align__(16) double xs[NB][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * 32 + warp * 4;
    const int k_begin = blockIdx.y * kchunk, k_end = min(kdim, k_begin + kchunk);
    double acc[4][NB];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[i][b] = 0.0;
    for (int k0 = k_begin; k0 < k_end; k0 += 64) {
        __syncthreads();
        for (int e = threadIdx.x; e < 64 * NB; e += 256) xs[e % NB][e / NB] = X[(size_t)k0 * NB + e];
        __syncthreads();
        double2 av[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            av[i] = r0 + i < R ? *reinterpret_cast<const double2 *>(A + (size_t)(r0 + i) * ld + k0 + 2 * lane) : make_double2(0.0, 0.0);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const double2 xv = *reinterpret_cast<const double2 *>(&xs[b][2 * lane]);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i][b] = fma(av[i].y, xv.y, fma(av[i].x, xv.x, acc[i][b]));
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[i][b] += __shfl_xor_sync(0xffffffffu, acc[i][b], o);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (r0 + i < R) {
                double *out = C + (size_t)blockIdx.y * slab + (size_t)(r0 + i) * NB;
#pragma unroll
                for (int b = 0; b < NB; ++b) out[b] = acc[i][b];
            }
    }
}

// out[o][b] = sum_k AT[k][o] * x[k][b] with AT k-major ([kdim x ldat], the output index contiguous) for NB = 2, 4, 8 designs:
// a CTA owns 64 outputs, its four quarters take every 4th k of a 64-row tile, the x tile is broadcast from shared memory
// (NB FMAs per loaded matrix element), the quarters are summed through shared memory.  Measured on the cfg4 LP
// (7808 x 512): faster than the warp-per-row kernel above for NB >= 2 (there every lane fetches 2 NB values of x per
// matrix pair: 279 us per iteration at NB = 8), slower for NB <= 2.
template <int NB>
__global__ void __launch_bounds__(256)
thin_kernel(const double *__restrict__ AT, int ldat, const double *__restrict__ X, double *__restrict__ C, int kdim,
            int kchunk, long long slab)
{
    __shared__ double xs[64 * NB];
    __shared__ double red[4][64][NB];
    const int tid = threadIdx.x, ol = tid & 63, q = tid >> 6;
    const int o = blockIdx.x * 64 + ol;
    const int k_begin = blockIdx.y * kchunk, k_end = min(kdim, k_begin + kchunk);
    double acc[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = 0.0;
    for (int k0 = k_begin; k0 < k_end; k0 += 64) {
        const int nk = min(64, k_end - k0);
        for (int e = tid; e < nk * NB; e += 256) xs[e] = X[(size_t)k0 * NB + e];
        __syncthreads();
#pragma unroll 4
        for (int kk = q; kk < nk; kk += 4) {
            const double a = AT[(size_t)(k0 + kk) * ldat + o];
#pragma unroll
            for (int b = 0; b < NB; ++b) acc[b] = fma(a, xs[kk * NB + b], acc[b]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) red[q][ol][b] = acc[b];
    __syncthreads();
    if (q == 0) {
        double *out = C + (size_t)blockIdx.y * slab + (size_t)o * NB;
#pragma unroll
        for (int b = 0; b < NB; ++b) out[b] = red[0][ol][b] + red[1][ol][b] + red[2][ol][b] + red[3][ol][b];
    }
}

// which kernel serves a thin batch (measured on the cfg4 LP, 7808 x 512, us per PDHG iteration, rows / staged rows / columns):
//   NB = 1: 34 / -- / 65     NB = 2: 46 / 35 / 64     NB = 4: 93 / 38 / 43     NB = 8: 279 / 58 / 53
// MBRF_THIN_KERNEL = rows | staged | columns forces one (developer switch)
enum { THIN_ROWS = 0, THIN_STAGED = 1, THIN_COLUMNS = 2 };
static int thin_choice(int nb)
{
    static int forced = -2;
    if (forced == -2) {
        const char *e = getenv("MBRF_THIN_KERNEL");
        forced = !e ? -1 : !strcmp(e, "rows") ? THIN_ROWS : !strcmp(e, "staged") ? THIN_STAGED : !strcmp(e, "columns") ? THIN_COLUMNS : -1;
    }
    if (forced >= 0) return (forced == THIN_STAGED && nb == 1) ? THIN_ROWS : forced;
    return nb == 1 ? THIN_ROWS : nb <= 4 ? THIN_STAGED : THIN_COLUMNS;
}
// A: row-major [R x ld] for the rows kernels; AT: the same matrix k-major [kdim x ldat] for the column kernel
static void launch_thin(int nb, cudaStream_t st, const double *A, int ld, const double *AT, int ldat, const double *X, double *C,
                        int R, int kdim, int kchunk, int nslab, long long slab)
{
    const int which = thin_choice(nb);
    if (which == THIN_STAGED) {
        const dim3 grid((R + 31) / 32, nslab);
        switch (nb) {
        case 2: gemv_rows4_kernel<2><<<grid, 256, 0, st>>>(A, ld, X, C, R, kdim, kchunk, slab); break;
        case 4: gemv_rows4_kernel<4><<<grid, 256, 0, st>>>(A, ld, X, C, R, kdim, kchunk, slab); break;
        default: gemv_rows4_kernel<8><<<grid, 256, 0, st>>>(A, ld, X, C, R, kdim, kchunk, slab); break;
        }
    } else if (which == THIN_ROWS) {
        const dim3 grid(R / 8, nslab);
        switch (nb) {
        case 1: gemv_rows_kernel<1><<<grid, 256, 0, st>>>(A, ld, X, C, R, kdim, kchunk, slab); break;
        case 2: gemv_rows_kernel<2><<<grid, 256, 0, st>>>(A, ld, X, C, R, kdim, kchunk, slab); break;
        case 4: gemv_rows_kernel<4><<<grid, 256, 0, st>>>(A, ld, X, C, R, kdim, kchunk, slab); break;
        default: gemv_rows_kernel<8><<<grid, 256, 0, st>>>(A, ld, X, C, R, kdim, kchunk, slab); break;
        }
    } else {
        const dim3 grid(R / 64, nslab);
        switch (nb) {
        case 1: thin_kernel<1><<<grid, 256, 0, st>>>(AT, ldat, X, C, kdim, kchunk, slab); break;
        case 2: thin_kernel<2><<<grid, 256, 0, st>>>(AT, ldat, X, C, kdim, kchunk, slab); break;
        case 4: thin_kernel<4><<<grid, 256, 0, st>>>(AT, ldat, X, C, kdim, kchunk, slab); break;
        default: thin_kernel<8><<<grid, 256, 0, st>>>(AT, ldat, X, C, kdim, kchunk, slab); break;
        }
    }
}

}  // namespace pdhg
}  // namespace mbrf
