// tcgen05 (5th-generation tensor core) product for the batched PDHG iterations: split-integer ("Ozaki") GEMM.
//
//   C[R x Bp] (fp64) = A * X,    A: [R x kdim] fp64 (K or K^T, fixed per solve),  X: [kdim x Bp] fp64 (the iterate)
//
// tcgen05.mma has no fp64 kind, and an fp32-accumulate product is at the level of the residuals PDHG has to drive down
// (DESIGN.md section 6).  The int8 kind accumulates in int32 EXACTLY, so the product is rebuilt from integer digits:
//   every row of A and every column of X is scaled by a power of two to |.| <= 1 and written as signed base-256 digits
//       a / alpha_r = (1/64) * sum_s A_s * 256^-s ,   A_s in [-128, 127]                     (x likewise, beta_b, X_t)
//   C_rb = alpha_r beta_b / 4096 * sum_d 256^-d * sum_{s+t=d} (A_s X_t)_rb ,   d < ND
// The level sums  sum_{s+t=d} A_s X_t  are exact int32 GEMMs (|.| <= (d+1) * 2^14 * kdim < 2^31 for kdim < 2^17 / ND)
// that share one TMEM accumulator per level; dropping the levels d >= ND leaves an error of about 2^(-8 ND + 3) of
// alpha_r * beta_b per term (ND = 5: measured 1e-11, ND = 6: 1e-13) — set by ND, not by the tensor core.
//
// One tile is 128 designs (the MMA's M side: TMEM lanes, so that a warp stores 256 contiguous bytes of a C row) x 64
// matrix rows: ND accumulators of 64 int32 columns in TMEM; digit tiles of KB = 64 k-values (all ND digit planes of X:
// 128 designs, and of A: 64 rows) arrive with two 3-D TMA copies per stage (64-byte swizzle)
// in a 3-stage shared-memory ring; one elected thread issues ND(ND+1)/2 tcgen05.mma.kind::i8 (M=128, N=64, K=32, both
// operands K-major) per 32 k-values; tcgen05.commit frees ring slots and finally signals the sixteen epilogue warps, which
// read the level accumulators with tcgen05.ld, combine them (Horner in 1/256 in fp64: exact) and
// store scaled doubles.  The kernel is persistent (one CTA per SM walks the tiles).
// Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM allocator, 2..17 = epilogue (TMEM lane quarter = warp % 4).
// Descriptor encodings were brought up with tools/microbench/tc_dbg.cu (MN-major tf32 operands return zeros on this
// part, K-major operands are exact: hence the transposed digit planes).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mbrf {
namespace tc {

constexpr int TM = 128, TN = 64, KB = 64, STAGES = 3;
constexpr int EPI_WARPS = 16, EPI_COLS = TN / (EPI_WARPS / 4);   // epilogue: 4 warps per TMEM lane quarter, 16 columns each
constexpr int THREADS = 64 + 32 * EPI_WARPS;                       // warps: TMA, MMA, epilogue
constexpr int MAX_ND = 6;
constexpr int stage_bytes(int nd) { return nd * (TM + TN) * KB; }
constexpr int smem_bytes(int nd) { return STAGES * stage_bytes(nd) + 1024 /*align*/ + 256 /*barriers*/; }

struct Params {
    double *C;                // [R x ldc] fp64, slab z at C + z*slab
    long long slab;
    int ldc;
    int R;                    // output rows (multiple of TN); designs >= ldc of a tile are not stored
    int kdim_total, kchunk;   // reduction range of slab z: [z*kchunk, min(kdim_total, (z+1)*kchunk)), multiples of KB
    int nslab;                // split-K slabs
    const double *sa;         // [R]   alpha_r / 64
    const double *sx;         // [ldc] beta_b / 64
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of a converged warp; the compiler then knows the region is single-lane and feeds the uniform-register operands
// of UTMALDG / UTCIMMA without a per-instruction lane loop (with `lane == 0` every MMA sat in an ELECT/R2UR/BRA loop)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void bar_init(uint64_t *b, unsigned n)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void bar_expect(uint64_t *b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t *b, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tTCW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra TCD;\n\tbra TCW;\n\tTCD:\n\t}"
        ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_3d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(s32(dst)), "l"((uint64_t)map), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1), K-major, 64-byte swizzle: rows of 64 bytes,
// groups of 8 rows are 512 bytes apart (stride byte offset); the leading byte offset is unused for swizzled K-major.
__device__ __forceinline__ uint64_t smem_desc_k64(uint32_t addr)
{
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(16 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) |
           (4ull << 61);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}

// A_TMEM: every A digit plane of a k-block is copied once into TMEM (tcgen05.cp) and the ND - s products that use plane s
// read it from there: shared-memory operand traffic per 32 k-values drops from ND(ND+1)/2 * 6 KB to ND * 4 KB +
// ND(ND+1)/2 * 2 KB, which moves the tile from shared-memory-bound to tensor-pipe-bound (measured).
__device__ __forceinline__ void bar_arrive(uint64_t *b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// exact int32 -> double without the (slow) I2F.F64: flip the sign bit into the low mantissa word of 2^52, subtract 2^52 + 2^31
__device__ __forceinline__ double i2d(int a)
{
    return __hiloint2double(0x43300000, a ^ (int)0x80000000) - 4503601774854144.0;
}

#ifdef TC_TIMING   // developer instrumentation (tools/microbench/tc_gemm_test.cu): clock stamps of every CTA's first tile
__device__ long long *tc_timing;
#define TC_STAMP(i) do { if (tc_timing && tile == (int)blockIdx.x) tc_timing[(size_t)blockIdx.x * 8 + (i)] = clock64(); } while (0)
#else
#define TC_STAMP(i) do { } while (0)
#endif

// Persistent: gridDim.x CTAs (one per SM) walk the tiles  t = blockIdx.x, blockIdx.x + gridDim.x, ...;  tile t is
// (matrix-row tile t % tiles_x, design tile (t / tiles_x) % tiles_y, split-K slab t / (tiles_x * tiles_y)).  The shared-memory
// ring runs on across tiles (the producer prefetches the next tile during the epilogue); the epilogue warps pull the whole
// accumulator into registers (ND x 16 ints per thread), hand TMEM back to the MMA warp (tmem_empty) and only then
// convert and store, so conversion and stores of tile t overlap the MMAs of tile t + 1.
template <int ND>
__global__ void __launch_bounds__(THREADS, 1)
tc_i8_gemm_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mX, const Params p)
{
    constexpr int STAGE = ND * (TM + TN) * KB;
    constexpr int X_PLANE = TM * KB, A_PLANE = TN * KB;           // bytes of one digit plane in a stage (designs are the M side)
    constexpr uint32_t TMEM_COLS = ND * TN <= 256 ? 256u : 512u;  // power of two >= ND accumulators of TN columns
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // swizzle atoms need aligned planes
    uint64_t *full = (uint64_t *)(smem + STAGES * STAGE);
    uint64_t *empty = full + STAGES;
    uint64_t *tmem_full = empty + STAGES;
    uint64_t *tmem_empty = tmem_full + 1;
    uint32_t *tmem_slot = (uint32_t *)(tmem_empty + 1);

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int tiles_x = p.R / TN, tiles_y = (p.ldc + TM - 1) / TM;   // matrix-row tiles (fastest), design tiles
    const int ntiles = tiles_x * tiles_y * p.nslab;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { bar_init(&full[s], 1); bar_init(&empty[s], 1); }
        bar_init(tmem_full, 1);
        bar_init(tmem_empty, EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0 && elect_one()) {
        // ---------------- TMA producer: all digit planes of a k-block with one copy per operand ----------------
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int r0 = (tile % tiles_x) * TN, b0 = ((tile / tiles_x) % tiles_y) * TM, z = tile / (tiles_x * tiles_y);
            const int k_begin = z * p.kchunk, k_end = min(p.kdim_total, k_begin + p.kchunk);
            for (int k0 = k_begin; k0 < k_end; k0 += KB, ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                bar_wait(&empty[s], ph ^ 1);
                uint8_t *st = smem + s * STAGE;
                bar_expect(&full[s], STAGE);
                tma_3d(&mX, &full[s], st, k0, b0, 0);                    // 128 designs x KB x ND planes  (MMA operand A)
                tma_3d(&mA, &full[s], st + ND * X_PLANE, k0, r0, 0);     // 64 matrix rows x KB x ND planes (MMA operand B)
            }
        }
    } else if (warp == 1 && elect_one()) {
        // ---------------- MMA issuer ----------------
        // instruction descriptor (cute::UMMA::InstrDescriptor): D = S32, A = B = signed int8, both K-major, N = 64, M = 128
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
        int it = 0, tl = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
            const int z = tile / (tiles_x * tiles_y);
            const int k_begin = z * p.kchunk, k_end = min(p.kdim_total, k_begin + p.kchunk);
            const int nk = k_end > k_begin ? (k_end - k_begin) / KB : 0;
            bar_wait(tmem_empty, (tl & 1) ^ 1);          // the epilogue has pulled the previous tile out of TMEM
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            TC_STAMP(1);
            for (int kb = 0; kb < nk; ++kb, ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                bar_wait(&full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (kb == 0) TC_STAMP(2);
                const uint32_t base = s32(smem + s * STAGE);
#pragma unroll
                for (int kk = 0; kk < KB / 32; ++kk) {   // one MMA covers K = 32 int8 (32 bytes): advance inside the swizzle row
#pragma unroll
                    for (int sx = 0; sx < ND; ++sx) {
                        const uint64_t dX = smem_desc_k64(base + sx * X_PLANE + kk * 32);
#pragma unroll
                        for (int sa = 0; sa < ND - sx; ++sa) {
                            const uint64_t dA = smem_desc_k64(base + ND * X_PLANE + sa * A_PLANE + kk * 32);
                            // level sa+sx; its first product of the tile (kb = 0, kk = 0, sx = 0) overwrites the accumulator
                            mma_i8(tmem + (uint32_t)((sa + sx) * TN), dX, dA, idesc, (kb > 0 || kk > 0 || sx > 0) ? 1u : 0u);
                        }
                    }
                }
                mma_commit(&empty[s]);          // slot free once these MMAs have read it
            }
            mma_commit(tmem_full);              // accumulators complete (arrives immediately for an empty slab)
            TC_STAMP(3);
        }
    } else if (warp >= 2) {
        // ---------------- epilogue: TMEM -> registers (release TMEM) -> fp64 global ----------------
        // four warps per TMEM lane quarter (warp % 4), each takes EPI_COLS = 16 of the tile's columns
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int part = (warp - 2) >> 2;             // columns [part*16, part*16 + 16)
        int tl = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
            const int r0 = (tile % tiles_x) * TN + part * EPI_COLS, b0 = ((tile / tiles_x) % tiles_y) * TM, z = tile / (tiles_x * tiles_y);
            const int k_begin = z * p.kchunk;
            const bool any = k_begin < p.kdim_total;
            const int b = b0 + q * 32 + lane;              // TMEM lane = design: a warp stores 256 contiguous bytes per matrix row
            const bool live = b < p.ldc;
            const double xs = live ? p.sx[b] : 0.0;
            const double my_sa = p.sa[r0 + (lane & (EPI_COLS - 1))];
            double *out = p.C + (size_t)z * p.slab + (size_t)r0 * p.ldc + b;
            bar_wait(tmem_full, tl & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tid == 64) TC_STAMP(4);
            int acc[ND][EPI_COLS];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * EPI_COLS);
#pragma unroll
            for (int d = 0; d < ND; ++d) tmem_ld16(taddr + (uint32_t)(d * TN), acc[d]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) bar_arrive(tmem_empty);
            if (tid == 64) TC_STAMP(5);
            // sum_d acc_d 256^-d by Horner in fp64 (every step exact: < 2^53 significant bits)
            // branch-free and level-major, so that the 16 independent Horner chains of a thread interleave (the FP64 pipe has
            // a long latency: one chain after the other cost 4200 cycles per tile, measured)
            double t[EPI_COLS];
#pragma unroll
            for (int j = 0; j < EPI_COLS; ++j) t[j] = i2d(acc[ND - 1][j]);
#pragma unroll
            for (int d = ND - 2; d >= 0; --d)
#pragma unroll
                for (int j = 0; j < EPI_COLS; ++j) t[j] = fma(t[j], 0.00390625, i2d(acc[d][j]));
            const double xs_any = any ? xs : 0.0;      // an empty split-K slab stores zeros
#pragma unroll
            for (int j = 0; j < EPI_COLS; ++j) {
                const double v = t[j] * (xs_any * __shfl_sync(0xffffffffu, my_sa, j));
                if (live) out[(size_t)j * p.ldc] = v;
            }
            if (tid == 64) TC_STAMP(6);
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------
// digit planes
// ---------------------------------------------------------------------------
// signed base-256 digits of the fixed-point number W = rint(x / scale * 2^(8 ND - 2)), |x| <= scale: taken from the least
// significant end, d = ((W + 128) mod 256) - 128 in [-128, 127], W <- (W - d) / 256 (exact); the top digit is then within
// [-65, 65].  x = scale / 64 * sum_s d_s * 256^-s up to the rounding of W (2^-(8 ND - 1) of scale).
template <int ND>
__device__ __forceinline__ void digits(double xf, int8_t (&d)[ND])   // xf = x * 2^(8 ND - 2) / scale
{
    long long W = __double2ll_rn(xf);
#pragma unroll
    for (int s = ND - 1; s > 0; --s) {
        const int dd = (int)((W + 128) & 255) - 128;
        d[s] = (int8_t)dd;
        W = (W - dd) >> 8;
    }
    d[0] = (int8_t)W;
}
// power of two >= max|x| (1 for a zero vector): returns the factor 2^(8 ND - 2) / scale applied before the digits
template <int ND>
__device__ __forceinline__ double pow2_scale(double mx, double *scale_over_64)
{
    int e = 0;
    if (mx > 0.0 && mx < 1e300) frexp(mx, &e);      // mx = f * 2^e, f in [0.5, 1)
    *scale_over_64 = ldexp(1.0, e - 6);
    return ldexp(1.0, 8 * ND - 2 - e);
}

// Rows of a fixed fp64 matrix A [R x ld] (kdim live columns) -> planes [ND][R][kdim] int8 + sa[r].  One warp per row.
template <int ND>
__global__ void slice_rows_kernel(const double *__restrict__ A, int ld, int R, int kdim, int8_t *__restrict__ out,
                                  double *__restrict__ sa)
{
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= R) return;
    const double *a = A + (size_t)r * ld;
    double mx = 0.0;
    for (int k = lane; k < kdim; k += 32) mx = fmax(mx, fabs(a[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    double so64;
    const double f = pow2_scale<ND>(mx, &so64);
    if (lane == 0) sa[r] = so64;
    const size_t plane = (size_t)R * kdim;
    for (int k = lane * 4; k < kdim; k += 128) {     // kdim is a multiple of 64: char4 stores
        int8_t d[4][ND];
#pragma unroll
        for (int j = 0; j < 4; ++j) digits<ND>(a[k + j] * f, d[j]);
#pragma unroll
        for (int s = 0; s < ND; ++s)
            *reinterpret_cast<char4 *>(out + s * plane + (size_t)r * kdim + k) = make_char4(d[0][s], d[1][s], d[2][s], d[3][s]);
    }
}

// per-design max |x| of an iterate X [kdim x Bp] (design index fastest); mx must be zeroed before
__global__ void col_absmax_kernel(const double *__restrict__ X, int kdim, int Bp, double *__restrict__ mx)
{
    __shared__ double sh[4][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int b = blockIdx.x * 64 + tx;
    double m = 0.0;
    for (int k = blockIdx.y * 4 + ty; k < kdim; k += 4 * gridDim.y) m = fmax(m, fabs(X[(size_t)k * Bp + b]));
    sh[ty][tx] = m;
    __syncthreads();
    if (ty == 0) {
        m = fmax(fmax(sh[0][tx], sh[1][tx]), fmax(sh[2][tx], sh[3][tx]));
        if (m > 0.0) atomicMax(reinterpret_cast<unsigned long long *>(mx + b), (unsigned long long)__double_as_longlong(m));
    }
}

// Iterate X [kdim x Bp] -> planes [ND][Bp][kdim] int8 (transposed: k fastest) + sx[b].  A CTA of 256 threads
// transposes a 64 (k) x 64 (designs) tile through shared memory: coalesced 512-byte reads; each thread packs the
// digits of 4 consecutive k into one 32-bit word per plane (row pitch 17 words: conflict-free), 64-byte row segments out.
// zero_other (optional, [Bp]): the max-accumulator of the OTHER iterate is cleared here for its next producer.
template <int ND>
__global__ void __launch_bounds__(256) slice_cols_kernel(const double *__restrict__ X, int kdim, int Bp,
                                                         const double *__restrict__ mx, int8_t *__restrict__ out,
                                                         double *__restrict__ sx, double *__restrict__ zero_other)
{
    __shared__ uint32_t tile[ND][64][17];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int b0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
    double so64;
    const double f = pow2_scale<ND>(mx[b0 + tx], &so64);
    if (blockIdx.y == 0 && ty == 0) {
        sx[b0 + tx] = so64;
        if (zero_other) zero_other[b0 + tx] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int q = ty + 4 * j;                         // quad of k: k0 + 4q .. k0 + 4q + 3
        const double *src = X + (size_t)(k0 + 4 * q) * Bp + b0 + tx;
        double v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = src[(size_t)i * Bp] * f;
        int8_t d[4][ND];
#pragma unroll
        for (int i = 0; i < 4; ++i) digits<ND>(v[i], d[i]);
#pragma unroll
        for (int s = 0; s < ND; ++s)
            tile[s][tx][q] = (uint32_t)(uint8_t)d[0][s] | ((uint32_t)(uint8_t)d[1][s] << 8) | ((uint32_t)(uint8_t)d[2][s] << 16) |
                             ((uint32_t)(uint8_t)d[3][s] << 24);
    }
    __syncthreads();
    const size_t plane = (size_t)Bp * kdim;
    for (int c = threadIdx.x; c < ND * 64 * 4; c += 256) {     // 16-byte chunks: (plane, design, quarter of the 64 k)
        const int s = c / 256, bb = (c >> 2) & 63, qd = c & 3;
        const uint32_t *t = &tile[s][bb][qd * 4];
        *reinterpret_cast<uint4 *>(out + s * plane + (size_t)(b0 + bb) * kdim + k0 + qd * 16) = make_uint4(t[0], t[1], t[2], t[3]);
    }
}
#endif  // __CUDACC__

// ---- host: tensor maps (driver entry point fetched through the runtime: no -lcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// digit planes [nd][rows][kdim] int8 (k fastest): boxes of KB k-values x box_rows rows x nd planes, 64-byte swizzle
inline bool make_map(CUtensorMap *map, const int8_t *base, int kdim, int rows, int nd, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)kdim, (cuuint64_t)rows, (cuuint64_t)nd};
    cuuint64_t strides[2] = {(cuuint64_t)kdim, (cuuint64_t)kdim * (cuuint64_t)rows};
    cuuint32_t box[3] = {(cuuint32_t)KB, (cuuint32_t)box_rows, (cuuint32_t)nd};
    cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc
}  // namespace mbrf
