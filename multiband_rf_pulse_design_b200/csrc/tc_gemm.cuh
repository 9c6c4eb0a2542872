// tcgen05 (5th-gen tensor core) GEMM tile for the PDHG products, 3xTF32 split precision.
//
//   C[R x Bp] (fp64) = AT^T * X,   AT: [kdim x R] fp32,  X: [kdim x Bp] fp32, both k-major exactly like the fp64
//   kernels (the output index is the contiguous one), each given as a tf32-exact "hi" part and a tf32 "lo"
//   remainder:   a*x ~= a_hi*x_hi + a_hi*x_lo + a_lo*x_hi     (the dropped a_lo*x_lo term is ~2^-22 relative)
//
// One CTA computes a 128 x 128 output tile: accumulators live in TMEM (128 lanes x 128 fp32 columns); operand
// tiles of 32 k-rows arrive by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle, boxes of 32 k x 32 outputs = 4 KB) into a
// 3-stage shared-memory ring; one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=128, K=8, both operands
// MN-major), tcgen05.commit releases ring slots and finally signals the epilogue warps, which read the
// accumulator with tcgen05.ld (32 lanes x 32 columns per instruction) and store it widened to fp64.
// Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM allocator, 2..5 = epilogue (TMEM lane quarter = warp % 4).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mbrf {
namespace tc {

constexpr int TM = 128, TN = 128, TBK = 32, STAGES = 3, THREADS = 192;
constexpr int OP_BYTES = TBK * 128 * 4;            // one operand tile: 32 k-rows x 128 outputs x fp32 = 16 KB
constexpr int STAGE_BYTES = 4 * OP_BYTES;          // A_hi, A_lo, X_hi, X_lo
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct Params {
    double *C;          // [R x ldc] fp64, slab blockIdx.z at C + z*slab
    long long slab;
    int ldc;
    int kdim_total, kchunk;   // reduction range of slab z: [z*kchunk, min(kdim_total, (z+1)*kchunk)), multiples of TBK
    int passes;               // 3: split precision (hi/lo), 1: plain tf32 (tests)
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
__device__ __forceinline__ void tma_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int c_inner, int c_outer)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s32(dst)), "l"((uint64_t)map), "r"(s32(bar)), "r"(c_inner), "r"(c_outer) : "memory");
}
// shared-memory matrix descriptor, MN-major, 128-byte swizzle (cute::UMMA::SmemDescriptor, version 1):
// blocks of 32 outputs (128 B) are `lbo` bytes apart, groups of 8 k-rows are `sbo` bytes apart.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
               const __grid_constant__ CUtensorMap mXh, const __grid_constant__ CUtensorMap mXl, const Params p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // swizzle atoms need 1024-B alignment
    uint64_t *full = (uint64_t *)(smem + STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *accf = empty + STAGES;
    uint32_t *tmem_slot = (uint32_t *)(accf + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.y * TM, col0 = blockIdx.x * TN;
    const int k_begin = blockIdx.z * p.kchunk;
    const int k_end = min(p.kdim_total, k_begin + p.kchunk);
    const int nk = (k_end - k_begin) / TBK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { bar_init(&full[s], 1); bar_init(&empty[s], 1); }
        bar_init(accf, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: 128 columns (power of two >= 32) for the 128 x 128 fp32 accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(s32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer ----------------
        for (int it = 0; it < nk; ++it) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            bar_wait(&empty[s], ph ^ 1);
            uint8_t *st = smem + s * STAGE_BYTES;
            bar_expect(&full[s], p.passes == 3 ? 4 * OP_BYTES : 2 * OP_BYTES);
            const int k0 = k_begin + it * TBK;
#pragma unroll
            for (int b = 0; b < 4; ++b) {   // four 32-output blocks per operand tile
                tma_2d(&mAh, &full[s], st + 0 * OP_BYTES + b * 4096, row0 + b * 32, k0);
                tma_2d(&mXh, &full[s], st + 2 * OP_BYTES + b * 4096, col0 + b * 32, k0);
                if (p.passes == 3) {
                    tma_2d(&mAl, &full[s], st + 1 * OP_BYTES + b * 4096, row0 + b * 32, k0);
                    tma_2d(&mXl, &full[s], st + 3 * OP_BYTES + b * 4096, col0 + b * 32, k0);
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer ----------------
        // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both MN-major, N = 128, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TN >> 3) << 17) |
                               ((uint32_t)(TM >> 4) << 24);
        for (int it = 0; it < nk; ++it) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            bar_wait(&full[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t base = s32(smem + s * STAGE_BYTES);
#pragma unroll
            for (int kb = 0; kb < TBK / 8; ++kb) {   // one MMA covers K = 8 (32 bytes of tf32)
                const uint64_t dAh = smem_desc(base + 0 * OP_BYTES + kb * 1024, 4096, 1024);
                const uint64_t dAl = smem_desc(base + 1 * OP_BYTES + kb * 1024, 4096, 1024);
                const uint64_t dXh = smem_desc(base + 2 * OP_BYTES + kb * 1024, 4096, 1024);
                const uint64_t dXl = smem_desc(base + 3 * OP_BYTES + kb * 1024, 4096, 1024);
                const uint32_t first = (it > 0 || kb > 0) ? 1u : 0u;
                if (p.passes == 3) {   // small terms first
                    mma_tf32(tmem, dAl, dXh, idesc, first);
                    mma_tf32(tmem, dAh, dXl, idesc, 1u);
                    mma_tf32(tmem, dAh, dXh, idesc, 1u);
                } else {
                    mma_tf32(tmem, dAh, dXh, idesc, first);
                }
            }
            mma_commit(&empty[s]);          // slot free once these MMAs have read it
        }
        mma_commit(accf);                   // accumulator complete
    } else if (warp >= 2) {
        // ---------------- epilogue: TMEM -> registers -> fp64 global ----------------
        bar_wait(accf, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int row = row0 + q * 32 + lane;
        double *out = p.C + (size_t)blockIdx.z * p.slab + (size_t)row * p.ldc + col0;
#pragma unroll
        for (int c0 = 0; c0 < TN; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (nk > 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 2)
                    *reinterpret_cast<double2 *>(out + c0 + j) =
                        make_double2((double)__uint_as_float(r[j]), (double)__uint_as_float(r[j + 1]));
            } else {
#pragma unroll
                for (int j = 0; j < 32; j += 2) *reinterpret_cast<double2 *>(out + c0 + j) = make_double2(0.0, 0.0);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
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

// fp32 matrix [rows(k) x cols(outputs)] with leading dimension ld (elements): boxes of 32 k-rows x 32 outputs, 128B swizzle
inline bool make_map(CUtensorMap *map, const float *base, int rows, int cols, int ld)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {32, (cuuint32_t)TBK};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc
}  // namespace mbrf
