// Descriptor bring-up for tcgen05.mma.kind::tf32: one CTA, one K=8 MMA (M=128,N=128), operands scattered into shared
// memory by the threads in a chosen canonical layout.  Prints the max error per layout mode.
//   mode 0: MN-major, 128B swizzle   mode 1: K-major, 128B swizzle   mode 2: MN-major, no swizzle   mode 3: K-major no swizzle
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tc_dbg tc_dbg.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <stdint.h>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout)
{
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
           ((uint64_t)layout << 61);
}

__global__ void __launch_bounds__(128, 1) dbg_kernel(const float *A, const float *B, float *D, int mode, uint32_t lbo, uint32_t sbo)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem, *sB = smem + 16384;
    uint64_t *bar = (uint64_t *)(smem + 32768);
    uint32_t *slot = (uint32_t *)(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 8192; i += 128) ((uint32_t *)smem)[i] = 0;
    __syncthreads();
    // element (mn, k): A[mn*8+k]
    for (int i = tid; i < 128 * 8; i += 128) {
        const int mn = i >> 3, k = i & 7;
        uint32_t off;
        if (mode == 0) { off = (mn >> 5) * lbo + k * 128 + (mn & 31) * 4; off ^= ((off >> 7) & 7) << 4; }
        else if (mode == 1) { off = (mn >> 3) * sbo + (mn & 7) * 128 + k * 4; off ^= ((off >> 7) & 7) << 4; }
        else if (mode == 2) { off = (mn >> 2) * sbo + (mn & 3) * 4 + k * 16; }
        else { off = (mn >> 3) * sbo + (mn & 7) * 16 + (k >> 2) * lbo + (k & 3) * 4; }
        *(float *)(sA + off) = A[i];
        *(float *)(sB + off) = B[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(s32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const bool mn_major = (mode == 0 || mode == 2);
        const uint32_t layout = (mode < 2) ? 2u : 0u;
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        if (mn_major) idesc |= (1u << 15) | (1u << 16);
        const uint64_t da = mk_desc(s32(sA), lbo, sbo, layout), db = mk_desc(s32(sB), lbo, sbo, layout);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u)
                     : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra E;\n\tbra W;\n\tE:\n\t}" ::"r"(s32(bar))
                 : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 128; c0 += 8) {
        uint32_t r[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}


// ---- int8: K-major, 128B swizzle rows of 128 k-values; NK MMAs of K=32 advance the descriptor start address by 32 B ----
template <int N>
__global__ void __launch_bounds__(128, 1) dbg_i8(const int8_t *A, const int8_t *B, int *D, int nk)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem, *sB = smem + 16384;
    uint64_t *bar = (uint64_t *)(smem + 32768);
    uint32_t *slot = (uint32_t *)(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * 128; i += 128) {
        const int mn = i >> 7, k = i & 127;
        uint32_t off = (mn >> 3) * 1024 + (mn & 7) * 128 + k;
        off ^= ((off >> 7) & 7) << 4;
        sA[off] = (uint8_t)A[i];
        if (mn < N) sB[off] = (uint8_t)B[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(s32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int kb = 0; kb < nk; ++kb) {
            const uint64_t da = mk_desc(s32(sA) + kb * 32, 16, 1024, 2), db = mk_desc(s32(sB) + kb * 32, 16, 1024, 2);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc),
                         "r"(kb > 0 ? 1u : 0u)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW8:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra E8;\n\tbra W8;\n\tE8:\n\t}" ::"r"(s32(bar))
                 : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t r[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * N + c0 + j] = (int)r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

template <int N>
static int run_i8(int nk)
{
    std::vector<int8_t> A(128 * 128), B(128 * 128);
    std::vector<int> D(128 * N);
    for (auto &v : A) v = (int8_t)((rand() % 129) - 64);
    for (auto &v : B) v = (int8_t)((rand() % 129) - 64);
    int8_t *dA, *dB; int *dD;
    cudaMalloc(&dA, 16384); cudaMalloc(&dB, 16384); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, A.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 128 * N * 4);
    cudaFuncSetAttribute(dbg_i8<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    dbg_i8<N><<<1, 128, 40000>>>(dA, dB, dD, nk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("i8 N %d nk %d: %s\n", N, nk, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost);
    long long bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            int s = 0;
            for (int k = 0; k < nk * 32; ++k) s += (int)A[m * 128 + k] * (int)B[n * 128 + k];
            bad += s != D[m * N + n];
        }
    printf("i8 N %d nk %d: mismatches %lld of %d  D[0..2]=%d %d %d\n", N, nk, bad, 128 * N, D[0], D[1], D[2]);
    return 0;
}

// ---- A operand staged in TMEM: tcgen05.cp 128x256b (smem -> tmem) of a 64-byte-swizzled K-major tile, then
//      tcgen05.mma with [a_tmem].  Dumps the TMEM copy of A as well. ----
__device__ __forceinline__ uint64_t mk_desc64(uint32_t addr)
{
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__global__ void __launch_bounds__(128, 1) dbg_i8_ts(const int8_t *A, const int8_t *B, int *D, uint32_t *Adump, int variant)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem, *sB = smem + 8192;
    uint64_t *bar = (uint64_t *)(smem + 16384);
    uint32_t *slot = (uint32_t *)(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * 64; i += 128) {          // element (mn, k), k < 64
        const int mn = i >> 6, k = i & 63;
        uint32_t off = (mn >> 3) * 512 + (mn & 7) * 64 + k;
        off ^= ((off >> 7) & 3) << 4;
        sA[off] = (uint8_t)A[mn * 128 + k];
        if (mn < 64) sB[off] = (uint8_t)B[mn * 128 + k];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(s32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    const uint32_t tA = tmem + 64;                         // A staging: 16 columns (2 k-steps of 8 columns)
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int kb = 0; kb < 2; ++kb) {
            const uint64_t da = mk_desc64(s32(sA) + kb * 32);
            if (variant == 0)
                asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tA + kb * 8), "l"(da) : "memory");
        }
        for (int kb = 0; kb < 2; ++kb) {
            const uint64_t db = mk_desc64(s32(sB) + kb * 32);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(tA + kb * 8), "l"(db), "r"(idesc),
                         "r"(kb > 0 ? 1u : 0u)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tWT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra ET;\n\tbra WT;\n\tET:\n\t}" ::"r"(s32(bar))
                 : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 80; c0 += 8) {
        uint32_t r[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) {
            if (c0 < 64) D[(warp * 32 + lane) * 64 + c0 + j] = (int)r[j];
            else Adump[(warp * 32 + lane) * 16 + (c0 - 64) + j] = r[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}
static int run_ts(int variant)
{
    std::vector<int8_t> A(128 * 128), B(128 * 128);
    std::vector<int> D(128 * 64);
    std::vector<uint32_t> Ad(128 * 16);
    for (auto &v : A) v = (int8_t)((rand() % 255) - 127);
    for (auto &v : B) v = (int8_t)((rand() % 255) - 127);
    int8_t *dA, *dB; int *dD; uint32_t *dAd;
    cudaMalloc(&dA, 16384); cudaMalloc(&dB, 16384); cudaMalloc(&dD, 128 * 64 * 4); cudaMalloc(&dAd, 128 * 16 * 4);
    cudaMemcpy(dA, A.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 128 * 64 * 4); cudaMemset(dAd, 0xff, 128 * 16 * 4);
    cudaFuncSetAttribute(dbg_i8_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, 20000);
    dbg_i8_ts<<<1, 128, 20000>>>(dA, dB, dD, dAd, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ts variant %d: %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(Ad.data(), dAd, 128 * 16 * 4, cudaMemcpyDeviceToHost);
    long long bad = 0, badA = 0;
    for (int m = 0; m < 128; ++m) {
        for (int n = 0; n < 64; ++n) {
            int s = 0;
            for (int k = 0; k < 64; ++k) s += (int)A[m * 128 + k] * (int)B[n * 128 + k];
            bad += s != D[m * 64 + n];
        }
        for (int c = 0; c < 16; ++c) {
            uint32_t w = 0;
            for (int j = 0; j < 4; ++j) w |= (uint32_t)(uint8_t)A[m * 128 + 4 * c + j] << (8 * j);
            badA += w != Ad[m * 16 + c];
        }
    }
    printf("A-in-TMEM variant %d: D mismatches %lld of 8192, A-copy mismatches %lld of 2048; row0 A words tmem %08x %08x expect %02x%02x%02x%02x\n", variant, bad, badA,
           Ad[0], Ad[1], (uint8_t)A[3], (uint8_t)A[2], (uint8_t)A[1], (uint8_t)A[0]);
    return 0;
}

int main(int argc, char **argv)
{
    run_ts(0);
    run_i8<128>(1); run_i8<128>(4); run_i8<64>(4); run_i8<32>(2);
    std::vector<float> A(1024), B(1024), D(16384);
    srand(3);
    for (auto &v : A) v = (float)((rand() % 17) - 8) / 8.f;
    for (auto &v : B) v = (float)((rand() % 17) - 8) / 8.f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, 4096); cudaMalloc(&dB, 4096); cudaMalloc(&dD, 65536);
    cudaMemcpy(dA, A.data(), 4096, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), 4096, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(dbg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    struct Cfg { int mode; uint32_t lbo, sbo; } cfgs[] = {
        {0, 4096, 1024}, {0, 1024, 4096}, {0, 1024, 1024}, {1, 16, 1024}, {1, 0, 1024}, {2, 128, 128}, {2, 1024, 128}, {3, 128, 256}, {3, 2048, 128},
    };
    for (auto &c : cfgs) {
        // for layouts whose fill depends on which of lbo/sbo the hardware uses, the scatter above uses the SAME lbo/sbo
        // roles as documented in CUTLASS; alternative role assignments are tried through the swapped entries.
        cudaMemset(dD, 0xff, 65536);
        dbg_kernel<<<1, 128, 40000>>>(dA, dB, dD, c.mode, c.lbo, c.sbo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d lbo %u sbo %u: %s\n", c.mode, c.lbo, c.sbo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, 65536, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0; int nz = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 128; ++n) {
                double s = 0;
                for (int k = 0; k < 8; ++k) s += (double)A[m * 8 + k] * B[n * 8 + k];
                maxerr = fmax(maxerr, fabs(s - D[m * 128 + n])); maxref = fmax(maxref, fabs(s));
                nz += D[m * 128 + n] != 0.f;
            }
        printf("mode %d lbo %u sbo %u: max err %.3e (max ref %.3f) nonzero %d  D[0..3]=%g %g %g %g\n", c.mode, c.lbo, c.sbo, maxerr, maxref, nz,
               D[0], D[1], D[2], D[3]);
    }
    return 0;
}
