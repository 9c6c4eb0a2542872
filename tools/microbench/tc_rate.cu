// tcgen05.mma.kind::i8 issue/throughput probe: one CTA issues NMMA back-to-back MMAs (M=128, K=32, operands resident in
// shared memory, 64-byte swizzle K-major) for several N and reports cycles per MMA.  A second mode takes A from TMEM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tc_rate tc_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <stdint.h>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc64(uint32_t addr)
{
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int nmma, int a_tmem, int nacc, long long *out)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bar = (uint64_t *)(smem + 65536);
    uint32_t *slot = (uint32_t *)(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16384; i += 128) ((uint32_t *)smem)[i] = 0x01010101u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    long long t0 = 0, t1 = 0;
    uint32_t el = 0;
    if (__shfl_sync(0xffffffffu, warp, 0) == 0)
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
    if (el) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = mk_desc64(s32(smem)), db = mk_desc64(s32(smem) + 16384);
        const uint32_t tA = tmem + 480;
        if (a_tmem) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tA), "l"(da) : "memory");
        uint64_t dA[5], dB[5];
        for (int j = 0; j < 5; ++j) { dA[j] = da + (uint64_t)((j * 8192) >> 4); dB[j] = db + (uint64_t)((j * 4096) >> 4); }
        t0 = clock64();
        for (int i = 0; i < nmma / 15; ++i) {
#pragma unroll
            for (int sa = 0; sa < 5; ++sa)
#pragma unroll
                for (int sx = 0; sx < 5 - sa; ++sx) {
                    const uint32_t d = tmem + (uint32_t)((nacc > 1 ? (sa + sx) : 0) * N);
                    if (a_tmem)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(tA), "l"(dB[sx]), "r"(idesc), "r"(1u) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(dA[sa]), "l"(dB[sx]), "r"(idesc), "r"(1u) : "memory");
                }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
        t1 = clock64();
    }
    asm volatile("{\n\t.reg .pred p;\n\tWR:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra ER;\n\tbra WR;\n\tER:\n\t}" ::"r"(s32(bar)) : "memory");
    if (el) { out[0] = t1 - t0; out[1] = clock64() - t0; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
int main()
{
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    const int nm = 3000;
    for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
        for (int N : {32, 64, 96, 128, 256})
            for (int nacc : {1, 5}) {
                if (nacc * N > (a_tmem ? 480 : 512)) continue;
                rate_kernel<<<1, 128, 70000>>>(N, nm, a_tmem, nacc, d);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
                rate_kernel<<<1, 128, 70000>>>(N, nm, a_tmem, nacc, d);
                cudaDeviceSynchronize();
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("A %s N %3d accumulators %d: issue %.1f clk/MMA, complete %.1f clk/MMA  (math-bound would be %d)\n", a_tmem ? "tmem" : "smem", N, nacc,
                       (double)h[0] / nm, (double)h[1] / nm, N / 2);
            }
    return 0;
}
