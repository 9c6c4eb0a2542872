// Microbenchmark: does a non-FP64 instruction issue "for free" next to DFMAs on sm_100a?
// Each variant runs 8 independent DFMA chains per thread; variants add k other instructions
// per 8 DFMAs (integer ALU, FP32 FMA, shared-memory broadcast load, constant-bank operand).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int VARIANT>
__global__ void __launch_bounds__(256) k(double *out, int iters, double seed, const int *ip)
{
    __shared__ double sm[256];
    sm[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed + threadIdx.x + j;
    const double m = 0.999999, c = 1e-9;
    int x = ip[0], y = threadIdx.x;
    float f = (float)seed, g = 1.0001f;
    double acc2 = 0;
    int idx = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fma(a[j], m, c);
            if (VARIANT == 1) { y = y * 3 + x; }                                  // 1 IMAD per 8 DFMA
            if (VARIANT == 2) { y = y * 3 + x; x = x ^ y; y += x; x = x * 5 + y; } // 4 int per 8 DFMA
            if (VARIANT == 3) { f = fmaf(f, g, 0.5f); }                           // 1 FFMA per 8
            if (VARIANT == 4) { f = fmaf(f, g, 0.5f); g = fmaf(g, f, 0.25f); f = fmaf(f, g, 0.5f); g = fmaf(g, f, 0.25f); }
            if (VARIANT == 5) { idx = (idx + 1) & 255; acc2 += sm[idx]; }         // LDS (+ 1 DADD) per 8
            if (VARIANT == 6) { idx = (idx + 2) & 255; a[0] += sm[idx]; a[1] += sm[idx + 1]; }  // LDS.128-ish
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 123.456 || y == 12345 || f == 7.0f || acc2 == 3.3) out[0] = s + y + f + x + g + acc2;
}

template <int V>
void run(const char *name, double *d, int *ip, double extra_dfma)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 2048, blocks = 148 * 8, threads = 256;
    float best = 1e9;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        k<V><<<blocks, threads>>>(d, iters, 1.0, ip);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float t;
        cudaEventElapsedTime(&t, e0, e1);
        if (r && t < best) best = t;
    }
    double dfma = (8.0 + extra_dfma) * 16 * iters * (double)blocks * threads;
    printf("%-40s %8.3f ms   %6.2f TFLOP/s (DFMA only counted)\n", name, best, 2 * dfma / best / 1e9);
}

int main()
{
    double *d;
    int *ip;
    cudaMalloc(&d, 8);
    cudaMalloc(&ip, 4);
    cudaMemset(ip, 0, 4);
    run<0>("8 DFMA", d, ip, 0);
    run<1>("8 DFMA + 1 IMAD", d, ip, 0);
    run<2>("8 DFMA + 4 int", d, ip, 0);
    run<3>("8 DFMA + 1 FFMA", d, ip, 0);
    run<4>("8 DFMA + 4 FFMA", d, ip, 0);
    run<5>("8 DFMA + 1 LDS + 1 DADD + 2 int", d, ip, 1);
    run<6>("8 DFMA + 2 LDS + 2 DADD + 2 int", d, ip, 2);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
