// Standalone correctness/speed test of the tcgen05 3xTF32 GEMM tile (csrc/tc_gemm.cuh).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tc_gemm_test tc_gemm_test.cu
#include "../../multiband_rf_pulse_design_b200/csrc/tc_gemm.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace mbrf::tc;

static float tf32_round(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

int main(int argc, char **argv)
{
    const int kdim = argc > 1 ? atoi(argv[1]) : 512, R = argc > 2 ? atoi(argv[2]) : 256, Bp = argc > 3 ? atoi(argv[3]) : 128;
    const int passes = argc > 4 ? atoi(argv[4]) : 3;
    std::vector<double> A((size_t)kdim * R), X((size_t)kdim * Bp);
    std::vector<float> Ah(A.size()), Al(A.size()), Xh(X.size()), Xl(X.size());
    srand(1);
    for (size_t i = 0; i < A.size(); ++i) { A[i] = (rand() / (double)RAND_MAX - 0.5); Ah[i] = tf32_round((float)A[i]); Al[i] = tf32_round((float)(A[i] - Ah[i])); }
    for (size_t i = 0; i < X.size(); ++i) { X[i] = (rand() / (double)RAND_MAX - 0.5); Xh[i] = tf32_round((float)X[i]); Xl[i] = tf32_round((float)(X[i] - Xh[i])); }
    float *dAh, *dAl, *dXh, *dXl;
    double *dC;
    cudaMalloc(&dAh, A.size() * 4); cudaMalloc(&dAl, A.size() * 4); cudaMalloc(&dXh, X.size() * 4); cudaMalloc(&dXl, X.size() * 4);
    cudaMalloc(&dC, (size_t)R * Bp * 8);
    cudaMemcpy(dAh, Ah.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dAl, Al.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dXh, Xh.data(), X.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dXl, Xl.data(), X.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xff, (size_t)R * Bp * 8);
    CUtensorMap mAh, mAl, mXh, mXl;
    if (!make_map(&mAh, dAh, kdim, R, R) || !make_map(&mAl, dAl, kdim, R, R) || !make_map(&mXh, dXh, kdim, Bp, Bp) ||
        !make_map(&mXl, dXl, kdim, Bp, Bp)) { printf("tensor map failed\n"); return 1; }
    cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    Params p; p.C = dC; p.slab = 0; p.ldc = Bp; p.kdim_total = kdim; p.kchunk = kdim; p.passes = passes;
    dim3 grid(Bp / TN, R / TM, 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    tc_gemm_kernel<<<grid, THREADS, SMEM_BYTES>>>(mAh, mAl, mXh, mXl, p);
    cudaError_t err = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(err));
    if (err != cudaSuccess) return 2;
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) tc_gemm_kernel<<<grid, THREADS, SMEM_BYTES>>>(mAh, mAl, mXh, mXl, p);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    std::vector<double> C((size_t)R * Bp);
    cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int r = 0; r < R; r += 7)
        for (int b = 0; b < Bp; b += 5) {
            double s = 0;
            for (int k = 0; k < kdim; ++k) s += (passes == 3 ? A[(size_t)k * R + r] * X[(size_t)k * Bp + b] : (double)Ah[(size_t)k * R + r] * Xh[(size_t)k * Bp + b]);
            maxerr = fmax(maxerr, fabs(s - C[(size_t)r * Bp + b])); maxref = fmax(maxref, fabs(s));
        }
    printf("kdim %d R %d Bp %d passes %d: max err %.3e (max |ref| %.3f)  %.3f ms  %.1f TFLOP/s effective\n", kdim, R, Bp, passes, maxerr,
           maxref, ms, 2.0 * kdim * R * Bp / ms / 1e9);
    return maxerr < 1e-4 * (passes == 3 ? 0.02 : 1.0) * fmax(maxref, 1.0) + 1e-4 ? 0 : 3;
}
