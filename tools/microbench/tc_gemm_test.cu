// Standalone correctness/speed test of the tcgen05 split-integer GEMM (csrc/tc_gemm.cuh): fp64 operands -> digit
// planes -> int8 tensor-core level products -> fp64 result, compared with a long-double reference.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tc_gemm_test tc_gemm_test.cu
// usage: tc_gemm_test kdim R Bp [P split-K slabs]
#define TC_TIMING
#include "../../multiband_rf_pulse_design_b200/csrc/tc_gemm.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace mbrf::tc;

template <int ND>
static int run(int kdim, int R, int Bp, int P)
{
    std::vector<double> A((size_t)R * kdim), X((size_t)kdim * Bp);
    srand(1);
    for (int r = 0; r < R; ++r) {
        const double amp = ldexp(1.0, (r % 7) - 3);   // rows of different magnitude
        for (int k = 0; k < kdim; ++k) A[(size_t)r * kdim + k] = amp * cos(0.001 * (r + 1) * (k + 1));
    }
    for (int k = 0; k < kdim; ++k)
        for (int b = 0; b < Bp; ++b) {
            const double u = rand() / (double)RAND_MAX - 0.5;
            X[(size_t)k * Bp + b] = (b % 5 == 0 && k % 3) ? 0.0 : u * ldexp(1.0, (b % 11) - 5) * ((k % 17 == 0) ? 1e-6 : 1.0);
        }
    double *dA, *dX, *dC, *dsa, *dsx, *dmx;
    int8_t *pA, *pX;
    cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dX, X.size() * 8); cudaMalloc(&dC, (size_t)P * R * Bp * 8);
    cudaMalloc(&dsa, R * 8); cudaMalloc(&dsx, Bp * 8); cudaMalloc(&dmx, Bp * 8);
    cudaMalloc(&pA, (size_t)ND * R * kdim); cudaMalloc(&pX, (size_t)ND * Bp * kdim);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xff, (size_t)P * R * Bp * 8);
    slice_rows_kernel<ND><<<(R * 32 + 255) / 256, 256>>>(dA, kdim, R, kdim, pA, dsa);
    CUtensorMap mA, mX;
    if (!make_map(&mA, pA, kdim, R, ND, TN) || !make_map(&mX, pX, kdim, Bp, ND, TM)) { printf("tensor map failed\n"); return 1; }
    cudaFuncSetAttribute(tc_i8_gemm_kernel<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(ND));
    Params p;
    p.C = dC; p.slab = (long long)R * Bp; p.ldc = Bp; p.R = R; p.kdim_total = kdim;
    p.kchunk = ((kdim + P - 1) / P + KB - 1) / KB * KB; p.sa = dsa; p.sx = dsx;
    p.nslab = P;
    const int ntiles = (R / TN) * ((Bp + TM - 1) / TM) * P;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    dim3 grid(ntiles < nsm ? ntiles : nsm, 1, 1);
    auto product = [&]() {
        cudaMemsetAsync(dmx, 0, Bp * 8);
        col_absmax_kernel<<<dim3(Bp / 64, 16), 256>>>(dX, kdim, Bp, dmx);
        slice_cols_kernel<ND><<<dim3(Bp / 64, kdim / 64), 256>>>(dX, kdim, Bp, dmx, pX, dsx, nullptr);
        tc_i8_gemm_kernel<ND><<<grid, THREADS, smem_bytes(ND)>>>(mA, mX, p);
    };
    product();
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("ND %d launch: %s\n", ND, cudaGetErrorString(err)); return 2; }
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) product();
    cudaEventRecord(e1);
    for (int i = 0; i < 20; ++i) tc_i8_gemm_kernel<ND><<<grid, THREADS, smem_bytes(ND)>>>(mA, mX, p);
    cudaEventRecord(e2); cudaEventSynchronize(e2);
    float ms_all, ms_mm; cudaEventElapsedTime(&ms_all, e0, e1); cudaEventElapsedTime(&ms_mm, e1, e2);
    ms_all /= 20; ms_mm /= 20;
    std::vector<double> C((size_t)P * R * Bp);
    cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost);
    double maxrel = 0, maxabs = 0;
    for (int r = 0; r < R; r += 5)
        for (int b = 0; b < Bp; b += 3) {
            long double s = 0, sabs = 0;
            double amax = 0, xmax = 0;
            for (int k = 0; k < kdim; ++k) {
                const double a = A[(size_t)r * kdim + k], x = X[(size_t)k * Bp + b];
                s += (long double)a * x; sabs += fabsl((long double)a * x);
                amax = fmax(amax, fabs(a)); xmax = fmax(xmax, fabs(x));
            }
            double c = 0;
            for (int z = 0; z < P; ++z) c += C[(size_t)z * R * Bp + (size_t)r * Bp + b];
            const double e = fabs((double)(s - c));
            maxabs = fmax(maxabs, e);
            if (amax * xmax > 0) maxrel = fmax(maxrel, e / (amax * xmax * sqrt((double)kdim)));
            if (!(e == e)) { printf("NaN at %d %d\n", r, b); return 3; }
        }
    {   // phase timeline of a few CTAs (clock64 deltas in cycles; stamps: 0 start, 1 setup done, 2 first stage landed, 3 MMAs issued,
        // 4 accumulators complete, 5 TMEM drained, 6 tile stored) -- first tile of every CTA
        const size_t nct = (size_t)grid.x;
        long long *dt;
        cudaMalloc(&dt, nct * 64); cudaMemset(dt, 0, nct * 64);
        cudaMemcpyToSymbol(mbrf::tc::tc_timing, &dt, sizeof dt);
        tc_i8_gemm_kernel<ND><<<grid, THREADS, smem_bytes(ND)>>>(mA, mX, p);
        cudaDeviceSynchronize();
        std::vector<long long> ht(nct * 8);
        cudaMemcpy(ht.data(), dt, nct * 64, cudaMemcpyDeviceToHost);
        long long *nul = nullptr;
        cudaMemcpyToSymbol(mbrf::tc::tc_timing, &nul, sizeof nul);
        double avg[7] = {0};
        for (size_t c = 0; c < nct; ++c)
            for (int i = 2; i < 7; ++i) avg[i] += (double)(ht[c * 8 + i] - ht[c * 8 + i - 1]) / nct;
        printf("   first tile, mean cycles over %zu CTAs: first stage %.0f | issue %.0f | drain %.0f | TMEM->regs %.0f | convert+store %.0f\n", nct,
               avg[2], avg[3], avg[4], avg[5], avg[6]);
        cudaFree(dt);
    }
    const double flops = 2.0 * kdim * R * Bp;
    printf("ND %d kdim %d R %d Bp %d P %d: max abs err %.3e, err/(amax*xmax*sqrt(k)) %.3e | product %.3f ms (gemm %.3f ms = %.1f TFLOP/s fp64-equivalent, %.0f TOP/s int8)\n",
           ND, kdim, R, Bp, P, maxabs, maxrel, ms_all, ms_mm, flops / ms_mm / 1e9, flops * (ND * (ND + 1) / 2) / ms_mm / 1e9);
    cudaFree(dA); cudaFree(dX); cudaFree(dC); cudaFree(dsa); cudaFree(dsx); cudaFree(dmx); cudaFree(pA); cudaFree(pX);
    return 0;
}

int main(int argc, char **argv)
{
    const int kdim = argc > 1 ? atoi(argv[1]) : 512, R = argc > 2 ? atoi(argv[2]) : 256, Bp = argc > 3 ? atoi(argv[3]) : 128;
    const int P = argc > 4 ? atoi(argv[4]) : 1;
    int rc = run<5>(kdim, R, Bp, P);
    if (!rc) rc = run<6>(kdim, R, Bp, P);
    if (!rc) rc = run<4>(kdim, R, Bp, P);
    return rc;
}
