// Microbenchmark (diagnostics): cost of feeding distinct FP64 constants to DFMA on sm_100a.
// Each "group" uses one fresh constant for REUSE independent DFMAs (the RK combinations use every
// tableau entry for 3..9 FMAs).  Variants: 64-bit literal (UMOV pair), shared memory (LDS),
// __constant__ (LDC), kernel parameter (LDCU).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NG = 256;   // groups per loop body
struct Tab { double c[NG]; };
__constant__ double ctab[NG];

template <int REUSE, int MODE>
__global__ void __launch_bounds__(256, 1) k(int iters, double* sink, long long* cyc, Tab tab) {
    __shared__ double stab[NG];
    for (int i = threadIdx.x; i < NG; i += blockDim.x) stab[i] = 1e-7 + 1e-9 * i;
    __syncthreads();
    double a[REUSE];
#pragma unroll
    for (int i = 0; i < REUSE; ++i) a[i] = 1.0 + 1e-3 * i + 1e-6 * threadIdx.x;
    const double m = 1.0 + 1e-9 * (threadIdx.x & 7);
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            double c;
            if (MODE == 0) c = 1e-7 + 1e-9 * (double)g;          // literal
            else if (MODE == 1) c = stab[g];                     // shared
            else if (MODE == 2) c = ctab[g];                     // __constant__
            else c = tab.c[g];                                   // kernel parameter
#pragma unroll
            for (int i = 0; i < REUSE; ++i) a[i] = fma(a[i], m, c);
        }
    }
    const long long c1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < REUSE; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 16 + (threadIdx.x >> 5)] = c1 - c0;
}

template <int REUSE, int MODE>
void run(const char* name, int warps) {
    double* sink; long long* cyc;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 148 * 16 * 8);
    Tab tab; for (int i = 0; i < NG; ++i) tab.c[i] = 1e-7 + 1e-9 * i;
    cudaMemcpyToSymbol(ctab, tab.c, sizeof tab.c);
    const int iters = 2048;
    k<REUSE, MODE><<<148, warps * 32>>>(8, sink, cyc, tab);
    k<REUSE, MODE><<<148, warps * 32>>>(iters, sink, cyc, tab);
    long long h[16];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    const double per = (double)h[0] / ((double)iters * NG * REUSE);
    printf("%-10s reuse %d warps/SMSP %d: %.2f cycles/DFMA/warp -> FP64 pipe %.0f%% %s\n", name, REUSE, warps / 4, per, 100.0 * 2.0 * (warps / 4.0) / per,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(sink); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8}) {
        run<3, 0>("literal", w); run<3, 1>("shared", w); run<3, 2>("constant", w); run<3, 3>("param", w);
        run<9, 0>("literal", w); run<9, 1>("shared", w); run<9, 2>("constant", w); run<9, 3>("param", w);
    }
    return 0;
}
