// Microbenchmark (diagnostics, not product code): does a stream of SHFL.32 instructions interleaved with DFMAs
// slow the FP64 pipe on sm_100a?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_shfl fp64_shfl.cu && ./fp64_shfl
#include <cstdio>
#include <cuda_runtime.h>

template <int NSHFL, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1) k(int iters, double* sink, long long* cyc) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-3 * i + 1e-6 * threadIdx.x;
    const double m = 1.0 + 1e-9 * (threadIdx.x & 7);
    const int lane = threadIdx.x & 31;
    const int base = lane - lane % 3;
    const int s1 = (lane < 30) ? base + (lane - base + 1) % 3 : lane, s2 = (lane < 30) ? base + (lane - base + 2) % 3 : lane;
    double x = a[0], acc = 0.0;
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int b = 0; b < 5; ++b) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, 1e-7);          // 40 DFMA per iteration
            if (b < NSHFL) {                                                   // 2 SHFL.32 per double shuffle
                const double y = __shfl_sync(0xffffffffu, a[b], (b & 1) ? s1 : s2);
                acc += y;                                                      // +1 DADD
            }
        }
    }
    const long long c1 = clock64();
    double s = acc + x;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;
    if (lane == 0) cyc[blockIdx.x * 32 + (threadIdx.x >> 5)] = c1 - c0;
}

template <int NSHFL, int WARPS>
void run(int n_sm, double* sink, long long* cyc) {
    const int iters = 20000;
    k<NSHFL, WARPS><<<n_sm, 32 * WARPS>>>(iters, sink, cyc);
    cudaDeviceSynchronize();
    long long h[32 * 148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double mean = 0; for (int w = 0; w < WARPS; ++w) mean += (double)h[w];
    mean /= WARPS;
    const double fp64 = 40.0 + NSHFL;     // DFMA + DADD per iteration per warp
    printf("warps/SM %2d  shfl-doubles/iter %d : %.1f cycles/iter/warp, FP64 pipe use %.1f%% (2 cycles per warp FP64 instr per SMSP)\n", WARPS, NSHFL,
           mean / iters, 100.0 * fp64 * 2.0 * (WARPS / 4.0) / (mean / iters));
}

int main() {
    int n_sm = 0; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    double* sink; long long* cyc;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, sizeof(long long) * 32 * 148);
    run<0, 8>(n_sm, sink, cyc); run<2, 8>(n_sm, sink, cyc); run<5, 8>(n_sm, sink, cyc);
    run<0, 16>(n_sm, sink, cyc); run<2, 16>(n_sm, sink, cyc); run<5, 16>(n_sm, sink, cyc);
    run<0, 12>(n_sm, sink, cyc); run<5, 12>(n_sm, sink, cyc);
    return 0;
}
