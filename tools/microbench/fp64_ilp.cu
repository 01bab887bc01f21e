// Microbenchmark (diagnostics): DFMA issue rate of ONE warp as a function of its instruction-level parallelism (independent
// accumulator chains) and of the number of resident warps per SM sub-partition, straight-line body of ~1024 DFMAs (16 KB).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_ilp fp64_ilp.cu && ./fp64_ilp
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__(512, 1) k(int iters, double* sink, long long* cyc) {
    constexpr int BODY = 1056 / ILP * ILP;
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = 1.0 + 1e-3 * i + 1e-6 * threadIdx.x;
    const double m = 1.0 + 1e-9 * (threadIdx.x & 7);
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int b = 0; b < BODY / ILP; ++b) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, 1e-7);
        }
    }
    const long long c1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 16 + (threadIdx.x >> 5)] = c1 - c0;
}

template <int ILP>
void run(int warps) {
    constexpr int BODY = 1056 / ILP * ILP;
    double* sink; long long* cyc;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 148 * 16 * 8);
    const int iters = 2000;
    k<ILP><<<148, warps * 32>>>(8, sink, cyc);
    k<ILP><<<148, warps * 32>>>(iters, sink, cyc);
    long long h[16];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    const double per = (double)h[0] / ((double)iters * BODY);
    printf("ILP %2d  warps/SMSP %d: %.2f cycles/DFMA/warp -> FP64 pipe %.0f%%  %s\n", ILP, warps / 4, per, 100.0 * 2.0 * (warps / 4.0) / per,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(sink); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 12, 16}) { run<1>(w); run<2>(w); run<3>(w); run<4>(w); run<6>(w); run<8>(w); run<12>(w); }
    return 0;
}
