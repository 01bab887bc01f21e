// Microbenchmark (diagnostics, not product code): per-warp issue rate of straight-line FP64 code on sm_100a
// as a function of resident warps per SM sub-partition and of the loop-body size.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_frontend fp64_frontend.cu && ./fp64_frontend
#include <cstdio>
#include <cuda_runtime.h>

template <int BODY, int ILP, bool IMM>
__global__ void __launch_bounds__(512, 1) k(int iters, double* sink, long long* cyc) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = 1.0 + 1e-3 * i + 1e-6 * threadIdx.x;
    const double m = 1.0 + 1e-9 * (threadIdx.x & 7);
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int b = 0; b < BODY / ILP; ++b) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (IMM) a[i] = fma(a[i], m, 1e-7 + 1e-9 * (double)(b * ILP + i));   // distinct 64-bit literal per instruction
                else a[i] = fma(a[i], m, 1e-7);
            }
        }
    }
    const long long c1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 16 + (threadIdx.x >> 5)] = c1 - c0;
}

template <int BODY, int ILP, bool IMM>
void run(const char* name, int warps) {
    double* sink; long long* cyc;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 148 * 16 * 8);
    const int iters = (1 << 22) / BODY;
    k<BODY, ILP, IMM><<<148, warps * 32>>>(8, sink, cyc);
    k<BODY, ILP, IMM><<<148, warps * 32>>>(iters, sink, cyc);
    long long h[16];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    const double per = (double)h[0] / ((double)iters * BODY);
    printf("%-28s warps/SM %2d (%.2f per SMSP): %.2f cycles/DFMA/warp  -> SMSP FP64 pipe %.0f%%  %s\n", name, warps, warps / 4.0, per,
           100.0 * 2.0 * (warps / 4.0) / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(sink); cudaFree(cyc);
}

int main() {
    // instruction-cache capacity: loop bodies of 24..96 KB at 2 warps per SM sub-partition
    run<1536, 8, false>("body 24 KB", 8); run<2048, 8, false>("body 32 KB", 8); run<2560, 8, false>("body 40 KB", 8);
    run<3072, 8, false>("body 48 KB", 8); run<3584, 8, false>("body 56 KB", 8); run<4096, 8, false>("body 64 KB", 8);
    run<5120, 8, false>("body 80 KB", 8); run<6144, 8, false>("body 96 KB", 8); run<8192, 8, false>("body 128 KB", 8);
    const int ws[] = {4, 8, 12, 16};
    for (int w : ws) run<4096, 8, false>("body 4096 (64 KB) ilp8", w);
    for (int w : ws) run<2048, 2, false>("body 2048 ilp2", w);
    return 0;
}
