// Microbenchmark (diagnostics): the column warps' attempt body of the half-column K3 (lto_indirect_hc.cu col_attempt) ALONE -- no
// state warps, no hand-off, fake stage records in shared memory -- as a function of resident warps per SM sub-partition.
// Separates what the column code costs by itself from what it loses to the rest of the kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I ../../include -o hc_colbench hc_colbench.cu
#include "../../lowthrustopt_b200/csrc/lto_indirect_hc.cu"
#include <cstdio>
using namespace lto;
using namespace lto::ihc;

// POLLUTE: the warps above the first 8 run a DIFFERENT straight-line body (POLLUTE x 1024 dependent integer multiply-adds, 16 KB
// each; no FP64, one instruction in flight) for as long as the column warps work: a second instruction stream per SM sub-partition.
template <int KB> __device__ __noinline__ unsigned pollute(unsigned x, volatile int* stop) {
    while (!*stop) {
#pragma unroll
        for (int i = 0; i < KB * 64; ++i) x = x * 1664525u + (unsigned)(1013904223u + i);
    }
    return x;
}

// POLLUTE < 0: the extra warps are "state-like": FP64 work with little instruction-level parallelism.
//   -1: one dependent DFMA chain;  -2: two chains;  -3: the state right-hand side itself (sc_eval2<true>, inlined) on a dependent input
__device__ __noinline__ double pollute_fp64(int mode, double x, volatile int* stop) {
    double a = x, b = x + 1.0;
    const double m = 1.0 + 1e-12, c = 1e-13;
    hcm::Law lw; lw.aL = 1e-3; lw.rho_inv = 1.0; lw.rq = 2.5e-4;
    while (!*stop) {
        if (mode == -1) {
#pragma unroll
            for (int i = 0; i < 256; ++i) a = fma(a, m, c);
        } else if (mode == -2) {
#pragma unroll
            for (int i = 0; i < 128; ++i) { a = fma(a, m, c); b = fma(b, m, c); }
        } else {
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
                const double R[3] = {1.1 + 1e-9 * a, 0.01, 0.05}, V[3] = {0.01, 0.17, 0.0}, M[3] = {0.5, -0.7 + 1e-9 * b, 0.4}, N[3] = {0.1, 0.2, 0.0};
                double kr[3], kl[3], U[6], W[6], G[6];
                hcm::sc_eval2<true>(R, V, M, N, 0.01215, 0.98785, 2.0, 1.0, lw, kr, kl, U, W, G);
                a = kr[0] + W[3] + G[4]; b = kl[1] + U[2];
            }
        }
    }
    return a + b;
}

template <bool ERR, int POLLUTE = 0>
__global__ void __launch_bounds__(384, 1) k_colbench(int iters, double* sink, long long* cyc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* rec = reinterpret_cast<double2*>(smem_raw);
    __shared__ int stop, n_done;
    if (threadIdx.x == 0) { stop = 0; n_done = 0; }
    for (int i = threadIdx.x; i < 13 * NC2 * TS; i += blockDim.x) rec[i] = make_double2(0.3 + 1e-3 * (i % 17), -0.2 + 1e-3 * (i % 13));
    __syncthreads();
    if (POLLUTE > 0 && threadIdx.x >= 256) {
        const unsigned r = pollute<POLLUTE>(threadIdx.x, &stop);
        if (r == 12345u) sink[1] = r;
        return;
    }
    if (POLLUTE < 0 && threadIdx.x >= 256) {
        const double r = pollute_fp64(POLLUTE, 1.0 + 1e-6 * threadIdx.x, &stop);
        if (r == 12345.0) sink[1] = r;
        return;
    }
    const int lane = threadIdx.x & 31, g = lane >> 3, half = g & 1, s8 = lane & 7;
    double p[3] = {1.0 + 1e-3 * lane, 0.5, -0.25}, pd[3] = {0.1, -0.3, 0.2 + 1e-3 * lane};
    const double h = 1e-2, w2 = 2.0;
    double es = 0.0;
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double pn[3], pdn[3];
        es += col_attempt<ERR>(p, pd, h, w2, rec + s8, half, 1e-13, 1e-13, pn, pdn);
#pragma unroll
        for (int q = 0; q < 3; ++q) { p[q] = pn[q]; pd[q] = pdn[q]; }
    }
    const long long c1 = clock64();
    if (es == 123.456 || p[0] == 77.0) sink[0] = es + p[0] + pd[1];
    if (lane == 0) cyc[blockIdx.x * 16 + (threadIdx.x >> 5)] = c1 - c0;
    if (POLLUTE != 0 && lane == 0 && atomicAdd(&n_done, 1) == 7) stop = 1;
}

template <bool ERR, int POLLUTE = 0>
void run(int warps) {
    double* sink; long long* cyc;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 148 * 16 * 8);
    cudaFuncSetAttribute(k_colbench<ERR, POLLUTE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)REC_BYTES);
    const int iters = 400;
    k_colbench<ERR, POLLUTE><<<148, warps * 32, REC_BYTES>>>(4, sink, cyc);
    k_colbench<ERR, POLLUTE><<<148, warps * 32, REC_BYTES>>>(iters, sink, cyc);
    long long h[16];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("col_attempt<%d>  warps %2d  other-stream %3d: %.0f cycles per attempt per warp %s\n", (int)ERR, warps, POLLUTE < 0 ? POLLUTE : POLLUTE * 16, (double)h[0] / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(sink); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 12}) run<true>(w);
    for (int w : {4, 8, 12}) run<false>(w);
    // 8 column warps + 4 warps (one per sub-partition) that run another instruction stream of 16 / 32 / 64 KB
    run<true, 1>(12); run<true, 2>(12); run<true, 4>(12);
    // 8 column warps + 4 "state-like" FP64 warps (one per sub-partition): 1 dependent DFMA chain | 2 chains | the state right-hand side
    run<true, -1>(12); run<true, -2>(12); run<true, -3>(12);
    return 0;
}
