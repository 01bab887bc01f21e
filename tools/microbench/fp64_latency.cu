// fp64_latency.cu -- dependent-issue latencies that bound a state warp's chain (one warp, clock64 around a dependent chain).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0); const double p = fma(0.375, e, 0.5); return fma(y * e, p, y);
}
__device__ __forceinline__ double fast_rcp(double x) {
    double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0); return fma(y, fma(e, e, e), y);
}
template <int OP>
__global__ void k(double* out, long long* cyc, double seed, int busy_warps) {
    double x = seed + threadIdx.x * 1e-9, m = 1.0 + 1e-12, c = 1e-13;
    // busy_warps = bit mask of the background warps (warp w sits on SM sub-partition w % 4) that saturate their FP64 pipe
    if ((threadIdx.x >> 5) > 0 && !((busy_warps >> (threadIdx.x >> 5)) & 1)) return;
    if ((threadIdx.x >> 5) > 0) {
        double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3;
        for (int i = 0; i < 40000; ++i) { a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c); }
        if (a0 + a1 + a2 + a3 == 1.2345) out[1] = a0;
        return;
    }
    const int N = 512;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = fma(x, m, c);
        if (OP == 1) x = x * m;
        if (OP == 2) x = x + c;
        if (OP == 3) x = fast_rsqrt(x) + 1.0;
        if (OP == 4) x = fast_rcp(x) + 1.0;
        if (OP == 5) x = exp(x) * 1e-1;
        if (OP == 6) x = sqrt(x) + 1.0;
        if (OP == 7) x = 1.0 / x + 1.0;
        if (OP == 8) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = (t1 - t0); }
    out[threadIdx.x & 31] = x;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 4096); cudaMalloc(&c, 64);
    const char* names[] = {"DFMA", "DMUL", "DADD", "fast_rsqrt+add", "fast_rcp+add", "exp()*c", "IEEE sqrt+add", "IEEE div+add", "shfl.f64"};
    const int masks[4] = {0, 0xfe, 0x10, 0xee};
    const char* what[4] = {"--- one warp alone", "--- with 7 background warps issuing DFMA on the same SM (all sub-partitions)",
                           "--- with ONE background warp on the SAME sub-partition (warp 4)", "--- with 6 background warps on the OTHER three sub-partitions only"};
    for (int bi = 0; bi < 4; ++bi) {
        const int busy = masks[bi];
        printf("%s\n", what[bi]);
        for (int op = 0; op < 9; ++op) {
            long long h = 0; int threads = busy ? 256 : 32;
            for (int rep = 0; rep < 2; ++rep) {
                switch (op) {
                    case 0: k<0><<<1, threads>>>(d, c, 1.0, busy); break; case 1: k<1><<<1, threads>>>(d, c, 1.0, busy); break;
                    case 2: k<2><<<1, threads>>>(d, c, 1.0, busy); break; case 3: k<3><<<1, threads>>>(d, c, 1.5, busy); break;
                    case 4: k<4><<<1, threads>>>(d, c, 1.5, busy); break; case 5: k<5><<<1, threads>>>(d, c, 0.5, busy); break;
                    case 6: k<6><<<1, threads>>>(d, c, 1.5, busy); break; case 7: k<7><<<1, threads>>>(d, c, 1.5, busy); break;
                    case 8: k<8><<<1, threads>>>(d, c, 1.5, busy); break;
                }
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            printf("%-16s %.1f cycles per dependent op\n", names[op], h / 512.0);
        }
    }
    return 0;
}
