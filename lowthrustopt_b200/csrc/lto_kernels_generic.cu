// lto_kernels_generic.cu -- general kernels: one thread per leg (direct) / per segment
// (indirect), every mode and system.  Stage storage lives in local memory, so these are
// the coverage path (ADAPTIVE direct, 14-dim, odd sizes) and the on-device cross-check
// of the throughput kernels, not the headline.
#include "lto_internal.h"

namespace lto {

// ---------------------------------------------------------------------------
// direct: defectCalc / jacobianCalc (multiShoot_CRTBP_direct.jl:66-143).
// Thread 2s   = forward  leg of segment s (node a, td = +1)   (:82-86)
// Thread 2s+1 = backward leg of segment s (node b, td = -1)   (:88-98)
// The pair sits in adjacent lanes; the defect (:101) is formed with one shuffle.
// ---------------------------------------------------------------------------
template <int NS, bool SENS>
__global__ void __launch_bounds__(64) k_direct_generic(DirectArgs a) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long s = tid >> 1;
    const int back = (int)(tid & 1);
    const bool active = s < a.n_seg;
    double xend[NS];
    double S[SENS ? NS * (NS + 3) : 1];
    double me = 0.0;
    int st = 0, natt = 0;
    if (active) {
        const long long ia = lto_node_a(s, a.npt);
        const double* X = back ? a.Xb : a.Xa;
        const double* U = back ? a.ub : a.ua;
        double x0[NS], u[3];
#pragma unroll
        for (int i = 0; i < NS; ++i) x0[i] = X[ia * NS + i];
#pragma unroll
        for (int i = 0; i < 3; ++i) u[i] = U[ia * 3 + i];
        const double ta = a.ta[ia], tb = a.tb[ia];
        const double tmid = ta + (tb - ta) / 2.0;                 // :70
        st = ep_leg<NS, SENS>(x0, u, back, ta, tmid, a.cfg, a.c, xend, S, &me, &natt);
    } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) xend[i] = 0.0;
    }
    // pair exchange
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        const double other = __shfl_xor_sync(full, xend[i], 1);
        if (active && !back) a.defect[s * NS + i] = xend[i] - other;      // :101
    }
    const double me_o = __shfl_xor_sync(full, me, 1);
    const int st_o = __shfl_xor_sync(full, st, 1);
    if (active && !back) {
        if (a.errors) a.errors[s] = fmax(me, me_o);                        // :104
        if (a.status) a.status[s] = st ? st : st_o;
    }
    if (SENS && active) {
        constexpr int NV = 2 * (NS + 3);
        double* J = a.jac + s * (long long)(NS * NV);
        const double sg = back ? -1.0 : 1.0;
        for (int j = 0; j < NS + 3; ++j) {
            const int col = (j < NS) ? (back ? NS + j : j) : (2 * NS + (back ? 3 : 0) + (j - NS));
            for (int i = 0; i < NS; ++i) J[col * NS + i] = sg * S[j * NS + i];
        }
    }
}

// ---------------------------------------------------------------------------
// indirect: defectCalc / jacobianCalc (multiShoot_CRTBP_indirect.jl:63-124), thread per segment.
// ---------------------------------------------------------------------------
template <int ND, bool SENS>
__global__ void __launch_bounds__(64) k_indirect_generic(IndirectArgs a) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n_seg) return;
    const long long ia = lto_node_a(s, a.npt);
    const long long it = lto_traj_of(s, a.npt);
    double x0[ND], xend[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) x0[i] = a.x0[ia * ND + i];
    const double tl = a.thrustLimit_arr ? a.thrustLimit_arr[it] : a.c.thrustLimit;
    const double rho = a.rho_arr ? a.rho_arr[it] : a.c.rho;
    int na = 0, nt = 0;
    double* Phi = SENS ? a.phi + s * (long long)(ND * ND) : nullptr;
    double PhiL[SENS ? ND * ND : 1];
    int st = sc_seg<ND, SENS>(x0, a.t0[ia], a.t1[ia], a.cfg, a.c, tl, rho, xend, PhiL, &na, &nt);
    if (SENS) for (int i = 0; i < ND * ND; ++i) Phi[i] = PhiL[i];
#pragma unroll
    for (int i = 0; i < ND; ++i)
        a.defect[s * ND + i] = a.x_target ? xend[i] - a.x_target[ia * ND + i] : xend[i];   // :82
    if (a.status) a.status[s] = st;
    if (a.nsteps_out) { a.nsteps_out[2 * s] = na; a.nsteps_out[2 * s + 1] = nt; }
}

cudaError_t launch_direct_generic(const DirectArgs& a, int nstate, cudaStream_t st, int* n_launch) {
    *n_launch = 0;
    if (a.n_seg <= 0) return cudaSuccess;
    const int block = 64;
    const long long threads = 2 * a.n_seg;
    const unsigned grid = (unsigned)((threads + block - 1) / block);
    const bool sens = a.jac != nullptr;
    if (nstate == 6) { if (sens) k_direct_generic<6, true><<<grid, block, 0, st>>>(a); else k_direct_generic<6, false><<<grid, block, 0, st>>>(a); }
    else if (nstate == 7) { if (sens) k_direct_generic<7, true><<<grid, block, 0, st>>>(a); else k_direct_generic<7, false><<<grid, block, 0, st>>>(a); }
    else return cudaErrorInvalidValue;
    *n_launch = 1;
    return cudaGetLastError();
}

cudaError_t launch_indirect_generic(const IndirectArgs& a, int ndim, cudaStream_t st, int* n_launch) {
    *n_launch = 0;
    if (a.n_seg <= 0) return cudaSuccess;
    const int block = 64;
    const unsigned grid = (unsigned)((a.n_seg + block - 1) / block);
    const bool sens = a.phi != nullptr;
    if (ndim == 12) { if (sens) k_indirect_generic<12, true><<<grid, block, 0, st>>>(a); else k_indirect_generic<12, false><<<grid, block, 0, st>>>(a); }
    else if (ndim == 14) { if (sens) k_indirect_generic<14, true><<<grid, block, 0, st>>>(a); else k_indirect_generic<14, false><<<grid, block, 0, st>>>(a); }
    else return cudaErrorInvalidValue;
    *n_launch = 1;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Row sums of squares: out[r] = sum_i v[r*len + i]^2 -- the line searches' merit value
// er[ind] = sum(defect[:].^2) (multiShoot_CRTBP_indirect.jl:241, multiShoot_CRTBP_direct.jl:425) per trial
// trajectory, so that only one double per trajectory leaves the GPU.  One warp per row, fixed summation
// order (deterministic).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sumsq_rows(const double* __restrict__ v, long long n_rows, long long len, double* __restrict__ out) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const double* p = v + row * len;
    double s = 0.0;
    for (long long i = lane; i < len; i += 32) s = fma(p[i], p[i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = s;
}

cudaError_t launch_sumsq_rows(const double* v, long long n_rows, long long len, double* out, cudaStream_t st) {
    if (n_rows <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n_rows * 32 + 255) / 256);
    k_sumsq_rows<<<grid, 256, 0, st>>>(v, n_rows, len, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// FP64 issue-rate probe: CHAINS independent DFMA chains per thread, register resident.
// ---------------------------------------------------------------------------
constexpr int PROBE_CHAINS = 8;
__global__ void __launch_bounds__(256) k_fp64_probe(int iters, double* sink) {
    double a[PROBE_CHAINS];
    const double m = 1.0 + 1e-9 * (threadIdx.x & 7), c = 1e-7;
#pragma unroll
    for (int i = 0; i < PROBE_CHAINS; ++i) a[i] = 1.0 + i * 1e-3 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < PROBE_CHAINS; ++i) a[i] = fma(a[i], m, c);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < PROBE_CHAINS; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;   // never true; keeps the chains alive
}

cudaError_t launch_fp64_probe(int iters, double* d_sink, int n_sm, cudaStream_t st, long long* n_threads, int* chains) {
    const int block = 256, per_sm = 8;   // 64 warps / SM
    const unsigned grid = (unsigned)(n_sm * per_sm);
    k_fp64_probe<<<grid, block, 0, st>>>(iters, d_sink);
    *n_threads = (long long)grid * block;
    *chains = PROBE_CHAINS * 8;          // FMAs per thread per iteration
    return cudaGetLastError();
}

}  // namespace lto
