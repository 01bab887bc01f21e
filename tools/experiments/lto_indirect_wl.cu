// lto_indirect_wl.cu -- throughput kernel of the indirect method (K3), ndim = 12, "warp-local" layout:
// defectCalc + jacobianCalc of multiShoot_CRTBP_indirect.jl:63-124 for a whole batch in one launch.  Each segment integrates
// [x | Phi] (12 + 144 components) with the adaptive order-8 pair and the OrdinaryDiffEq-style controller of lto_prop_generic.cuh
// (drive_rk8), joint error norm over x and Phi (LTO_NORM_STATE_SENS, the ForwardDiff semantics); Phi replaces
// ForwardDiff.jacobian(f, x0) (:121).
//
// Why this layout (round 2).  K3 (lto_indirect_cw.cu) and K3-hc (lto_indirect_hc.cu) split the work by ROLE:
// state warps produce the stage linearisations, column warps consume them.  Both lose to what the roles cost -- the state chain is
// serial with the column phase of its own tile (K3), or the two instruction streams evict each other from the instruction cache
// and share FP64 pipes (K3-hc).  Here there are no roles and no inter-warp protocol at all:
//   * a WARP owns 8 segment slots outright.  One attempted step of its 8 segments is 7 passes of the same warp:
//       pass S   lane = (copy, half, slot): the state in the second-order variables (lto_hc_math.cuh), r-half and lv-half of a slot
//                on two lanes 8 apart; 13 stages ROLLED into a loop, stage derivatives in the warp's private shared memory (so this
//                pass is ~7 KB of code instead of ~40 KB); publishes U, W, G per stage and slot into the warp's private records.
//       pass 0-5 lane = (column of the pair, half, slot): the 24 half-columns of each slot, 4 per lane-quad and pass, 13 stages
//                unrolled with the 39 stage derivatives in registers (the hot loop, ~19 KB).
//     Then the 26 error partials of a slot are summed by two shuffles and every lane of the slot takes the same accept/reject
//     decision.  No mbarrier, no named barrier, no flag in shared memory; __syncwarp only.
//   * 8 warps per SM, two per sub-partition: while one warp runs its (latency-bound) state pass the other one on the same
//     sub-partition is, most of the time, in a column pass and has the FP64 pipe to itself.
//   * current / candidate half-columns live in an L2-resident scratch (two buffers per warp, the slot's parity bit says which one
//     is current): an accepted step is a flip, a rejected one re-reads.
//   * a finished segment's STM is assembled in the (then idle) record area in the output's own layout and leaves as ONE
//     1152-byte bulk store -- full packets when `phi` is NVLink peer memory of the solver rank.
// Work queue: a slot that finishes pulls the next segment from a global counter; its Hairer-Norsett-Wanner initial step is computed
// by the slot's own lanes (two right-hand sides).
#include "lto_internal.h"
#include "lto_cw_common.cuh"
#include "lto_hc_math.cuh"
#include "lto_prop_generic.cuh"
#include <algorithm>
#include <cstdlib>

namespace lto {
namespace iwl {

using namespace cwc;
using namespace hcm;

constexpr int ND = 12;
constexpr int NWARP = 8;          // warps per CTA, two per SM sub-partition
constexpr int SL = 8;             // segment slots per warp
constexpr int NPASS = 6;          // column passes per attempted step: 2 columns x 2 halves x 8 slots each
constexpr int NTHREADS = 32 * NWARP;
constexpr int NC2 = 9;            // double2 per stage record: U[6] W[6] G[6]

constexpr size_t REC_BYTES = (size_t)13 * NC2 * SL * sizeof(double2);      // stage records of the warp's 8 slots (later: STM staging)
constexpr size_t KS_BYTES = (size_t)13 * 3 * 32 * sizeof(double);          // the state pass's stage derivatives, [stage][component][lane]
constexpr size_t WARP_BYTES = REC_BYTES + KS_BYTES;
constexpr size_t SMEM = NWARP * WARP_BYTES;
static_assert(SMEM <= 232448, "shared-memory plan exceeds 227 KB");
static_assert(WARP_BYTES % 16 == 0 && REC_BYTES % 16 == 0, "16-byte alignment of the per-warp areas");
static_assert((size_t)SL * ND * ND * sizeof(double) <= REC_BYTES, "the STM staging of 8 slots must fit the record area");
constexpr size_t SCR_DOUBLES_PER_WARP = (size_t)2 * NPASS * 6 * 32;        // [parity][pass][component][lane]

__constant__ double T_B[13][13] = LTO_TAB_B_INIT;
__constant__ double T_G[13][13] = LTO_TAB_G_INIT;
__constant__ double T_C[13] = LTO_TAB_C_INIT;
__constant__ double T_CHI[13] = LTO_TAB_CHI_INIT;
__constant__ double T_CHIB[13] = LTO_TAB_CHIB_INIT;

// ---------------------------------------------------------------------------
// Column pass: one attempted RK step of one half-column (p, pd) (lto_hc_math.cuh):
//   k_J = U P_J + X Po_J + C Pd_J,   X = G, Po = dlv-half's position (half 0)  |  X = W, Po = dr-half's position (half 1)
// ---------------------------------------------------------------------------
template <int J>
__device__ __forceinline__ void col_stage(K3& K, const double (&p)[3], const double (&pd)[3], double h, double h2, double w2,
                                          const double2* __restrict__ rec, int xoff) {
    double P[3], Pd[3], Po[3];
    stage_in<J>(K, p, pd, h, h2, P, Pd);
#pragma unroll
    for (int q = 0; q < 3; ++q) Po[q] = __shfl_xor_sync(0xffffffffu, P[q], 8);
    const double2* w = rec + J * NC2 * SL;
    double U[6], X[6];
    { const double2 a = w[0 * SL], b = w[1 * SL], c = w[2 * SL]; U[0] = a.x; U[1] = a.y; U[2] = b.x; U[3] = b.y; U[4] = c.x; U[5] = c.y; }
    { const double2 a = w[xoff], b = w[xoff + SL], c = w[xoff + 2 * SL]; X[0] = a.x; X[1] = a.y; X[2] = b.x; X[3] = b.y; X[4] = c.x; X[5] = c.y; }
    double k[3];
    col_rhs(U, X, w2, P, Pd, Po, k);
#pragma unroll
    for (int q = 0; q < 3; ++q) K.k[J][q] = k[q];
}

__device__ __forceinline__ double col_attempt(const double (&p)[3], const double (&pd)[3], double h, double w2, const double2* __restrict__ rec,
                                              int half, double atol, double rtol, double (&pn)[3], double (&pdn)[3]) {
    const double h2 = h * h;
    const int xoff = half ? 3 * SL : 6 * SL;                            // W for the dlv-half, G for the dr-half
    K3 K;
    col_stage<0>(K, p, pd, h, h2, w2, rec, xoff);  col_stage<1>(K, p, pd, h, h2, w2, rec, xoff);  col_stage<2>(K, p, pd, h, h2, w2, rec, xoff);
    col_stage<3>(K, p, pd, h, h2, w2, rec, xoff);  col_stage<4>(K, p, pd, h, h2, w2, rec, xoff);  col_stage<5>(K, p, pd, h, h2, w2, rec, xoff);
    col_stage<6>(K, p, pd, h, h2, w2, rec, xoff);  col_stage<7>(K, p, pd, h, h2, w2, rec, xoff);  col_stage<8>(K, p, pd, h, h2, w2, rec, xoff);
    col_stage<9>(K, p, pd, h, h2, w2, rec, xoff);  col_stage<10>(K, p, pd, h, h2, w2, rec, xoff); col_stage<11>(K, p, pd, h, h2, w2, rec, xoff);
    col_stage<12>(K, p, pd, h, h2, w2, rec, xoff);
    step_update(K, p, pd, h, h2, pn, pdn);
    double ep[3], epd[3];
    step_error(K, h, h2, ep, epd);
    return col_err_sumsq(half, w2, p, pd, pn, pdn, ep, epd, atol, rtol);
}

// ---------------------------------------------------------------------------
// State pass: one attempted step of the slot's state half (p, pd) = (r, v) [half 0] or (lv, lv') [half 1]; the partner half sits 8
// lanes away.  Stages rolled, stage derivatives in the warp's shared memory (ks, lane-strided).  Publishes the 13 stage records.
// Returns the half's sum of (error / scale)^2 in the reference's coordinates (col_err_sumsq: half 0 -> (r, v), half 1 -> (lr, lv)).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double state_attempt(const double (&p)[3], const double (&pd)[3], double h, const SCConst& c, double w2, const Law& lw,
                                                int half, bool pub, double* __restrict__ ks, double2* __restrict__ rec, double atol, double rtol,
                                                double (&pn)[3], double (&pdn)[3]) {
    const double h2 = h * h;
#pragma unroll 1
    for (int J = 0; J < 13; ++J) {
        double sb[3] = {0.0, 0.0, 0.0}, sg[3] = {0.0, 0.0, 0.0};
#pragma unroll 4
        for (int l = 0; l < J; ++l) {
            const double b = T_B[J][l], gq = T_G[J][l];
#pragma unroll
            for (int q = 0; q < 3; ++q) { const double kq = ks[(l * 3 + q) * 32]; sb[q] = fma(b, kq, sb[q]); sg[q] = fma(gq, kq, sg[q]); }
        }
        double P[3], Pd[3], Q[3], Qd[3];
        const double hc = h * T_C[J];
#pragma unroll
        for (int q = 0; q < 3; ++q) { Pd[q] = fma(h, sb[q], pd[q]); P[q] = fma(h2, sg[q], fma(hc, pd[q], p[q])); }
#pragma unroll
        for (int q = 0; q < 3; ++q) { Q[q] = __shfl_xor_sync(0xffffffffu, P[q], 8); Qd[q] = __shfl_xor_sync(0xffffffffu, Pd[q], 8); }
        double R[3], V[3], M[3], N[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) { R[q] = half ? Q[q] : P[q]; V[q] = half ? Qd[q] : Pd[q]; M[q] = half ? P[q] : Q[q]; N[q] = half ? Pd[q] : Qd[q]; }
        double kr[3], kl[3], U[6], W[6], G[6];
        sc_eval2<true>(R, V, M, N, c.mu, c.m1, w2, c.p, lw, kr, kl, U, W, G);
#pragma unroll
        for (int q = 0; q < 3; ++q) ks[(J * 3 + q) * 32] = half ? kl[q] : kr[q];
        if (pub) {                                                       // both halves hold the whole record: each writes half of it
            double2* w = rec + J * NC2 * SL;
            if (half == 0) {
                w[0 * SL] = make_double2(U[0], U[1]); w[1 * SL] = make_double2(U[2], U[3]); w[2 * SL] = make_double2(U[4], U[5]);
                w[3 * SL] = make_double2(W[0], W[1]); w[4 * SL] = make_double2(W[2], W[3]);
            } else {
                w[5 * SL] = make_double2(W[4], W[5]);
                w[6 * SL] = make_double2(G[0], G[1]); w[7 * SL] = make_double2(G[2], G[3]); w[8 * SL] = make_double2(G[4], G[5]);
            }
        }
    }
    // 8th-order update (ode.jl:937) and the embedded error estimate (ode.jl:940)
    double sv[3] = {0.0, 0.0, 0.0}, sr[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
    for (int l = 0; l < 13; ++l) {
        const double cv = T_CHI[l], cr = T_CHIB[l];
#pragma unroll
        for (int q = 0; q < 3; ++q) { const double kq = ks[(l * 3 + q) * 32]; sv[q] = fma(cv, kq, sv[q]); sr[q] = fma(cr, kq, sr[q]); }
    }
    double ep[3], epd[3];
    const double ce = h * lto_tab::ERRC, ce2 = h2 * lto_tab::ERRC;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        pdn[q] = fma(h, sv[q], pd[q]);
        pn[q] = fma(h2, sr[q], fma(h, pd[q], p[q]));
        const double k0 = ks[(0 * 3 + q) * 32], k10 = ks[(10 * 3 + q) * 32], k11 = ks[(11 * 3 + q) * 32], k12 = ks[(12 * 3 + q) * 32];
        ep[q] = ce2 * (k0 - k11);
        epd[q] = ce * ((k0 + k10) - (k11 + k12));
    }
    return col_err_sumsq(half, w2, p, pd, pn, pdn, ep, epd, atol, rtol);
}

__device__ __forceinline__ double rms12(const double (&e)[ND], const double (&y)[ND], double atol, double rtol) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ND; ++i) { const double q = e[i] * f_rcp(fma(rtol, fabs(y[i]), atol)); s = fma(q, q, s); }
    return sqrt(s * (1.0 / (double)ND));
}

__global__ void __launch_bounds__(NTHREADS, 1) k_indirect_wl(const __grid_constant__ IndirectArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned fullmask = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 3, half = g & 1, csel = g >> 1, s8 = lane & 7;
    unsigned char* const wbase = smem_raw + (size_t)warp * WARP_BYTES;
    double2* const rec = reinterpret_cast<double2*>(wbase) + s8;
    double* const ks = reinterpret_cast<double*>(wbase + REC_BYTES) + lane;
    double* const stg = reinterpret_cast<double*>(wbase) + (size_t)s8 * (ND * ND);          // STM staging of this lane's slot (record area)
    double* const scr = a.scratch + ((size_t)blockIdx.x * NWARP + warp) * SCR_DOUBLES_PER_WARP + lane;
    const double w2 = 2.0 * a.c.omega;
    const double atol = a.cfg.atol, rtol = a.cfg.rtol;
    const double inv_ne = 1.0 / (double)(ND * (ND + 1));
    const bool bulk = (reinterpret_cast<uintptr_t>(a.phi) & 15u) == 0;

    // per-slot control, replicated on the slot's four lanes
    double zp[3] = {0.0, 0.0, 0.0}, zpd[3] = {0.0, 0.0, 0.0};              // this lane's half of the state z = (r, v | lv, lv')
    double tcur = 0.0, tf = 0.0, h = 0.0, span = 1.0;
    Law lw; lw.aL = 0.0; lw.rho_inv = 1.0; lw.rq = 0.0;
    long long seg = -1, ia = 0;
    int na = 0, nt = 0, status = 0, par = 0;
    bool active = false, exhausted = false, lastrej = false, fresh = false;
    long long c_refill = 0, c_state = 0, c_cols = 0, c_dec = 0, n_att = 0, n_refill = 0;     // diagnostics (a.prof)
    const long long c_begin = clock64();

    while (true) {
        const long long k0 = clock64();
        // ---- refill: every idle slot pulls its next segment from the work queue (lane s8 is the slot's g = 0 lane)
        const bool want = !active && !exhausted;
        if (__any_sync(fullmask, want)) {
            ++n_refill;
            long long idx = -1;
            if (want && g == 0) idx = (long long)atomicAdd(a.counter, 1ull);
            idx = __shfl_sync(fullmask, idx, s8);
            const bool got = want && idx >= 0 && idx < a.n_seg;
            if (want && !got) exhausted = true;
            double x[ND];
            double t0 = 0.0, t1 = 0.0;
            Law nl; nl.aL = 0.0; nl.rho_inv = 1.0; nl.rq = 0.0;
            long long nia = 0;
            if (got) {
                nia = lto_node_a(idx, a.npt);
                const long long it = lto_traj_of(idx, a.npt);
#pragma unroll
                for (int i = 0; i < ND; ++i) x[i] = a.x0[nia * ND + i];
                t0 = a.t0[nia]; t1 = a.t1[nia];
                if (!(t0 < t1)) t1 = t0;                                  // empty span: one zero-length step, Phi = I
                const double tl = a.thrustLimit_arr ? a.thrustLimit_arr[it] : a.c.thrustLimit;
                const double rho = a.rho_arr ? a.rho_arr[it] : a.c.rho;
                nl.aL = tl * a.c.kthr / a.c.mass;                         // CRTBP_stateCostate_deriv.jl:33
                nl.rho_inv = 1.0 / rho;
                nl.rq = nl.aL / (4.0 * rho);
            } else {
#pragma unroll
                for (int i = 0; i < ND; ++i) x[i] = 0.0;
            }
            // Hairer-Norsett-Wanner initial step over the state components (drive_rk8 in lto_prop_generic.cuh), in the reference's
            // variables [r v lr lv]: f0 = f(x), f1 = f(x + h0 f0)
            const double nspan = t1 - t0;
            double f0[ND], y[ND], d1 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
            for (int i = 0; i < ND; ++i) { y[i] = x[i]; f0[i] = 0.0; }
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                double r[3], v[3], lv[3], lvd[3], kr[3], kl[3], U[6], W[6], G[6], cn[3], f[ND];
                to_z(w2, y, r, v, lv, lvd);
                sc_eval2<false>(r, v, lv, lvd, a.c.mu, a.c.m1, w2, a.c.p, nl, kr, kl, U, W, G);
                coriolis(w2, lvd, cn);
#pragma unroll
                for (int q = 0; q < 3; ++q) { f[q] = v[q]; f[3 + q] = kr[q]; f[6 + q] = -(kl[q] - cn[q]); f[9 + q] = lvd[q]; }   // lr' = -U lv
                if (pass == 0) {
                    const double d0 = rms12(x, x, atol, rtol);
                    d1 = rms12(f, x, atol, rtol);
                    h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
                    h0 = fmin(h0, nspan);
#pragma unroll
                    for (int i = 0; i < ND; ++i) { f0[i] = f[i]; y[i] = fma(h0, f[i], x[i]); }
                } else {
#pragma unroll
                    for (int i = 0; i < ND; ++i) f[i] -= f0[i];
                    const double d2 = rms12(f, x, atol, rtol) / h0;
                    const double dm = fmax(d1, d2);
                    h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : inv_eighth_root(dm / 0.01);
                }
            }
            if (got) {
                seg = idx; ia = nia;
                double r[3], v[3], lv[3], lvd[3];
                to_z(w2, x, r, v, lv, lvd);
#pragma unroll
                for (int q = 0; q < 3; ++q) { zp[q] = half ? lv[q] : r[q]; zpd[q] = half ? lvd[q] : v[q]; }
                tcur = t0; tf = t1; span = nspan; lw = nl;
                h = fmin(fmin(100.0 * h0, h1), nspan);
                na = 0; nt = 0; status = 0; lastrej = false;
                active = true; fresh = true;
            }
        }
        if (!__any_sync(fullmask, active)) break;
        const long long k1 = clock64();
        c_refill += k1 - k0;

        // ---- one attempted step of the warp's 8 segments
        bool last = false;
        if (active) {
            if (tcur + h >= tf) { h = tf - tcur; last = true; }
            ++nt;
        }
        const double hh = active ? h : 0.0;
        bulk_store_wait_read();                                          // the STM staging of the previous round has been read out of the record area
        __syncwarp();
        double zn[3], znd[3];
        double es = state_attempt(zp, zpd, hh, a.c, w2, lw, half, csel == 0, ks, rec, atol, rtol, zn, znd);
        if (csel != 0) es = 0.0;                                         // the copy lanes carry the same state: counted once
        __syncwarp();                                                    // the records are complete
        const long long k2 = clock64();
        c_state += k2 - k1;
#pragma unroll 1
        for (int c = 0; c < NPASS; ++c) {
            const int col = 2 * c + csel;
            double* const cur = scr + (size_t)((par * NPASS + c) * 6) * 32;
            double* const cnd = scr + (size_t)(((par ^ 1) * NPASS + c) * 6) * 32;
            double p[3], pd[3];
            if (fresh) {
                col_init(col, half, w2, p, pd);
#pragma unroll
                for (int q = 0; q < 3; ++q) { __stcg(cur + q * 32, p[q]); __stcg(cur + (3 + q) * 32, pd[q]); }
            } else {
#pragma unroll
                for (int q = 0; q < 3; ++q) { p[q] = __ldcg(cur + q * 32); pd[q] = __ldcg(cur + (3 + q) * 32); }
            }
            double pn[3], pdn[3];
            es += col_attempt(p, pd, hh, w2, rec, half, atol, rtol, pn, pdn);
#pragma unroll
            for (int q = 0; q < 3; ++q) { __stcg(cnd + q * 32, pn[q]); __stcg(cnd + (3 + q) * 32, pdn[q]); }
        }
        fresh = false;
        const long long k3 = clock64();
        c_cols += k3 - k2; ++n_att;

        // ---- accept / reject, per slot (the slot's four lanes take the same decision from the same sum)
        es += __shfl_xor_sync(fullmask, es, 8);
        es += __shfl_xor_sync(fullmask, es, 16);
        bool finished = false;
        if (active) {
            const double u = es * inv_ne;                               // eest^2: eest <= 1 <=> u <= 1, eest^(-1/8) = u^(-1/16)
            if (!(u == u)) { status = LTO_ST_NAN; finished = true; }
            else {
                double q = (u == 0.0) ? 5.0 : ((u < 1e300) ? 0.9 * inv_sixteenth_root(u) : 0.2);
                q = fmin(5.0, fmax(0.2, q));
                if (u <= 1.0) {
                    ++na; par ^= 1;                                     // the candidates become z and the current columns
#pragma unroll
                    for (int q3 = 0; q3 < 3; ++q3) { zp[q3] = zn[q3]; zpd[q3] = znd[q3]; }
                    if (last) { tcur = tf; finished = true; }
                    else { tcur += h; if (lastrej) q = fmin(q, 1.0); lastrej = false; }
                } else {
                    lastrej = true; q = fmin(q, 1.0);
                }
                h *= q;
            }
            if (!finished) {                                            // drive_rk8's loop-top checks
                if (h < span * 1e-12) { status = LTO_ST_HMIN; finished = true; }
                else if (nt >= a.cfg.max_attempts) { status = LTO_ST_MAXSTEPS; finished = true; }
            }
        }
        if (__any_sync(fullmask, finished)) {
            // ---- defect = x(t1) - XC_all[:, i+1] (multiShoot_CRTBP_indirect.jl:82), back in the reference's variables:
            // half 0 holds (r, v) = rows 0..5, half 1 (lv, lv') -> (lr, lv) = rows 6..11
            double o[6];
            col_out(half, w2, zp, zpd, o);
            int nan = 0;
#pragma unroll
            for (int i = 0; i < 6; ++i) nan |= !(o[i] == o[i]);
            nan |= __shfl_xor_sync(fullmask, nan, 8);
            if (finished) {
                if (nan && status == 0) status = LTO_ST_NAN;
                if (csel == 0) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const int row = 6 * half + i;
                        a.defect[seg * ND + row] = a.x_target ? o[i] - a.x_target[ia * ND + row] : o[i];
                    }
                    if (g == 0) {
                        if (a.status) a.status[seg] = status;
                        if (a.nsteps_out) { a.nsteps_out[2 * seg] = na; a.nsteps_out[2 * seg + 1] = nt; }
                    }
                }
            }
            // ---- Phi: the slot's current half-columns (accepted: the candidates just written; ended in error: the last accepted ones)
            __syncwarp();                                                // nobody reads the records any more
            if (finished) {
#pragma unroll 1
                for (int c = 0; c < NPASS; ++c) {
                    const double* cur = scr + (size_t)((par * NPASS + c) * 6) * 32;
                    double p[3], pd[3], oc[6];
#pragma unroll
                    for (int q = 0; q < 3; ++q) { p[q] = __ldcg(cur + q * 32); pd[q] = __ldcg(cur + (3 + q) * 32); }
                    col_out(half, w2, p, pd, oc);                        // rows 6 half .. 6 half + 5 of column `col` of ForwardDiff.jacobian(f, x0) (:121)
                    const int col = 2 * c + csel;
                    if (bulk) {
                        double2* d = reinterpret_cast<double2*>(stg + col * ND + 6 * half);
                        d[0] = make_double2(oc[0], oc[1]); d[1] = make_double2(oc[2], oc[3]); d[2] = make_double2(oc[4], oc[5]);
                    } else {
                        double* out = a.phi + seg * (long long)(ND * ND) + col * ND + 6 * half;
#pragma unroll
                        for (int i = 0; i < 6; ++i) out[i] = oc[i];
                    }
                }
                fence_proxy_async();
            }
            __syncwarp();
            if (finished && bulk && g == 0) bulk_store(a.phi + seg * (long long)(ND * ND), smem_u32(stg), ND * ND * sizeof(double));
            if (finished) active = false;
        }
        c_dec += clock64() - k3;
    }
    if (a.prof && lane == 0) {                                           // [CTA][warp][8]: refill, state pass, column passes, decision + output, attempts, refills, alive
        unsigned long long* o = a.prof + ((size_t)blockIdx.x * NWARP + warp) * 8;
        o[0] = c_refill; o[1] = c_state; o[2] = c_cols; o[3] = c_dec; o[4] = n_att; o[5] = n_refill; o[6] = clock64() - c_begin; o[7] = 0;
    }
    bulk_store_wait_all();                                               // the last bulk stores must have completed before the CTA retires
}

}  // namespace iwl

size_t indirect_wl_scratch_bytes(int n_sm) { return (size_t)n_sm * iwl::NWARP * iwl::SCR_DOUBLES_PER_WARP * sizeof(double); }

cudaError_t launch_indirect_wl(const IndirectArgs& a, cudaStream_t st, int* n_launch) {
    *n_launch = 0;
    if (a.phi == nullptr || a.counter == nullptr || a.scratch == nullptr || a.cfg.controller != 0 || a.cfg.err_norm == 0 || a.n_seg <= 0 ||
        a.n_seg > 0x7fffffffll)
        return cudaErrorNotSupported;
    // per device: a single process may drive several GPUs (lto_init_devices)
    static int n_sm_dev[64] = {0};
    static bool attr_dev[64] = {false};
    int dev = 0;
    { cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e; }
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!attr_dev[dev]) {
        cudaError_t e = cudaDeviceGetAttribute(&n_sm_dev[dev], cudaDevAttrMultiProcessorCount, dev); if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(iwl::k_indirect_wl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)iwl::SMEM);
        if (e != cudaSuccess) return e;
        attr_dev[dev] = true;
    }
    const int n_sm = n_sm_dev[dev];
    cudaError_t e = cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    const long long per_cta = (long long)iwl::NWARP * iwl::SL;
    const int grid = (int)std::min<long long>((a.n_seg + per_cta - 1) / per_cta, (long long)n_sm);
    iwl::k_indirect_wl<<<grid, iwl::NTHREADS, iwl::SMEM, st>>>(a);
    e = cudaGetLastError();
    if (e == cudaSuccess) *n_launch = 1;
    return e;
}

}  // namespace lto
