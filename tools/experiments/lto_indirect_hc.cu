// lto_indirect_hc.cu -- throughput kernel of the indirect method (K3), ndim = 12, "half-column" layout:
// defectCalc + jacobianCalc of multiShoot_CRTBP_indirect.jl:63-124 for a whole batch in one launch.  Each segment integrates
// [x | Phi] (12 + 144 components) with the adaptive order-8 pair and the OrdinaryDiffEq-style controller of lto_prop_generic.cuh
// (drive_rk8); Phi replaces ForwardDiff.jacobian(f, x0) (:121).
//
// Why this layout (round 2).  The round-1 kernel gave a thread a whole STM column: 117 stage derivatives ->
// 255 registers -> 8 warps per SM, two per sub-partition, the two state warps alone on the fourth sub-partition (a quarter of the
// SM's FP64 pipe idle) and the state warp's chain exposed (FP64 pipe 48 % active).  Here the system is integrated in the
// second-order variables (r, v, lv, lv') (lto_hc_math.cuh): every STM column splits into two 3-vector second-order halves of the
// SAME shape, each advanced in Nystrom form with 39 stored doubles.  One thread = one half-column:
//   * ~150 registers -> 12 warps per SM, three per sub-partition; all four FP64 pipes carry column work;
//   * setmaxnreg moves registers from the 8 column warps (128) to the state warps (248), which keep their 78 stage derivatives
//     in registers;
//   * 3 tiles of 32 segment slots in flight (stage records 3 x 59.9 KB of shared memory), one state warp per tile, so a tile's
//     next attempt has two column visits of the other tiles to be ready: the state warp's dependent chain is off the critical
//     path.
// Mapping:
//   tile        = 32 segment slots owned by one state warp (lane = slot): nonlinear 12-dim system, step-size controller, work
//                 queue (a slot that finishes its segment pulls the next one from a global counter).  Publishes per attempted
//                 step the 13 stage linearisations U, W, G (18 doubles per stage and slot) + h + flags.
//   column warp = 8 per CTA.  A tile visit is 24 warp-tasks = 6 column pairs x 4 slot octets, three per column warp.  In a task
//                 lane = (g, s): g = 2 * (column of the pair) + half, s = slot of the octet; the two halves of a column sit 8 lanes
//                 apart and swap their stage positions with one shfl.xor per component.  The four groups read the same U record
//                 (one shared-memory wavefront per 128-bit load), G (dr-halves) and W (dlv-halves) two.
//   columns between visits live in an L2-resident scratch, two buffers per (tile, task, lane); the state warp publishes which one
//   holds the current (last accepted) column, so an accepted step is a flip and a rejected one re-reads the same buffer.
// Step control: joint norm over x and Phi (LTO_NORM_STATE_SENS, the ForwardDiff semantics) or x alone (robust estimate + sharp-law
// safeguard of the state-only controller, lto_prop_generic.cuh).  Errors are scaled in the reference's variables (r, v, lr, lv).
#include "lto_internal.h"
#include "lto_cw_common.cuh"
#include "lto_hc_math.cuh"
#include <algorithm>
#include <cstdlib>

namespace lto {
namespace ihc {

using namespace cwc;
using namespace hcm;

constexpr int ND = 12;
constexpr int NTILE = 3;          // tiles in flight = state warps at work
constexpr int TS = 32;            // segment slots per tile
#ifdef LTO_IHC_ISO                // experiment: state warps alone on SM sub-partition 3 (warps 3, 7, 11), nine column warps on the other three,
constexpr int NCW = 9;            // no register reallocation; a tile visit's 24 tasks are dealt round-robin over the 9 column warps ACROSS visits
#else
constexpr int NCW = 8;            // column warps (warp groups 0 and 1)
#endif
constexpr int NCT = 32 * NCW;     // column threads
constexpr int NW = 12;            // + warp group 2: 3 state warps and one idle warp
constexpr int NTHREADS = 32 * NW;
constexpr int NTASK = 24;         // warp-tasks per tile visit: 6 column pairs x 4 slot octets
constexpr int NPH = NTASK / NCW;  // tasks per column warp and tile visit (default layout)
constexpr int NC2 = 9;            // double2 per stage record: U[6] W[6] G[6]
constexpr int REG_COL = 128, REG_STATE = 248;   // setmaxnreg: launched at 168; the 8 column warps release 8 x 40 registers per lane, the 4 warps of the state group claim 4 x 80 -- only what was released inside the CTA can be claimed (a larger claim spins forever)

enum { F_STORE = 2, F_RESET = 4, F_ACTIVE = 8, F_PAR = 16 };

// ---- shared-memory plan ------------------------------------------------------
constexpr size_t REC_BYTES = (size_t)13 * NC2 * TS * sizeof(double2);      // stage records of one tile
constexpr size_t HDR_BYTES = (size_t)TS * (sizeof(double) + sizeof(int2)); // h, {flags, segment}
constexpr size_t ERR_BYTES = (size_t)ND * TS * sizeof(double);             // error partials per column and slot (both halves summed)
constexpr int NXW = ND + 7;                                                // next-segment stash: x0, t0, tf, aL, 1/rho, aL/(4 rho), h0, tol scale
constexpr size_t NXT_BYTES = (size_t)NXW * TS * sizeof(double);
constexpr size_t ZN_BYTES = (size_t)2 * ND * TS * sizeof(double);           // the state z and the candidate of the attempt in flight (double buffer; not live in registers across the right-hand-side calls)
constexpr size_t TILE_BYTES = REC_BYTES + HDR_BYTES + ERR_BYTES + ZN_BYTES + NXT_BYTES;
constexpr size_t BAR_BYTES = 32;                                           // full, done, tile_done
constexpr size_t SMEM = NTILE * TILE_BYTES + NTILE * BAR_BYTES;
static_assert(SMEM <= 232448, "three tiles must fit the 227 KB of shared memory per CTA");
static_assert(TILE_BYTES % 16 == 0, "tiles must stay 16-byte aligned");
constexpr size_t SCR_DOUBLES_PER_CTA = (size_t)NTILE * 2 * NTASK * 6 * 32; // [tile][parity][task][component][lane]

struct TileSmem {
    double2* rec; double* hval; int2* hctl; double* errp; double* zn; double* nx;
    unsigned bar_full, bar_done; volatile int* tile_done;
};

__device__ __forceinline__ TileSmem tile_smem(unsigned char* base, int t) {
    unsigned char* p = base + (size_t)t * TILE_BYTES;
    TileSmem s;
    s.rec = reinterpret_cast<double2*>(p); p += REC_BYTES;
    s.hval = reinterpret_cast<double*>(p); p += TS * sizeof(double);
    s.hctl = reinterpret_cast<int2*>(p); p += TS * sizeof(int2);
    s.errp = reinterpret_cast<double*>(p); p += ERR_BYTES;
    s.zn = reinterpret_cast<double*>(p); p += ZN_BYTES;
    s.nx = reinterpret_cast<double*>(p);
    unsigned char* b = base + (size_t)NTILE * TILE_BYTES + (size_t)t * BAR_BYTES;
    s.bar_full = smem_u32(b); s.bar_done = smem_u32(b + 8);
    s.tile_done = reinterpret_cast<volatile int*>(b + 16);
    return s;
}

template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ---------------------------------------------------------------------------
// Column thread: one attempted RK step of one half-column (p, pd) (lto_hc_math.cuh):
//   k_J = U P_J + X Po_J + C Pd_J,   X = G, Po = dlv-half's position (half 0)  |  X = W, Po = dr-half's position (half 1)
// Stage 11 (index 10) enters only the error estimate: skipped when the columns are not part of the step-control norm.
// ---------------------------------------------------------------------------
template <int J>
__device__ __forceinline__ void col_stage(K3& K, const double (&p)[3], const double (&pd)[3], double h, double h2, double w2,
                                          const double2* __restrict__ rec, int xoff) {
    double P[3], Pd[3], Po[3];
    stage_in<J>(K, p, pd, h, h2, P, Pd);
#pragma unroll
    for (int q = 0; q < 3; ++q) Po[q] = __shfl_xor_sync(0xffffffffu, P[q], 8);
    const double2* w = rec + J * NC2 * TS;
    double U[6], X[6];
    { const double2 a = w[0 * TS], b = w[1 * TS], c = w[2 * TS]; U[0] = a.x; U[1] = a.y; U[2] = b.x; U[3] = b.y; U[4] = c.x; U[5] = c.y; }
    { const double2 a = w[xoff], b = w[xoff + TS], c = w[xoff + 2 * TS]; X[0] = a.x; X[1] = a.y; X[2] = b.x; X[3] = b.y; X[4] = c.x; X[5] = c.y; }
    double k[3];
    col_rhs(U, X, w2, P, Pd, Po, k);
#pragma unroll
    for (int q = 0; q < 3; ++q) K.k[J][q] = k[q];
}

// (Measured and switched off: re-converging the 8 column warps a few times per attempt with a named barrier, which kept the round-1
// kernel's instruction-fetch windows together, costs 18 % here -- the instruction-cache misses came from the state warps' code size.)
#if defined(LTO_IHC_LOCKSTEP_PAIR)        // the two column warps of one SM sub-partition (cw, cw + 4) only: they then share their instruction fetches
#define LTO_IHC_LOCKSTEP() asm volatile("bar.sync %0, 64;" ::"r"(lsid) : "memory")
#elif defined(LTO_IHC_LOCKSTEP_ON)
#define LTO_IHC_LOCKSTEP() asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory")
#else
#define LTO_IHC_LOCKSTEP()
#endif

template <bool ERR>
__device__ __forceinline__ double col_attempt(const double (&p)[3], const double (&pd)[3], double h, double w2, const double2* __restrict__ rec,
                                              int half, double atol, double rtol, double (&pn)[3], double (&pdn)[3], int lsid = 1) {
    const double h2 = h * h;
    (void)lsid;
    const int xoff = half ? 3 * TS : 6 * TS;                            // W for the dlv-half, G for the dr-half
    K3 K;
    LTO_IHC_LOCKSTEP();
    col_stage<0>(K, p, pd, h, h2, w2, rec, xoff);  col_stage<1>(K, p, pd, h, h2, w2, rec, xoff);  col_stage<2>(K, p, pd, h, h2, w2, rec, xoff);
    col_stage<3>(K, p, pd, h, h2, w2, rec, xoff);  col_stage<4>(K, p, pd, h, h2, w2, rec, xoff);  col_stage<5>(K, p, pd, h, h2, w2, rec, xoff);
    LTO_IHC_LOCKSTEP();
    col_stage<6>(K, p, pd, h, h2, w2, rec, xoff);  col_stage<7>(K, p, pd, h, h2, w2, rec, xoff);  col_stage<8>(K, p, pd, h, h2, w2, rec, xoff);
    LTO_IHC_LOCKSTEP();
    col_stage<9>(K, p, pd, h, h2, w2, rec, xoff);
    if (ERR) col_stage<10>(K, p, pd, h, h2, w2, rec, xoff);
    LTO_IHC_LOCKSTEP();
    col_stage<11>(K, p, pd, h, h2, w2, rec, xoff); col_stage<12>(K, p, pd, h, h2, w2, rec, xoff);
    step_update(K, p, pd, h, h2, pn, pdn);
    if (!ERR) return 0.0;
    double ep[3], epd[3];
    step_error(K, h, h2, ep, epd);
    return col_err_sumsq(half, w2, p, pd, pn, pdn, ep, epd, atol, rtol);
}

template <bool JOINT>
__device__ __forceinline__ void column_warp(const IndirectArgs& a, int cw, int lane, unsigned char* smem) {
    const int g = lane >> 3, half = g & 1, csel = g >> 1, s8 = lane & 7;
    const double w2 = 2.0 * a.c.omega;
    const double atol = a.cfg.atol, rtol = a.cfg.rtol;
    double* const scr = a.scratch + (size_t)blockIdx.x * SCR_DOUBLES_PER_CTA + lane;
    const bool wide = (reinterpret_cast<uintptr_t>(a.phi) & 15u) == 0;
    unsigned alive = (1u << NTILE) - 1u;
    unsigned visit = 0;
    long long c_wait = 0, c_work = 0, n_work = 0, c_att = 0;
    const long long c_begin = clock64();
    while (alive) {
#pragma unroll 1
        for (int t = 0; t < NTILE; ++t) {
            if (!(alive & (1u << t))) continue;
            const TileSmem S = tile_smem(smem, t);
            const long long c0 = clock64();
            mbar_wait_parked(S.bar_full, visit & 1);
            const long long c1 = clock64();
            c_wait += c1 - c0;
            const bool done = *S.tile_done != 0;
#ifdef LTO_IHC_ISO
            const int gv = (int)(visit * NTILE + t);
            const int k0 = ((cw - (NTASK * gv) % NCW) % NCW + NCW) % NCW;
            int n_my = 0;
#pragma unroll 1
            for (int task = k0; task < NTASK; task += NCW) {
                ++n_my;
#else
#pragma unroll 1
            for (int ph = 0; ph < NPH; ++ph) {
                const int task = ph * NCW + cw;
#endif
#ifdef LTO_IHC_PACE_TASK                  // one barrier per task: the column warps start every task's instruction stream together
                asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
#endif
                const int col = 2 * (task >> 2) + csel;
                const int slot = (task & 3) * 8 + s8;
                const int2 hc = S.hctl[slot];
                const double h = S.hval[slot];
                const int par = (hc.x & F_PAR) ? 1 : 0;
                double* const cur = scr + (size_t)((t * 2 + par) * NTASK + task) * (6 * 32);
                double* const cnd = scr + (size_t)((t * 2 + (par ^ 1)) * NTASK + task) * (6 * 32);
                double p[3], pd[3];
                if ((hc.x & F_STORE) || ((hc.x & F_ACTIVE) && !(hc.x & F_RESET))) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) { p[q] = __ldcg(cur + q * 32); pd[q] = __ldcg(cur + (3 + q) * 32); }
                } else {
#pragma unroll
                    for (int q = 0; q < 3; ++q) { p[q] = 0.0; pd[q] = 0.0; }
                }
                if (hc.x & F_STORE) {                                  // rows 6 half .. 6 half + 5 of column `col` of ForwardDiff.jacobian(f, x0) (:121)
                    double o[6];
                    col_out(half, w2, p, pd, o);
                    double* out = a.phi + (long long)hc.y * (ND * ND) + col * ND + 6 * half;
                    if (wide) {
#pragma unroll
                        for (int i = 0; i < 6; i += 2)
                            asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(out + i), "d"(o[i]), "d"(o[i + 1]) : "memory");
                    } else {
#pragma unroll
                        for (int i = 0; i < 6; ++i) out[i] = o[i];
                    }
                }
                if (hc.x & F_RESET) {
                    col_init(col, half, w2, p, pd);
#pragma unroll
                    for (int q = 0; q < 3; ++q) { __stcg(cur + q * 32, p[q]); __stcg(cur + (3 + q) * 32, pd[q]); }
                }
                if (done) continue;
                double pn[3], pdn[3];
                const long long ca = a.prof ? clock64() : 0;
                double es = col_attempt<JOINT>(p, pd, h, w2, S.rec + slot, half, atol, rtol, pn, pdn, 1 + (cw & 3));
                if (a.prof) c_att += clock64() - ca;
#pragma unroll
                for (int q = 0; q < 3; ++q) { __stcg(cnd + q * 32, pn[q]); __stcg(cnd + (3 + q) * 32, pdn[q]); }
                if (JOINT) {
                    es += __shfl_xor_sync(0xffffffffu, es, 8);
                    if (half == 0) S.errp[col * TS + slot] = es;
                }
            }
#ifdef LTO_IHC_ISO
            if (done) alive &= ~(1u << t);
            else { for (int i = 0; i < n_my; ++i) mbar_arrive(S.bar_done); c_work += clock64() - c1; n_work += n_my; }
#else
            if (done) alive &= ~(1u << t);
            else { mbar_arrive(S.bar_done); c_work += clock64() - c1; n_work += NPH; }
#endif
        }
        ++visit;
    }
    if (a.prof && lane == 0) {
        unsigned long long* o = a.prof + ((size_t)blockIdx.x * NW + cw) * 4;
        o[0] = c_work; o[1] = c_wait; o[2] = n_work; o[3] = clock64() - c_begin;
        a.prof[(size_t)gridDim.x * NW * 4 + (size_t)gridDim.x * NTILE + (size_t)blockIdx.x * NCW + cw] = c_att;   // cycles inside col_attempt
    }
}

// ---------------------------------------------------------------------------
// State warp.  The 78 stage derivatives of z = (r, v, lv, lvd) stay in registers (13 unrolled stages, compile-time tableau
// sparsity); ONE out-of-line copy of the right-hand side serves all 13 stages, and the claim of the next segment is inlined at
// its one call site -- the state warps' code competes with the column warps' ~19 KB loop body for the instruction caches
// (sm__icc_request_hit_rate, stall_no_instruction in profiles/), so every duplicate counts.  Measured alternative (kept in
// tools/experiments/lto_indirect_hc_tmem.cu): stages rolled into a loop with the derivatives in the warp's lane quadrant of TENSOR
// MEMORY (tcgen05.st / tcgen05.ld) -- a third of the code, correct results, but the dependent load -> wait -> FMA rounds of the
// stage-input loop cost 24 k cycles per attempt (48 k in all against 28 k here), so the state warps became the bound.
// ---------------------------------------------------------------------------
struct Out6 { double v[6]; };
__device__ __noinline__ Out6 sc_eval2_call(double r0, double r1, double r2, double v0, double v1, double m0, double m1, double m2, double n0, double n1,
                                           double mu, double mu1, double w2, double pexp, double aL, double rho_inv, double rq, double2* w) {
    const double R[3] = {r0, r1, r2}, V[3] = {v0, v1, 0.0}, M[3] = {m0, m1, m2}, N[3] = {n0, n1, 0.0};
    Law lw; lw.aL = aL; lw.rho_inv = rho_inv; lw.rq = rq;
    double kr[3], kl[3], U[6], W[6], G[6];
    sc_eval2<true>(R, V, M, N, mu, mu1, w2, pexp, lw, kr, kl, U, W, G);
    w[0 * TS] = make_double2(U[0], U[1]); w[1 * TS] = make_double2(U[2], U[3]); w[2 * TS] = make_double2(U[4], U[5]);
    w[3 * TS] = make_double2(W[0], W[1]); w[4 * TS] = make_double2(W[2], W[3]); w[5 * TS] = make_double2(W[4], W[5]);
    w[6 * TS] = make_double2(G[0], G[1]); w[7 * TS] = make_double2(G[2], G[3]); w[8 * TS] = make_double2(G[4], G[5]);
    Out6 o;
#pragma unroll
    for (int q = 0; q < 3; ++q) { o.v[q] = kr[q]; o.v[3 + q] = kl[q]; }
    return o;
}

// one stage of the state: inputs from the stage derivatives in REGISTERS (compile-time tableau sparsity), right-hand side + record
template <int J>
__device__ __forceinline__ void state_stage(K3& Kr, K3& Kl, const double* __restrict__ z, double h, double h2, const SCConst& c, double w2,
                                            const Law& lw, double2* __restrict__ rec) {
    const double r[3] = {z[0 * TS], z[1 * TS], z[2 * TS]}, v[3] = {z[3 * TS], z[4 * TS], z[5 * TS]}, lv[3] = {z[6 * TS], z[7 * TS], z[8 * TS]},
                 lvd[3] = {z[9 * TS], z[10 * TS], z[11 * TS]};
    double R[3], V[3], M[3], N[3];
    stage_in<J>(Kr, r, v, h, h2, R, V);
    stage_in<J>(Kl, lv, lvd, h, h2, M, N);
    const Out6 o = sc_eval2_call(R[0], R[1], R[2], V[0], V[1], M[0], M[1], M[2], N[0], N[1], c.mu, c.m1, w2, c.p, lw.aL, lw.rho_inv, lw.rq,
                                 rec + J * NC2 * TS);
#pragma unroll
    for (int q = 0; q < 3; ++q) { Kr.k[J][q] = o.v[q]; Kl.k[J][q] = o.v[3 + q]; }
}

__device__ __forceinline__ double rms12(const double (&e)[ND], const double (&y)[ND], double atol, double rtol) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ND; ++i) { const double q = e[i] * f_rcp(fma(rtol, fabs(y[i]), atol)); s = fma(q, q, s); }
    return sqrt(s * (1.0 / (double)ND));
}

// Off the critical path: while the column warps work on the attempt just published, every slot that has no successor yet claims
// its NEXT segment from the work queue, loads it, and runs the Hairer-Norsett-Wanner initial-step estimate over the state
// components (drive_rk8 in lto_prop_generic.cuh).  The result waits in the tile's shared-memory stash until the slot's current
// segment finishes.  A slot claims just in time (`soon`: it is idle, or the attempt just published reaches t1), so no segment is
// hoarded while other slots run dry.  Inlined at its ONE call site; the two right-hand sides of the estimate share one inlined copy
// (a two-pass loop): no out-of-line call in the state warp, so nothing it keeps in registers is saved and restored around one.
struct Claim { long long nseg; int exhausted; };
template <bool JOINT>
__device__ __forceinline__ Claim prepare_next(const IndirectArgs& a, double* nxs, Claim c, bool soon) {
    const unsigned fullmask = 0xffffffffu;
    const bool want = soon && !c.exhausted && c.nseg < 0;
    if (!__any_sync(fullmask, want)) return c;
    const double w2 = 2.0 * a.c.omega;
    double x[ND];
    double t0 = 0.0, tf = 0.0, ts = 1.0;
    Law lw; lw.aL = 0.0; lw.rho_inv = 1.0; lw.rq = 0.0;
    bool got = false;
    if (want) {
        const long long idx = (long long)atomicAdd(a.counter, 1ull);
        if (idx < a.n_seg) {
            got = true; c.nseg = idx;
            const long long ia = lto_node_a(idx, a.npt), it = lto_traj_of(idx, a.npt);
#pragma unroll
            for (int i = 0; i < ND; ++i) x[i] = a.x0[ia * ND + i];
            t0 = a.t0[ia]; tf = a.t1[ia];
            if (!(t0 < tf)) tf = t0;                                      // empty span: one zero-length step, Phi = I
            const double tl = a.thrustLimit_arr ? a.thrustLimit_arr[it] : a.c.thrustLimit;
            const double rho = a.rho_arr ? a.rho_arr[it] : a.c.rho;
            lw.aL = tl * a.c.kthr / a.c.mass;                             // CRTBP_stateCostate_deriv.jl:33
            lw.rho_inv = 1.0 / rho;
            lw.rq = lw.aL / (4.0 * rho);
            if (!JOINT) ts = state_tol_scale(a.c.p, rho);
        } else {
            c.exhausted = 1;
        }
    }
    if (!got) {
#pragma unroll
        for (int i = 0; i < ND; ++i) x[i] = 0.0;
    }
    const double atol = a.cfg.atol * ts, rtol = a.cfg.rtol * ts;
    const double span = tf - t0;
    // f0 = f(x), then f1 = f(x + h0 f0), in the reference's variables [r v lr lv]
    double f0[ND], y[ND], d1 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
    for (int i = 0; i < ND; ++i) { y[i] = x[i]; f0[i] = 0.0; }
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        double r[3], v[3], lv[3], lvd[3], kr[3], kl[3], U[6], W[6], G[6], cn[3], f[ND];
        to_z(w2, y, r, v, lv, lvd);
        sc_eval2<false>(r, v, lv, lvd, a.c.mu, a.c.m1, w2, a.c.p, lw, kr, kl, U, W, G);
        coriolis(w2, lvd, cn);
#pragma unroll
        for (int q = 0; q < 3; ++q) { f[q] = v[q]; f[3 + q] = kr[q]; f[6 + q] = -(kl[q] - cn[q]); f[9 + q] = lvd[q]; }   // lr' = -U lv = -(lv'' - C lv')
        if (pass == 0) {
            const double d0 = rms12(x, x, atol, rtol);
            d1 = rms12(f, x, atol, rtol);
            h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
            h0 = fmin(h0, span);
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
#pragma unroll
        for (int i = 0; i < ND; ++i) nxs[i * TS] = x[i];
        nxs[(ND + 0) * TS] = t0; nxs[(ND + 1) * TS] = tf; nxs[(ND + 2) * TS] = lw.aL; nxs[(ND + 3) * TS] = lw.rho_inv; nxs[(ND + 4) * TS] = lw.rq;
        nxs[(ND + 5) * TS] = fmin(fmin(100.0 * h0, h1), span);
        nxs[(ND + 6) * TS] = ts;
    }
    return c;
}

template <bool JOINT>
__device__ __forceinline__ void state_warp(const IndirectArgs& a, int t, int lane, unsigned char* smem) {
    const TileSmem S = tile_smem(smem, t);
    const int slot = lane;
    const unsigned fullmask = 0xffffffffu;
    const double w2 = 2.0 * a.c.omega;
    double atol = a.cfg.atol, rtol = a.cfg.rtol;                          // per slot when the norm is the state's alone (state_tol_scale)
    const double inv_ne = JOINT ? 1.0 / (double)(ND * (ND + 1)) : 1.0 / (double)ND;
    int par = 0;                                                          // which scratch buffer holds the slot's current columns
    int zi = 0;                                                           // which half of the double buffer holds z = (r, v, lv, lvd)
    double* const zbuf = S.zn + slot;
#pragma unroll
    for (int i = 0; i < ND; ++i) { zbuf[i * TS] = 0.0; zbuf[(ND + i) * TS] = 0.0; }
    double tcur = 0.0, tf = 0.0, h = 0.0, span = 1.0, esum = 0.0;
    Law lw; lw.aL = 0.0; lw.rho_inv = 1.0; lw.rq = 0.0;
    long long seg = -1, ia = 0;
    int na = 0, nt = 0, status = 0;
    bool active = false, lastrej = false, last = false, have = false;
    Claim nxt; nxt.nseg = -1; nxt.exhausted = 0;                          // successor segment claimed and staged by prepare_next()
    unsigned visit = 0;
    double2* rec = S.rec + slot;
    double* const nxs = S.nx + slot;
    long long c_wait = 0, c_work = 0, c_pre = 0;
    const long long c_begin = clock64();
    bool soon = true, retried = false;
    int flags = 0, store_seg = 0;                                         // published with the next attempt
    while (true) {
        nxt = prepare_next<JOINT>(a, nxs, nxt, soon);                     // (overlaps the column warps' work on the attempt just published)
        bool finished = false;
        const long long c0 = clock64();
        long long c1 = c0;
        if (have) {
            have = false;
            mbar_wait_parked(S.bar_done, (visit - 1) & 1);
            c1 = clock64();
            c_wait += c1 - c0;
            if (active) {
                double s2 = esum;
                if (JOINT) {
#pragma unroll
                    for (int c = 0; c < ND; ++c) s2 += S.errp[c * TS + slot];
                }
                const double u = s2 * inv_ne;                           // eest^2: eest <= 1 <=> u <= 1, eest^(-1/8) = u^(-1/16)
                if (!(u == u)) { status = LTO_ST_NAN; finished = true; }
                else {
                    double q = (u == 0.0) ? 5.0 : ((u < 1e300) ? 0.9 * inv_sixteenth_root(u) : 0.2);
                    q = fmin(5.0, fmax(0.2, q));
                    if (u <= 1.0) {
                        ++na;
                        par ^= 1; zi ^= 1;                                // the candidates become z and the current columns
                        if (last) { tcur = tf; finished = true; }
                        else { tcur += h; if (lastrej) q = fmin(q, 1.0); lastrej = false; }
                    } else {
                        lastrej = true; q = fmin(q, 1.0);
                    }
                    h *= q;
                }
            }
        }
        if (active && !finished) {                                       // drive_rk8's loop-top checks
            if (h < span * 1e-12) { status = LTO_ST_HMIN; finished = true; }
            else if (nt >= a.cfg.max_attempts) { status = LTO_ST_MAXSTEPS; finished = true; }
        }
        if (active && finished) {
            // ---- defect = x(t1) - XC_all[:, i+1] (multiShoot_CRTBP_indirect.jl:82), back in the reference's variables
            const double* zs = zbuf + zi * ND * TS;
            double xv[ND];
#pragma unroll
            for (int q = 0; q < 3; ++q) { xv[q] = zs[q * TS]; xv[3 + q] = zs[(3 + q) * TS]; xv[9 + q] = zs[(6 + q) * TS]; }
            xv[6] = fma(w2, xv[10], -zs[9 * TS]); xv[7] = fma(-w2, xv[9], -zs[10 * TS]); xv[8] = -zs[11 * TS];
            bool nan = false;
#pragma unroll
            for (int i = 0; i < ND; ++i) {
                nan |= !(xv[i] == xv[i]);
                a.defect[seg * ND + i] = a.x_target ? xv[i] - a.x_target[ia * ND + i] : xv[i];
            }
            if (nan && status == 0) status = LTO_ST_NAN;
            if (a.status) a.status[seg] = status;
            if (a.nsteps_out) { a.nsteps_out[2 * seg] = na; a.nsteps_out[2 * seg + 1] = nt; }
            flags |= F_STORE; store_seg = (int)seg;
            active = false;
        }
        if (!active && nxt.nseg >= 0) {                                  // take the successor prepared by prepare_next()
            seg = nxt.nseg; nxt.nseg = -1; ia = lto_node_a(seg, a.npt);
            double* zs = zbuf + zi * ND * TS;
#pragma unroll
            for (int q = 0; q < 3; ++q) { zs[q * TS] = nxs[q * TS]; zs[(3 + q) * TS] = nxs[(3 + q) * TS]; zs[(6 + q) * TS] = nxs[(9 + q) * TS]; }
            zs[9 * TS] = fma(w2, nxs[10 * TS], -nxs[6 * TS]); zs[10 * TS] = fma(-w2, nxs[9 * TS], -nxs[7 * TS]); zs[11 * TS] = -nxs[8 * TS];
            tcur = nxs[(ND + 0) * TS]; tf = nxs[(ND + 1) * TS];
            span = tf - tcur;
            lw.aL = nxs[(ND + 2) * TS]; lw.rho_inv = nxs[(ND + 3) * TS]; lw.rq = nxs[(ND + 4) * TS];
            h = nxs[(ND + 5) * TS];
            if (!JOINT) { const double ts = nxs[(ND + 6) * TS]; atol = a.cfg.atol * ts; rtol = a.cfg.rtol * ts; }
            na = 0; nt = 0; status = 0; lastrej = false;
            active = true; flags |= F_RESET;
        }
        if (!__any_sync(fullmask, active) && !retried) {                 // nobody has work (only after segments ended in error): claim on demand, once
            retried = true; soon = true;
            continue;
        }
        retried = false;
        if (!__any_sync(fullmask, active)) {
            S.hval[slot] = 0.0; S.hctl[slot] = make_int2(flags | (par ? F_PAR : 0), store_seg);
            if (lane == 0) *S.tile_done = 1;
            mbar_arrive(S.bar_full);
            break;
        }
        // ---- one attempted step: 13 stages, rolled
        const long long c2 = clock64();
        c_pre += c2 - c1;
        last = false;
        if (tcur + h >= tf) { h = tf - tcur; last = true; }
        if (active) ++nt;
        S.hval[slot] = h; S.hctl[slot] = make_int2(flags | (active ? F_ACTIVE : 0) | (par ? F_PAR : 0), store_seg);
        const double h2 = h * h;
        K3 Kr, Kl;
        const double* z = zbuf + zi * ND * TS;
        state_stage<0>(Kr, Kl, z, h, h2, a.c, w2, lw, rec);   state_stage<1>(Kr, Kl, z, h, h2, a.c, w2, lw, rec);   state_stage<2>(Kr, Kl, z, h, h2, a.c, w2, lw, rec);
        state_stage<3>(Kr, Kl, z, h, h2, a.c, w2, lw, rec);   state_stage<4>(Kr, Kl, z, h, h2, a.c, w2, lw, rec);   state_stage<5>(Kr, Kl, z, h, h2, a.c, w2, lw, rec);
        state_stage<6>(Kr, Kl, z, h, h2, a.c, w2, lw, rec);   state_stage<7>(Kr, Kl, z, h, h2, a.c, w2, lw, rec);   state_stage<8>(Kr, Kl, z, h, h2, a.c, w2, lw, rec);
        state_stage<9>(Kr, Kl, z, h, h2, a.c, w2, lw, rec);   state_stage<10>(Kr, Kl, z, h, h2, a.c, w2, lw, rec);  state_stage<11>(Kr, Kl, z, h, h2, a.c, w2, lw, rec);
        state_stage<12>(Kr, Kl, z, h, h2, a.c, w2, lw, rec);
        {
            const double r[3] = {z[0 * TS], z[1 * TS], z[2 * TS]}, v[3] = {z[3 * TS], z[4 * TS], z[5 * TS]}, lv[3] = {z[6 * TS], z[7 * TS], z[8 * TS]},
                         lvd[3] = {z[9 * TS], z[10 * TS], z[11 * TS]};
            double rn[3], vn[3], lvn[3], lvdn[3];
            step_update(Kr, r, v, h, h2, rn, vn);
            step_update(Kl, lv, lvd, h, h2, lvn, lvdn);
            esum = state_err_sumsq<!JOINT>(Kr, Kl, w2, h, h2, r, v, lv, lvd, rn, vn, lvn, lvdn, atol, rtol);
            double* zc = zbuf + (zi ^ 1) * ND * TS;
#pragma unroll
            for (int q = 0; q < 3; ++q) { zc[q * TS] = rn[q]; zc[(3 + q) * TS] = vn[q]; zc[(6 + q) * TS] = lvn[q]; zc[(9 + q) * TS] = lvdn[q]; }
        }
        mbar_arrive(S.bar_full);                                         // the whole attempt's record
        c_work += clock64() - c2;
        have = true; ++visit;
        flags = 0; store_seg = 0;
        soon = last || !active;
    }
    if (a.prof && lane == 0) {
        unsigned long long* o = a.prof + ((size_t)blockIdx.x * NW + NCW + t) * 4;
        o[0] = c_work; o[1] = c_wait; o[2] = visit; o[3] = clock64() - c_begin;
        a.prof[(size_t)gridDim.x * NW * 4 + (size_t)blockIdx.x * NTILE + t] = c_pre;
    }
}

template <bool JOINT, int RC = REG_COL, int RS = REG_STATE>
__global__ void __launch_bounds__(NTHREADS, 1) k_indirect_hc(const __grid_constant__ IndirectArgs a) {   // (grid constant: out-of-line callees take its address without a local copy)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < NTILE) {
        const TileSmem S = tile_smem(smem_raw, threadIdx.x);
        mbar_init(S.bar_full, 32);
#ifdef LTO_IHC_ISO
        mbar_init(S.bar_done, NTASK * 32);
#else
        mbar_init(S.bar_done, NCT);
#endif
        *S.tile_done = 0;
    }
    __syncthreads();
    // warp groups 0 and 1 (warps 0..7, two per SM sub-partition): column warps, give registers away;
    // warp group 2 (warps 8..11, one per sub-partition): state warps of tiles 0..2 take them (warp 11 has no tile)
#ifdef LTO_IHC_ISO
    if ((warp & 3) == 3) state_warp<JOINT>(a, warp >> 2, lane, smem_raw);
    else column_warp<JOINT>(a, warp - (warp >> 2), lane, smem_raw);
#else
    if (warp < NCW) {
        reg_dec<RC>();
        column_warp<JOINT>(a, warp, lane, smem_raw);
    } else {
        reg_inc<RS>();
        if (warp - NCW < NTILE) state_warp<JOINT>(a, warp - NCW, lane, smem_raw);
    }
#endif
}

}  // namespace ihc

size_t indirect_hc_scratch_bytes(int n_sm) { return (size_t)n_sm * ihc::SCR_DOUBLES_PER_CTA * sizeof(double); }

template <bool JOINT, int RC = ihc::REG_COL, int RS = ihc::REG_STATE>
static cudaError_t launch_ihc(const IndirectArgs& a, cudaStream_t st) {
    // per device: a single process may drive several GPUs (lto_init_devices)
    static int n_sm_dev[64] = {0};
    static bool attr_dev[64] = {false};
    int dev = 0;
    { cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e; }
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!attr_dev[dev]) {
        cudaError_t e = cudaDeviceGetAttribute(&n_sm_dev[dev], cudaDevAttrMultiProcessorCount, dev); if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(ihc::k_indirect_hc<JOINT, RC, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ihc::SMEM);
        if (e != cudaSuccess) return e;
        attr_dev[dev] = true;
    }
    const int n_sm = n_sm_dev[dev];
    cudaError_t e = cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    const long long per_cta = (long long)ihc::NTILE * ihc::TS;
    const int grid = (int)std::min<long long>((a.n_seg + per_cta - 1) / per_cta, (long long)n_sm);
    ihc::k_indirect_hc<JOINT, RC, RS><<<grid, ihc::NTHREADS, ihc::SMEM, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_indirect_hc(const IndirectArgs& a, cudaStream_t st, int* n_launch) {
    *n_launch = 0;
    if (a.phi == nullptr || a.counter == nullptr || a.scratch == nullptr || a.cfg.controller != 0 || a.n_seg <= 0 || a.n_seg > 0x7fffffffll)
        return cudaErrorNotSupported;
    cudaError_t e;
#ifdef LTO_HC_VARIANTS                                                  // development: register splits side by side (LTO_HC_REGS=152 | 144)
    static int v = -1;
    if (v < 0) { const char* s = getenv("LTO_HC_REGS"); v = s ? atoi(s) : 0; }
    if (v == 152 && a.cfg.err_norm != 0) e = launch_ihc<true, 152, 200>(a, st);
    else if (v == 144 && a.cfg.err_norm != 0) e = launch_ihc<true, 144, 216>(a, st);
    else if (v == 136 && a.cfg.err_norm != 0) e = launch_ihc<true, 136, 232>(a, st);
    else
#endif
    e = (a.cfg.err_norm != 0) ? launch_ihc<true>(a, st) : launch_ihc<false>(a, st);
    if (e == cudaSuccess) *n_launch = 1;
    return e;
}

}  // namespace lto
