// lto_indirect_hc2.cu -- K3-hc with what rounds of measurement asked for: the half-column layout of lto_indirect_hc.cu
// (second-order variables, one thread per 3-vector Nystrom half-column, three tiles in flight, current / candidate columns in two L2 buffers)
// with
//   * the state warps' 13 stages ROLLED, their stage derivatives in SHARED memory (the 60 KB of unrolled state code was what evicted the
//     column loop from the instruction caches; the tensor-memory version of this loop, lto_indirect_hc_tmem.cu, waited on tcgen05.ld),
//   * tiles of 24 slots (the shared memory the stage derivatives need), 18 tasks per tile visit = 2 per column warp with NINE column warps,
//   * no setmaxnreg (nothing needs more than 168 registers any more),
//   * one named barrier per task that keeps the column warps on the same instructions (the K1 nstate-6 fix).
// Experimental: built only by tools/experiments/build_variant.sh, selected with LTO_K3=hc2.
#include "lto_internal.h"
#include "lto_cw_common.cuh"
#include "lto_hc_math.cuh"
#include <algorithm>
#include <cstdlib>

namespace lto {
namespace ihc2 {

using namespace cwc;
using namespace hcm;

constexpr int ND = 12;
constexpr int NTILE = 3;          // tiles in flight = state warps at work
constexpr int TS = 24;            // segment slots per tile (lanes 24..31 of a state warp stay idle: the stage derivatives need the shared memory)
constexpr int TL = 32;            // lanes: the small per-lane arrays are sized for the whole warp
constexpr int NCW = 9;            // column warps
constexpr int NCT = 32 * NCW;     // column threads
constexpr int NW = 12;            // + warp group 2: 3 state warps and one idle warp
constexpr int NTHREADS = 32 * NW;
constexpr int NTASK = 18;         // warp-tasks per tile visit: 6 column pairs x 3 slot octets
constexpr int NPH = NTASK / NCW;  // tasks per column warp and tile visit
constexpr int NC2 = 9;            // double2 per stage record: U[6] W[6] G[6]

enum { F_STORE = 2, F_RESET = 4, F_ACTIVE = 8, F_PAR = 16 };

// ---- shared-memory plan ------------------------------------------------------
constexpr size_t REC_BYTES = (size_t)13 * NC2 * TS * sizeof(double2);      // stage records of one tile
constexpr size_t HDR_BYTES = (size_t)TL * (sizeof(double) + sizeof(int2)); // h, {flags, segment}
constexpr size_t ERR_BYTES = (size_t)ND * TL * sizeof(double);             // error partials per column and slot (both halves summed)
constexpr int NXW = ND + 7;                                                // next-segment stash: x0, t0, tf, aL, 1/rho, aL/(4 rho), h0, tol scale
constexpr size_t NXT_BYTES = (size_t)NXW * TL * sizeof(double);
constexpr int KST = TS + 1;                                                // stage derivatives of the state: [stage][component][slot], one spare column for the idle lanes
constexpr size_t KS_BYTES = (size_t)13 * 6 * KST * sizeof(double);
constexpr size_t TILE_BYTES = REC_BYTES + HDR_BYTES + ERR_BYTES + NXT_BYTES + KS_BYTES;
constexpr size_t BAR_BYTES = 32;                                           // full, done, tile_done
constexpr size_t SMEM = NTILE * TILE_BYTES + NTILE * BAR_BYTES;
static_assert(SMEM <= 232448, "three tiles must fit the 227 KB of shared memory per CTA");
static_assert(TILE_BYTES % 16 == 0, "tiles must stay 16-byte aligned");
constexpr size_t SCR_DOUBLES_PER_CTA = (size_t)NTILE * 2 * NTASK * 6 * 32; // [tile][parity][task][component][lane]

struct TileSmem {
    double2* rec; double* hval; int2* hctl; double* errp; double* nx; double* ks;
    unsigned bar_full, bar_done; volatile int* tile_done;
};

__device__ __forceinline__ TileSmem tile_smem(unsigned char* base, int t) {
    unsigned char* p = base + (size_t)t * TILE_BYTES;
    TileSmem s;
    s.rec = reinterpret_cast<double2*>(p); p += REC_BYTES;
    s.hval = reinterpret_cast<double*>(p); p += TL * sizeof(double);
    s.hctl = reinterpret_cast<int2*>(p); p += TL * sizeof(int2);
    s.errp = reinterpret_cast<double*>(p); p += ERR_BYTES;
    s.nx = reinterpret_cast<double*>(p); p += NXT_BYTES;
    s.ks = reinterpret_cast<double*>(p);
    unsigned char* b = base + (size_t)NTILE * TILE_BYTES + (size_t)t * BAR_BYTES;
    s.bar_full = smem_u32(b); s.bar_done = smem_u32(b + 8);
    s.tile_done = reinterpret_cast<volatile int*>(b + 16);
    return s;
}


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
#ifdef LTO_IHC_LOCKSTEP_ON
#define LTO_IHC_LOCKSTEP() asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory")
#else
#define LTO_IHC_LOCKSTEP()
#endif

template <bool ERR>
__device__ __forceinline__ double col_attempt(const double (&p)[3], const double (&pd)[3], double h, double w2, const double2* __restrict__ rec,
                                              int half, double atol, double rtol, double (&pn)[3], double (&pdn)[3]) {
    const double h2 = h * h;
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
#pragma unroll 1
            for (int ph = 0; ph < NPH; ++ph) {
                asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");   // pace: the column warps start every task's instruction stream together
                const int task = ph * NCW + cw;
                const int col = 2 * (task / 3) + csel;
                const int slot = (task % 3) * 8 + s8;
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
                double es = col_attempt<JOINT>(p, pd, h, w2, S.rec + slot, half, atol, rtol, pn, pdn);
                if (a.prof) c_att += clock64() - ca;
#pragma unroll
                for (int q = 0; q < 3; ++q) { __stcg(cnd + q * 32, pn[q]); __stcg(cnd + (3 + q) * 32, pdn[q]); }
                if (JOINT) {
                    es += __shfl_xor_sync(0xffffffffu, es, 8);
                    if (half == 0) S.errp[col * TL + slot] = es;
                }
            }
            if (done) alive &= ~(1u << t);
            else { mbar_arrive(S.bar_done); c_work += clock64() - c1; n_work += NPH; }
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
// State warp.
//
// Code size matters as much as arithmetic here: the column warps stream a ~19 KB straight-line body, and every other instruction
// stream on the SM competes with it for the instruction caches (round 2, first version: 13 unrolled stages with the 78 stage
// derivatives in registers = 33 KB of state code, sm__icc_request_hit_rate 66 %, stall_no_instruction the top stall of BOTH warp
// kinds, the columns 55 % slower than with idle state warps; profiles/r02_*).  So the state warp runs ROLLED loops over stages:
//   * the stage derivatives k_J = (r'', lv'') live in the warp's own lane quadrant of TENSOR MEMORY (tcgen05.st / tcgen05.ld,
//     16 columns per stage and lane; 12-cycle loads, dynamically addressed) -- TMEM is otherwise unused by this FP64 kernel and is
//     the only on-chip store left: registers and shared memory are full;
//   * tableau coefficients come from __constant__ tables (uniform loads), zero entries skipped by uniform branches;
//   * the 8th-order update and the error estimate are accumulated stage by stage;
//   * one out-of-line copy of the right-hand side serves all 13 stages, arguments and results in registers.
// ---------------------------------------------------------------------------
__constant__ double tB[13][13] = LTO_TAB_B_INIT;
__constant__ double tG[13][13] = LTO_TAB_G_INIT;
__constant__ double tB11[13] = {lto_tab::Bf(11, 0), lto_tab::Bf(11, 1), lto_tab::Bf(11, 2), lto_tab::Bf(11, 3), lto_tab::Bf(11, 4), lto_tab::Bf(11, 5), lto_tab::Bf(11, 6),
                                lto_tab::Bf(11, 7), lto_tab::Bf(11, 8), lto_tab::Bf(11, 9), lto_tab::Bf(11, 10), 0.0, 0.0};
__constant__ double tC[13] = LTO_TAB_C_INIT;
__constant__ double tCHI[13] = LTO_TAB_CHI_INIT;
__constant__ double tCHIB[13] = LTO_TAB_CHIB_INIT;
__constant__ double tPSI[13] = LTO_TAB_PSI_INIT;
__constant__ double tPSIB[13] = LTO_TAB_PSIB_INIT;

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
        for (int i = 0; i < ND; ++i) nxs[i * TL] = x[i];
        nxs[(ND + 0) * TL] = t0; nxs[(ND + 1) * TL] = tf; nxs[(ND + 2) * TL] = lw.aL; nxs[(ND + 3) * TL] = lw.rho_inv; nxs[(ND + 4) * TL] = lw.rq;
        nxs[(ND + 5) * TL] = fmin(fmin(100.0 * h0, h1), span);
        nxs[(ND + 6) * TL] = ts;
    }
    return c;
}

template <bool JOINT>
__device__ __forceinline__ void state_warp(const IndirectArgs& a, int t, int lane, unsigned char* smem) {
    const TileSmem S = tile_smem(smem, t);
    const int slot = lane;
    const unsigned fullmask = 0xffffffffu;
    const double w2 = 2.0 * a.c.omega;
    double* const ks = S.ks + ((lane < TS) ? lane : TS);                  // this slot's stage derivatives (idle lanes share the spare column)
    double atol = a.cfg.atol, rtol = a.cfg.rtol;                          // per slot when the norm is the state's alone (state_tol_scale)
    const double inv_ne = JOINT ? 1.0 / (double)(ND * (ND + 1)) : 1.0 / (double)ND;
    int par = 0;                                                          // which scratch buffer holds the slot's current columns
    double z[ND], zn[ND];                                                 // z = (r, v, lv, lvd) and the candidate of the attempt in flight
#pragma unroll
    for (int i = 0; i < ND; ++i) { z[i] = 0.0; zn[i] = 0.0; }
    double tcur = 0.0, tf = 0.0, h = 0.0, span = 1.0, esum = 0.0;
    Law lw; lw.aL = 0.0; lw.rho_inv = 1.0; lw.rq = 0.0;
    long long seg = -1, ia = 0;
    int na = 0, nt = 0, status = 0;
    bool active = false, lastrej = false, last = false, have = false;
    Claim nxt; nxt.nseg = -1; nxt.exhausted = (lane >= TS) ? 1 : 0;   // (lanes without a slot never claim)
    //                          // successor segment claimed and staged by prepare_next()
    unsigned visit = 0;
    double2* rec = S.rec + slot;
    double* const nxs = S.nx + slot;
    long long c_wait = 0, c_work = 0, c_pre = 0, c_qin = 0, c_qev = 0, c_qst = 0;
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
                    for (int c = 0; c < ND; ++c) s2 += S.errp[c * TL + slot];
                }
                const double u = s2 * inv_ne;                           // eest^2: eest <= 1 <=> u <= 1, eest^(-1/8) = u^(-1/16)
                if (!(u == u)) { status = LTO_ST_NAN; finished = true; }
                else {
                    double q = (u == 0.0) ? 5.0 : ((u < 1e300) ? 0.9 * inv_sixteenth_root(u) : 0.2);
                    q = fmin(5.0, fmax(0.2, q));
                    if (u <= 1.0) {
                        ++na;
                        par ^= 1;                                         // the candidates become z and the current columns
#pragma unroll
                        for (int i = 0; i < ND; ++i) z[i] = zn[i];
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
            double xv[ND];
#pragma unroll
            for (int q = 0; q < 3; ++q) { xv[q] = z[q]; xv[3 + q] = z[3 + q]; xv[9 + q] = z[6 + q]; }
            xv[6] = fma(w2, z[7], -z[9]); xv[7] = fma(-w2, z[6], -z[10]); xv[8] = -z[11];
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
#pragma unroll
            for (int q = 0; q < 3; ++q) { z[q] = nxs[q * TL]; z[3 + q] = nxs[(3 + q) * TL]; z[6 + q] = nxs[(9 + q) * TL]; }
            z[9] = fma(w2, nxs[10 * TL], -nxs[6 * TL]); z[10] = fma(-w2, nxs[9 * TL], -nxs[7 * TL]); z[11] = -nxs[8 * TL];
            tcur = nxs[(ND + 0) * TL]; tf = nxs[(ND + 1) * TL];
            span = tf - tcur;
            lw.aL = nxs[(ND + 2) * TL]; lw.rho_inv = nxs[(ND + 3) * TL]; lw.rq = nxs[(ND + 4) * TL];
            h = nxs[(ND + 5) * TL];
            if (!JOINT) { const double ts = nxs[(ND + 6) * TL]; atol = a.cfg.atol * ts; rtol = a.cfg.rtol * ts; }
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
        long long q_in = 0, q_ev = 0, q_st = 0;
        double kp[6];                                                    // k_{J-1}: used from registers, every older one comes back from tensor memory
#pragma unroll
        for (int c = 0; c < 6; ++c) kp[c] = 0.0;
#pragma unroll 1
        for (int J = 0; J < 13; ++J) {
            double sb[6], sg[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) { sb[c] = 0.0; sg[c] = 0.0; }
            const long long q0 = a.prof ? clock64() : 0;
#pragma unroll 4
            for (int l = 0; l < J; ++l) {
                const double b = tB[J][l], g = tG[J][l];                 // (zero coefficients cost an FMA, not a branch)
#pragma unroll
                for (int c = 0; c < 6; ++c) { const double k = ks[(l * 6 + c) * KST]; sb[c] = fma(b, k, sb[c]); sg[c] = fma(g, k, sg[c]); }
            }
            const long long q1 = a.prof ? clock64() : 0;
            const double ch = h * tC[J];
            const double R[3] = {fma(h2, sg[0], fma(ch, z[3], z[0])), fma(h2, sg[1], fma(ch, z[4], z[1])), fma(h2, sg[2], fma(ch, z[5], z[2]))};
            const double M[3] = {fma(h2, sg[3], fma(ch, z[9], z[6])), fma(h2, sg[4], fma(ch, z[10], z[7])), fma(h2, sg[5], fma(ch, z[11], z[8]))};
            const double V[3] = {fma(h, sb[0], z[3]), fma(h, sb[1], z[4]), 0.0}, N[3] = {fma(h, sb[3], z[9]), fma(h, sb[4], z[10]), 0.0};
            double U[6], W[6], G[6];
            {   // right-hand side + stage record: ONE inlined copy (the stage loop is rolled)
                double kr[3], kl[3];
                sc_eval2<true>(R, V, M, N, a.c.mu, a.c.m1, w2, a.c.p, lw, kr, kl, U, W, G);
#pragma unroll
                for (int q = 0; q < 3; ++q) { kp[q] = kr[q]; kp[3 + q] = kl[q]; }
            }
            const long long q2 = a.prof ? clock64() : 0;
            double2* w = rec + J * NC2 * TS;
            if (lane < TS) {
            w[0 * TS] = make_double2(U[0], U[1]); w[1 * TS] = make_double2(U[2], U[3]); w[2 * TS] = make_double2(U[4], U[5]);
            w[3 * TS] = make_double2(W[0], W[1]); w[4 * TS] = make_double2(W[2], W[3]); w[5 * TS] = make_double2(W[4], W[5]);
            w[6 * TS] = make_double2(G[0], G[1]); w[7 * TS] = make_double2(G[2], G[3]); w[8 * TS] = make_double2(G[4], G[5]);
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) ks[(J * 6 + c) * KST] = kp[c];
            if (a.prof) { const long long q3 = clock64(); q_in += q1 - q0; q_ev += q2 - q1; q_st += q3 - q2; }
        }
        if (a.prof) { c_qin += q_in; c_qev += q_ev; c_qst += q_st; }
        // update (chi, chi^T B) and error (psi, psi^T B; ROB: -B[11], k_1 - k_12) sums, from the stage derivatives in shared memory: keeping
        // them as running sums through the stage loop costs 48 registers the right-hand side needs
        double su[6], sp[6], e1[6], e2[6], g1[6], g2[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) { su[c] = 0.0; sp[c] = 0.0; e1[c] = 0.0; e2[c] = 0.0; g1[c] = 0.0; g2[c] = 0.0; }
#pragma unroll 1
        for (int l = 0; l < 13; ++l) {
            const double wc = tCHI[l], wb = tCHIB[l], wp = tPSI[l], wq = tPSIB[l];
            const double wa = -tB11[l], wd = (l == 0) ? 1.0 : (l == 11) ? -1.0 : 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const double k = ks[(l * 6 + c) * KST];
                su[c] = fma(wc, k, su[c]); sp[c] = fma(wb, k, sp[c]); e2[c] = fma(wp, k, e2[c]); e1[c] = fma(wq, k, e1[c]);
                if (!JOINT) { g1[c] = fma(wa, k, g1[c]); g2[c] = fma(wd, k, g2[c]); }
            }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {                                    // 8th-order update (ode.jl:937)
            zn[q] = fma(h2, sp[q], fma(h, z[3 + q], z[q]));
            zn[3 + q] = fma(h, su[q], z[3 + q]);
            zn[6 + q] = fma(h2, sp[3 + q], fma(h, z[9 + q], z[6 + q]));
            zn[9 + q] = fma(h, su[3 + q], z[9 + q]);
        }
        esum = state_err_sumsq_acc<!JOINT>(e1, e2, g1, g2, w2, h, h2, z, zn, atol, rtol);
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
        unsigned long long* q = a.prof + (size_t)gridDim.x * (NW * 4 + NTILE + NCW) + ((size_t)blockIdx.x * NTILE + t) * 3;
        q[0] = c_qin; q[1] = c_qev; q[2] = c_qst;                        // stage-input loop | right-hand side | record + tensor-memory store + sums
    }
}

template <bool JOINT>
__global__ void __launch_bounds__(NTHREADS, 1) k_indirect_hc2(const __grid_constant__ IndirectArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < NTILE) {
        const TileSmem S = tile_smem(smem_raw, threadIdx.x);
        mbar_init(S.bar_full, 32);
        mbar_init(S.bar_done, NCT);
        *S.tile_done = 0;
    }
    __syncthreads();
    // warps 0..8: column warps (sub-partitions 0 1 2 3 0 1 2 3 0); warps 9..11: the state warps of tiles 0..2 (sub-partitions 1, 2, 3)
    if (warp < NCW) column_warp<JOINT>(a, warp, lane, smem_raw);
    else state_warp<JOINT>(a, warp - NCW, lane, smem_raw);
}

}  // namespace ihc2

size_t indirect_hc2_scratch_bytes(int n_sm) { return (size_t)n_sm * ihc2::SCR_DOUBLES_PER_CTA * sizeof(double); }

template <bool JOINT>
static cudaError_t launch_ihc2(const IndirectArgs& a, cudaStream_t st) {
    // per device: a single process may drive several GPUs (lto_init_devices)
    static int n_sm_dev[64] = {0};
    static bool attr_dev[64] = {false};
    int dev = 0;
    { cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e; }
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!attr_dev[dev]) {
        cudaError_t e = cudaDeviceGetAttribute(&n_sm_dev[dev], cudaDevAttrMultiProcessorCount, dev); if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(ihc2::k_indirect_hc2<JOINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ihc2::SMEM);
        if (e != cudaSuccess) return e;
        attr_dev[dev] = true;
    }
    const int n_sm = n_sm_dev[dev];
    cudaError_t e = cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    const long long per_cta = (long long)ihc2::NTILE * ihc2::TS;
    const int grid = (int)std::min<long long>((a.n_seg + per_cta - 1) / per_cta, (long long)n_sm);
    ihc2::k_indirect_hc2<JOINT><<<grid, ihc2::NTHREADS, ihc2::SMEM, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_indirect_hc2(const IndirectArgs& a, cudaStream_t st, int* n_launch) {
    *n_launch = 0;
    if (a.phi == nullptr || a.counter == nullptr || a.scratch == nullptr || a.cfg.controller != 0 || a.n_seg <= 0 || a.n_seg > 0x7fffffffll)
        return cudaErrorNotSupported;
    cudaError_t e;
    e = (a.cfg.err_norm != 0) ? launch_ihc2<true>(a, st) : launch_ihc2<false>(a, st);
    if (e == cudaSuccess) *n_launch = 1;
    return e;
}

}  // namespace lto
