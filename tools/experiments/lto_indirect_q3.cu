// lto_indirect_q3.cu -- throughput kernel for the indirect method, 12-dim, defect + STM (K3, "q-split" layout):
// defectCalc + jacobianCalc of multiShoot_CRTBP_indirect.jl:63-124 for a whole batch in one launch, the
// variational equations standing in for ForwardDiff-through-the-solver (:103-121).
//
// Why a second layout.  In lto_indirect_cw.cu a thread owns a whole STM column (12 components x 13 stage
// derivatives = 117 doubles): 255 registers, 8 warps per SM, two per sub-partition -- the FP64 pipe idles on
// dependent-issue latency and the state warps' long chain cannot be hidden (ncu: profiles/r01_k_indirect_cw_v2b).
// Here every 12-vector y = [r v lr lv] is split over THREE adjacent lanes by Cartesian component q in {x, y, z}:
// lane q holds (r_q, v_q, lr_q, lv_q) and the q-components of the stage derivatives (39 doubles).  The RK
// combinations are purely local; per stage a lane needs the other two lanes' r and lv (two shuffles each) and one
// v for the Coriolis term.  ~130 registers per thread -> 15 warps per SM.
//
//   CTA    = 2 tiles in flight x 30 segment slots
//   state warps  (3 per tile): lane triple = slot.  Evaluate CRTBP_stateCostate_deriv! and ROW q of its linearisation
//                 (U = U_xx, W = d(U lv)/dr, G = du/dlv) once per stage, publish them (symmetric storage) in shared memory.
//   column warps (9): lane triple = (slot, STM column); 360 column tasks per tile visit = 4 phases x 9 warps x 10 triples.
//   hand-off     = two mbarriers per tile, as in lto_indirect_cw.cu; slots are refilled from a global work queue.
// Step control, controller constants and the Hairer initial step are those of lto_indirect_cw.cu / drive_rk8.
#include "lto_internal.h"
#include "lto_cw_common.cuh"
#include <algorithm>

namespace lto {
namespace iq3 {

using namespace cwc;

constexpr int ND = 12;
constexpr int NTILE = 3;               // tiles in flight: the columns never wait as long as a state attempt takes < 2 column visits
constexpr int TS = 20;                 // slots per tile
constexpr int TSP = 20;
constexpr int NSW = 2;                 // state warps per tile (10 slots each)
constexpr int NCW = 8;                 // column warps
constexpr int NPH = 3;                 // phases per tile visit
constexpr int NTASK = TS * ND;         // 240 = NPH * NCW * 10
constexpr int CSTR = NTASK * 3;        // 720 lanes' worth of one component
constexpr int NW = NTILE * NSW + NCW;  // 14
constexpr int NTHREADS = 32 * NW;      // 448
constexpr int RS = 13 * 18;            // doubles per slot of the stage records: 13 stages x 3 matrices x [xx xy zz xz yy yz]
static_assert(NPH * NCW * 10 == NTASK, "phases must tile the column tasks");

enum { F_ACCEPT = 1, F_STORE = 2, F_RESET = 4, F_ACTIVE = 8 };

// The tableau as __constant__ data: a q-split lane uses every coefficient for only 1..3 FMAs, so materialising the 64-bit
// literals (two UMOVs each) would cost more issue slots than the arithmetic; a constant-bank operand costs none.
// Sparsity still comes from the constexpr tables (zero entries are skipped at compile time).
__constant__ double qB[13][13] = LTO_TAB_B_INIT;
__constant__ double qG[13][13] = LTO_TAB_G_INIT;
__constant__ double qC[13] = LTO_TAB_C_INIT;
__constant__ double qCHI[13] = LTO_TAB_CHI_INIT;
__constant__ double qCHIB[13] = LTO_TAB_CHIB_INIT;

constexpr size_t REC_BYTES = (size_t)TS * RS * sizeof(double);                 // 56,400
constexpr size_t HDR_BYTES = (size_t)TSP * (sizeof(double) + sizeof(int2));    // h, {flags, segment}
constexpr size_t CUR_BYTES = (size_t)4 * CSTR * sizeof(double);                // current columns
constexpr size_t ERR_BYTES = (size_t)36 * TSP * sizeof(double);                // error partials per (column, q) and slot
constexpr size_t TILE_BYTES = REC_BYTES + HDR_BYTES + CUR_BYTES + ERR_BYTES;
constexpr size_t BAR_BYTES = 64;                                               // full, done, alive[3]
constexpr size_t SMEM = NTILE * TILE_BYTES + NTILE * BAR_BYTES;
constexpr size_t SCRATCH_BYTES_PER_CTA = (size_t)NTILE * 4 * CSTR * sizeof(double);   // candidate columns (global, L2 resident)

struct TileSmem {
    double* rec; double* hval; int2* hctl; double* cur; double* errp;
    unsigned bar_full, bar_done; int* alive;
};

__device__ __forceinline__ TileSmem tile_smem(unsigned char* base, int t) {
    unsigned char* p = base + (size_t)t * TILE_BYTES;
    TileSmem s;
    s.rec = reinterpret_cast<double*>(p); p += REC_BYTES;
    s.hval = reinterpret_cast<double*>(p); p += TSP * sizeof(double);
    s.hctl = reinterpret_cast<int2*>(p); p += TSP * sizeof(int2);
    s.cur = reinterpret_cast<double*>(p); p += CUR_BYTES;
    s.errp = reinterpret_cast<double*>(p);
    unsigned char* b = base + (size_t)NTILE * TILE_BYTES + (size_t)t * BAR_BYTES;
    s.bar_full = smem_u32(b); s.bar_done = smem_u32(b + 8);
    s.alive = reinterpret_cast<int*>(b + 16);
    return s;
}

// lane geometry of a triple
struct Tri {
    int q, srcA, srcB, srcV, base;     // srcA / srcB: lanes holding component (q+1)%3 / (q+2)%3
    bool valid;
};
__device__ __forceinline__ Tri make_tri(int lane) {
    Tri t;
    const int g = lane / 3;
    t.q = lane - 3 * g; t.base = 3 * g; t.valid = lane < 30;
    if (t.valid) {
        t.srcA = t.base + (t.q + 1) % 3; t.srcB = t.base + (t.q + 2) % 3;
        t.srcV = (t.q < 2) ? t.base + (t.q ^ 1) : lane;
    } else { t.srcA = lane; t.srcB = lane; t.srcV = lane; }
    return t;
}
__device__ __forceinline__ double shf(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
// sum over the triple in a fixed order (bitwise identical in its three lanes)
__device__ __forceinline__ double tri_sum(double v, const Tri& t) {
    const double a = shf(v, t.valid ? t.base : t.srcA), b = shf(v, t.valid ? t.base + 1 : t.srcA), c = shf(v, t.valid ? t.base + 2 : t.srcA);
    return (a + b) + c;
}

// ---------------------------------------------------------------------------
// q-components of the 13 stage derivatives of one 12-vector; Nystrom form for (r, v) as in lto_indirect_cw.cu.
// ---------------------------------------------------------------------------
struct KQ { double kv[13], kl[13], km[13]; };

template <int J>
__device__ __forceinline__ void stage_input(const KQ& K, const double (&y)[4], double h, double h2, double& R, double& V, double& L, double& M) {
    if (J == 0) { R = y[0]; V = y[1]; L = y[2]; M = y[3]; return; }
    double av = 0.0, ar = 0.0, al = 0.0, am = 0.0;
#pragma unroll
    for (int l = 0; l < J; ++l) {
        if (lto_tab::Bf(J, l) != 0.0) {
            av = fma(qB[J][l], K.kv[l], av);
            al = fma(qB[J][l], K.kl[l], al);
            am = fma(qB[J][l], K.km[l], am);
        }
        if (lto_tab::Gf(J, l) != 0.0) ar = fma(qG[J][l], K.kv[l], ar);
    }
    V = fma(h, av, y[1]);
    R = fma(h2, ar, fma(h * qC[J], y[1], y[0]));
    L = fma(h, al, y[2]);
    M = fma(h, am, y[3]);
}

// 8th-order update (ode.jl:937) and, if ERR, this lane's share of the scaled squared error (ode.jl:940)
template <bool ERR>
__device__ __forceinline__ double step_finish(const KQ& K, const double (&y)[4], double h, double h2, double atol, double rtol, double (&yn)[4]) {
    double sv = 0.0, sr = 0.0, sl = 0.0, sm = 0.0;
#pragma unroll
    for (int l = 0; l < 13; ++l) {
        if (lto_tab::CHIf(l) != 0.0) {
            sv = fma(qCHI[l], K.kv[l], sv);
            sl = fma(qCHI[l], K.kl[l], sl);
            sm = fma(qCHI[l], K.km[l], sm);
        }
        if (lto_tab::CHIBf(l) != 0.0) sr = fma(qCHIB[l], K.kv[l], sr);
    }
    yn[0] = fma(h2, sr, fma(h, y[1], y[0]));
    yn[1] = fma(h, sv, y[1]);
    yn[2] = fma(h, sl, y[2]);
    yn[3] = fma(h, sm, y[3]);
    double esum = 0.0;
    if (ERR) {
        const double ce = h * lto_tab::ERRC, ce2 = h2 * lto_tab::ERRC;
        double e[4];
        e[0] = ce2 * (K.kv[0] - K.kv[11]);                                   // psi^T B = e_1 - e_12
        e[1] = ce * ((K.kv[0] + K.kv[10]) - (K.kv[11] + K.kv[12]));
        e[2] = ce * ((K.kl[0] + K.kl[10]) - (K.kl[11] + K.kl[12]));
        e[3] = ce * ((K.km[0] + K.km[10]) - (K.km[11] + K.km[12]));
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double sc = fma(rtol, fmax(fabs(y[b]), fabs(yn[b])), atol);
            const double r = e[b] * fast_rcp(sc);
            esum = fma(r, r, esum);
        }
    }
    return esum;
}

// ---------------------------------------------------------------------------
// Column triple: one attempted RK step of one STM column.
//   kv_q = U_q.pr + (C pv)_q + G_q.plv ;  kl_q = -(W_q.pr + U_q.plv) ;  km_q = -plr_q - (C^T plv)_q
// ---------------------------------------------------------------------------
struct ColGeo { int pb, sb; double wq, wA, wB; };     // LDS.128 at pb -> (M_qq, M_qA), LDS.64 at sb -> M_qB

template <int J>
__device__ __forceinline__ void col_stage(KQ& K, const double (&p)[4], double h, double h2, const double* __restrict__ rec, const Tri& t, const ColGeo& g) {
    double R, V, L, M;
    stage_input<J>(K, p, h, h2, R, V, L, M);
    const double RA = shf(R, t.srcA), RB = shf(R, t.srcB);
    const double MA = shf(M, t.srcA), MB = shf(M, t.srcB);
    const double Vo = shf(V, t.srcV);
    const double* w = rec + J * 18;
    const double2 u2 = *reinterpret_cast<const double2*>(w + g.pb), w2v = *reinterpret_cast<const double2*>(w + 6 + g.pb),
                  g2 = *reinterpret_cast<const double2*>(w + 12 + g.pb);
    const double U0 = u2.x, UA = u2.y, UB = w[g.sb];
    const double W0 = w2v.x, WA = w2v.y, WB = w[6 + g.sb];
    const double G0 = g2.x, GA = g2.y, GB = w[12 + g.sb];
    K.kv[J] = fma(U0, R, fma(UA, RA, fma(UB, RB, fma(G0, M, fma(GA, MA, fma(GB, MB, g.wq * Vo))))));
    K.kl[J] = -fma(W0, R, fma(WA, RA, fma(WB, RB, fma(U0, M, fma(UA, MA, UB * MB)))));
    K.km[J] = fma(g.wA, MA, fma(g.wB, MB, -L));               // -L - (C^T M)_q
}

// The column warps run the same ~20 KB straight-line body; re-converging them a few times per attempt keeps their
// instruction-fetch windows together (one miss stream instead of eight).
#ifndef LTO_Q3_NOLOCKSTEP
#define LTO_Q3_LOCKSTEP() asm volatile("bar.sync 1, %0;" ::"n"(32 * NCW) : "memory")
#else
#define LTO_Q3_LOCKSTEP()
#endif

template <bool ERR>
__device__ __forceinline__ double col_attempt(const double (&p)[4], double h, const double* __restrict__ rec, const Tri& t, const ColGeo& g,
                                              double atol, double rtol, double (&pn)[4]) {
    const double h2 = h * h;
    KQ K;
    col_stage<0>(K, p, h, h2, rec, t, g);  col_stage<1>(K, p, h, h2, rec, t, g);  col_stage<2>(K, p, h, h2, rec, t, g);
    col_stage<3>(K, p, h, h2, rec, t, g);  col_stage<4>(K, p, h, h2, rec, t, g);  col_stage<5>(K, p, h, h2, rec, t, g);
    LTO_Q3_LOCKSTEP();
    col_stage<6>(K, p, h, h2, rec, t, g);  col_stage<7>(K, p, h, h2, rec, t, g);  col_stage<8>(K, p, h, h2, rec, t, g);
    LTO_Q3_LOCKSTEP();
    col_stage<9>(K, p, h, h2, rec, t, g);
    if (ERR) col_stage<10>(K, p, h, h2, rec, t, g);       // stage 11 enters only the error estimate
    else { K.kv[10] = 0.0; K.kl[10] = 0.0; K.km[10] = 0.0; }
    LTO_Q3_LOCKSTEP();
    col_stage<11>(K, p, h, h2, rec, t, g); col_stage<12>(K, p, h, h2, rec, t, g);
    return step_finish<ERR>(K, p, h, h2, atol, rtol, pn);
}

__device__ __forceinline__ double sel4(const double (&p)[4], int i) { return i == 0 ? p[0] : (i == 1 ? p[1] : (i == 2 ? p[2] : p[3])); }

template <bool JOINT>
__device__ __forceinline__ void column_warp(const IndirectArgs& a, int cw, int lane, unsigned char* smem) {
    const Tri t = make_tri(lane);
    const int q = t.q;
    ColGeo g;
    g.pb = (q == 0) ? 0 : (q == 1 ? 4 : 2); g.sb = (q == 0) ? 3 : (q == 1 ? 1 : 5);
    const double w2 = 2.0 * a.c.omega;
    g.wq = (q == 0) ? w2 : (q == 1 ? -w2 : 0.0); g.wA = (q == 0) ? w2 : 0.0; g.wB = (q == 1) ? -w2 : 0.0;
    const double atol = a.cfg.atol, rtol = a.cfg.rtol;
    const int gi = t.valid ? lane / 3 : 9;                                  // ghost lanes shadow triple 9 (never store)
    double* cand_cta = a.scratch + (size_t)blockIdx.x * (SCRATCH_BYTES_PER_CTA / sizeof(double));
    const bool wide = (reinterpret_cast<uintptr_t>(a.phi) & 31u) == 0;
    // output exchange tables (lane q writes elements [4q, 4q+4) of the column; element i = component i/3 of lane i%3)
    const int sup1 = (q == 0) ? 3 : (q == 1 ? 0 : 1), sup2 = (q == 0) ? 2 : (q == 1 ? 3 : 0), sup3 = q + 1;
    const int src1 = t.valid ? t.base + (q + 1) % 3 : lane, src2 = t.valid ? t.base + (q + 2) % 3 : lane;
    unsigned alive = (1u << NTILE) - 1u;
    unsigned visit = 0;
    int v3 = 0;
    while (alive) {
#pragma unroll 1
        for (int tl = 0; tl < NTILE; ++tl) {
            if (!(alive & (1u << tl))) continue;
            const TileSmem S = tile_smem(smem, tl);
            mbar_wait_parked(S.bar_full, visit & 1);
            const bool done = S.alive[v3] == 0;
#pragma unroll 1
            for (int ph = 0; ph < NPH; ++ph) {
                // warp-phases 0..29: one slot, columns 0..9 (every lane of equal q reads the same record words: one broadcast
                // wavefront per load); warp-phases 30..35: columns 10, 11 of five slots each
                const int wp = ph * NCW + cw;
                const int col = (wp < TS) ? gi : 10 + (gi & 1);
                const int slot = (wp < TS) ? wp : 5 * (wp - TS) + (gi >> 1);
                const int ci = (wp * 10 + gi) * 3 + q;
                const int2 hc = S.hctl[slot];
                const double h = S.hval[slot];
                double* sc = S.cur + ci;
                double* sn = cand_cta + (size_t)tl * 4 * CSTR + ci;
                double p[4];
                if (hc.x & F_ACCEPT) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) { p[c] = __ldcg(sn + c * CSTR); if (t.valid) sc[c * CSTR] = p[c]; }
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) p[c] = sc[c * CSTR];
                }
                if (__any_sync(0xffffffffu, (hc.x & F_STORE) != 0)) {
                    // column `col` of ForwardDiff.jacobian(f, x0) (:121): regroup the triple's 12 values so that every lane
                    // owns 4 consecutive elements and writes them with one 32-byte store
                    double o[4];
                    o[0] = shf(sel4(p, q), lane);
                    o[1] = shf(sel4(p, sup1), src1);
                    o[2] = shf(sel4(p, sup2), src2);
                    o[3] = shf(sel4(p, sup3), lane);
                    if ((hc.x & F_STORE) && t.valid) {
                        double* out = a.phi + (long long)hc.y * (ND * ND) + col * ND + 4 * q;
                        if (wide) asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(out), "d"(o[0]), "d"(o[1]), "d"(o[2]), "d"(o[3]) : "memory");
                        else { out[0] = o[0]; out[1] = o[1]; out[2] = o[2]; out[3] = o[3]; }
                    }
                }
                if (hc.x & F_RESET) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) { p[c] = (col == 3 * c + q) ? 1.0 : 0.0; if (t.valid) sc[c * CSTR] = p[c]; }
                }
#ifdef LTO_Q3_NOLOCKSTEP
                if (done || !__any_sync(0xffffffffu, (hc.x & F_ACTIVE) != 0)) continue;
#else
                if (done) continue;                                        // (uniform over the column warps: the lock-step barriers need all of them)
#endif
                double pn[4];
                const double es = col_attempt<JOINT>(p, h, S.rec + slot * RS, t, g, atol, rtol, pn);
                if (t.valid) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) __stcg(sn + c * CSTR, pn[c]);
                    if (JOINT) S.errp[(col * 3 + q) * TSP + slot] = es;
                }
            }
            if (done) alive &= ~(1u << tl);
            else mbar_arrive(S.bar_done);
        }
        ++visit; v3 = (v3 + 1) % 3;
    }
}

// ---------------------------------------------------------------------------
// State triple: right-hand side CRTBP_stateCostate_deriv! (src/CRTBP_stateCostate_deriv.jl:9-90), component q,
// and ROW q of its linearisation.  Every lane of the triple evaluates the scalar part (1/r^3, 1/r^5, control law)
// redundantly; d_b = (x + mu [- 1], y, z) is held in the triple's rotated order (q, q+1, q+2).
// ---------------------------------------------------------------------------
struct LawConst { double aL, rho_inv, rho_inv_quarter_aL; };
struct StGeo { double m1q, m1A, m1B, oq, oA, oB, wq, cq; bool q0; };     // mu / 1 on the x-component of each rotated slot

template <bool LIN>
__device__ __forceinline__ void sc_eval_q(double R, double RA, double RB, double Vo, double L, double M, double MA, double MB,
                                          const SCConst& c, const LawConst& lw, const StGeo& g, int q, bool publish,
                                          double& kv, double& kl, double& km, double* __restrict__ w) {
    // ---- gravity (:69-70, :78-81)
    const double d1q = R + g.m1q, d1A = RA + g.m1A, d1B = RB + g.m1B;
    const double d2q = d1q - g.oq, d2A = d1A - g.oA, d2B = d1B - g.oB;
    const double i1 = fast_rsqrt(fma(d1q, d1q, fma(d1A, d1A, d1B * d1B))), i2 = fast_rsqrt(fma(d2q, d2q, fma(d2A, d2A, d2B * d2B)));
    const double i1s = i1 * i1, i2s = i2 * i2;
    const double a31 = c.m1 * i1s * i1, a32 = c.mu * i2s * i2;
    const double a51 = 3.0 * a31 * i1s, a52 = 3.0 * a32 * i2s;
    const double gg = -(a31 + a32);
    // ---- control law (:36-64): u_acc = -umag * lv/|lv| = -uon * lv
    const double n2 = fma(M, M, fma(MA, MA, MB * MB));
    const bool dead = !(n2 > 0.0);                                     // :59-64 NaN guard -> zero control
    const double in = dead ? 0.0 : fast_rsqrt(n2);
    const double n = n2 * in;
    double umag, dn = 0.0;
    if (c.p == 1.0) {                                                  // :41-43
        const double y = fmin(fmax((n - 1.0) * lw.rho_inv, -700.0), 700.0);
        const double ey = exp(y);                                      // tanh(y/2) = 1 - 2/(e^y + 1)
        const double th = fma(-2.0, fast_rcp(ey + 1.0), 1.0);
        umag = fma(0.5 * lw.aL, th, 0.5 * lw.aL);
        dn = lw.rho_inv_quarter_aL * fma(-th, th, 1.0);
    } else if (c.p == 0.0) {                                           // :36-39
        umag = lw.aL;
    } else {                                                           // :45-50
        const double e = 1.0 / (c.p - 1.0);
        const double wv = (c.p == 2.0) ? 0.5 * n : pow(n / c.p, e);
        if (wv > lw.aL) umag = lw.aL;
        else { umag = wv; dn = dead ? 0.0 : e * wv * in; }
    }
    if (dead) { umag = 0.0; dn = 0.0; }
    if (!(n2 == n2)) umag = n2;                                        // a NaN costate stays NaN (reported through status[])
    const double uon = umag * in;
    // ---- row q of U = U_xx, derivatives (:78-88)
    const double p1 = a51 * d1q, p2 = a52 * d2q;
    const double Uqq = fma(p1, d1q, fma(p2, d2q, g.cq + gg));
    const double UqA = fma(p1, d1A, p2 * d2A);
    const double UqB = fma(p1, d1B, p2 * d2B);
    kv = fma(-uon, M, fma(-a31, d1q, fma(-a32, d2q, fma(g.wq, Vo, g.cq * R))));
    kl = -fma(Uqq, M, fma(UqA, MA, UqB * MB));
    km = fma(g.wq, g.q0 ? MA : MB, -L);
    if (LIN) {
        // row q of W = d(U lv)/dr and of G = du_acc/dlv = -uon I + (uon - dn) lh lh^T
        const double e1 = a51 * fma(d1q, M, fma(d1A, MA, d1B * MB)), e2 = a52 * fma(d2q, M, fma(d2A, MA, d2B * MB));
        const double h1 = -5.0 * e1 * i1s, h2 = -5.0 * e2 * i2s;
        const double ee = e1 + e2;
        const double hq1 = h1 * d1q, hq2 = h2 * d2q;
        const double pm = (p1 + p2) * 1.0;
        const double Wqq = fma(hq1, d1q, fma(hq2, d2q, fma(2.0 * pm, M, ee)));
        const double WqA = fma(hq1, d1A, fma(hq2, d2A, fma(p1, MA, fma(p2, MA, M * fma(a51, d1A, a52 * d2A)))));
        const double WqB = fma(hq1, d1B, fma(hq2, d2B, fma(p1, MB, fma(p2, MB, M * fma(a51, d1B, a52 * d2B)))));
        const double cd = uon - dn;
        const double lq = M * in, lA = MA * in, lB = MB * in;
        const double cl = cd * lq;
        const double Gqq = fma(cl, lq, -uon), GqA = cl * lA, GqB = cl * lB;
        if (publish) {
            // symmetric storage [xx xy zz xz yy yz]: lane 0 owns xx, xy, xz; lane 1 yy, yz; lane 2 zz
            const int dq = (q == 0) ? 0 : (q == 1 ? 4 : 2);
            w[dq] = Uqq; w[6 + dq] = Wqq; w[12 + dq] = Gqq;
            if (q == 0) { w[1] = UqA; w[3] = UqB; w[7] = WqA; w[9] = WqB; w[13] = GqA; w[15] = GqB; }
            if (q == 1) { w[5] = UqA; w[11] = WqA; w[17] = GqA; }
        }
    }
}

// One out-of-line copy of the right-hand side serves all 13 stages (keeps the state warps' code small: they share the
// instruction caches with the column warps).  Everything travels in registers.
__device__ __forceinline__ StGeo make_stgeo(int q, double mu, double w2) {
    StGeo g;
    g.m1q = (q == 0) ? mu : 0.0; g.m1A = (q == 2) ? mu : 0.0; g.m1B = (q == 1) ? mu : 0.0;
    g.oq = (q == 0) ? 1.0 : 0.0; g.oA = (q == 2) ? 1.0 : 0.0; g.oB = (q == 1) ? 1.0 : 0.0;
    g.wq = (q == 0) ? w2 : (q == 1 ? -w2 : 0.0); g.cq = (q < 2) ? 1.0 : 0.0; g.q0 = (q == 0);
    return g;
}
__device__ __noinline__ double sc_eval_q_call(double R, double RA, double RB, double Vo, double L, double M, double MA, double MB,
                                              double mu, double p, double omega, double aL, double rho_inv, double rq, int q, bool publish,
                                              double* w, double* out2) {
    SCConst c; c.mu = mu; c.m1 = 1.0 - mu; c.p = p; c.omega = omega;
    LawConst lw; lw.aL = aL; lw.rho_inv = rho_inv; lw.rho_inv_quarter_aL = rq;
    const StGeo g = make_stgeo(q, mu, 2.0 * omega);
    double kv, kl, km;
    sc_eval_q<true>(R, RA, RB, Vo, L, M, MA, MB, c, lw, g, q, publish, kv, kl, km, w);
    out2[0] = kl; out2[1] = km;
    return kv;
}

template <int J>
__device__ __forceinline__ void state_stage(KQ& K, const double (&x)[4], double h, double h2, const SCConst& c, const LawConst& lw,
                                            const StGeo& g, const Tri& t, double* __restrict__ rec) {
    double R, V, L, M;
    stage_input<J>(K, x, h, h2, R, V, L, M);
    const double RA = shf(R, t.srcA), RB = shf(R, t.srcB);
    const double MA = shf(M, t.srcA), MB = shf(M, t.srcB);
    const double Vo = shf(V, t.srcV);
    double out2[2];
    K.kv[J] = sc_eval_q_call(R, RA, RB, Vo, L, M, MA, MB, c.mu, c.p, c.omega, lw.aL, lw.rho_inv, lw.rho_inv_quarter_aL, t.q, t.valid, rec + J * 18, out2);
    K.kl[J] = out2[0]; K.km[J] = out2[1];
}

// scaled RMS partial over this lane's 4 components
__device__ __forceinline__ double ssq4(const double (&e)[4], const double (&y)[4], double atol, double rtol) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const double r = e[i] * fast_rcp(fma(rtol, fabs(y[i]), atol)); s = fma(r, r, s); }
    return s;
}

template <bool JOINT>
__device__ __forceinline__ void state_warp(const IndirectArgs& a, int tl, int sw, int lane, unsigned char* smem) {
    const TileSmem S = tile_smem(smem, tl);
    const Tri t = make_tri(lane);
    const int q = t.q;
    const int slot = sw * 10 + (t.valid ? lane / 3 : 9);
    const StGeo g = make_stgeo(q, a.c.mu, 2.0 * a.c.omega);
    const unsigned fullmask = 0xffffffffu;
    const double atol = a.cfg.atol, rtol = a.cfg.rtol;
    const double inv_ne = JOINT ? 1.0 / (double)(ND * (ND + 1)) : 1.0 / (double)ND;
    double x[4] = {0.0, 0.0, 0.0, 0.0}, xn[4] = {0.0, 0.0, 0.0, 0.0};
    double tcur = 0.0, tf = 0.0, h = 0.0, span = 1.0, esum = 0.0;
    LawConst lw; lw.aL = 0.0; lw.rho_inv = 1.0; lw.rho_inv_quarter_aL = 0.0;
    long long seg = -1, ia = 0;
    int na = 0, nt = 0, status = 0;
    bool active = false, lastrej = false, last = false, have = false, exhausted = !t.valid;
    unsigned visit = 0;
    int v3 = 0;                                                           // visit % 3: which `alive` counter this visit uses
    double* rec = S.rec + slot * RS;
    while (true) {
        int flags = 0, store_seg = 0;
        bool finished = false;
        // (every shuffle below is executed by the whole warp: lanes differ in `active`, never in control flow around a shuffle)
        if (have) {
            mbar_wait_parked(S.bar_done, (visit - 1) & 1);
            if (sw == 0 && lane == 0) S.alive[(v3 + 1) % 3] = 0;           // last read during visit - 2; next used by visit + 1
            double s2 = esum;
            if (JOINT) {
#pragma unroll
                for (int c = 0; c < ND; ++c) s2 += S.errp[(c * 3 + q) * TSP + slot];
            }
            const double eest = sqrt(tri_sum(s2, t) * inv_ne);
            if (active) {
                if (!(eest == eest)) { status = LTO_ST_NAN; finished = true; }
                else {
                    double f = (eest == 0.0) ? 5.0 : 0.9 * inv_eighth_root(eest);
                    f = fmin(5.0, fmax(0.2, f));
                    if (eest <= 1.0) {
                        ++na; flags |= F_ACCEPT;
#pragma unroll
                        for (int i = 0; i < 4; ++i) x[i] = xn[i];
                        if (last) { tcur = tf; finished = true; }
                        else { tcur += h; if (lastrej) f = fmin(f, 1.0); lastrej = false; }
                    } else {
                        lastrej = true; f = fmin(f, 1.0);
                    }
                    h *= f;
                }
            }
        }
        if (active && !finished) {                                       // drive_rk8's loop-top checks
            if (h < span * 1e-12) { status = LTO_ST_HMIN; finished = true; }
            else if (nt >= a.cfg.max_attempts) { status = LTO_ST_MAXSTEPS; finished = true; }
        }
        {
            bool nanl = false;
#pragma unroll
            for (int i = 0; i < 4; ++i) nanl |= !(x[i] == x[i]);
            const int n3 = (int)tri_sum(nanl ? 1.0 : 0.0, t);
            if (active && finished) {
                // ---- defect = x(t1) - XC_all[:, i+1] (multiShoot_CRTBP_indirect.jl:82)
                if (n3 && status == 0) status = LTO_ST_NAN;
#pragma unroll
                for (int i = 0; i < 4; ++i) a.defect[seg * ND + 3 * i + q] = a.x_target ? x[i] - a.x_target[ia * ND + 3 * i + q] : x[i];
                if (q == 0) {
                    if (a.status) a.status[seg] = status;
                    if (a.nsteps_out) { a.nsteps_out[2 * seg] = na; a.nsteps_out[2 * seg + 1] = nt; }
                }
                flags |= F_STORE; store_seg = (int)seg;
                active = false;
            }
        }
        bool fresh = false;
        {
            const bool want = !active && !exhausted;
            long long idx = 0;
            if (want && q == 0) idx = (long long)atomicAdd(a.counter, 1ull);
            idx = __shfl_sync(fullmask, idx, t.valid ? t.base : lane);
            if (want) {
                if (idx < a.n_seg) {
                    seg = idx; ia = lto_node_a(seg, a.npt);
                    const long long it = lto_traj_of(seg, a.npt);
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = a.x0[ia * ND + 3 * i + q];
                    tcur = a.t0[ia]; tf = a.t1[ia];
                    if (!(tcur < tf)) tf = tcur;                          // empty span: one zero-length step, Phi = I
                    span = tf - tcur;
                    const double tlim = a.thrustLimit_arr ? a.thrustLimit_arr[it] : a.c.thrustLimit;
                    const double rho = a.rho_arr ? a.rho_arr[it] : a.c.rho;
                    lw.aL = tlim * a.c.kthr / a.c.mass;                   // :33
                    lw.rho_inv = 1.0 / rho;
                    lw.rho_inv_quarter_aL = lw.aL / (4.0 * rho);
                    na = 0; nt = 0; status = 0; lastrej = false;
                    active = true; fresh = true; flags |= F_RESET;
                } else {
                    exhausted = true;
                }
            }
        }
        const bool warp_active = __any_sync(fullmask, active);
        if (warp_active && lane == 0) atomicAdd(&S.alive[v3], 1);
        if (!warp_active) {
            // nothing to integrate in this warp: deliver the flags and see whether the whole tile is finished
            if (t.valid && q == 0) { S.hval[slot] = 0.0; S.hctl[slot] = make_int2(flags, store_seg); }
            mbar_arrive(S.bar_full);
            mbar_wait_parked(S.bar_full, visit & 1);
            if (S.alive[v3] == 0) break;
            have = true; ++visit; v3 = (v3 + 1) % 3;
            continue;
        }
        // ---- one attempted step (13 stages); a fresh slot first picks its initial step
        KQ K;
        state_stage<0>(K, x, 0.0, 0.0, a.c, lw, g, t, rec);
        if (__any_sync(fullmask, fresh)) {
            // Hairer-Norsett-Wanner initial step over the state components (drive_rk8 in lto_prop_generic.cuh)
            double f0[4] = {x[1], K.kv[0], K.kl[0], K.km[0]}, y1[4];
            const double d0 = sqrt(tri_sum(ssq4(x, x, atol, rtol), t) * (1.0 / (double)ND));
            const double d1 = sqrt(tri_sum(ssq4(f0, x, atol, rtol), t) * (1.0 / (double)ND));
            double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
            h0 = fmin(h0, span);
#pragma unroll
            for (int i = 0; i < 4; ++i) y1[i] = fma(h0, f0[i], x[i]);
            {
                const double RA = shf(y1[0], t.srcA), RB = shf(y1[0], t.srcB), MA = shf(y1[3], t.srcA), MB = shf(y1[3], t.srcB), Vo = shf(y1[1], t.srcV);
                double kv1, kl1, km1;
                sc_eval_q<false>(y1[0], RA, RB, Vo, y1[2], y1[3], MA, MB, a.c, lw, g, q, false, kv1, kl1, km1, nullptr);
                y1[0] = y1[1] - f0[0]; y1[1] = kv1 - f0[1]; y1[2] = kl1 - f0[2]; y1[3] = km1 - f0[3];
            }
            const double d2 = sqrt(tri_sum(ssq4(y1, x, atol, rtol), t) * (1.0 / (double)ND)) / h0;
            const double dm = fmax(d1, d2);
            const double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : inv_eighth_root(dm / 0.01);
            if (fresh) h = fmin(fmin(100.0 * h0, h1), span);
        }
        last = false;
        if (tcur + h >= tf) { h = tf - tcur; last = true; }
        if (active) ++nt;
        if (t.valid && q == 0) { S.hval[slot] = h; S.hctl[slot] = make_int2(flags | (active ? F_ACTIVE : 0), store_seg); }
        const double h2 = h * h;
        state_stage<1>(K, x, h, h2, a.c, lw, g, t, rec);  state_stage<2>(K, x, h, h2, a.c, lw, g, t, rec);  state_stage<3>(K, x, h, h2, a.c, lw, g, t, rec);
        state_stage<4>(K, x, h, h2, a.c, lw, g, t, rec);  state_stage<5>(K, x, h, h2, a.c, lw, g, t, rec);  state_stage<6>(K, x, h, h2, a.c, lw, g, t, rec);
        state_stage<7>(K, x, h, h2, a.c, lw, g, t, rec);  state_stage<8>(K, x, h, h2, a.c, lw, g, t, rec);  state_stage<9>(K, x, h, h2, a.c, lw, g, t, rec);
        state_stage<10>(K, x, h, h2, a.c, lw, g, t, rec); state_stage<11>(K, x, h, h2, a.c, lw, g, t, rec); state_stage<12>(K, x, h, h2, a.c, lw, g, t, rec);
        esum = step_finish<true>(K, x, h, h2, atol, rtol, xn);
        mbar_arrive(S.bar_full);                                         // the whole attempt's record
        have = true; ++visit; v3 = (v3 + 1) % 3;
    }
}

template <bool JOINT>
__global__ void __launch_bounds__(NTHREADS, 1) k_indirect_q3(IndirectArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < NTILE) {
        const TileSmem S = tile_smem(smem_raw, threadIdx.x);
        mbar_init(S.bar_full, 32 * NSW);
        mbar_init(S.bar_done, 32 * NCW);
        S.alive[0] = 0; S.alive[1] = 0; S.alive[2] = 0;
    }
    __syncthreads();
    // warp w sits on SM sub-partition w % 4: state warps 0..5 (two each on sub-partitions 0 and 1, one each on 2 and 3),
    // column warps 6..13 (two per sub-partition) -- near-equal FP64 load (a column warp does ~2.8x a state warp's work).
    if (warp < NTILE * NSW) state_warp<JOINT>(a, warp / NSW, warp % NSW, lane, smem_raw);
    else column_warp<JOINT>(a, warp - NTILE * NSW, lane, smem_raw);
}

}  // namespace iq3

size_t indirect_q3_scratch_bytes(int n_sm) { return (size_t)n_sm * iq3::SCRATCH_BYTES_PER_CTA; }

template <bool JOINT>
static cudaError_t launch_iq3(const IndirectArgs& a, cudaStream_t st) {
    static int n_sm_dev[64] = {0};
    static bool attr_dev[64] = {false};
    int dev = 0;
    { cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e; }
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!attr_dev[dev]) {
        cudaError_t e = cudaDeviceGetAttribute(&n_sm_dev[dev], cudaDevAttrMultiProcessorCount, dev); if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(iq3::k_indirect_q3<JOINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)iq3::SMEM);
        if (e != cudaSuccess) return e;
        attr_dev[dev] = true;
    }
    const int n_sm = n_sm_dev[dev];
    cudaError_t e = cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    const long long per_cta = (long long)iq3::NTILE * iq3::TS;
    const int grid = (int)std::min<long long>((a.n_seg + per_cta - 1) / per_cta, (long long)n_sm);
    iq3::k_indirect_q3<JOINT><<<grid, iq3::NTHREADS, iq3::SMEM, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_indirect_q3(const IndirectArgs& a, int ndim, cudaStream_t st, int* n_launch) {
    *n_launch = 0;
    if (ndim != 12 || a.phi == nullptr || a.counter == nullptr || a.scratch == nullptr || a.cfg.controller != 0 || a.n_seg <= 0 ||
        a.n_seg > 0x7fffffffll)
        return cudaErrorNotSupported;
    cudaError_t e = (a.cfg.err_norm != 0) ? launch_iq3<true>(a, st) : launch_iq3<false>(a, st);
    if (e == cudaSuccess) *n_launch = 1;
    return e;
}

}  // namespace lto
