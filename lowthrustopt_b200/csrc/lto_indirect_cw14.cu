// lto_indirect_cw14.cu -- throughput kernels for the indirect method with mass, ndim = 14
// (state + costate + mass: [r v m lr lv lm], 14 x 14 STM -- the system BASELINE.json's north_star names):
// defectCalc + jacobianCalc of multiShoot_CRTBP_indirect.jl:63-124 for a whole batch in one launch.  The
// 14-dim right-hand side is the CRTBP_stateCostate_deriv! form (src/CRTBP_stateCostate_deriv.jl:9-90)
// extended by the mass / mass-costate equations of GeneralCode/twoBody_stateCostate_mass_deriv.jl:26,57,61,76
// exactly as lto_math.cuh sc_stage<14> / sc_col<14> state them.
//
// Same column-warp layout as lto_indirect_cw.cu (K3), with what the two extra components change:
//   * 14 columns -> 7 column warps (warp w carries columns 2w, 2w+1 of 16 slots per half-phase) + ONE state warp
//     that serves both tiles alternately = 8 warps, 255 registers per thread; the column warps sit 2-2-2-1 on the
//     four SM sub-partitions, the state warp shares the last one.
//   * the mass-costate lm enters no right-hand side, so lm (and the lm-row of every STM column) is a pure
//     quadrature: no stage values are stored for it, only the running 8th-order sum and error combination.
//   * G = du/dlv is published as its generators (lh, uon, cd): G plv = -uon plv + cd (lh.plv) lh, and the mass
//     couplings gm, ml, lml are all multiples of lh -> a stage record is 22 doubles (12-dim: 18).
//   * the stage records of two tiles take 147 KB of shared memory, so there is no room for K3's shared-memory
//     candidate stash: a column's current value and its candidate live in two L2-resident buffers per (tile,
//     half-phase, thread) and an accepted step just flips which one is current (no copy).
#include "lto_internal.h"
#include "lto_cw_common.cuh"
#include <algorithm>

namespace lto {
namespace icw14 {

using namespace cwc;

constexpr int ND = 14;
constexpr int NTILE = 2;
constexpr int TS = 32;
constexpr int HS = 16;
constexpr int NCW = 7;
constexpr int NCT = 32 * NCW;     // 224 column threads
constexpr int NC2 = 11;           // double2 per stage record: U[6] W[6] lh[3] uon cd cgm cml clml mm lmm
constexpr int NW = 1 + NCW;       // 8 warps: 7 column warps + ONE state warp serving both tiles
constexpr int NTHREADS = 32 * NW;

enum { F_ACCEPT = 1, F_STORE = 2, F_RESET = 4, F_ACTIVE = 8 };

constexpr int NSS = 4;            // double2 per published stage STATE: (r0,r1) (r2,lv0) (lv1,lv2) (m,-)
constexpr size_t SST_BYTES = (size_t)13 * NSS * TS * sizeof(double2);       // stage states of one tile (state warp -> column warps)
constexpr size_t HDR_BYTES = (size_t)TS * (4 * sizeof(double) + sizeof(int2));   // h, tk, 1/rho, 1/(4 rho), {flags, segment}
// Candidate stash, slot-major (as in K3): a slot's 14 columns are 1568 contiguous bytes in the output's layout, so the state warp sends a finished
// segment's STM as ONE bulk store.  Slot stride 1584 B: 16-byte aligned, and a quarter-warp's 128-bit accesses fall into 32 different banks.
constexpr int STASH_STRIDE = ND * ND + 2;
constexpr size_t CUR_BYTES = (size_t)TS * STASH_STRIDE * sizeof(double);
constexpr size_t ERR_BYTES = (size_t)ND * TS * sizeof(double);
constexpr size_t XN_BYTES = (size_t)2 * ND * TS * sizeof(double);
constexpr size_t TILE_BYTES = SST_BYTES + HDR_BYTES + CUR_BYTES + ERR_BYTES + XN_BYTES;
constexpr size_t LIN_BYTES = (size_t)13 * NC2 * HS * sizeof(double2);        // stage linearisations of ONE half-phase (built by the column threads)
constexpr size_t BAR_BYTES = 64;
constexpr size_t LIN_OFFSET = NTILE * TILE_BYTES;
constexpr size_t BAR_OFFSET = LIN_OFFSET + LIN_BYTES;
constexpr size_t SMEM = BAR_OFFSET + NTILE * BAR_BYTES;
static_assert(SMEM <= 232448, "shared-memory plan exceeds 227 KB");
constexpr size_t SCRATCH_DOUBLES_PER_CTA = (size_t)NTILE * 2 * ND * NCT;    // current columns: [tile][half][component][thread]
constexpr int NLIN = 13 * HS;     // (stage, slot) linearisation items per half-phase: 208 <= 224 column threads
static_assert(NLIN <= NCT, "one linearisation item per column thread");

struct TileSmem {
    double2* sst; double* hval; double* tk; double* rho_inv; double* rq; int2* hctl; double* cur; double* errp; double* xn;
    unsigned bar_full, bar_done; volatile int* tile_done;
};

__device__ __forceinline__ TileSmem tile_smem(unsigned char* base, int t) {
    unsigned char* p = base + (size_t)t * TILE_BYTES;
    TileSmem s;
    s.sst = reinterpret_cast<double2*>(p); p += SST_BYTES;
    s.hval = reinterpret_cast<double*>(p); p += TS * sizeof(double);
    s.tk = reinterpret_cast<double*>(p); p += TS * sizeof(double);
    s.rho_inv = reinterpret_cast<double*>(p); p += TS * sizeof(double);
    s.rq = reinterpret_cast<double*>(p); p += TS * sizeof(double);
    s.hctl = reinterpret_cast<int2*>(p); p += TS * sizeof(int2);
    s.cur = reinterpret_cast<double*>(p); p += CUR_BYTES;
    s.errp = reinterpret_cast<double*>(p); p += ERR_BYTES;
    s.xn = reinterpret_cast<double*>(p);
    unsigned char* b = base + BAR_OFFSET + (size_t)t * BAR_BYTES;
    s.bar_full = smem_u32(b); s.bar_done = smem_u32(b + 8);
    s.tile_done = reinterpret_cast<volatile int*>(b + 16);
    return s;
}

// y = [r(0..2) v(3..5) m(6) lr(7..9) lv(10..12) lm(13)].  Stage derivatives kept: v', lr', lv' (Nystrom form for
// (r, v) as in K3) and m'; lm' only as running sums.
struct KStore { double kv[13][3], kl[13][3], km[13][3], kq[13]; double slm, elm, ela; };   // ela: klm_1 - klm_12 (robust estimate)

template <int J, bool SPLIT = false>
__device__ __forceinline__ void stage_input(const KStore& K, const double (&y)[ND], double h, double h2,
                                            double (&R)[3], double (&V)[3], double (&L)[3], double (&M)[3], double& Q) {
    if (J == 0) {
#pragma unroll
        for (int q = 0; q < 3; ++q) { R[q] = y[q]; V[q] = y[3 + q]; L[q] = y[7 + q]; M[q] = y[10 + q]; }
        Q = y[6];
        return;
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        // SPLIT (state warps): two partial sums per combination (even / odd stage index) halve the dependent-FMA depth of the chain
        double av = 0.0, ar = 0.0, al = 0.0, am = 0.0, bv = 0.0, br = 0.0, bl = 0.0, bm = 0.0;
#pragma unroll
        for (int l = 0; l < J; ++l) {
            if (SPLIT && (l & 1)) {
                if (lto_tab::Bf(J, l) != 0.0) {
                    bv = fma(lto_tab::Bf(J, l), K.kv[l][q], bv);
                    bl = fma(lto_tab::Bf(J, l), K.kl[l][q], bl);
                    bm = fma(lto_tab::Bf(J, l), K.km[l][q], bm);
                }
                if (lto_tab::Gf(J, l) != 0.0) br = fma(lto_tab::Gf(J, l), K.kv[l][q], br);
            } else {
                if (lto_tab::Bf(J, l) != 0.0) {
                    av = fma(lto_tab::Bf(J, l), K.kv[l][q], av);
                    al = fma(lto_tab::Bf(J, l), K.kl[l][q], al);
                    am = fma(lto_tab::Bf(J, l), K.km[l][q], am);
                }
                if (lto_tab::Gf(J, l) != 0.0) ar = fma(lto_tab::Gf(J, l), K.kv[l][q], ar);
            }
        }
        if (SPLIT) { av += bv; ar += br; al += bl; am += bm; }
        V[q] = fma(h, av, y[3 + q]);
        R[q] = fma(h2, ar, fma(h * lto_tab::Cf(J), y[3 + q], y[q]));
        L[q] = fma(h, al, y[7 + q]);
        M[q] = fma(h, am, y[10 + q]);
    }
    double aq = 0.0;
#pragma unroll
    for (int l = 0; l < J; ++l)
        if (lto_tab::Bf(J, l) != 0.0) aq = fma(lto_tab::Bf(J, l), K.kq[l], aq);
    Q = fma(h, aq, y[6]);
}

template <int J>
__device__ __forceinline__ void lm_accumulate(KStore& K, double klm) {
    if (J == 0) { K.slm = 0.0; K.elm = 0.0; K.ela = klm; }
    if (J == 11) K.ela -= klm;
    if (lto_tab::CHIf(J) != 0.0) K.slm = fma(lto_tab::CHIf(J), klm, K.slm);
    if (lto_tab::PSIf(J) != 0.0) K.elm = fma(lto_tab::PSIf(J), klm, K.elm);
}


// State-only step control (LTO_NORM_STATE: K4, and K3 when the columns are not in the norm): the embedded estimate e = ga + gb is the
// sum of the differences ga ~ (k1 - k12) and gb ~ (k11 - k13), which can cancel; the controller then takes max(|e|, |ga|, |gb|)
// per component (lto_prop_generic.cuh drive_rk8 `robust`).  A NaN estimate stays NaN.
__device__ __forceinline__ double rob_est(double e, double ga) {
    const double m = fmax(fabs(e), fmax(fabs(ga), fabs(e - ga)));
    return (e == e) ? m : e;
}

// 8th-order update (ode.jl:937) and, if ERR, the scaled squared error of the embedded estimate (ode.jl:940 with the
// controller's scaling atol + rtol*max(|y|, |ynew|)).
template <bool ERR, bool ROB = false>
__device__ __forceinline__ double step_finish(const KStore& K, const double (&y)[ND], double h, double h2, double atol, double rtol,
                                              double (&yn)[ND]) {
    double esum = 0.0;
    const double ce = h * lto_tab::ERRC, ce2 = h2 * lto_tab::ERRC;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double sv = 0.0, sr = 0.0, sl = 0.0, sm = 0.0;
#pragma unroll
        for (int l = 0; l < 13; ++l) {
            if (lto_tab::CHIf(l) != 0.0) {
                sv = fma(lto_tab::CHIf(l), K.kv[l][q], sv);
                sl = fma(lto_tab::CHIf(l), K.kl[l][q], sl);
                sm = fma(lto_tab::CHIf(l), K.km[l][q], sm);
            }
            if (lto_tab::CHIBf(l) != 0.0) sr = fma(lto_tab::CHIBf(l), K.kv[l][q], sr);
        }
        yn[q] = fma(h2, sr, fma(h, y[3 + q], y[q]));
        yn[3 + q] = fma(h, sv, y[3 + q]);
        yn[7 + q] = fma(h, sl, y[7 + q]);
        yn[10 + q] = fma(h, sm, y[10 + q]);
        if (ERR) {
            const int idx[4] = {q, 3 + q, 7 + q, 10 + q};
            double e[4];
            e[0] = ce2 * (K.kv[0][q] - K.kv[11][q]);                                          // psi^T B = e_1 - e_12
            e[1] = ce * ((K.kv[0][q] + K.kv[10][q]) - (K.kv[11][q] + K.kv[12][q]));
            e[2] = ce * ((K.kl[0][q] + K.kl[10][q]) - (K.kl[11][q] + K.kl[12][q]));
            e[3] = ce * ((K.km[0][q] + K.km[10][q]) - (K.km[11][q] + K.km[12][q]));
            if (ROB) {
                double gr = 0.0;                                                              // (k1 - k12) of r' = v: -h B[11] . kv
#pragma unroll
                for (int l = 0; l < 11; ++l)
                    if (lto_tab::Bf(11, l) != 0.0) gr = fma(lto_tab::Bf(11, l), K.kv[l][q], gr);
                e[0] = rob_est(e[0], -ce2 * gr);
                e[1] = rob_est(e[1], ce * (K.kv[0][q] - K.kv[11][q]));
                e[2] = rob_est(e[2], ce * (K.kl[0][q] - K.kl[11][q]));
                e[3] = rob_est(e[3], ce * (K.km[0][q] - K.km[11][q]));
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double sc = fma(rtol, fmax(fabs(y[idx[b]]), fabs(yn[idx[b]])), atol);
                const double r = e[b] * fast_rcp(sc);
                esum = fma(r, r, esum);
            }
        }
    }
    {
        double sq = 0.0;
#pragma unroll
        for (int l = 0; l < 13; ++l)
            if (lto_tab::CHIf(l) != 0.0) sq = fma(lto_tab::CHIf(l), K.kq[l], sq);
        yn[6] = fma(h, sq, y[6]);
        yn[13] = fma(h, K.slm, y[13]);
        if (ERR) {
            double eq = ce * ((K.kq[0] + K.kq[10]) - (K.kq[11] + K.kq[12]));
            double el = ce * K.elm;
            if (ROB) { eq = rob_est(eq, ce * (K.kq[0] - K.kq[11])); el = rob_est(el, ce * K.ela); }
            const double r1 = eq * fast_rcp(fma(rtol, fmax(fabs(y[6]), fabs(yn[6])), atol));
            const double r2 = el * fast_rcp(fma(rtol, fmax(fabs(y[13]), fabs(yn[13])), atol));
            esum = fma(r1, r1, esum);
            esum = fma(r2, r2, esum);
        }
    }
    return esum;
}

// ---------------------------------------------------------------------------
// Column thread: one attempted RK step of one STM column phi = [pr pv pm plr plv plm]  (lto_math.cuh sc_col<14>):
//   q   = lh . plv
//   kv  = U pr + C pv - uon plv + (cd q + cgm pm) lh
//   kq  = mm pm + cml q
//   kl  = -(W pr + U plv);   km = -plr - C^T plv
//   klm = clml q + lmm pm
// ---------------------------------------------------------------------------
__device__ __forceinline__ void sym3_mul_nacc(const double M[6], const double v[3], double out[3]) {
    out[0] = fma(-M[0], v[0], fma(-M[3], v[1], fma(-M[4], v[2], out[0])));
    out[1] = fma(-M[3], v[0], fma(-M[1], v[1], fma(-M[5], v[2], out[1])));
    out[2] = fma(-M[4], v[0], fma(-M[5], v[1], fma(-M[2], v[2], out[2])));
}

template <int J>
__device__ __forceinline__ void col_stage(KStore& K, const double (&p)[ND], double h, double h2, double w2, const double2* __restrict__ rec) {
    double R[3], V[3], L[3], M[3], Q;
    stage_input<J>(K, p, h, h2, R, V, L, M, Q);
    const double2* w = rec + J * NC2 * HS;
    double U[6], W[6];
    { const double2 a = w[0 * HS], b = w[1 * HS], c = w[2 * HS]; U[0] = a.x; U[1] = a.y; U[2] = b.x; U[3] = b.y; U[4] = c.x; U[5] = c.y; }
    { const double2 a = w[3 * HS], b = w[4 * HS], c = w[5 * HS]; W[0] = a.x; W[1] = a.y; W[2] = b.x; W[3] = b.y; W[4] = c.x; W[5] = c.y; }
    const double2 g0 = w[6 * HS], g1 = w[7 * HS], g2 = w[8 * HS], g3 = w[9 * HS], g4 = w[10 * HS];
    const double lh[3] = {g0.x, g0.y, g1.x};
    const double uon = g1.y, cd = g2.x, cgm = g2.y, cml = g3.x, clml = g3.y, mm = g4.x, lmm = g4.y;
    const double qd = fma(lh[0], M[0], fma(lh[1], M[1], lh[2] * M[2]));
    const double sc = fma(cd, qd, cgm * Q);
    double a[3] = {fma(sc, lh[0], w2 * V[1]), fma(sc, lh[1], -w2 * V[0]), sc * lh[2]};
#pragma unroll
    for (int q = 0; q < 3; ++q) a[q] = fma(-uon, M[q], a[q]);
    sym3_mul_acc(U, R, a);
    double b[3] = {0.0, 0.0, 0.0};
    sym3_mul_nacc(W, R, b);
    sym3_mul_nacc(U, M, b);
#pragma unroll
    for (int q = 0; q < 3; ++q) { K.kv[J][q] = a[q]; K.kl[J][q] = b[q]; }
    K.km[J][0] = fma(w2, M[1], -L[0]);
    K.km[J][1] = fma(-w2, M[0], -L[1]);
    K.km[J][2] = -L[2];
    K.kq[J] = fma(mm, Q, cml * qd);
    lm_accumulate<J>(K, fma(clml, qd, lmm * Q));
}

#define LTO_ICW14_LOCKSTEP() asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory")

template <bool ERR>
__device__ __forceinline__ double col_attempt(const double (&p)[ND], double h, double w2, const double2* __restrict__ rec,
                                              double atol, double rtol, double (&pn)[ND]) {
    const double h2 = h * h;
    KStore K;
    col_stage<0>(K, p, h, h2, w2, rec);  col_stage<1>(K, p, h, h2, w2, rec);  col_stage<2>(K, p, h, h2, w2, rec);
    col_stage<3>(K, p, h, h2, w2, rec);  col_stage<4>(K, p, h, h2, w2, rec);  col_stage<5>(K, p, h, h2, w2, rec);
    LTO_ICW14_LOCKSTEP();
    col_stage<6>(K, p, h, h2, w2, rec);  col_stage<7>(K, p, h, h2, w2, rec);  col_stage<8>(K, p, h, h2, w2, rec);
    LTO_ICW14_LOCKSTEP();
    col_stage<9>(K, p, h, h2, w2, rec);
    if (ERR) col_stage<10>(K, p, h, h2, w2, rec);
    else { K.kq[10] = 0.0; }
    LTO_ICW14_LOCKSTEP();
    col_stage<11>(K, p, h, h2, w2, rec); col_stage<12>(K, p, h, h2, w2, rec);
    return step_finish<ERR>(K, p, h, h2, atol, rtol, pn);
}

template <bool JOINT>
__device__ __forceinline__ void column_warp(const IndirectArgs& a, int cw, int lane, unsigned char* smem);

// ---------------------------------------------------------------------------
// State warp: the 14-dim right-hand side and its linearisation (lto_math.cuh sc_stage<14> is the reference
// formulation; same arithmetic, arranged without divisions).
// ---------------------------------------------------------------------------
struct LawConst { double tk, rho_inv, rho_inv_quarter; };   // per slot: thrustLimit*k (aL = tk/m, :33), 1/rho, 1/(4 rho)

template <bool LIN>
__device__ __forceinline__ void sc_eval(const double (&R)[3], const double (&V)[3], const double (&L)[3], const double (&M)[3], double Q,
                                        const SCConst& c, const LawConst& lw, double (&kv)[3], double (&kl)[3], double (&km)[3],
                                        double& kq, double& klm, double2* __restrict__ w) {
    const double w2 = 2.0 * c.omega;
    const double dx1 = R[0] + c.mu, dx2 = dx1 - 1.0;
    const double yz = fma(R[1], R[1], R[2] * R[2]);
    const double i1 = fast_rsqrt(fma(dx1, dx1, yz)), i2 = fast_rsqrt(fma(dx2, dx2, yz));
    const double i1s = i1 * i1, i2s = i2 * i2;
    const double a31 = c.m1 * i1s * i1, a32 = c.mu * i2s * i2;
    const double a51 = 3.0 * a31 * i1s, a52 = 3.0 * a32 * i2s;
    const double gg = -(a31 + a32), s5 = a51 + a52;
    const double p1 = a51 * dx1, p2 = a52 * dx2, t = p1 + p2;
    double U[6];
    U[0] = fma(p1, dx1, fma(p2, dx2, 1.0 + gg));
    U[1] = fma(s5 * R[1], R[1], 1.0 + gg);
    U[2] = fma(s5 * R[2], R[2], gg);
    U[3] = t * R[1]; U[4] = t * R[2]; U[5] = s5 * R[1] * R[2];
    // ---- control law (:36-64) with aL = thrustLimit k / m(t)
    const double im = fast_rcp(Q);
    const double aL = lw.tk * im;
    const double n2 = fma(M[0], M[0], fma(M[1], M[1], M[2] * M[2]));
    const bool dead = !(n2 > 0.0);
    const double in = dead ? 0.0 : fast_rsqrt(n2);
    const double n = n2 * in;
    double umag, dn = 0.0;
    bool prop = true;                                                  // umag proportional to aL -> d(umag)/dm = -umag/m
    if (c.p == 1.0) {
        const double y = fmin(fmax((n - 1.0) * lw.rho_inv, -700.0), 700.0);
        const double ey = exp(y);
        const double th = fma(-2.0, fast_rcp(ey + 1.0), 1.0);
        umag = fma(0.5 * aL, th, 0.5 * aL);
        dn = aL * lw.rho_inv_quarter * fma(-th, th, 1.0);
    } else if (c.p == 0.0) {
        umag = aL;
    } else {
        const double e = 1.0 / (c.p - 1.0);
        const double wv = (c.p == 2.0) ? 0.5 * n : pow(n / c.p, e);
        if (wv > aL) umag = aL;
        else { umag = wv; prop = false; dn = dead ? 0.0 : e * wv * in; }
    }
    if (dead) { umag = 0.0; dn = 0.0; }
    if (!(n2 == n2)) umag = n2;
    const double uon = umag * in;
    kv[0] = fma(-uon, M[0], fma(-a31, dx1, fma(-a32, dx2, fma(w2, V[1], R[0]))));
    kv[1] = fma(-uon, M[1], fma(gg, R[1], fma(-w2, V[0], R[1])));
    kv[2] = fma(-uon, M[2], gg * R[2]);
    kl[0] = -fma(U[0], M[0], fma(U[3], M[1], U[4] * M[2]));
    kl[1] = -fma(U[3], M[0], fma(U[1], M[1], U[5] * M[2]));
    kl[2] = -fma(U[4], M[0], fma(U[5], M[1], U[2] * M[2]));
    km[0] = fma(w2, M[1], -L[0]);
    km[1] = fma(-w2, M[0], -L[1]);
    km[2] = -L[2];
    kq = -c.cm * umag * Q;                                             // m'
    klm = -umag * n * im;                                              // lm' = (lv . u_acc)/m
    if (LIN) {
        const double dm = prop ? -umag * im : 0.0;                     // d(umag)/dm
        const double cd = uon - dn;
        const double l0 = M[0] * in, l1 = M[1] * in, l2 = M[2] * in;
        const double ylz = fma(R[1], M[1], R[2] * M[2]);
        const double e1 = a51 * fma(dx1, M[0], ylz), e2 = a52 * fma(dx2, M[0], ylz);
        const double h1 = -5.0 * e1 * i1s, h2 = -5.0 * e2 * i2s;
        const double ee = e1 + e2, hs = h1 + h2;
        const double hx = fma(h1, dx1, h2 * dx2);
        const double sM0 = s5 * M[0];
        w[0 * HS] = make_double2(U[0], U[1]);
        w[1 * HS] = make_double2(U[2], U[3]);
        w[2 * HS] = make_double2(U[4], U[5]);
        w[3 * HS] = make_double2(fma(h1 * dx1, dx1, fma(h2 * dx2, dx2, fma(2.0 * t, M[0], ee))),
                                 fma(hs * R[1], R[1], fma(2.0 * s5 * R[1], M[1], ee)));
        w[4 * HS] = make_double2(fma(hs * R[2], R[2], fma(2.0 * s5 * R[2], M[2], ee)),
                                 fma(hx, R[1], fma(t, M[1], sM0 * R[1])));
        w[5 * HS] = make_double2(fma(hx, R[2], fma(t, M[2], sM0 * R[2])),
                                 fma(hs * R[1], R[2], s5 * fma(R[1], M[2], M[1] * R[2])));
        w[6 * HS] = make_double2(l0, l1);
        w[7 * HS] = make_double2(l2, uon);
        w[8 * HS] = make_double2(cd, -dm);                                              // cd, cgm
        w[9 * HS] = make_double2(-c.cm * Q * dn, -fma(dn, n, umag) * im);               // cml, clml
        w[10 * HS] = make_double2(-c.cm * fma(Q, dm, umag), fma(-n, dm, umag * n * im) * im);   // mm, lmm
    }
}

struct Out11 { double v[11]; };
// The state warp evaluates the right-hand side only; the linearisation is built by the column threads (sc_lin below).
__device__ __noinline__ Out11 sc_eval_call(double r0, double r1, double r2, double v0, double v1, double v2, double l0, double l1, double l2,
                                           double m0, double m1, double m2, double q, double mu, double mu1, double omega, double pexp, double cm,
                                           double tk, double rho_inv, double rq) {
    const double R[3] = {r0, r1, r2}, V[3] = {v0, v1, v2}, L[3] = {l0, l1, l2}, M[3] = {m0, m1, m2};
    SCConst c; c.mu = mu; c.m1 = mu1; c.omega = omega; c.p = pexp; c.cm = cm;
    LawConst lw; lw.tk = tk; lw.rho_inv = rho_inv; lw.rho_inv_quarter = rq;
    double kv[3], kl[3], km[3], kq, klm;
    sc_eval<false>(R, V, L, M, q, c, lw, kv, kl, km, kq, klm, nullptr);
    Out11 o;
#pragma unroll
    for (int i = 0; i < 3; ++i) { o.v[i] = kv[i]; o.v[3 + i] = kl[i]; o.v[6 + i] = km[i]; }
    o.v[9] = kq; o.v[10] = klm;
    return o;
}

// One (stage, slot) item of the linearisation phase: stage state (r, lv, m) -> the 22-double record the column arithmetic uses.
__device__ __forceinline__ void sc_lin(const double2* __restrict__ sst, const SCConst& c, const LawConst& lw, double2* __restrict__ w) {
    const double2 a0 = sst[0 * TS], a1 = sst[1 * TS], a2 = sst[2 * TS], a3 = sst[3 * TS];
    const double R[3] = {a0.x, a0.y, a1.x}, M[3] = {a1.y, a2.x, a2.y}, Z[3] = {0.0, 0.0, 0.0};
    double kv[3], kl[3], km[3], kq, klm;                                // right-hand-side values: unused here, eliminated by the compiler
    sc_eval<true>(R, Z, Z, M, a3.x, c, lw, kv, kl, km, kq, klm, w);
}

template <bool JOINT>
__device__ __forceinline__ void column_warp(const IndirectArgs& a, int cw, int lane, unsigned char* smem) {
    const int col = 2 * cw + (lane >> 4);
    const int ct = cw * 32 + lane;
    const double w2 = 2.0 * a.c.omega;
    const double atol = a.cfg.atol, rtol = a.cfg.rtol;
    double* const cur_base = a.scratch + (size_t)blockIdx.x * SCRATCH_DOUBLES_PER_CTA + ct;   // current (last accepted) columns, L2 resident
    double2* const lin = reinterpret_cast<double2*>(smem + LIN_OFFSET);
    const int lstage = ct >> 4, ls16 = ct & (HS - 1);                   // this thread's (stage, slot) item of the linearisation phase
    unsigned alive = (1u << NTILE) - 1u;
    unsigned visit = 0;
    long long c_wait = 0, n_work = 0;
    const long long c_begin = clock64();
    while (alive) {
#pragma unroll 1
        for (int t = 0; t < NTILE; ++t) {
            if (!(alive & (1u << t))) continue;
            const TileSmem S = tile_smem(smem, t);
            const long long c0 = clock64();
            mbar_wait_parked(S.bar_full, visit & 1);
            c_wait += clock64() - c0;
            const bool done = *S.tile_done != 0;
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                const int slot = hf * HS + (lane & (HS - 1));
                const int2 hc = S.hctl[slot];
                const double h = S.hval[slot];
                double2* sc = reinterpret_cast<double2*>(S.cur + (size_t)slot * STASH_STRIDE + col * ND);   // candidate of the previous attempt (shared memory)
                double* sn = cur_base + (size_t)(t * 2 + hf) * ND * NCT;
                double p[ND];
                if (hc.x & F_ACCEPT) {                                 // an accepted step reads shared memory and refreshes the L2 copy
#pragma unroll
                    for (int i = 0; i < ND; i += 2) {
                        const double2 v = sc[i >> 1];
                        p[i] = v.x; p[i + 1] = v.y; __stcg(sn + i * NCT, v.x); __stcg(sn + (i + 1) * NCT, v.y);
                    }
                } else {                                               // a rejected one reloads the last accepted column
#pragma unroll
                    for (int i = 0; i < ND; ++i) p[i] = __ldcg(sn + i * NCT);
                }
                if (hc.x & F_STORE) {                                  // column `col` of ForwardDiff.jacobian(f, x0) (:121)
                    double* out = a.phi + (long long)hc.y * (ND * ND) + col * ND;
#pragma unroll
                    for (int i = 0; i < ND; i += 2)
                        asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(out + i), "d"(p[i]), "d"(p[i + 1]) : "memory");
                }
                if (hc.x & F_RESET) {
#pragma unroll
                    for (int i = 0; i < ND; ++i) { p[i] = (i == col) ? 1.0 : 0.0; __stcg(sn + i * NCT, p[i]); }
                }
                if (done) continue;
                // ---- linearisation phase: one (stage, slot) record per thread, straight from the state warp's stage states
                if (ct < NLIN) {
                    const int lslot = hf * HS + ls16;
                    LawConst lw; lw.tk = S.tk[lslot]; lw.rho_inv = S.rho_inv[lslot]; lw.rho_inv_quarter = S.rq[lslot];
                    sc_lin(S.sst + lstage * NSS * TS + lslot, a.c, lw, lin + lstage * NC2 * HS + ls16);
                }
                LTO_ICW14_LOCKSTEP();                                  // the half-phase's records are complete
                double pn[ND];
                const double es = col_attempt<JOINT>(p, h, w2, lin + (lane & (HS - 1)), atol, rtol, pn);
#pragma unroll
                for (int i = 0; i < ND; i += 2) sc[i >> 1] = make_double2(pn[i], pn[i + 1]);
                if (JOINT) S.errp[col * TS + slot] = es;
                LTO_ICW14_LOCKSTEP();                                  // all reads of the records done before the next phase rewrites them
            }
            if (done) alive &= ~(1u << t);
            else {
                fence_proxy_async();                                   // the stash may leave through the async proxy (state warp's bulk store)
                mbar_arrive(S.bar_done); n_work += 2;
            }
        }
        ++visit;
    }
    if (a.prof && lane == 0) {                                          // [CTA][8 warps][work, wait, count, alive]: warp 0 = state, 1..7 = columns
        unsigned long long* o = a.prof + ((size_t)blockIdx.x * NW + 1 + cw) * 4;
        const long long tot = clock64() - c_begin;
        o[0] = tot - c_wait; o[1] = c_wait; o[2] = n_work; o[3] = tot;
    }
}

template <int J>
__device__ __forceinline__ void state_stage(KStore& K, const double (&x)[ND], double h, double h2, const SCConst& c, const LawConst& lw,
                                            double2* __restrict__ sst) {
    double R[3], V[3], L[3], M[3], Q;
    stage_input<J, true>(K, x, h, h2, R, V, L, M, Q);
    if (sst) {                                                         // stage state for the column warps' linearisation phase
        double2* w = sst + J * NSS * TS;
        w[0 * TS] = make_double2(R[0], R[1]);
        w[1 * TS] = make_double2(R[2], M[0]);
        w[2 * TS] = make_double2(M[1], M[2]);
        w[3 * TS] = make_double2(Q, 0.0);
    }
    const Out11 o = sc_eval_call(R[0], R[1], R[2], V[0], V[1], V[2], L[0], L[1], L[2], M[0], M[1], M[2], Q, c.mu, c.m1, c.omega, c.p, c.cm,
                                 lw.tk, lw.rho_inv, lw.rho_inv_quarter);
#pragma unroll
    for (int q = 0; q < 3; ++q) { K.kv[J][q] = o.v[q]; K.kl[J][q] = o.v[3 + q]; K.km[J][q] = o.v[6 + q]; }
    K.kq[J] = o.v[9];
    lm_accumulate<J>(K, o.v[10]);
}

template <int J>
__device__ __forceinline__ void state_stage_s(KStore& K, const double* __restrict__ xs, double h, double h2, const SCConst& c, const LawConst& lw,
                                              double2* __restrict__ sst) {
    double x[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) x[i] = xs[i * TS];
    state_stage<J>(K, x, h, h2, c, lw, sst);
}

__device__ __forceinline__ double rms14(const double (&e)[ND], const double (&y)[ND], double atol, double rtol) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ND; ++i) { const double q = e[i] * fast_rcp(fma(rtol, fabs(y[i]), atol)); s = fma(q, q, s); }
    return sqrt(s * (1.0 / (double)ND));
}

// Hairer-Norsett-Wanner initial step over the state components (drive_rk8 in lto_prop_generic.cuh); K holds stage 0
__device__ __forceinline__ double initial_step(const KStore& K, const double (&x)[ND], double span, const SCConst& c, const LawConst& lw,
                                               double atol, double rtol) {
    double f0[ND], y1[ND];
#pragma unroll
    for (int q = 0; q < 3; ++q) { f0[q] = x[3 + q]; f0[3 + q] = K.kv[0][q]; f0[7 + q] = K.kl[0][q]; f0[10 + q] = K.km[0][q]; }
    f0[6] = K.kq[0]; f0[13] = K.elm;                                     // after stage 0: elm = psi_0 * klm_0 = klm_0
    const double d0 = rms14(x, x, atol, rtol), d1 = rms14(f0, x, atol, rtol);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
    h0 = fmin(h0, span);
#pragma unroll
    for (int i = 0; i < ND; ++i) y1[i] = fma(h0, f0[i], x[i]);
    {
        const double R1[3] = {y1[0], y1[1], y1[2]}, V1[3] = {y1[3], y1[4], y1[5]}, L1[3] = {y1[7], y1[8], y1[9]}, M1[3] = {y1[10], y1[11], y1[12]};
        const Out11 o = sc_eval_call(R1[0], R1[1], R1[2], V1[0], V1[1], V1[2], L1[0], L1[1], L1[2], M1[0], M1[1], M1[2], y1[6], c.mu, c.m1, c.omega,
                                     c.p, c.cm, lw.tk, lw.rho_inv, lw.rho_inv_quarter);
#pragma unroll
        for (int q = 0; q < 3; ++q) { y1[q] = V1[q] - f0[q]; y1[3 + q] = o.v[q] - f0[3 + q]; y1[7 + q] = o.v[3 + q] - f0[7 + q]; y1[10 + q] = o.v[6 + q] - f0[10 + q]; }
        y1[6] = o.v[9] - f0[6]; y1[13] = o.v[10] - f0[13];
    }
    const double d2 = rms14(y1, x, atol, rtol) / h0;
    const double dm = fmax(d1, d2);
    const double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : inv_eighth_root(dm / 0.01);
    return fmin(fmin(100.0 * h0, h1), span);
}

// Persistent per-slot controller state of one tile.  ONE state warp serves both tiles alternately: in steady state a tile's next
// attempt is computed while the column warps work on the other tile, so a second state warp would never run concurrently with
// the first (it only shortens the start-up) -- and with 8 warps instead of 9 every thread keeps 255 registers (the register
// file is split per SM sub-partition: a third warp on one sub-partition caps all threads at 168).
struct SlotCtl {
    double tcur, tf, h, span, esum, tk, rho_inv, rq;
    long long seg, ia;
    int na, nt, status, xi;
    bool active, lastrej, last, have;
    unsigned visit;
};

template <bool JOINT>
__device__ __forceinline__ void state_warp(const IndirectArgs& a, int lane, unsigned char* smem) {
    const int slot = lane;
    const unsigned fullmask = 0xffffffffu;
    const double atol = a.cfg.atol, rtol = a.cfg.rtol;
    const double inv_ne = JOINT ? 1.0 / (double)(ND * (ND + 1)) : 1.0 / (double)ND;
    SlotCtl ctl[NTILE];
#pragma unroll
    for (int t = 0; t < NTILE; ++t) {
        SlotCtl& c = ctl[t];
        c.tcur = 0.0; c.tf = 0.0; c.h = 0.0; c.span = 1.0; c.esum = 0.0; c.tk = 0.0; c.rho_inv = 1.0; c.rq = 0.25;
        c.seg = -1; c.ia = 0; c.na = 0; c.nt = 0; c.status = 0; c.xi = 0;
        c.active = false; c.lastrej = false; c.last = false; c.have = false; c.visit = 0;
        double* const xb = tile_smem(smem, t).xn + slot;
#pragma unroll
        for (int i = 0; i < ND; ++i) { xb[i * TS] = (i == 6) ? 1.0 : 0.0; xb[(ND + i) * TS] = (i == 6) ? 1.0 : 0.0; }
    }
    bool exhausted = false;
    const bool bulk = (reinterpret_cast<uintptr_t>(a.phi) & 15u) == 0;
    unsigned alive = (1u << NTILE) - 1u;
    long long c_wait = 0, c_work = 0, c_pre = 0, n_att = 0;
    const long long c_begin = clock64();
    while (alive) {
#pragma unroll 1
        for (int t = 0; t < NTILE; ++t) {
            if (!(alive & (1u << t))) continue;
            const TileSmem S = tile_smem(smem, t);
            SlotCtl c = ctl[t];
            double* const xbuf = S.xn + slot;
            double2* const rec = S.sst + slot;
            int flags = 0, store_seg = 0;
            bool finished = false;
            const long long c0 = clock64();
            long long c1 = c0;
            if (c.have) {
                mbar_wait_parked(S.bar_done, (c.visit - 1) & 1);
                c1 = clock64();
                c_wait += c1 - c0;
                if (c.active) {
                    double s2 = c.esum;
                    if (JOINT) {
#pragma unroll
                        for (int k = 0; k < ND; ++k) s2 += S.errp[k * TS + slot];
                    }
                    const double eest = sqrt(s2 * inv_ne);
                    if (!(eest == eest)) { c.status = LTO_ST_NAN; finished = true; }
                    else {
                        double q = (eest == 0.0) ? 5.0 : 0.9 * inv_eighth_root(eest);
                        q = fmin(5.0, fmax(0.2, q));
                        if (eest <= 1.0) {
                            ++c.na; flags |= F_ACCEPT;
                            c.xi ^= 1;
                            if (c.last) { c.tcur = c.tf; finished = true; }
                            else { c.tcur += c.h; if (c.lastrej) q = fmin(q, 1.0); c.lastrej = false; }
                        } else {
                            c.lastrej = true; q = fmin(q, 1.0);
                        }
                        c.h *= q;
                    }
                }
            }
            if (c.active && !finished) {
                if (c.h < c.span * 1e-12) { c.status = LTO_ST_HMIN; finished = true; }
                else if (c.nt >= a.cfg.max_attempts) { c.status = LTO_ST_MAXSTEPS; finished = true; }
            }
            if (c.active && finished) {
                bool nan = false;
                const double* xs = xbuf + c.xi * ND * TS;
#pragma unroll
                for (int i = 0; i < ND; ++i) {
                    const double xv = xs[i * TS];
                    nan |= !(xv == xv);
                    a.defect[c.seg * ND + i] = a.x_target ? xv - a.x_target[c.ia * ND + i] : xv;     // :82
                }
                if (nan && c.status == 0) c.status = LTO_ST_NAN;
                if (a.status) a.status[c.seg] = c.status;
                if (a.nsteps_out) { a.nsteps_out[2 * c.seg] = c.na; a.nsteps_out[2 * c.seg + 1] = c.nt; }
                // after an accepted last step the slot's stash IS the STM (column-major, :121): one bulk store; otherwise the column threads store
                if (bulk && (flags & F_ACCEPT)) bulk_store(a.phi + c.seg * (long long)(ND * ND), smem_u32(S.cur + (size_t)slot * STASH_STRIDE), ND * ND * sizeof(double));
                else flags |= F_STORE;
                store_seg = (int)c.seg;
                c.active = false;
            }
            bool fresh = false;
            if (!c.active && !exhausted) {
                const long long idx = (long long)atomicAdd(a.counter, 1ull);
                if (idx < a.n_seg) {
                    c.seg = idx; c.ia = lto_node_a(c.seg, a.npt);
                    const long long it = lto_traj_of(c.seg, a.npt);
#pragma unroll
                    for (int i = 0; i < ND; ++i) xbuf[(c.xi * ND + i) * TS] = a.x0[c.ia * ND + i];
                    c.tcur = a.t0[c.ia]; c.tf = a.t1[c.ia];
                    if (!(c.tcur < c.tf)) c.tf = c.tcur;
                    c.span = c.tf - c.tcur;
                    const double tl = a.thrustLimit_arr ? a.thrustLimit_arr[it] : a.c.thrustLimit;
                    const double rho = a.rho_arr ? a.rho_arr[it] : a.c.rho;
                    c.tk = tl * a.c.kthr;
                    c.rho_inv = 1.0 / rho;
                    c.rq = 0.25 / rho;
                    c.na = 0; c.nt = 0; c.status = 0; c.lastrej = false;
                    c.active = true; fresh = true; flags |= F_RESET;
                } else {
                    exhausted = true;
                }
            }
            if (!__any_sync(fullmask, c.active)) {
                S.hval[slot] = 0.0; S.hctl[slot] = make_int2(flags, store_seg);
                if (lane == 0) *S.tile_done = 1;
                mbar_arrive(S.bar_full);
                alive &= ~(1u << t);
                ctl[t] = c;
                continue;
            }
            LawConst lw; lw.tk = c.tk; lw.rho_inv = c.rho_inv; lw.rho_inv_quarter = c.rq;
            KStore K;
            const double* xs = xbuf + c.xi * ND * TS;
            const long long c2 = clock64();
            c_pre += c2 - c1;
            state_stage_s<0>(K, xs, 0.0, 0.0, a.c, lw, rec);
            if (__any_sync(fullmask, fresh)) {
                double x[ND];
#pragma unroll
                for (int i = 0; i < ND; ++i) x[i] = xs[i * TS];
                const double h1 = initial_step(K, x, c.span, a.c, lw, atol, rtol);
                if (fresh) c.h = h1;
            }
            c.last = false;
            if (c.tcur + c.h >= c.tf) { c.h = c.tf - c.tcur; c.last = true; }
            if (c.active) ++c.nt;
            const double h = c.h;
            S.hval[slot] = h; S.hctl[slot] = make_int2(flags | (c.active ? F_ACTIVE : 0), store_seg);
            S.tk[slot] = c.tk; S.rho_inv[slot] = c.rho_inv; S.rq[slot] = c.rq;
            const double h2 = h * h;
            state_stage_s<1>(K, xs, h, h2, a.c, lw, rec);  state_stage_s<2>(K, xs, h, h2, a.c, lw, rec);  state_stage_s<3>(K, xs, h, h2, a.c, lw, rec);
            state_stage_s<4>(K, xs, h, h2, a.c, lw, rec);  state_stage_s<5>(K, xs, h, h2, a.c, lw, rec);  state_stage_s<6>(K, xs, h, h2, a.c, lw, rec);
            state_stage_s<7>(K, xs, h, h2, a.c, lw, rec);  state_stage_s<8>(K, xs, h, h2, a.c, lw, rec);  state_stage_s<9>(K, xs, h, h2, a.c, lw, rec);
            state_stage_s<10>(K, xs, h, h2, a.c, lw, rec); state_stage_s<11>(K, xs, h, h2, a.c, lw, rec); state_stage_s<12>(K, xs, h, h2, a.c, lw, rec);
            {
                double x[ND], xn[ND];
#pragma unroll
                for (int i = 0; i < ND; ++i) x[i] = xs[i * TS];
                c.esum = step_finish<true, !JOINT>(K, x, h, h2, atol, rtol, xn);
                double* xc = xbuf + (c.xi ^ 1) * ND * TS;
#pragma unroll
                for (int i = 0; i < ND; ++i) xc[i * TS] = xn[i];
            }
            bulk_store_wait_read();                                     // the column warps overwrite the stash in this visit
            mbar_arrive(S.bar_full);
            c_work += clock64() - c2; ++n_att;
            c.have = true; ++c.visit;
            ctl[t] = c;
        }
    }
    bulk_store_wait_all();                                              // the last bulk stores must have completed before the CTA retires
    if (a.prof && lane == 0) {
        unsigned long long* o = a.prof + (size_t)blockIdx.x * NW * 4;
        o[0] = c_work; o[1] = c_wait; o[2] = n_att; o[3] = clock64() - c_begin;
        a.prof[(size_t)gridDim.x * NW * 4 + blockIdx.x] = c_pre;
    }
}

template <bool JOINT>
__global__ void __launch_bounds__(NTHREADS, 1) k_indirect_cw14(IndirectArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < NTILE) {
        const TileSmem S = tile_smem(smem_raw, threadIdx.x);
        mbar_init(S.bar_full, 32);
        mbar_init(S.bar_done, NCT);
        *S.tile_done = 0;
    }
    __syncthreads();
    // warps 0..6: column warps (sub-partitions 0,1,2,3,0,1,2); warp 7: the state warp (sub-partition 3, next to one column warp)
    if (warp == NCW) state_warp<JOINT>(a, lane, smem_raw);
    else column_warp<JOINT>(a, warp, lane, smem_raw);
}

// ---------------------------------------------------------------------------
// K4 (14-dim): defect-only, one lane per segment slot with the work queue; step control over the state.
// ---------------------------------------------------------------------------
constexpr int K4_THREADS = 128;

__global__ void __launch_bounds__(K4_THREADS, 2) k_indirect_state14(IndirectArgs a) {
    const unsigned fullmask = 0xffffffffu;
    double atol = a.cfg.atol, rtol = a.cfg.rtol;                          // per slot: state_tol_scale(p, rho) of the slot's segment
    double x[ND], xn[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) { x[i] = (i == 6) ? 1.0 : 0.0; xn[i] = x[i]; }
    double tcur = 0.0, tf = 0.0, h = 0.0, span = 1.0, esum = 0.0;
    LawConst lw; lw.tk = 0.0; lw.rho_inv = 1.0; lw.rho_inv_quarter = 0.25;
    long long seg = -1, ia = 0;
    int na = 0, nt = 0, status = 0;
    bool active = false, lastrej = false, last = false, have = false, exhausted = false;
    while (true) {
        bool finished = false;
        if (have && active) {
            const double eest = sqrt(esum * (1.0 / (double)ND));
            if (!(eest == eest)) { status = LTO_ST_NAN; finished = true; }
            else {
                double q = (eest == 0.0) ? 5.0 : 0.9 * inv_eighth_root(eest);
                q = fmin(5.0, fmax(0.2, q));
                if (eest <= 1.0) {
                    ++na;
#pragma unroll
                    for (int i = 0; i < ND; ++i) x[i] = xn[i];
                    if (last) { tcur = tf; finished = true; }
                    else { tcur += h; if (lastrej) q = fmin(q, 1.0); lastrej = false; }
                } else {
                    lastrej = true; q = fmin(q, 1.0);
                }
                h *= q;
            }
        }
        if (active && !finished) {
            if (h < span * 1e-12) { status = LTO_ST_HMIN; finished = true; }
            else if (nt >= a.cfg.max_attempts) { status = LTO_ST_MAXSTEPS; finished = true; }
        }
        if (active && finished) {
            bool nan = false;
#pragma unroll
            for (int i = 0; i < ND; ++i) nan |= !(x[i] == x[i]);
            if (nan && status == 0) status = LTO_ST_NAN;
#pragma unroll
            for (int i = 0; i < ND; ++i) a.defect[seg * ND + i] = a.x_target ? x[i] - a.x_target[ia * ND + i] : x[i];   // :82
            if (a.status) a.status[seg] = status;
            if (a.nsteps_out) { a.nsteps_out[2 * seg] = na; a.nsteps_out[2 * seg + 1] = nt; }
            active = false;
        }
        bool fresh = false;
        if (!active && !exhausted) {
            const long long idx = (long long)atomicAdd(a.counter, 1ull);
            if (idx < a.n_seg) {
                seg = idx; ia = lto_node_a(seg, a.npt);
                const long long it = lto_traj_of(seg, a.npt);
#pragma unroll
                for (int i = 0; i < ND; ++i) x[i] = a.x0[ia * ND + i];
                tcur = a.t0[ia]; tf = a.t1[ia];
                if (!(tcur < tf)) tf = tcur;
                span = tf - tcur;
                const double tl = a.thrustLimit_arr ? a.thrustLimit_arr[it] : a.c.thrustLimit;
                const double rho = a.rho_arr ? a.rho_arr[it] : a.c.rho;
                lw.tk = tl * a.c.kthr;
                lw.rho_inv = 1.0 / rho;
                lw.rho_inv_quarter = 0.25 / rho;
                { const double ts = state_tol_scale(a.c.p, rho); atol = a.cfg.atol * ts; rtol = a.cfg.rtol * ts; }
                na = 0; nt = 0; status = 0; lastrej = false;
                active = true; fresh = true;
            } else {
                exhausted = true;
            }
        }
        if (!__any_sync(fullmask, active)) break;
        KStore K;
        state_stage<0>(K, x, 0.0, 0.0, a.c, lw, nullptr);
        if (__any_sync(fullmask, fresh)) {
            const double h1 = initial_step(K, x, span, a.c, lw, atol, rtol);
            if (fresh) h = h1;
        }
        last = false;
        if (tcur + h >= tf) { h = tf - tcur; last = true; }
        if (active) ++nt;
        const double h2 = h * h;
        state_stage<1>(K, x, h, h2, a.c, lw, nullptr);  state_stage<2>(K, x, h, h2, a.c, lw, nullptr);  state_stage<3>(K, x, h, h2, a.c, lw, nullptr);
        state_stage<4>(K, x, h, h2, a.c, lw, nullptr);  state_stage<5>(K, x, h, h2, a.c, lw, nullptr);  state_stage<6>(K, x, h, h2, a.c, lw, nullptr);
        state_stage<7>(K, x, h, h2, a.c, lw, nullptr);  state_stage<8>(K, x, h, h2, a.c, lw, nullptr);  state_stage<9>(K, x, h, h2, a.c, lw, nullptr);
        state_stage<10>(K, x, h, h2, a.c, lw, nullptr); state_stage<11>(K, x, h, h2, a.c, lw, nullptr); state_stage<12>(K, x, h, h2, a.c, lw, nullptr);
        esum = step_finish<true, true>(K, x, h, h2, atol, rtol, xn);
        have = true;
    }
}

}  // namespace icw14

size_t indirect_cw14_scratch_bytes(int n_sm) { return (size_t)n_sm * icw14::SCRATCH_DOUBLES_PER_CTA * sizeof(double); }

template <bool JOINT>
static cudaError_t launch_icw14(const IndirectArgs& a, cudaStream_t st) {
    static int n_sm_dev[64] = {0};
    static bool attr_dev[64] = {false};
    int dev = 0;
    { cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e; }
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!attr_dev[dev]) {
        cudaError_t e = cudaDeviceGetAttribute(&n_sm_dev[dev], cudaDevAttrMultiProcessorCount, dev); if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(icw14::k_indirect_cw14<JOINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)icw14::SMEM);
        if (e != cudaSuccess) return e;
        attr_dev[dev] = true;
    }
    const int n_sm = n_sm_dev[dev];
    cudaError_t e = cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    const long long per_cta = (long long)icw14::NTILE * icw14::TS;
    const int grid = (int)std::min<long long>((a.n_seg + per_cta - 1) / per_cta, (long long)n_sm);
    icw14::k_indirect_cw14<JOINT><<<grid, icw14::NTHREADS, icw14::SMEM, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_indirect_cw14(const IndirectArgs& a, cudaStream_t st, int* n_launch) {
    *n_launch = 0;
    if (a.counter == nullptr || a.cfg.controller != 0 || a.n_seg <= 0 || a.n_seg > 0x7fffffffll) return cudaErrorNotSupported;
    if (a.phi == nullptr) {
        int dev = 0, n_sm = 0;
        cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), st); if (e != cudaSuccess) return e;
        const long long blocks = (a.n_seg + icw14::K4_THREADS - 1) / icw14::K4_THREADS;
        const int grid = (int)std::min<long long>(blocks, (long long)n_sm * 2);
        icw14::k_indirect_state14<<<grid, icw14::K4_THREADS, 0, st>>>(a);
        e = cudaGetLastError();
        if (e == cudaSuccess) *n_launch = 1;
        return e;
    }
    if (a.scratch == nullptr) return cudaErrorNotSupported;
    const cudaError_t e = (a.cfg.err_norm != 0) ? launch_icw14<true>(a, st) : launch_icw14<false>(a, st);
    if (e == cudaSuccess) *n_launch = 1;
    return e;
}

}  // namespace lto
