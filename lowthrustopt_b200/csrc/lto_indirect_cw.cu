// lto_indirect_cw.cu -- throughput kernel for the indirect method (K3), ndim = 12:
// defectCalc + jacobianCalc of multiShoot_CRTBP_indirect.jl:63-124 for a whole batch in one
// launch.  Each segment integrates [x | Phi] (12 + 144 components) with the adaptive
// order-8 pair and the OrdinaryDiffEq-style controller of lto_prop_generic.cuh
// (drive_rk8); Phi replaces ForwardDiff.jacobian(f, x0) (:121).
//
// Mapping:
//   tile        = 16 segment SLOTS; a CTA keeps NTILE tiles in flight
//   state warp  = one per tile; lane & 15 = slot.  Runs the nonlinear 12-dim system, the
//                 step-size controller and the per-slot work queue (a slot that finishes
//                 its segment pulls the next one from a global counter, so slots advance
//                 independently and no lane waits for the slowest segment of a batch).
//                 Publishes, per attempted step, the 13 stage linearisations U, W, G
//                 (18 doubles per stage and slot) + h + control flags in shared memory.
//   column warp = 6 per CTA, shared by all tiles; warp w carries STM columns 2w (lanes
//                 0-15) and 2w+1 (lanes 16-31) of the 16 slots, so the two half-warps read
//                 the same stage record (one shared-memory wavefront per 128-bit load).
//                 A thread owns one whole column (12 components): every RK combination
//                 and the structured product A*phi stay in its registers.
//   hand-off    = two mbarriers per tile (record full / columns done); the column warps
//                 cycle over the tiles, so a state warp's dependent chain for one tile is
//                 hidden behind the column work of the others.
// Step control uses the joint norm over x and Phi (LTO_NORM_STATE_SENS, the ForwardDiff
// semantics) or x alone: each column thread returns its partial sum of squared scaled
// errors and the state warp decides.  Between visits a column's current and candidate
// values live in a shared-memory stash, so rejecting a step costs nothing extra.
// The initial step is Hairer's estimate over the state components (the generic kernel
// and the oracle take it over x and Phi): the two paths may choose different step
// sequences and agree to the integration tolerance, not to rounding.
#include "lto_internal.h"
#include "lto_cw_common.cuh"
#include <algorithm>
#include <cstdlib>

namespace lto {
namespace icw {

using namespace cwc;

constexpr int ND = 12;
constexpr int TS = 16;            // segment slots per tile
constexpr int NCW = 6;            // column warps
constexpr int NCT = 32 * NCW;     // column threads
constexpr int NC2 = 9;            // double2 per stage record: U[6] W[6] G[6]

enum { F_ACCEPT = 1, F_STORE = 2, F_RESET = 4, F_ACTIVE = 8 };

template <int NTILE>
struct Cfg {
    static constexpr int NW = NTILE + NCW;
    static constexpr int NTHREADS = 32 * NW;
    static constexpr size_t REC_BYTES = (size_t)13 * NC2 * TS * sizeof(double2);      // stage records
    static constexpr size_t HDR_BYTES = (size_t)TS * (sizeof(double) + sizeof(int2)); // h, {flags, segment}
    static constexpr size_t STASH_BYTES = (size_t)2 * ND * NCT * sizeof(double);      // current / candidate columns
    static constexpr size_t ERR_BYTES = (size_t)ND * TS * sizeof(double);             // error partials per column and slot
    static constexpr size_t TILE_BYTES = REC_BYTES + HDR_BYTES + STASH_BYTES + ERR_BYTES;
    static constexpr size_t SMEM = NTILE * TILE_BYTES + NTILE * 32;                   // + {full, done, tile_done} per tile
};

struct TileSmem {
    double2* rec; double* hval; int2* hctl; double* stash; double* errp;
    unsigned bar_full, bar_done; volatile int* tile_done;
};

template <int NTILE>
__device__ __forceinline__ TileSmem tile_smem(unsigned char* base, int t) {
    typedef Cfg<NTILE> C;
    unsigned char* p = base + (size_t)t * C::TILE_BYTES;
    TileSmem s;
    s.rec = reinterpret_cast<double2*>(p); p += C::REC_BYTES;
    s.hval = reinterpret_cast<double*>(p); p += TS * sizeof(double);
    s.hctl = reinterpret_cast<int2*>(p); p += TS * sizeof(int2);
    s.stash = reinterpret_cast<double*>(p); p += C::STASH_BYTES;
    s.errp = reinterpret_cast<double*>(p);
    unsigned char* b = base + (size_t)NTILE * C::TILE_BYTES + (size_t)t * 32;
    s.bar_full = smem_u32(b); s.bar_done = smem_u32(b + 8);
    s.tile_done = reinterpret_cast<volatile int*>(b + 16);
    return s;
}

// ---------------------------------------------------------------------------
// Stage inputs of one 12-vector y = [r v lr lv] in Nystrom form for the (r, v) pair:
// only v', lr', lv' of every stage are kept (kv, kl, km); positions are rebuilt with
// G = B*B (lto_tableau.h).
// ---------------------------------------------------------------------------
struct KStore { double kv[13][3], kl[13][3], km[13][3]; };

template <int J>
__device__ __forceinline__ void stage_input(const KStore& K, const double (&y)[ND], double h, double h2,
                                            double (&R)[3], double (&V)[3], double (&L)[3], double (&M)[3]) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        if (J == 0) { R[q] = y[q]; V[q] = y[3 + q]; L[q] = y[6 + q]; M[q] = y[9 + q]; continue; }
        double av = 0.0, ar = 0.0, al = 0.0, am = 0.0;
#pragma unroll
        for (int l = 0; l < J; ++l) {
            if (lto_tab::Bf(J, l) != 0.0) {
                av = fma(lto_tab::Bf(J, l), K.kv[l][q], av);
                al = fma(lto_tab::Bf(J, l), K.kl[l][q], al);
                am = fma(lto_tab::Bf(J, l), K.km[l][q], am);
            }
            if (lto_tab::Gf(J, l) != 0.0) ar = fma(lto_tab::Gf(J, l), K.kv[l][q], ar);
        }
        V[q] = fma(h, av, y[3 + q]);
        R[q] = fma(h2, ar, fma(h * lto_tab::Cf(J), y[3 + q], y[q]));
        L[q] = fma(h, al, y[6 + q]);
        M[q] = fma(h, am, y[9 + q]);
    }
}

// 8th-order update (ode.jl:937) and the scaled squared error of the embedded estimate
// (ode.jl:940 with the controller's scaling atol + rtol*max(|y|, |ynew|)).
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
        yn[6 + q] = fma(h, sl, y[6 + q]);
        yn[9 + q] = fma(h, sm, y[9 + q]);
        double e[4];
        e[0] = ce2 * (K.kv[0][q] - K.kv[11][q]);                                              // psi^T B = e_1 - e_12
        e[1] = ce * ((K.kv[0][q] + K.kv[10][q]) - (K.kv[11][q] + K.kv[12][q]));
        e[2] = ce * ((K.kl[0][q] + K.kl[10][q]) - (K.kl[11][q] + K.kl[12][q]));
        e[3] = ce * ((K.km[0][q] + K.km[10][q]) - (K.km[11][q] + K.km[12][q]));
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double sc = fma(rtol, fmax(fabs(y[3 * b + q]), fabs(yn[3 * b + q])), atol);
            const double r = e[b] * fast_rcp(sc);
            esum = fma(r, r, esum);
        }
    }
    return esum;
}

// ---------------------------------------------------------------------------
// Column thread: one attempted RK step of one STM column phi = [pr pv plr plv].
//   kv = U pr + C pv + G plv;  kl = -(W pr + U plv);  km = -plr - C^T plv   (lto_math.cuh sc_col)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void sym3_mul_nacc(const double M[6], const double v[3], double out[3]) {
    out[0] = fma(-M[0], v[0], fma(-M[3], v[1], fma(-M[4], v[2], out[0])));
    out[1] = fma(-M[3], v[0], fma(-M[1], v[1], fma(-M[5], v[2], out[1])));
    out[2] = fma(-M[4], v[0], fma(-M[5], v[1], fma(-M[2], v[2], out[2])));
}

template <int J>
__device__ __forceinline__ void col_stage(KStore& K, const double (&p)[ND], double h, double h2, double w2, const double2* __restrict__ rec) {
    double R[3], V[3], L[3], M[3];
    stage_input<J>(K, p, h, h2, R, V, L, M);
    const double2* w = rec + J * NC2 * TS;
    double U[6], W[6], G[6];
    { const double2 a = w[0 * TS], b = w[1 * TS], c = w[2 * TS]; U[0] = a.x; U[1] = a.y; U[2] = b.x; U[3] = b.y; U[4] = c.x; U[5] = c.y; }
    { const double2 a = w[3 * TS], b = w[4 * TS], c = w[5 * TS]; W[0] = a.x; W[1] = a.y; W[2] = b.x; W[3] = b.y; W[4] = c.x; W[5] = c.y; }
    { const double2 a = w[6 * TS], b = w[7 * TS], c = w[8 * TS]; G[0] = a.x; G[1] = a.y; G[2] = b.x; G[3] = b.y; G[4] = c.x; G[5] = c.y; }
    double a[3] = {w2 * V[1], -w2 * V[0], 0.0};
    sym3_mul_acc(U, R, a);
    sym3_mul_acc(G, M, a);
    double b[3] = {0.0, 0.0, 0.0};
    sym3_mul_nacc(W, R, b);
    sym3_mul_nacc(U, M, b);
#pragma unroll
    for (int q = 0; q < 3; ++q) { K.kv[J][q] = a[q]; K.kl[J][q] = b[q]; }
    K.km[J][0] = fma(w2, M[1], -L[0]);
    K.km[J][1] = fma(-w2, M[0], -L[1]);
    K.km[J][2] = -L[2];
}

__device__ __forceinline__ double col_attempt(const double (&p)[ND], double h, double w2, const double2* __restrict__ rec,
                                              double atol, double rtol, double (&pn)[ND]) {
    const double h2 = h * h;
    KStore K;
    col_stage<0>(K, p, h, h2, w2, rec);  col_stage<1>(K, p, h, h2, w2, rec);  col_stage<2>(K, p, h, h2, w2, rec);
    col_stage<3>(K, p, h, h2, w2, rec);  col_stage<4>(K, p, h, h2, w2, rec);  col_stage<5>(K, p, h, h2, w2, rec);
    col_stage<6>(K, p, h, h2, w2, rec);  col_stage<7>(K, p, h, h2, w2, rec);  col_stage<8>(K, p, h, h2, w2, rec);
    col_stage<9>(K, p, h, h2, w2, rec);  col_stage<10>(K, p, h, h2, w2, rec); col_stage<11>(K, p, h, h2, w2, rec);
    col_stage<12>(K, p, h, h2, w2, rec);
    return step_finish(K, p, h, h2, atol, rtol, pn);
}

template <int NTILE>
__device__ __forceinline__ void column_warp(const IndirectArgs& a, int cw, int lane, unsigned char* smem) {
    const int slot = lane & (TS - 1);
    const int col = 2 * cw + (lane >> 4);
    const int ct = cw * 32 + lane;
    const double w2 = 2.0 * a.c.omega;
    const double atol = a.cfg.atol, rtol = a.cfg.rtol;
    unsigned alive = (1u << NTILE) - 1u, cur = 0u;        // cur bit t: which stash buffer holds the current column of tile t
    unsigned visit = 0;
    while (alive) {
#pragma unroll
        for (int t = 0; t < NTILE; ++t) {
            if (!(alive & (1u << t))) continue;
            const TileSmem S = tile_smem<NTILE>(smem, t);
            mbar_wait(S.bar_full, visit & 1);
            const int2 hc = S.hctl[slot];
            const double h = S.hval[slot];
            if (hc.x & F_ACCEPT) cur ^= (1u << t);
            const unsigned cb = (cur >> t) & 1u;
            double* sc = S.stash + (size_t)cb * ND * NCT + ct;
            double* sn = S.stash + (size_t)(cb ^ 1u) * ND * NCT + ct;
            double p[ND];
#pragma unroll
            for (int i = 0; i < ND; ++i) p[i] = sc[i * NCT];
            if (hc.x & F_STORE) {                                      // column `col` of ForwardDiff.jacobian(f, x0) (:121)
                double* out = a.phi + (long long)hc.y * (ND * ND) + col * ND;
#pragma unroll
                for (int i = 0; i < ND; ++i) out[i] = p[i];
            }
            if (hc.x & F_RESET) {
#pragma unroll
                for (int i = 0; i < ND; ++i) { p[i] = (i == col) ? 1.0 : 0.0; sc[i * NCT] = p[i]; }
            }
            if (*S.tile_done) { alive &= ~(1u << t); continue; }
            double pn[ND];
            const double es = col_attempt(p, h, w2, S.rec + slot, atol, rtol, pn);
#pragma unroll
            for (int i = 0; i < ND; ++i) sn[i * NCT] = pn[i];
            S.errp[col * TS + slot] = es;
            mbar_arrive(S.bar_done);
        }
        ++visit;
    }
}

// ---------------------------------------------------------------------------
// State warp.
// ---------------------------------------------------------------------------
template <int J, bool REC>
__device__ __forceinline__ int state_stage(KStore& K, const double (&x)[ND], double h, double h2, const SCConst& c, double tl, double rho,
                                           double2* __restrict__ rec, bool writer) {
    double s[ND], f[ND];
    {
        double R[3], V[3], L[3], M[3];
        stage_input<J>(K, x, h, h2, R, V, L, M);
#pragma unroll
        for (int q = 0; q < 3; ++q) { s[q] = R[q]; s[3 + q] = V[q]; s[6 + q] = L[q]; s[9 + q] = M[q]; }
    }
    SCStage st;
    const int bad = sc_stage<ND>(s, c, tl, rho, f, st);
#pragma unroll
    for (int q = 0; q < 3; ++q) { K.kv[J][q] = f[3 + q]; K.kl[J][q] = f[6 + q]; K.km[J][q] = f[9 + q]; }
    if (REC && writer) {
        double2* w = rec + J * NC2 * TS;
        w[0 * TS] = make_double2(st.U[0], st.U[1]); w[1 * TS] = make_double2(st.U[2], st.U[3]); w[2 * TS] = make_double2(st.U[4], st.U[5]);
        w[3 * TS] = make_double2(st.W[0], st.W[1]); w[4 * TS] = make_double2(st.W[2], st.W[3]); w[5 * TS] = make_double2(st.W[4], st.W[5]);
        w[6 * TS] = make_double2(st.G[0], st.G[1]); w[7 * TS] = make_double2(st.G[2], st.G[3]); w[8 * TS] = make_double2(st.G[4], st.G[5]);
    }
    return bad;
}

__device__ __forceinline__ double rms12(const double (&e)[ND], const double (&y)[ND], double atol, double rtol) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ND; ++i) { const double q = e[i] / fma(rtol, fabs(y[i]), atol); s = fma(q, q, s); }
    return sqrt(s / (double)ND);
}

template <int NTILE>
__device__ __forceinline__ void state_warp(const IndirectArgs& a, int t, int lane, unsigned char* smem) {
    const TileSmem S = tile_smem<NTILE>(smem, t);
    const int slot = lane & (TS - 1);
    const bool writer = lane < TS;                 // lanes 16-31 mirror lanes 0-15 (same values, no stores)
    const unsigned fullmask = 0xffffffffu;
    const double atol = a.cfg.atol, rtol = a.cfg.rtol;
    const bool joint = (a.phi != nullptr) && (a.cfg.err_norm != 0);
    const double inv_ne = joint ? 1.0 / (double)(ND * (ND + 1)) : 1.0 / (double)ND;
    double x[ND], xn[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) { x[i] = 0.0; xn[i] = 0.0; }
    double tcur = 0.0, tf = 0.0, h = 0.0, span = 1.0, tl = a.c.thrustLimit, rho = a.c.rho, esum = 0.0;
    long long seg = -1, ia = 0;
    int na = 0, nt = 0, status = 0;
    bool active = false, lastrej = false, last = false, have = false, exhausted = false;
    unsigned visit = 0;
    while (true) {
        int flags = 0, store_seg = 0;
        bool finished = false;
        if (have) {
            mbar_wait(S.bar_done, (visit - 1) & 1);
            if (active) {
                double s2 = esum;
                if (joint) {
#pragma unroll
                    for (int c = 0; c < ND; ++c) s2 += S.errp[c * TS + slot];
                }
                const double eest = sqrt(s2 * inv_ne);
                if (!(eest == eest)) { status = LTO_ST_NAN; finished = true; }
                else {
                    double q = (eest == 0.0) ? 5.0 : 0.9 * inv_eighth_root(eest);
                    q = fmin(5.0, fmax(0.2, q));
                    if (eest <= 1.0) {
                        ++na; flags |= F_ACCEPT;
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
        }
        if (active && !finished) {                                       // drive_rk8's loop-top checks
            if (h < span * 1e-12) { status = LTO_ST_HMIN; finished = true; }
            else if (nt >= a.cfg.max_attempts) { status = LTO_ST_MAXSTEPS; finished = true; }
        }
        if (active && finished) {
            // ---- defect = x(t1) - XC_all[:, i+1] (multiShoot_CRTBP_indirect.jl:82)
            bool nan = false;
#pragma unroll
            for (int i = 0; i < ND; ++i) nan |= !(x[i] == x[i]);
            if (nan && status == 0) status = LTO_ST_NAN;
            if (writer) {
#pragma unroll
                for (int i = 0; i < ND; ++i) a.defect[seg * ND + i] = a.x_target ? x[i] - a.x_target[ia * ND + i] : x[i];
                if (a.status) a.status[seg] = status;
                if (a.nsteps_out) { a.nsteps_out[2 * seg] = na; a.nsteps_out[2 * seg + 1] = nt; }
            }
            flags |= F_STORE; store_seg = (int)seg;
            active = false;
        }
        bool fresh = false;
        {
            const bool want = !active && !exhausted;
            long long idx = -1;
            if (writer && want) idx = (long long)atomicAdd(a.counter, 1ull);
            idx = __shfl_sync(fullmask, idx, slot);
            if (want) {
                if (idx < a.n_seg) {
                    seg = idx; ia = lto_node_a(seg, a.npt);
                    const long long it = lto_traj_of(seg, a.npt);
#pragma unroll
                    for (int i = 0; i < ND; ++i) x[i] = a.x0[ia * ND + i];
                    tcur = a.t0[ia]; tf = a.t1[ia];
                    if (!(tcur < tf)) tf = tcur;                          // empty span: one zero-length step, Phi = I
                    span = tf - tcur;
                    tl = a.thrustLimit_arr ? a.thrustLimit_arr[it] : a.c.thrustLimit;
                    rho = a.rho_arr ? a.rho_arr[it] : a.c.rho;
                    na = 0; nt = 0; status = 0; lastrej = false;
                    active = true; fresh = true; flags |= F_RESET;
                } else {
                    exhausted = true;
                }
            }
        }
        const bool any_active = __any_sync(fullmask, active);
        if (!any_active) {
            if (writer) { S.hval[slot] = 0.0; S.hctl[slot] = make_int2(flags, store_seg); }
            if (lane == 0) *S.tile_done = 1;
            mbar_arrive(S.bar_full);
            break;
        }
        // ---- one attempted step (13 stages); a fresh slot first picks its initial step
        KStore K;
        int bad = state_stage<0, true>(K, x, 0.0, 0.0, a.c, tl, rho, S.rec + slot, writer);
        if (__any_sync(fullmask, fresh)) {
            // Hairer-Norsett-Wanner initial step over the state components (drive_rk8 in lto_prop_generic.cuh)
            double f0[ND], f1[ND], y1[ND];
#pragma unroll
            for (int q = 0; q < 3; ++q) { f0[q] = x[3 + q]; f0[3 + q] = K.kv[0][q]; f0[6 + q] = K.kl[0][q]; f0[9 + q] = K.km[0][q]; }
            const double d0 = rms12(x, x, atol, rtol), d1 = rms12(f0, x, atol, rtol);
            double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
            h0 = fmin(h0, span);
#pragma unroll
            for (int i = 0; i < ND; ++i) y1[i] = fma(h0, f0[i], x[i]);
            SCStage st1;
            bad |= sc_stage<ND>(y1, a.c, tl, rho, f1, st1);
#pragma unroll
            for (int i = 0; i < ND; ++i) f1[i] -= f0[i];
            const double d2 = rms12(f1, x, atol, rtol) / h0;
            const double dm = fmax(d1, d2);
            const double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : inv_eighth_root(dm / 0.01);
            if (fresh) h = fmin(fmin(100.0 * h0, h1), span);
        }
        last = false;
        if (tcur + h >= tf) { h = tf - tcur; last = true; }
        if (active) ++nt;
        const double h2 = h * h;
        bad |= state_stage<1, true>(K, x, h, h2, a.c, tl, rho, S.rec + slot, writer);
        bad |= state_stage<2, true>(K, x, h, h2, a.c, tl, rho, S.rec + slot, writer);
        bad |= state_stage<3, true>(K, x, h, h2, a.c, tl, rho, S.rec + slot, writer);
        bad |= state_stage<4, true>(K, x, h, h2, a.c, tl, rho, S.rec + slot, writer);
        bad |= state_stage<5, true>(K, x, h, h2, a.c, tl, rho, S.rec + slot, writer);
        bad |= state_stage<6, true>(K, x, h, h2, a.c, tl, rho, S.rec + slot, writer);
        bad |= state_stage<7, true>(K, x, h, h2, a.c, tl, rho, S.rec + slot, writer);
        bad |= state_stage<8, true>(K, x, h, h2, a.c, tl, rho, S.rec + slot, writer);
        bad |= state_stage<9, true>(K, x, h, h2, a.c, tl, rho, S.rec + slot, writer);
        bad |= state_stage<10, true>(K, x, h, h2, a.c, tl, rho, S.rec + slot, writer);
        bad |= state_stage<11, true>(K, x, h, h2, a.c, tl, rho, S.rec + slot, writer);
        bad |= state_stage<12, true>(K, x, h, h2, a.c, tl, rho, S.rec + slot, writer);
        esum = step_finish(K, x, h, h2, atol, rtol, xn);
        if (active && bad) status = LTO_ST_BADP;
        if (writer) { S.hval[slot] = h; S.hctl[slot] = make_int2(flags | (active ? F_ACTIVE : 0), store_seg); }
        mbar_arrive(S.bar_full);
        have = true; ++visit;
    }
}

template <int NTILE>
__global__ void __launch_bounds__(Cfg<NTILE>::NTHREADS, 1) k_indirect_cw(IndirectArgs a) {
    typedef Cfg<NTILE> C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < NTILE) {
        const TileSmem S = tile_smem<NTILE>(smem_raw, threadIdx.x);
        mbar_init(S.bar_full, 32);
        mbar_init(S.bar_done, NCT);
        *S.tile_done = 0;
    }
    __syncthreads();
    if (warp < NTILE) state_warp<NTILE>(a, warp, lane, smem_raw);
    else column_warp<NTILE>(a, warp - NTILE, lane, smem_raw);
}

}  // namespace icw

template <int NTILE>
static cudaError_t launch_icw(const IndirectArgs& a, cudaStream_t st) {
    typedef icw::Cfg<NTILE> C;
    static int n_sm = 0;
    static bool attr = false;
    if (!attr) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(icw::k_indirect_cw<NTILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    cudaError_t e = cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    const long long per_cta = (long long)NTILE * icw::TS;
    const int grid = (int)std::min<long long>((a.n_seg + per_cta - 1) / per_cta, (long long)n_sm);
    icw::k_indirect_cw<NTILE><<<grid, C::NTHREADS, C::SMEM, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_indirect_cw(const IndirectArgs& a, int ndim, cudaStream_t st, int* n_launch) {
    *n_launch = 0;
    if (ndim != 12 || a.phi == nullptr || a.counter == nullptr || a.cfg.controller != 0 || a.n_seg <= 0 || a.n_seg > 0x7fffffffll)
        return cudaErrorNotSupported;
    static int ntile = 0;
    if (!ntile) { const char* v = getenv("LTO_ICW_NTILE"); ntile = (v && v[0] == '3') ? 3 : 2; }
    cudaError_t e = (ntile == 2) ? launch_icw<2>(a, st) : launch_icw<3>(a, st);
    if (e == cudaSuccess) *n_launch = 1;
    return e;
}

}  // namespace lto
