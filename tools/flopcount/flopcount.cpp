// flopcount.cpp -- the authoritative algorithmic-FLOP figures of bench.py's roofline ("the oracle's
// FLOP-counting scalar type produces the authoritative number").
//
// The kernels' own __host__ __device__ arithmetic (lowthrustopt_b200/csrc/lto_math.cuh, lto_prop_generic.cuh, lto_hc_math.cuh --
// the sparsity-exploiting formulations: structural zeros of A and of the tableau skipped, base-state right-hand side once per
// segment-stage) is compiled here with `double` replaced by a scalar that tallies every arithmetic operation it performs:
//     add / sub / mul / div / sqrt = 1,   fma = 2,   exp / tanh / pow = 1 each (also listed separately),
//     comparisons, fabs, fmax / fmin, negation, selects = 0
// (the rule of SURVEY 8(d)).  The counted runs are checked against plain-double runs of the same code (tests/test_flopcount.py),
// so what is counted is what computes the right answer.  Output: one JSON object on stdout (tools/flopcount/count.py writes it
// to profiles/flops_per_unit.json, which bench.py reads).
#include <math.h>
#include <cstdio>
#include <cstring>
#include <cstdint>
#include "../../lowthrustopt_b200/csrc/lto_tableau.h"      // stays in plain double (pragma once)

struct Tally { long long add, mul, div, sqrt_, fma_, special; };
static Tally g_t;
static inline long long flops(const Tally& t) { return t.add + t.mul + t.div + t.sqrt_ + 2 * t.fma_ + t.special; }

struct Counted {
    double v;
    Counted() : v(0.0) {}
    Counted(double x) : v(x) {}
    Counted(int x) : v((double)x) {}
    explicit operator bool() const { return v != 0.0; }
};
static inline Counted operator+(Counted a, Counted b) { ++g_t.add; return Counted(a.v + b.v); }
static inline Counted operator-(Counted a, Counted b) { ++g_t.add; return Counted(a.v - b.v); }
static inline Counted operator*(Counted a, Counted b) { ++g_t.mul; return Counted(a.v * b.v); }
static inline Counted operator/(Counted a, Counted b) { ++g_t.div; return Counted(a.v / b.v); }
static inline Counted operator-(Counted a) { return Counted(-a.v); }
static inline Counted& operator+=(Counted& a, Counted b) { a = a + b; return a; }
static inline Counted& operator-=(Counted& a, Counted b) { a = a - b; return a; }
static inline Counted& operator*=(Counted& a, Counted b) { a = a * b; return a; }
static inline bool operator<(Counted a, Counted b) { return a.v < b.v; }
static inline bool operator>(Counted a, Counted b) { return a.v > b.v; }
static inline bool operator<=(Counted a, Counted b) { return a.v <= b.v; }
static inline bool operator>=(Counted a, Counted b) { return a.v >= b.v; }
static inline bool operator==(Counted a, Counted b) { return a.v == b.v; }
static inline bool operator!=(Counted a, Counted b) { return a.v != b.v; }
static inline Counted fma(Counted a, Counted b, Counted c) { ++g_t.fma_; return Counted(::fma(a.v, b.v, c.v)); }
static inline Counted sqrt(Counted a) { ++g_t.sqrt_; return Counted(::sqrt(a.v)); }
static inline Counted fabs(Counted a) { return Counted(::fabs(a.v)); }
static inline Counted fmax(Counted a, Counted b) { return Counted(::fmax(a.v, b.v)); }
static inline Counted fmin(Counted a, Counted b) { return Counted(::fmin(a.v, b.v)); }
static inline Counted exp(Counted a) { ++g_t.special; return Counted(::exp(a.v)); }
static inline Counted tanh(Counted a) { ++g_t.special; return Counted(::tanh(a.v)); }
static inline Counted pow(Counted a, Counted b) { ++g_t.special; return Counted(::pow(a.v, b.v)); }

#ifdef LTO_FLOPCOUNT_PLAIN
typedef double real;
#else
#define double Counted
#endif
#include "../../lowthrustopt_b200/csrc/lto_prop_generic.cuh"
#include "../../lowthrustopt_b200/csrc/lto_hc_math.cuh"
#ifndef LTO_FLOPCOUNT_PLAIN
typedef double real;                     // = Counted
#undef double
#endif

using namespace lto;

static inline double val(double x) { return x; }
static inline double val(const Counted& x) { return x.v; }

static const double MU = 0.012150585609624037, DU = 384747.96285603708, TU = 375699.81732246041;

static EPConst make_ep() { EPConst c; c.mu = MU; c.m1 = 1.0 - MU; c.kthr = TU * TU / DU / 1e3; c.cmdot = TU / (2000.0 * 9.81); c.default_mass = 1000.0; return c; }
static SCConst make_sc(double p, double rho, double tl) {
    SCConst c; c.mu = MU; c.m1 = 1.0 - MU; c.kthr = TU * TU / DU / 1e3; c.thrustLimit = tl; c.mass = 1000.0; c.omega = 1.0; c.p = p; c.rho = rho;
    c.cm = TU / (TU * TU / DU / 1e3 * 2000.0 * 9.81); return c;
}

// ---- direct: one segment = forward + backward leg, FIXED grid nsteps = 10, state + [Phi | Gamma] (lto_prop_generic.cuh ep_leg)
template <int NS>
static long long count_direct(double* checksum) {
    const EPConst c = make_ep(); DirectCfg cfg; cfg.mode = 0; cfg.nsteps = 10; cfg.tol = 1e-13; cfg.err_norm = 0; cfg.max_attempts = 100000;
    real xa[7] = {1.12, 0.01, 0.05, 0.01, 0.17, 0.02, 950.0}, xb[7] = {1.13, 0.04, 0.05, 0.02, 0.16, 0.01, 949.95};
    real ua[3] = {0.03, -0.02, 0.05}, ub[3] = {0.01, 0.04, -0.03};
    real xe[2][NS], S[2][NS * (NS + 3)], me[2]; int natt[2];
    memset(&g_t, 0, sizeof g_t);
    const real t0 = 0.0, t1 = 0.15, tm = (t0 + t1) * 0.5;      // (the mid-point is part of the hot path: multiShoot_CRTBP_direct.jl:84)
    ep_leg<NS, true>(xa, ua, 0, t0, tm, cfg, c, xe[0], S[0], &me[0], &natt[0]);
    ep_leg<NS, true>(xb, ub, 1, tm, t1, cfg, c, xe[1], S[1], &me[1], &natt[1]);
    real d[NS];
    for (int i = 0; i < NS; ++i) d[i] = xe[0][i] - xe[1][i];    // :101
    double cs = 0.0;
    for (int i = 0; i < NS; ++i) cs += val(d[i]);
    for (int i = 0; i < NS * (NS + 3); ++i) cs += val(S[0][i]) - val(S[1][i]);
    *checksum = cs;
    return flops(g_t);
}

// ---- indirect, generic first-order formulation: ONE attempted RKF7(8) step of [x | Phi] + the joint scaled error norm
template <int ND>
static long long count_indirect_generic(double p, double rho, double tl, double* checksum) {
    const SCConst c = make_sc(p, rho, tl);
    typedef SCRhs<ND, true> R;
    constexpr int NT = R::NT;
    static real y[NT], yn[NT], gam[NT], ytmp[NT], k[13 * NT];
    const double x12[12] = {1.12, 0.01, 0.05, 0.01, 0.17, 0.02, 0.03, -0.06, 0.1, 0.5, -0.7, 0.4};
    if (ND == 12) for (int i = 0; i < 12; ++i) y[i] = x12[i];
    else { for (int i = 0; i < 6; ++i) y[i] = x12[i]; y[6] = 950.0; for (int i = 0; i < 6; ++i) y[7 + i] = x12[6 + i]; y[13] = 0.02; }
    for (int i = ND; i < NT; ++i) y[i] = 0.0;
    for (int j = 0; j < ND; ++j) y[ND * (1 + j) + j] = 1.0;
    R rhs(c, tl, rho);
    memset(&g_t, 0, sizeof g_t);
    rkf78_step<NT>(rhs, y, real(0.02), yn, gam, k, ytmp);
    const real e = scaled_rms<0>(gam, y, yn, real(1e-13), real(1e-13), NT);
    double cs = val(e);
    for (int i = 0; i < NT; ++i) cs += val(yn[i]);
    *checksum = cs;
    return flops(g_t);
}

// ---- indirect 12, half-column formulation (lto_hc_math.cuh): ONE attempted step of the state and the 24 half-columns
template <int J>
static void st_stage(hcm::K3& Kr, hcm::K3& Kl, const real (&r)[3], const real (&v)[3], const real (&lv)[3], const real (&lvd)[3], real h, real h2,
                     const SCConst& c, const hcm::Law& lw, real (*U)[6], real (*W)[6], real (*G)[6]) {
    real R[3], V[3], M[3], N[3], kr[3], kl[3];
    hcm::stage_in<J>(Kr, r, v, h, h2, R, V);
    hcm::stage_in<J>(Kl, lv, lvd, h, h2, M, N);
    hcm::sc_eval2<true>(R, V, M, N, c.mu, c.m1, real(2.0) * c.omega, c.p, lw, kr, kl, U[J], W[J], G[J]);
    for (int q = 0; q < 3; ++q) { Kr.k[J][q] = kr[q]; Kl.k[J][q] = kl[q]; }
}
template <int J>
static void cl_stage(hcm::K3& Ka, hcm::K3& Kc, const real (&a)[3], const real (&ad)[3], const real (&cc)[3], const real (&cd)[3], real h, real h2, real w2,
                     const real (*U)[6], const real (*W)[6], const real (*G)[6]) {
    real Pa[3], Pad[3], Pc[3], Pcd[3], ka[3], kc[3];
    hcm::stage_in<J>(Ka, a, ad, h, h2, Pa, Pad);
    hcm::stage_in<J>(Kc, cc, cd, h, h2, Pc, Pcd);
    hcm::col_rhs(U[J], G[J], w2, Pa, Pad, Pc, ka);
    hcm::col_rhs(U[J], W[J], w2, Pc, Pcd, Pa, kc);
    for (int q = 0; q < 3; ++q) { Ka.k[J][q] = ka[q]; Kc.k[J][q] = kc[q]; }
}
#define ALL13(F, ...) F<0>(__VA_ARGS__); F<1>(__VA_ARGS__); F<2>(__VA_ARGS__); F<3>(__VA_ARGS__); F<4>(__VA_ARGS__); F<5>(__VA_ARGS__); F<6>(__VA_ARGS__); \
    F<7>(__VA_ARGS__); F<8>(__VA_ARGS__); F<9>(__VA_ARGS__); F<10>(__VA_ARGS__); F<11>(__VA_ARGS__); F<12>(__VA_ARGS__)

static long long count_indirect12_halfcol(double p, double rho, double tl, double* checksum) {
    const SCConst c = make_sc(p, rho, tl);
    const real w2 = real(2.0) * c.omega;
    hcm::Law lw; lw.aL = c.thrustLimit * c.kthr / c.mass; lw.rho_inv = real(1.0) / c.rho; lw.rq = lw.aL / (real(4.0) * c.rho);
    const real x[12] = {1.12, 0.01, 0.05, 0.01, 0.17, 0.02, 0.03, -0.06, 0.1, 0.5, -0.7, 0.4};
    real r[3], v[3], lv[3], lvd[3];
    hcm::to_z(w2, x, r, v, lv, lvd);
    real ca[12][2][3], cad[12][2][3];
    for (int j = 0; j < 12; ++j) for (int hf = 0; hf < 2; ++hf) hcm::col_init(j, hf, w2, ca[j][hf], cad[j][hf]);
    const real h = 0.02, h2 = h * h, atol = 1e-13, rtol = 1e-13;
    memset(&g_t, 0, sizeof g_t);
    hcm::K3 Kr, Kl;
    real U[13][6], W[13][6], G[13][6];
    ALL13(st_stage, Kr, Kl, r, v, lv, lvd, h, h2, c, lw, U, W, G);
    real rn[3], vn[3], lvn[3], lvdn[3];
    hcm::step_update(Kr, r, v, h, h2, rn, vn);
    hcm::step_update(Kl, lv, lvd, h, h2, lvn, lvdn);
    real s2 = hcm::state_err_sumsq<false>(Kr, Kl, w2, h, h2, r, v, lv, lvd, rn, vn, lvn, lvdn, atol, rtol);
    double cs = 0.0;
    for (int j = 0; j < 12; ++j) {
        hcm::K3 Ka, Kc;
        real an[3], adn[3], cn[3], cdn[3], ep[3], epd[3];
        ALL13(cl_stage, Ka, Kc, ca[j][0], cad[j][0], ca[j][1], cad[j][1], h, h2, w2, U, W, G);
        hcm::step_update(Ka, ca[j][0], cad[j][0], h, h2, an, adn);
        hcm::step_update(Kc, ca[j][1], cad[j][1], h, h2, cn, cdn);
        hcm::step_error(Ka, h, h2, ep, epd);
        s2 += hcm::col_err_sumsq(0, w2, ca[j][0], cad[j][0], an, adn, ep, epd, atol, rtol);
        hcm::step_error(Kc, h, h2, ep, epd);
        s2 += hcm::col_err_sumsq(1, w2, ca[j][1], cad[j][1], cn, cdn, ep, epd, atol, rtol);
        for (int q = 0; q < 3; ++q) cs += val(an[q]) + val(adn[q]) + val(cn[q]) + val(cdn[q]);
    }
    const real u = s2 * real(1.0 / 156.0);
    cs += val(u);
    for (int q = 0; q < 3; ++q) cs += val(rn[q]) + val(vn[q]) + val(lvn[q]) + val(lvdn[q]);
    *checksum = cs;
    return flops(g_t);
}

static void emit(const char* key, long long f, double cs, bool last = false) {
    printf("  \"%s\": {\"flops\": %lld, \"add\": %lld, \"mul\": %lld, \"div\": %lld, \"sqrt\": %lld, \"fma\": %lld, \"special_calls\": %lld, \"checksum\": %.17g}%s\n",
           key, f, g_t.add, g_t.mul, g_t.div, g_t.sqrt_, g_t.fma_, g_t.special, cs, last ? "" : ",");
}

int main() {
    double cs;
    long long f;
    printf("{\n");
    f = count_direct<7>(&cs); emit("direct7_segment", f, cs);
    f = count_direct<6>(&cs); emit("direct6_segment", f, cs);
    f = count_indirect_generic<12>(1.0, 1.0, 0.05, &cs); emit("indirect12_step_first_order", f, cs);
    f = count_indirect12_halfcol(1.0, 1.0, 0.05, &cs); emit("indirect12_step_half_column", f, cs);
    f = count_indirect_generic<14>(1.0, 1.0, 0.05, &cs); emit("indirect14_step_first_order", f, cs, true);
    printf("}\n");
    return 0;
}
