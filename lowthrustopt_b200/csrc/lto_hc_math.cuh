// lto_hc_math.cuh -- per-thread arithmetic of the "half-column" throughput kernel of the indirect method
// (lto_indirect_hc.cu; K3, ndim = 12).  __host__ __device__, so that tests/native/lto_hostcheck.cpp runs the very same
// arithmetic on the CPU against the CPU checker's dual numbers (tests/test_host_math.py) before any GPU time is spent.
//
// Formulation.  CRTBP_stateCostate_deriv! (src/CRTBP_stateCostate_deriv.jl:78-88) is
//     r'  = v                       v'  = grad(Omega)(r) + C v - uon(|lv|) lv
//     lr' = -U(r) lv                lv' = -lr + C lv            C = 2 w [[0,1,0],[-1,0,0],[0,0,0]],  C^T = -C
// (U = gravity gradient + centrifugal part; w = time_direction).  With the CONSTANT linear change of variables
//     (r, v, lr, lv)  ->  (r, r' = v, lv, lv' = -lr + C lv)
// both halves become second-order systems of the same shape,
//     r''  = grad(Omega)(r) + C r' - uon lv           lv'' = U(r) lv + C lv'
// and so do the two halves of every STM column (a = d r, c = d lv):
//     a''  = U a + C a' + G c                          c''  = U c + C c' + W a        (G = du/dlv, W = d(U lv)/dr)
// A Runge-Kutta method commutes with a constant linear change of variables, so integrating (r, r', lv, lv') and mapping back
// gives the discrete solution of the original system up to rounding.  Every 3-vector second-order system is advanced in
// Nystrom form: only the 13 x 3 stage second derivatives are stored (39 doubles instead of 117 per column), positions are rebuilt
// with G = B*B (lto_tableau.h).  One thread owns ONE such half-column; the pair (a, c) of a column sits on two lanes of a warp and
// swaps stage positions by shuffle.  Error estimates are mapped back to (r, v, lr, lv) before scaling, so the controller sees the
// reference's components (dlr = -c' + C c).
#pragma once
#include "lto_math.cuh"
#include "lto_tableau.h"

namespace lto {
namespace hcm {

struct K3 { double k[13][3]; };                 // stage second derivatives of one 3-vector second-order system

// branch-free 1/sqrt(x), 1/x for normal positive x: hardware seed (MUFU.RSQ64H / RCP64H) + one cubically convergent correction on
// the device (error ~2^-60, no slow-path subroutine on the state warp's dependent chain); IEEE on the host test build
LTO_HD double f_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    const double p = fma(0.375, e, 0.5);
    return fma(y * e, p, y);
#else
    return 1.0 / sqrt(x);
#endif
}
LTO_HD double f_rcp(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
#else
    return 1.0 / x;
#endif
}

// stage input: P = p + c_J h pd + h^2 sum_l G_Jl k_l,   Pd = pd + h sum_l B_Jl k_l
// SPLIT (state warps): two partial sums per combination (even / odd stage index) halve the dependent-FMA depth of the chain
template <int J, bool SPLIT = false>
LTO_HD void stage_in(const K3& K, const double (&p)[3], const double (&pd)[3], double h, double h2, double (&P)[3], double (&Pd)[3]) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        if (J == 0) { P[q] = p[q]; Pd[q] = pd[q]; continue; }
        double ab = 0.0, ag = 0.0, bb = 0.0, bg = 0.0;
#pragma unroll
        for (int l = 0; l < J; ++l) {
            if (SPLIT && (l & 1)) {
                if (lto_tab::Bf(J, l) != 0.0) bb = fma(lto_tab::Bf(J, l), K.k[l][q], bb);
                if (lto_tab::Gf(J, l) != 0.0) bg = fma(lto_tab::Gf(J, l), K.k[l][q], bg);
            } else {
                if (lto_tab::Bf(J, l) != 0.0) ab = fma(lto_tab::Bf(J, l), K.k[l][q], ab);
                if (lto_tab::Gf(J, l) != 0.0) ag = fma(lto_tab::Gf(J, l), K.k[l][q], ag);
            }
        }
        if (SPLIT) { ab += bb; ag += bg; }
        Pd[q] = fma(h, ab, pd[q]);
        P[q] = fma(h2, ag, fma(h * lto_tab::Cf(J), pd[q], p[q]));
    }
}

// 8th-order update (ode.jl:937) of (p, pd)
LTO_HD void step_update(const K3& K, const double (&p)[3], const double (&pd)[3], double h, double h2, double (&pn)[3], double (&pdn)[3]) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double sv = 0.0, sr = 0.0;
#pragma unroll
        for (int l = 0; l < 13; ++l) {
            if (lto_tab::CHIf(l) != 0.0) sv = fma(lto_tab::CHIf(l), K.k[l][q], sv);
            if (lto_tab::CHIBf(l) != 0.0) sr = fma(lto_tab::CHIBf(l), K.k[l][q], sr);
        }
        pdn[q] = fma(h, sv, pd[q]);
        pn[q] = fma(h2, sr, fma(h, pd[q], p[q]));
    }
}

// embedded error estimate (ode.jl:940) of the position-like and the velocity-like components:
//   ep = h^2 (41/840) (psi^T B) k = h^2 (41/840) (k_1 - k_12),   epd = h (41/840) psi^T k = h (41/840) (k_1 + k_11 - k_12 - k_13)
LTO_HD void step_error(const K3& K, double h, double h2, double (&ep)[3], double (&epd)[3]) {
    const double ce = h * lto_tab::ERRC, ce2 = h2 * lto_tab::ERRC;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        ep[q] = ce2 * (K.k[0][q] - K.k[11][q]);
        epd[q] = ce * ((K.k[0][q] + K.k[10][q]) - (K.k[11][q] + K.k[12][q]));
    }
}
// the (k_1 - k_12) parts alone (robust estimate of the state-only controller, lto_prop_generic.cuh drive_rk8)
LTO_HD void step_error_a(const K3& K, double h, double h2, double (&gp)[3], double (&gpd)[3]) {
    const double ce = h * lto_tab::ERRC, ce2 = h2 * lto_tab::ERRC;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double gr = 0.0;
#pragma unroll
        for (int l = 0; l < 11; ++l)
            if (lto_tab::Bf(11, l) != 0.0) gr = fma(lto_tab::Bf(11, l), K.k[l][q], gr);
        gp[q] = -ce2 * gr;
        gpd[q] = ce * (K.k[0][q] - K.k[11][q]);
    }
}

// C v for the Coriolis matrix C = w2 [[0,1,0],[-1,0,0],[0,0,0]]
LTO_HD void coriolis(double w2, const double (&v)[3], double (&out)[3]) { out[0] = w2 * v[1]; out[1] = -w2 * v[0]; out[2] = 0.0; }

// stage second derivative of a half-column:  k = U P + X Po + C Pd   (X = G for the (dr, dv) half, W for the (dlv, dlv') half)
LTO_HD void col_rhs(const double (&U)[6], const double (&X)[6], double w2, const double (&P)[3], const double (&Pd)[3],
                    const double (&Po)[3], double (&k)[3]) {
    k[0] = w2 * Pd[1]; k[1] = -w2 * Pd[0]; k[2] = 0.0;
    sym3_mul_acc(U, P, k);
    sym3_mul_acc(X, Po, k);
}

// sum over the 6 components of one half-column of (error / scale)^2 in the REFERENCE's coordinates.
//   half 0 (dr, dv): components are (p, pd) themselves.
//   half 1 (dlv, dlv'): dlv = p, dlr = -pd + C p  -> the pair (dlr, dlv) is scaled and summed.
LTO_HD double col_err_sumsq(int half, double w2, const double (&p)[3], const double (&pd)[3], const double (&pn)[3], const double (&pdn)[3],
                            const double (&ep)[3], const double (&epd)[3], double atol, double rtol) {
    const double sg = half ? -1.0 : 1.0, sc = half ? w2 : 0.0;
    // second triple: half 0 -> pd;  half 1 -> -pd + C p
    const double y2[3] = {fma(sc, p[1], sg * pd[0]), fma(-sc, p[0], sg * pd[1]), sg * pd[2]};
    const double n2[3] = {fma(sc, pn[1], sg * pdn[0]), fma(-sc, pn[0], sg * pdn[1]), sg * pdn[2]};
    const double e2[3] = {fma(sc, ep[1], sg * epd[0]), fma(-sc, ep[0], sg * epd[1]), sg * epd[2]};
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const double r1 = ep[q] * f_rcp(fma(rtol, fmax(fabs(p[q]), fabs(pn[q])), atol));
        const double r2 = e2[q] * f_rcp(fma(rtol, fmax(fabs(y2[q]), fabs(n2[q])), atol));
        s = fma(r1, r1, s);
        s = fma(r2, r2, s);
    }
    return s;
}

// initial condition of half `half` of STM column `col` (d/dx0[col] in the reference's ordering [r v lr lv])
LTO_HD void col_init(int col, int half, double w2, double (&p)[3], double (&pd)[3]) {
    const int blk = col / 3, qq = col - 3 * blk;
#pragma unroll
    for (int q = 0; q < 3; ++q) {                                     // (no run-time indexing: the arrays live in registers)
        const bool hit = (q == qq);
        p[q] = (hit && ((half == 0 && blk == 0) || (half == 1 && blk == 3))) ? 1.0 : 0.0;
        pd[q] = (hit && half == 0 && blk == 1) ? 1.0 : (hit && half == 1 && blk == 2) ? -1.0 : 0.0;   // dlv' = -dlr + C dlv
    }
    if (half == 1 && blk == 3) { if (qq == 0) pd[1] = -w2; if (qq == 1) pd[0] = w2; }                  // C e_q
}

// the 6 entries of STM column `col` a half-column thread owns, in the reference's ordering: half 0 -> rows 0..5 (dr, dv),
// half 1 -> rows 6..11 (dlr, dlv)
LTO_HD void col_out(int half, double w2, const double (&p)[3], const double (&pd)[3], double (&o)[6]) {
    if (half == 0) {
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = pd[0]; o[4] = pd[1]; o[5] = pd[2];
    } else {
        o[0] = fma(w2, p[1], -pd[0]); o[1] = fma(-w2, p[0], -pd[1]); o[2] = -pd[2];
        o[3] = p[0]; o[4] = p[1]; o[5] = p[2];
    }
}

// ---------------------------------------------------------------------------
// State side.  z = (r, v, lv, lvd) with lvd = lv' = -lr + C lv.
// ---------------------------------------------------------------------------
struct Law { double aL, rho_inv, rq; };       // per segment: aL = thrustLimit k / mass (:33), 1 / rho, aL / (4 rho)

// CRTBP_stateCostate_deriv! (src/CRTBP_stateCostate_deriv.jl:9-90) in the second-order variables:
//   kr = r'' = grad(Omega)(R) + C V - uon M          (:78-81 with the control law :36-64)
//   kl = lv'' = U(R) M + C N                          (:83-88:  lr' = -U lv,  lv' = -lr + C lv)
// and, if LIN, the stage linearisation U (6), W = d(U M)/dR (6), G = d(-uon M)/dM (6), symmetric storage xx yy zz xy xz yz.
// Same arithmetic as lto_math.cuh sc_stage<12>, arranged for a short dependent chain: no divisions, exp instead of tanh.
template <bool LIN>
LTO_HD void sc_eval2(const double (&R)[3], const double (&V)[3], const double (&M)[3], const double (&N)[3], double mu, double m1, double w2,
                     double pexp, const Law& lw, double (&kr)[3], double (&kl)[3], double (&U)[6], double (&W)[6], double (&G)[6]) {
    // ---- gravity (:69-70, :78-81) and its gradient
    const double dx1 = R[0] + mu, dx2 = dx1 - 1.0;
    const double yz = fma(R[1], R[1], R[2] * R[2]);
    const double i1 = f_rsqrt(fma(dx1, dx1, yz)), i2 = f_rsqrt(fma(dx2, dx2, yz));
    const double i1s = i1 * i1, i2s = i2 * i2;
    const double a31 = m1 * i1s * i1, a32 = mu * i2s * i2;
    const double a51 = 3.0 * a31 * i1s, a52 = 3.0 * a32 * i2s;
    const double gg = -(a31 + a32), s5 = a51 + a52;
    const double p1 = a51 * dx1, p2 = a52 * dx2, t = p1 + p2;
    U[0] = fma(p1, dx1, fma(p2, dx2, 1.0 + gg));
    U[1] = fma(s5 * R[1], R[1], 1.0 + gg);
    U[2] = fma(s5 * R[2], R[2], gg);
    U[3] = t * R[1]; U[4] = t * R[2]; U[5] = s5 * R[1] * R[2];
    // ---- control law (:36-64): u_acc = -umag * lv/|lv| = -uon * lv
    const double n2 = fma(M[0], M[0], fma(M[1], M[1], M[2] * M[2]));
    const bool dead = !(n2 > 0.0);                                     // :59-64 NaN guard -> zero control
    const double in = dead ? 0.0 : f_rsqrt(n2);
    const double n = n2 * in;
    double umag, dn = 0.0;
    if (pexp == 1.0) {                                                 // :41-43  0.5 (1 + tanh((n-1)/(2 rho))) aL
        const double y = fmin(fmax((n - 1.0) * lw.rho_inv, -700.0), 700.0);
        const double ey = exp(y);                                      // tanh(y/2) = 1 - 2/(e^y + 1)
        const double th = fma(-2.0, f_rcp(ey + 1.0), 1.0);
        umag = fma(0.5 * lw.aL, th, 0.5 * lw.aL);
        dn = lw.rq * fma(-th, th, 1.0);
    } else if (pexp == 0.0) {                                          // :36-39
        umag = lw.aL;
    } else {                                                           // :45-50
        const double e = 1.0 / (pexp - 1.0);
        const double wv = (pexp == 2.0) ? 0.5 * n : pow(n / pexp, e);
        if (wv > lw.aL) umag = lw.aL;
        else { umag = wv; dn = dead ? 0.0 : e * wv * in; }
    }
    if (dead) { umag = 0.0; dn = 0.0; }
    if (!(n2 == n2)) umag = n2;                                        // a NaN costate stays NaN (reported through status[])
    const double uon = umag * in;
    kr[0] = fma(-uon, M[0], fma(-a31, dx1, fma(-a32, dx2, fma(w2, V[1], R[0]))));
    kr[1] = fma(-uon, M[1], fma(gg, R[1], fma(-w2, V[0], R[1])));
    kr[2] = fma(-uon, M[2], gg * R[2]);
    kl[0] = fma(U[0], M[0], fma(U[3], M[1], fma(U[4], M[2], w2 * N[1])));
    kl[1] = fma(U[3], M[0], fma(U[1], M[1], fma(U[5], M[2], -w2 * N[0])));
    kl[2] = fma(U[4], M[0], fma(U[5], M[1], U[2] * M[2]));
    if (LIN) {
        // G = -uon I + (uon - dn) lh lh^T,  lh = lv/|lv|
        const double cd = uon - dn;
        const double l0 = M[0] * in, l1 = M[1] * in, l2 = M[2] * in;
        const double c0 = cd * l0, c1 = cd * l1;
        // W = d(U lv)/dr = sum_b [ h_b d_b d_b^T + a5_b (d_b lv^T + lv d_b^T) ] + (e1 + e2) I,  d_b = (dx_b, y, z)
        const double ylz = fma(R[1], M[1], R[2] * M[2]);
        const double e1 = a51 * fma(dx1, M[0], ylz), e2 = a52 * fma(dx2, M[0], ylz);
        const double h1 = -5.0 * e1 * i1s, h2 = -5.0 * e2 * i2s;
        const double ee = e1 + e2, hs = h1 + h2;
        const double hx = fma(h1, dx1, h2 * dx2);
        const double sM0 = s5 * M[0];
        W[0] = fma(h1 * dx1, dx1, fma(h2 * dx2, dx2, fma(2.0 * t, M[0], ee)));          // xx
        W[1] = fma(hs * R[1], R[1], fma(2.0 * s5 * R[1], M[1], ee));                   // yy
        W[2] = fma(hs * R[2], R[2], fma(2.0 * s5 * R[2], M[2], ee));                   // zz
        W[3] = fma(hx, R[1], fma(t, M[1], sM0 * R[1]));                                // xy
        W[4] = fma(hx, R[2], fma(t, M[2], sM0 * R[2]));                                // xz
        W[5] = fma(hs * R[1], R[2], s5 * fma(R[1], M[2], M[1] * R[2]));                // yz
        G[0] = fma(c0, l0, -uon); G[1] = fma(c1, l1, -uon); G[2] = fma(cd * l2, l2, -uon);
        G[3] = c0 * l1; G[4] = c0 * l2; G[5] = c1 * l2;
    }
}

// (r, v, lr, lv) -> z = (r, v, lv, lvd)  and back
LTO_HD void to_z(double w2, const double* x, double (&r)[3], double (&v)[3], double (&lv)[3], double (&lvd)[3]) {
#pragma unroll
    for (int q = 0; q < 3; ++q) { r[q] = x[q]; v[q] = x[3 + q]; lv[q] = x[9 + q]; }
    lvd[0] = fma(w2, x[10], -x[6]); lvd[1] = fma(-w2, x[9], -x[7]); lvd[2] = -x[8];
}
LTO_HD void lr_of(double w2, const double (&lv)[3], const double (&lvd)[3], double (&lr)[3]) {
    lr[0] = fma(w2, lv[1], -lvd[0]); lr[1] = fma(-w2, lv[0], -lvd[1]); lr[2] = -lvd[2];
}

LTO_HD double rob_abs(double e, double ga) {               // max(|e|, |ga|, |e - ga|); a NaN estimate stays NaN
    const double m = fmax(fabs(e), fmax(fabs(ga), fabs(e - ga)));
    return (e == e) ? m : e;
}

// sum over the 12 state components (r, v, lr, lv) of (error / scale)^2 for the attempt that took z = (r, v, lv, lvd) to zn.
// ROB: the cancellation-free estimate of the state-only controller.
template <bool ROB>
LTO_HD double state_err_sumsq(const K3& Kr, const K3& Kl, double w2, double h, double h2, const double (&r)[3], const double (&v)[3],
                              const double (&lv)[3], const double (&lvd)[3], const double (&rn)[3], const double (&vn)[3],
                              const double (&lvn)[3], const double (&lvdn)[3], double atol, double rtol) {
    double er[3], ev[3], el[3], eld[3], elr[3], lr[3], lrn[3];
    step_error(Kr, h, h2, er, ev);
    step_error(Kl, h, h2, el, eld);
    lr_of(w2, el, eld, elr);
    lr_of(w2, lv, lvd, lr);
    lr_of(w2, lvn, lvdn, lrn);
    if (ROB) {
        double gr[3], gv[3], gl[3], gld[3], glr[3];
        step_error_a(Kr, h, h2, gr, gv);
        step_error_a(Kl, h, h2, gl, gld);
        lr_of(w2, gl, gld, glr);
#pragma unroll
        for (int q = 0; q < 3; ++q) { er[q] = rob_abs(er[q], gr[q]); ev[q] = rob_abs(ev[q], gv[q]); el[q] = rob_abs(el[q], gl[q]); elr[q] = rob_abs(elr[q], glr[q]); }
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const double q0 = er[q] * f_rcp(fma(rtol, fmax(fabs(r[q]), fabs(rn[q])), atol));
        const double q1 = ev[q] * f_rcp(fma(rtol, fmax(fabs(v[q]), fabs(vn[q])), atol));
        const double q2 = elr[q] * f_rcp(fma(rtol, fmax(fabs(lr[q]), fabs(lrn[q])), atol));
        const double q3 = el[q] * f_rcp(fma(rtol, fmax(fabs(lv[q]), fabs(lvn[q])), atol));
        s = fma(q0, q0, s); s = fma(q1, q1, s); s = fma(q2, q2, s); s = fma(q3, q3, s);
    }
    return s;
}

// Same sum from error vectors that were accumulated stage by stage (the rolled state loop of lto_indirect_hc.cu):
//   e1 = (psi^T B) . k, e2 = psi . k per component c = [r-part (3) | lv-part (3)]; g1 = -(B[11] . k), g2 = k_1 - k_12 (ROB only)
template <bool ROB>
LTO_HD double state_err_sumsq_acc(const double (&e1)[6], const double (&e2)[6], const double (&g1)[6], const double (&g2)[6], double w2, double h,
                                  double h2, const double (&z)[12], const double (&zn)[12], double atol, double rtol) {
    const double ce = h * lto_tab::ERRC, ce2 = h2 * lto_tab::ERRC;
    double er[3], ev[3], el[3], eld[3], elr[3], lr[3], lrn[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) { er[q] = ce2 * e1[q]; ev[q] = ce * e2[q]; el[q] = ce2 * e1[3 + q]; eld[q] = ce * e2[3 + q]; }
    const double lv[3] = {z[6], z[7], z[8]}, lvd[3] = {z[9], z[10], z[11]}, lvn[3] = {zn[6], zn[7], zn[8]}, lvdn[3] = {zn[9], zn[10], zn[11]};
    lr_of(w2, el, eld, elr);
    lr_of(w2, lv, lvd, lr);
    lr_of(w2, lvn, lvdn, lrn);
    if (ROB) {
        double gr[3], gv[3], gl[3], gld[3], glr[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) { gr[q] = ce2 * g1[q]; gv[q] = ce * g2[q]; gl[q] = ce2 * g1[3 + q]; gld[q] = ce * g2[3 + q]; }
        lr_of(w2, gl, gld, glr);
#pragma unroll
        for (int q = 0; q < 3; ++q) { er[q] = rob_abs(er[q], gr[q]); ev[q] = rob_abs(ev[q], gv[q]); el[q] = rob_abs(el[q], gl[q]); elr[q] = rob_abs(elr[q], glr[q]); }
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const double q0 = er[q] * f_rcp(fma(rtol, fmax(fabs(z[q]), fabs(zn[q])), atol));
        const double q1 = ev[q] * f_rcp(fma(rtol, fmax(fabs(z[3 + q]), fabs(zn[3 + q])), atol));
        const double q2 = elr[q] * f_rcp(fma(rtol, fmax(fabs(lr[q]), fabs(lrn[q])), atol));
        const double q3 = el[q] * f_rcp(fma(rtol, fmax(fabs(lv[q]), fabs(lvn[q])), atol));
        s = fma(q0, q0, s); s = fma(q1, q1, s); s = fma(q2, q2, s); s = fma(q3, q3, s);
    }
    return s;
}

}  // namespace hcm
}  // namespace lto
