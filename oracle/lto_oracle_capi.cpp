// lto_oracle_capi.cpp -- ctypes-callable entry points of the CPU ORACLE.
// TEST INFRASTRUCTURE ONLY (see lto_oracle.hpp header).  Built by oracle/Makefile
// into oracle/liblto_oracle.so.  Never linked into liblto_b200.so.
#include "lto_oracle.hpp"
#include "../lowthrustopt_b200/csrc/lto_prop_generic.cuh"   // equal-algorithm CPU baseline only (oracle_direct_variational_host)
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace lto_oracle;

namespace {

// ---- a3. defectCalc (direct), src/multiShoot_CRTBP_direct.jl:66-109, one segment.
// mode 0 = FIXED (ode7_8, the reference path); mode 1 = ADAPTIVE (ode78 controller
// on each leg, ode.jl:479-534).  T = double or Dual.
template <class T>
int direct_segment(int n, int nsteps, const T* Xa, const T* Xb, const T* ua, const T* ub,
                   double ta, double tb, const DirectParams& P, int mode, double tol, bool with_partials,
                   T* defect, double* err, int* steps /*[2] attempted per leg, may be null*/) {
    double tmid = ta + (tb - ta) / 2.0;                      // :70  t_TU[1:end-1] + diff(t_TU)/2
    T x0[MAXN], xf[MAXN], xb[MAXN];
    double me_f = 0.0, me_b = 0.0;
    int status = 0;
    // ---- forward leg (:82-86)
    for (int c = 0; c < n; ++c) x0[c] = Xa[c];
    auto rhs_f = [&](const T* y, T* dy) { crtbp_ep_rhs<T>(y, n, P, ua, 1.0, dy); return true; };
    if (mode == 0) {
        ode7_8<T>(rhs_f, n, ta, tmid, nsteps, x0, xf, &me_f);
    } else {
        int na, nt; int st = ode78<T>(rhs_f, n, ta, tmid, tol, with_partials, x0, xf, &na, &nt);
        if (st) status = st; if (steps) steps[0] = nt;
    }
    // ---- backward leg (:88-98): flip velocity, td = -1, SAME positive time grid
    for (int c = 0; c < n; ++c) x0[c] = Xb[c];
    for (int c = 3; c < 6; ++c) x0[c] = -x0[c];              // :92
    auto rhs_b = [&](const T* y, T* dy) { crtbp_ep_rhs<T>(y, n, P, ub, -1.0, dy); return true; };
    if (mode == 0) {
        ode7_8<T>(rhs_b, n, ta, tmid, nsteps, x0, xb, &me_b);
    } else {
        int na, nt; int st = ode78<T>(rhs_b, n, ta, tmid, tol, with_partials, x0, xb, &na, &nt);
        if (st && !status) status = st; if (steps) steps[1] = nt;
    }
    for (int c = 3; c < 6; ++c) xb[c] = -xb[c];              // :98
    for (int c = 0; c < n; ++c) defect[c] = xf[c] - xb[c];   // :101
    *err = std::max(me_f, me_b);                             // :104
    for (int c = 0; c < n; ++c) if (o_isnan(defect[c])) status = 1;
    return status;
}

template <int NP>
int direct_segment_var(int n, int nsteps, const double* Xa, const double* Xb, const double* ua, const double* ub,
                       double ta, double tb, const DirectParams& P, int mode, double tol, bool with_partials,
                       double* defect, double* err, double* jac) {
    typedef Dual<NP> D;
    D dXa[MAXN], dXb[MAXN], dua[3], dub[3], ddef[MAXN];
    // column order [X_i, X_{i+1}, u_i, u_{i+1}] (multiShoot_CRTBP_direct.jl:125)
    for (int c = 0; c < n; ++c) { dXa[c] = D(Xa[c]); dXa[c].d[c] = 1.0; dXb[c] = D(Xb[c]); dXb[c].d[n + c] = 1.0; }
    for (int c = 0; c < 3; ++c) { dua[c] = D(ua[c]); dua[c].d[2 * n + c] = 1.0; dub[c] = D(ub[c]); dub[c].d[2 * n + 3 + c] = 1.0; }
    int st = direct_segment<D>(n, nsteps, dXa, dXb, dua, dub, ta, tb, P, mode, tol, with_partials, ddef, err, nullptr);
    for (int c = 0; c < n; ++c) defect[c] = ddef[c].v;
    for (int j = 0; j < NP; ++j) for (int c = 0; c < n; ++c) jac[j * n + c] = ddef[c].d[j];
    return st;
}

IndirectParams make_ip(const double* ip) {
    IndirectParams P; P.MU = ip[0]; P.DU = ip[1]; P.TU = ip[2]; P.thrustLimit = ip[3]; P.mass = ip[4];
    P.time_direction = ip[5]; P.p = ip[6]; P.rho = ip[7]; P.Isp = ip[8]; return P;
}

template <class T>
bool sc_rhs(int ndim, const T* s, const IndirectParams& P, T* d) {
    return (ndim == 12) ? crtbp_sc_rhs12<T>(s, P, d) : crtbp_sc_rhs14<T>(s, P, d);
}

template <int ND>
int indirect_segment_jac(const double* x0, double t0, double t1, const IndirectParams& P, double atol, double rtol,
                         int controller, double* xend, double* phi, int* nacc, int* natt) {
    typedef Dual<ND> D;
    D s0[MAXN], s1[MAXN];
    for (int c = 0; c < ND; ++c) { s0[c] = D(x0[c]); s0[c].d[c] = 1.0; }
    auto rhs = [&](const D* y, D* dy) { return sc_rhs<D>(ND, y, P, dy); };
    int st;
    if (controller == 0) st = rk8_adaptive<D>(rhs, ND, t0, t1, atol, rtol, true, s0, s1, nacc, natt);
    else                 st = ode78<D>(rhs, ND, t0, t1, rtol, true, s0, s1, nacc, natt);
    for (int c = 0; c < ND; ++c) { xend[c] = s1[c].v; if (std::isnan(s1[c].v)) st = st ? st : 1; }
    for (int j = 0; j < ND; ++j) for (int c = 0; c < ND; ++c) phi[j * ND + c] = s1[c].d[j];
    return st;
}

}  // namespace

extern "C" {

int oracle_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// a1
int oracle_ep_rhs(int n, const double* state, const double* dp, const double* control, double td, double* out) {
    DirectParams P{dp[0], dp[1], dp[2], dp[3]};
    crtbp_ep_rhs<double>(state, n, P, control, td, out);
    return 0;
}
// a5 (+ 14-dim extension)
int oracle_sc_rhs(int ndim, const double* s, const double* ip, double* out) {
    IndirectParams P = make_ip(ip);
    return sc_rhs<double>(ndim, s, P, out) ? 0 : -1;
}
// d(rhs)/d(state) by dual numbers, column-major ndim x ndim (tests the kernels' hand-derived A)
int oracle_sc_rhs_jac(int ndim, const double* s, const double* ip, double* A) {
    IndirectParams P = make_ip(ip);
    if (ndim == 12) { typedef Dual<12> D; D x[MAXN], d[MAXN]; for (int c = 0; c < 12; ++c) { x[c] = D(s[c]); x[c].d[c] = 1; }
        if (!crtbp_sc_rhs12<D>(x, P, d)) return -1; for (int j = 0; j < 12; ++j) for (int c = 0; c < 12; ++c) A[j * 12 + c] = d[c].d[j]; return 0; }
    if (ndim == 14) { typedef Dual<14> D; D x[MAXN], d[MAXN]; for (int c = 0; c < 14; ++c) { x[c] = D(s[c]); x[c].d[c] = 1; }
        if (!crtbp_sc_rhs14<D>(x, P, d)) return -1; for (int j = 0; j < 14; ++j) for (int c = 0; c < 14; ++c) A[j * 14 + c] = d[c].d[j]; return 0; }
    return -2;
}
// a2: ode7_8 on the direct RHS; returns full trajectory Xout (n x nsteps, column-major) like the reference
int oracle_ode7_8_ep(int n, double t0, double t1, int nsteps, const double* x0, const double* dp,
                     const double* control, double td, double* Xout, double* maxErr) {
    DirectParams P{dp[0], dp[1], dp[2], dp[3]};
    auto rhs = [&](const double* y, double* dy) { crtbp_ep_rhs<double>(y, n, P, control, td, dy); return true; };
    double x[MAXN], xn[MAXN], g[MAXN]; double me = 0.0;
    for (int c = 0; c < n; ++c) { x[c] = x0[c]; Xout[c] = x0[c]; }
    for (int ind = 1; ind < nsteps; ++ind) {
        double hi = linrange_at(t0, t1, nsteps, ind) - linrange_at(t0, t1, nsteps, ind - 1);
        rkf78_step<double>(rhs, n, x, hi, xn, g);
        double delta = 0.0; for (int c = 0; c < n; ++c) delta = std::max(delta, std::fabs(g[c]));
        if (delta > me) me = delta;
        for (int c = 0; c < n; ++c) { x[c] = xn[c]; Xout[ind * n + c] = xn[c]; }
    }
    *maxErr = me; return 0;
}
// 3b: ode78 on the direct RHS
int oracle_ode78_ep(int n, double t0, double t1, double tol, const double* x0, const double* dp,
                    const double* control, double td, double* xend, int* nacc, int* natt) {
    DirectParams P{dp[0], dp[1], dp[2], dp[3]};
    auto rhs = [&](const double* y, double* dy) { crtbp_ep_rhs<double>(y, n, P, control, td, dy); return true; };
    return ode78<double>(rhs, n, t0, t1, tol, false, x0, xend, nacc, natt);
}

// a3: defectCalc (direct) over a batch of independent (a,b) node pairs.
// Arrays are per-segment AoS: Xa[s*n + c], ua[s*3 + c], ...
int oracle_direct_defect(long long n_seg, int n, int nsteps, const double* Xa, const double* Xb,
                         const double* ua, const double* ub, const double* ta, const double* tb,
                         const double* dp, int mode, double tol, double* defect, double* errors, int* status,
                         int* steps, int nthreads) {
    DirectParams P{dp[0], dp[1], dp[2], dp[3]};
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1)
    for (long long s = 0; s < n_seg; ++s) {
        int st = direct_segment<double>(n, nsteps, Xa + s * n, Xb + s * n, ua + s * 3, ub + s * 3, ta[s], tb[s], P,
                                        mode, tol, false, defect + s * n, errors + s, steps ? steps + 2 * s : nullptr);
        if (status) status[s] = st;
    }
    return 0;
}

// a4: jacobianCalc (direct), reference-faithful forward FD (multiShoot_CRTBP_direct.jl:111-143):
// for each of nvar = 2(n+3) variables perturb by +pert and RE-PROPAGATE BOTH LEGS (:136).
// jac[s] is n x nvar column-major = rows (s-1)n+1..sn of Jac_temp.
int oracle_direct_jac_fd(long long n_seg, int n, int nsteps, const double* Xa, const double* Xb,
                         const double* ua, const double* ub, const double* ta, const double* tb,
                         const double* dp, int mode, double tol, double pert, const double* defect_nom,
                         double* jac, int nthreads) {
    DirectParams P{dp[0], dp[1], dp[2], dp[3]};
    const int nvar = 2 * (n + 3);
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1)
    for (long long s = 0; s < n_seg; ++s) {
        double XU[2 * MAXN + 6], dtmp[MAXN], err;
        for (int j = 0; j < nvar; ++j) {
            for (int c = 0; c < n; ++c) { XU[c] = Xa[s * n + c]; XU[n + c] = Xb[s * n + c]; }        // :125
            for (int c = 0; c < 3; ++c) { XU[2 * n + c] = ua[s * 3 + c]; XU[2 * n + 3 + c] = ub[s * 3 + c]; }
            XU[j] = XU[j] + pert;                                                                     // :129
            direct_segment<double>(n, nsteps, XU, XU + n, XU + 2 * n, XU + 2 * n + 3, ta[s], tb[s], P, mode, tol,
                                   false, dtmp, &err, nullptr);                                       // :136
            for (int c = 0; c < n; ++c) jac[s * n * nvar + j * n + c] = (dtmp[c] - defect_nom[s * n + c]) / pert; // :140
        }
    }
    return 0;
}

// Variational (dual-number) Jacobian of the SAME discrete map: what the FD approximates.
int oracle_direct_jac_var(long long n_seg, int n, int nsteps, const double* Xa, const double* Xb,
                          const double* ua, const double* ub, const double* ta, const double* tb,
                          const double* dp, int mode, double tol, int with_partials,
                          double* defect, double* errors, double* jac, int* status, int nthreads) {
    DirectParams P{dp[0], dp[1], dp[2], dp[3]};
    const int nvar = 2 * (n + 3);
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1)
    for (long long s = 0; s < n_seg; ++s) {
        int st;
        if (n == 6) st = direct_segment_var<18>(n, nsteps, Xa + s * n, Xb + s * n, ua + s * 3, ub + s * 3, ta[s], tb[s], P, mode, tol,
                                                with_partials != 0, defect + s * n, errors + s, jac + s * n * nvar);
        else        st = direct_segment_var<20>(n, nsteps, Xa + s * n, Xb + s * n, ua + s * 3, ub + s * 3, ta[s], tb[s], P, mode, tol,
                                                with_partials != 0, defect + s * n, errors + s, jac + s * n * nvar);
        if (status) status[s] = st;
    }
    return 0;
}

// a6: per-segment propagation of the indirect path (state only).  controller 0 = the
// OrdinaryDiffEq-style controller documented in lto_oracle.hpp, 1 = ode78's.
// thrustLimit_arr / rho_arr: optional per-segment overrides (continuation batches).
int oracle_indirect_prop(long long n_seg, int ndim, const double* x0, const double* t0, const double* t1,
                         const double* ip, const double* thrustLimit_arr, const double* rho_arr,
                         double atol, double rtol, int controller, double* xend, int* status, int* nacc, int* natt,
                         int nthreads, int est /* -1: default (1, the state-only controller of the kernels); 0: plain Fehlberg estimate */) {
    if (est < 0) est = 1;
    IndirectParams P0 = make_ip(ip);
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1)
    for (long long s = 0; s < n_seg; ++s) {
        IndirectParams P = P0;
        if (thrustLimit_arr) P.thrustLimit = thrustLimit_arr[s];
        if (rho_arr) P.rho = rho_arr[s];
        auto rhs = [&](const double* y, double* dy) { return sc_rhs<double>(ndim, y, P, dy); };
        int na = 0, nt = 0, st;
        const double ts = (est == 1) ? state_tol_scale(P.p, P.rho) : 1.0;      // est = 1: the kernels' state-only controller
        if (controller == 0) st = rk8_adaptive<double>(rhs, ndim, t0[s], t1[s], atol * ts, rtol * ts, false, x0 + s * ndim, xend + s * ndim, &na, &nt, 100000, est);
        else                 st = ode78<double>(rhs, ndim, t0[s], t1[s], rtol, false, x0 + s * ndim, xend + s * ndim, &na, &nt);
        for (int c = 0; c < ndim; ++c) if (std::isnan(xend[s * ndim + c]) && !st) st = 1;
        if (status) status[s] = st; if (nacc) nacc[s] = na; if (natt) natt[s] = nt;
    }
    return 0;
}

// Extended-precision truth for the end states: the same right-hand side and RKF7(8) controller in 80-bit long double at
// a tolerance below the double-precision runs' (SURVEY 7.1 "long double truth for error budgets").  xend is rounded to double.
int oracle_indirect_prop_ld(long long n_seg, int ndim, const double* x0, const double* t0, const double* t1,
                            const double* ip, const double* thrustLimit_arr, const double* rho_arr,
                            double atol, double rtol, double* xend, int* status, int nthreads) {
    IndirectParams P0 = make_ip(ip);
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1)
    for (long long s = 0; s < n_seg; ++s) {
        IndirectParams P = P0;
        if (thrustLimit_arr) P.thrustLimit = thrustLimit_arr[s];
        if (rho_arr) P.rho = rho_arr[s];
        auto rhs = [&](const long double* y, long double* dy) { return sc_rhs<long double>(ndim, y, P, dy); };
        long double xi[MAXN], xo[MAXN];
        for (int c = 0; c < ndim; ++c) xi[c] = x0[s * ndim + c];
        int na = 0, nt = 0;
        int st = rk8_adaptive<long double>(rhs, ndim, t0[s], t1[s], atol, rtol, false, xi, xo, &na, &nt);
        for (int c = 0; c < ndim; ++c) { xend[s * ndim + c] = (double)xo[c]; if (std::isnan(xend[s * ndim + c]) && !st) st = 1; }
        if (status) status[s] = st;
    }
    return 0;
}

// a7: ForwardDiff.jacobian through the adaptive solver (multiShoot_CRTBP_indirect.jl:103-121):
// Dual<ndim> numbers through the same integrator, controller norm INCLUDING partials.
int oracle_indirect_prop_jac(long long n_seg, int ndim, const double* x0, const double* t0, const double* t1,
                             const double* ip, const double* thrustLimit_arr, const double* rho_arr,
                             double atol, double rtol, int controller, double* xend, double* phi, int* status,
                             int* nacc, int* natt, int nthreads) {
    IndirectParams P0 = make_ip(ip);
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 0 ? nthreads : 1)
    for (long long s = 0; s < n_seg; ++s) {
        IndirectParams P = P0;
        if (thrustLimit_arr) P.thrustLimit = thrustLimit_arr[s];
        if (rho_arr) P.rho = rho_arr[s];
        int na = 0, nt = 0, st;
        if (ndim == 12) st = indirect_segment_jac<12>(x0 + s * 12, t0[s], t1[s], P, atol, rtol, controller, xend + s * 12, phi + s * 144, &na, &nt);
        else            st = indirect_segment_jac<14>(x0 + s * 14, t0[s], t1[s], P, atol, rtol, controller, xend + s * 14, phi + s * 196, &na, &nt);
        if (status) status[s] = st; if (nacc) nacc[s] = na; if (natt) natt[s] = nt;
    }
    return 0;
}

// ---------------------------------------------------------------------------
// Equal-algorithm CPU baseline (mode 2): the GPU path's OWN arithmetic -- state + variational equations
// [Phi | Gamma] in one RKF7(8) integration per leg, structural zeros skipped (lowthrustopt_b200/csrc/lto_math.cuh and
// lto_prop_generic.cuh are __host__ __device__) -- compiled for the host cores and run with OpenMP over segments.  A baseline
// only: nothing here feeds a parity check, and the product never links this file.
// ---------------------------------------------------------------------------
int oracle_direct_variational_host(long long n_seg, int n, int nsteps, const double* Xa, const double* Xb, const double* ua,
                                   const double* ub, const double* ta, const double* tb, const double* dp, double* defect,
                                   double* jac, int nthreads) {
    lto::EPConst c; c.mu = dp[0]; c.m1 = 1.0 - dp[0]; c.kthr = dp[2] * dp[2] / dp[1] / 1e3; c.cmdot = dp[2] / (dp[3] * 9.81); c.default_mass = 1000.0;
    lto::DirectCfg cfg; cfg.mode = 0; cfg.nsteps = nsteps; cfg.tol = 1e-13; cfg.err_norm = 0; cfg.max_attempts = 100000;
    const int nv = 2 * (n + 3), nc = n + 3;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1)
    for (long long s = 0; s < n_seg; ++s) {
        double xf[7], xb[7], Sf[7 * 10], Sb[7 * 10], ef, eb; int nf, nb;
        const double tm = 0.5 * (ta[s] + tb[s]);
        if (n == 7) {
            lto::ep_leg<7, true>(Xa + s * 7, ua + s * 3, 0, ta[s], tm, cfg, c, xf, Sf, &ef, &nf);
            lto::ep_leg<7, true>(Xb + s * 7, ub + s * 3, 1, tm, tb[s], cfg, c, xb, Sb, &eb, &nb);
        } else {
            lto::ep_leg<6, true>(Xa + s * 6, ua + s * 3, 0, ta[s], tm, cfg, c, xf, Sf, &ef, &nf);
            lto::ep_leg<6, true>(Xb + s * 6, ub + s * 3, 1, tm, tb[s], cfg, c, xb, Sb, &eb, &nb);
        }
        for (int i = 0; i < n; ++i) defect[s * n + i] = xf[i] - xb[i];
        // block columns [X_a | X_b | u_a | u_b] (multiShoot_CRTBP_direct.jl:125), column-major n x nv
        double* J = jac + s * (long long)n * nv;
        for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) { J[j * n + i] = Sf[j * n + i]; J[(n + j) * n + i] = -Sb[j * n + i]; }
        for (int j = 0; j < 3; ++j) for (int i = 0; i < n; ++i) { J[(2 * n + j) * n + i] = Sf[(n + j) * n + i]; J[(2 * n + 3 + j) * n + i] = -Sb[(n + j) * n + i]; }
        (void)nc;
    }
    return 0;
}

}  // extern "C"
