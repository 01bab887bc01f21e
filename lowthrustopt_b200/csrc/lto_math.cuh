// lto_math.cuh -- CRTBP dynamics and their analytic Jacobians (variational equations)
// for the segment-propagation kernels.  Sparsity-exploiting formulation: the base
// state produces a small set of coefficients (U_xx, W, G, ...) once per RK stage and
// every sensitivity column is advanced with the structured product A*s.
//
// Reference equations:
//   direct   : src/CRTBP_prop_EP_deriv.jl:8-61        (state [r v (m)], thrust vector u [N])
//   indirect : src/CRTBP_stateCostate_deriv.jl:9-90   (state [r v lr lv], control law from lv)
//   14-dim   : extension, mass terms after GeneralCode/twoBody_stateCostate_mass_deriv.jl:26,57,61,76
// The reference has no variational equations (it uses forward FD / ForwardDiff);
// A = d(rhs)/d(state) below is derived from those right-hand sides.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define LTO_HD __host__ __device__ __forceinline__
#else
#define LTO_HD inline
#endif

namespace lto {

// ----------------------------------------------------------------------------
// Constants shared by all kernels (one struct in kernel-parameter space).
// ----------------------------------------------------------------------------
struct EPConst {        // direct path
    double mu;          // MU
    double m1;          // 1 - MU
    double kthr;        // TU^2/(DU*1e3): N/kg -> DU/TU^2   (CRTBP_prop_EP_deriv.jl:32)
    double cmdot;       // TU/(Isp*g0)                     (CRTBP_prop_EP_deriv.jl:41-42)
    double default_mass;// 1000 kg                         (CRTBP_prop_EP_deriv.jl:20)
};

struct SCConst {        // indirect path
    double mu, m1;
    double kthr;        // TU^2/(DU*1e3)
    double thrustLimit; // N   (may be overridden per segment)
    double mass;        // kg  (12-dim only)
    double omega;       // time_direction
    double p, rho;
    double cm;          // TU/(kthr*Isp*g0) : m' = -cm*umag*m   (14-dim only)
};

// Gravity-gradient coefficients of one RK stage (symmetric 3x3: xx,yy,zz,xy,xz,yz)
struct Grav {
    double dx1, dx2, y, z;
    double a3_1, a3_2;      // mu_b / r_b^3
    double a5_1, a5_2;      // 3 mu_b / r_b^5
    double U[6];
};

// 1/sqrt(x): hardware-assisted on the device (<= 1 ulp), IEEE division on the host test build
LTO_HD double lto_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

LTO_HD void grav_eval(double x, double y, double z, double mu, double m1, Grav& g) {
    g.dx1 = x + mu;
    g.dx2 = g.dx1 - 1.0;
    g.y = y; g.z = z;
    const double q = y * y + z * z;
    const double r1s = fma(g.dx1, g.dx1, q);
    const double r2s = fma(g.dx2, g.dx2, q);
    const double i1 = lto_rsqrt(r1s);
    const double i2 = lto_rsqrt(r2s);
    const double i1s = i1 * i1, i2s = i2 * i2;
    const double i13 = i1s * i1, i23 = i2s * i2;
    g.a3_1 = m1 * i13;
    g.a3_2 = mu * i23;
    g.a5_1 = 3.0 * g.a3_1 * i1s;
    g.a5_2 = 3.0 * g.a3_2 * i2s;
    const double gg = -(g.a3_1 + g.a3_2);
    const double s5 = g.a5_1 + g.a5_2;
    const double t = fma(g.a5_1, g.dx1, g.a5_2 * g.dx2);
    g.U[0] = 1.0 + gg + fma(g.a5_1 * g.dx1, g.dx1, g.a5_2 * g.dx2 * g.dx2);
    g.U[1] = 1.0 + gg + s5 * y * y;
    g.U[2] = gg + s5 * z * z;
    g.U[3] = t * y;
    g.U[4] = t * z;
    g.U[5] = s5 * y * z;
}

// acceleration from gravity + centrifugal + Coriolis (no thrust)
LTO_HD void grav_accel(const Grav& g, double x, double vx, double vy, double omega, double a[3]) {
    a[0] = fma(-g.a3_1, g.dx1, fma(-g.a3_2, g.dx2, fma(2.0 * omega, vy, x)));
    a[1] = fma(-(g.a3_1 + g.a3_2), g.y, fma(-2.0 * omega, vx, g.y));
    a[2] = -(g.a3_1 + g.a3_2) * g.z;
}

LTO_HD void sym3_mul(const double M[6], const double v[3], double out[3]) {
    out[0] = fma(M[0], v[0], fma(M[3], v[1], M[4] * v[2]));
    out[1] = fma(M[3], v[0], fma(M[1], v[1], M[5] * v[2]));
    out[2] = fma(M[4], v[0], fma(M[5], v[1], M[2] * v[2]));
}
LTO_HD void sym3_mul_acc(const double M[6], const double v[3], double out[3]) {
    out[0] = fma(M[0], v[0], fma(M[3], v[1], fma(M[4], v[2], out[0])));
    out[1] = fma(M[3], v[0], fma(M[1], v[1], fma(M[5], v[2], out[1])));
    out[2] = fma(M[4], v[0], fma(M[5], v[1], fma(M[2], v[2], out[2])));
}

// ----------------------------------------------------------------------------
// DIRECT path.  State x = [r v (m)], NS = 6 or 7.  Per-leg constants:
//   u[3] thrust [N], omega = +-1, unorm = |u|, mdot = -omega*|u|*cmdot.
// Stage coefficients for the sensitivity columns s = [s_r s_v (s_m)]:
//   s_r' = s_v
//   s_v' = U s_r + C s_v + am*s_m + bv        C s_v = (2w s_vy, -2w s_vx, 0)
//   s_m' = bm
// with am = -kthr*u/m^2 (A[4:6,7]), and for the control column c:
//   bv = (kthr/m) e_c  (B[4:6,:]),  bm = -omega*cmdot*uhat_c  (B[7,:]);
// uhat_c := 1 when |u| = 0 (one-sided slope the reference's forward FD produces).
// ----------------------------------------------------------------------------
struct EPStage {
    double U[6];
    double kom;     // kthr / m
    double am[3];   // -kthr*u/m^2
};

template <int NS>
LTO_HD void ep_stage(const double* x, const double u[3], double omega, double mdot, const EPConst& c,
                     double* f, EPStage& st) {
    Grav g;
    grav_eval(x[0], x[1], x[2], c.mu, c.m1, g);
    const double m = (NS == 7) ? x[6] : c.default_mass;
    const double im = 1.0 / m;
    st.kom = c.kthr * im;
    double a[3];
    grav_accel(g, x[0], x[3], x[4], omega, a);
    f[0] = x[3]; f[1] = x[4]; f[2] = x[5];
    f[3] = fma(u[0], st.kom, a[0]);
    f[4] = fma(u[1], st.kom, a[1]);
    f[5] = fma(u[2], st.kom, a[2]);
    if (NS == 7) f[6] = mdot;
#pragma unroll
    for (int i = 0; i < 6; ++i) st.U[i] = g.U[i];
    const double k2 = -st.kom * im;
    st.am[0] = k2 * u[0]; st.am[1] = k2 * u[1]; st.am[2] = k2 * u[2];
}

// column derivative; bvc = index (0..2) of the control column or -1; bm = its mass-row forcing
template <int NS>
LTO_HD void ep_col(const EPStage& st, double omega, const double* s, int bvc, double bm, double* ds) {
    ds[0] = s[3]; ds[1] = s[4]; ds[2] = s[5];
    double acc[3];
    acc[0] = 2.0 * omega * s[4];
    acc[1] = -2.0 * omega * s[3];
    acc[2] = 0.0;
    if (NS == 7) {
        acc[0] = fma(st.am[0], s[6], acc[0]);
        acc[1] = fma(st.am[1], s[6], acc[1]);
        acc[2] = fma(st.am[2], s[6], acc[2]);
    }
    if (bvc >= 0) acc[bvc] += st.kom;
    sym3_mul_acc(st.U, s, acc);
    ds[3] = acc[0]; ds[4] = acc[1]; ds[5] = acc[2];
    if (NS == 7) ds[6] = bm;
}

// ----------------------------------------------------------------------------
// INDIRECT path.  ND = 12: s = [r v lr lv];  ND = 14: s = [r v m lr lv lm].
// Control law (CRTBP_stateCostate_deriv.jl:36-64):  u_acc = -umag(n; aL) lv/n,  n = |lv|.
// Stage coefficients:
//   U (gravity gradient), W = d(U lv)/dr, G = d(u_acc)/d(lv)
//   14-dim extras: gm = d(u_acc)/dm (3), mm = d(m')/dm, ml = d(m')/d(lv) (3),
//                  lml = d(lm')/d(lv) (3), lmm = d(lm')/dm
// Column equations (A*phi):
//   pr' = pv
//   pv' = U pr + C pv + G plv (+ gm*pm)
//   pm' = mm*pm + ml.plv                               (14)
//   plr' = -W pr - U plv
//   plv' = -plr - C^T plv      C^T plv = (-2w plv_y, 2w plv_x, 0)
//   plm' = lml.plv + lmm*pm                            (14)
// ----------------------------------------------------------------------------
struct SCStage {
    double U[6], W[6], G[6];
    double gm[3], ml[3], lml[3];
    double mm, lmm;
};

// returns 0 ok, -1 invalid p
template <int ND>
LTO_HD int sc_stage(const double* s, const SCConst& c, double thrustLimit, double rho, double* f, SCStage& st) {
    constexpr int OL = (ND == 14) ? 7 : 6;      // offset of lr
    const double* lr = s + OL;
    const double* lv = s + OL + 3;
    const double omega = c.omega;
    Grav g;
    grav_eval(s[0], s[1], s[2], c.mu, c.m1, g);
    const double m = (ND == 14) ? s[6] : c.mass;
    const double aL = thrustLimit * c.kthr / m;                       // :33
    const double n2 = fma(lv[0], lv[0], fma(lv[1], lv[1], lv[2] * lv[2]));
    const double n = sqrt(n2);
    double umag, dn = 0.0;       // umag and d(umag)/dn at fixed aL
    bool prop_aL;                // umag proportional to aL (=> d(umag)/dm = -umag/m in the 14-dim system)
    if (c.p == 0.0) {                                                 // :36-39
        umag = aL; prop_aL = true;
    } else if (c.p == 1.0) {                                          // :41-43
        const double th = tanh((n - 1.0) / (2.0 * rho));
        umag = 0.5 * (1.0 + th) * aL; prop_aL = true;
        dn = aL * (1.0 - th * th) / (4.0 * rho);
    } else if (c.p > 1.0) {                                           // :45-50
        const double e = 1.0 / (c.p - 1.0);
        const double w = (c.p == 2.0) ? 0.5 * n : pow(n / c.p, e);
        if (w > aL) { umag = aL; prop_aL = true; }
        else { umag = w; prop_aL = false; dn = (n > 0.0) ? e * w / n : 0.0; }
    } else {
        return -1;                                                    // :52
    }
    double lh[3] = {0.0, 0.0, 0.0};
    double uon = 0.0;            // umag / n
    const bool dead = !(n > 0.0);                                     // :59-64  NaN guard -> zero control
    if (!dead) {
        const double in = 1.0 / n;
        lh[0] = lv[0] * in; lh[1] = lv[1] * in; lh[2] = lv[2] * in;
        uon = umag * in;
    } else {
        umag = 0.0; dn = 0.0;
    }
    double a[3];
    grav_accel(g, s[0], s[3], s[4], omega, a);
    f[0] = s[3]; f[1] = s[4]; f[2] = s[5];
    f[3] = fma(-umag, lh[0], a[0]);                                   // :79
    f[4] = fma(-umag, lh[1], a[1]);                                   // :80
    f[5] = fma(-umag, lh[2], a[2]);                                   // :81
    double Ul[3];
    sym3_mul(g.U, lv, Ul);
    f[OL + 0] = -Ul[0]; f[OL + 1] = -Ul[1]; f[OL + 2] = -Ul[2];       // :83-85
    f[OL + 3] = fma(2.0 * omega, lv[1], -lr[0]);                      // :86
    f[OL + 4] = fma(-2.0 * omega, lv[0], -lr[1]);                     // :87
    f[OL + 5] = -lr[2];                                               // :88
#pragma unroll
    for (int i = 0; i < 6; ++i) st.U[i] = g.U[i];
    // G = -[dn lh lh^T + (umag/n)(I - lh lh^T)]
    {
        const double cd = uon - dn;      // G = -uon I + (uon - dn) lh lh^T
        st.G[0] = fma(cd * lh[0], lh[0], -uon);
        st.G[1] = fma(cd * lh[1], lh[1], -uon);
        st.G[2] = fma(cd * lh[2], lh[2], -uon);
        st.G[3] = cd * lh[0] * lh[1];
        st.G[4] = cd * lh[0] * lh[2];
        st.G[5] = cd * lh[1] * lh[2];
    }
    // W = sum_b [ a5_b ((d.l) I + d l^T + l d^T) - 5 a5_b / r_b^2 (d.l) d d^T ],  a5_b = 3 mu_b / r_b^5
    {
        const double d1[3] = {g.dx1, g.y, g.z};
        const double d2[3] = {g.dx2, g.y, g.z};
        const double q = g.y * g.y + g.z * g.z;
        const double ir1s = 1.0 / fma(g.dx1, g.dx1, q);
        const double ir2s = 1.0 / fma(g.dx2, g.dx2, q);
        const double dl1 = fma(d1[0], lv[0], fma(d1[1], lv[1], d1[2] * lv[2]));
        const double dl2 = fma(d2[0], lv[0], fma(d2[1], lv[1], d2[2] * lv[2]));
        const double e1 = g.a5_1 * dl1, e2 = g.a5_2 * dl2;            // a5 (d.l)
        const double h1 = -5.0 * e1 * ir1s, h2 = -5.0 * e2 * ir2s;    // -15 mu (d.l)/r^7
        const int ia[6] = {0, 1, 2, 0, 0, 1}, ib[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int i = ia[k], j = ib[k];
            double w = fma(h1 * d1[i], d1[j], h2 * d2[i] * d2[j]);
            w = fma(g.a5_1, fma(d1[i], lv[j], lv[i] * d1[j]), w);
            w = fma(g.a5_2, fma(d2[i], lv[j], lv[i] * d2[j]), w);
            if (i == j) w += e1 + e2;
            st.W[k] = w;
        }
    }
    if (ND == 14) {
        const double im = 1.0 / m;
        const double dm = prop_aL ? -umag * im : 0.0;                 // d(umag)/dm
        f[6] = -c.cm * umag * m;                                      // m'
        f[13] = -umag * n * im;                                       // lm' = (lv . u_acc)/m
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            st.gm[i] = -dm * lh[i];
            st.ml[i] = -c.cm * m * dn * lh[i];
            st.lml[i] = -(fma(dn, n, umag)) * im * lh[i];
        }
        st.mm = -c.cm * fma(m, dm, umag);
        st.lmm = fma(-n * dm, im, umag * n * im * im);
    }
    return 0;
}

template <int ND>
LTO_HD void sc_col(const SCStage& st, double omega, const double* s, double* ds) {
    constexpr int OL = (ND == 14) ? 7 : 6;
    const double* pr = s; const double* pv = s + 3; const double* plr = s + OL; const double* plv = s + OL + 3;
    ds[0] = pv[0]; ds[1] = pv[1]; ds[2] = pv[2];
    double acc[3];
    acc[0] = 2.0 * omega * pv[1];
    acc[1] = -2.0 * omega * pv[0];
    acc[2] = 0.0;
    if (ND == 14) {
        acc[0] = fma(st.gm[0], s[6], acc[0]);
        acc[1] = fma(st.gm[1], s[6], acc[1]);
        acc[2] = fma(st.gm[2], s[6], acc[2]);
    }
    sym3_mul_acc(st.U, pr, acc);
    sym3_mul_acc(st.G, plv, acc);
    ds[3] = acc[0]; ds[4] = acc[1]; ds[5] = acc[2];
    double b[3] = {0.0, 0.0, 0.0};
    sym3_mul_acc(st.W, pr, b);
    sym3_mul_acc(st.U, plv, b);
    ds[OL + 0] = -b[0]; ds[OL + 1] = -b[1]; ds[OL + 2] = -b[2];
    ds[OL + 3] = fma(2.0 * omega, plv[1], -plr[0]);
    ds[OL + 4] = fma(-2.0 * omega, plv[0], -plr[1]);
    ds[OL + 5] = -plr[2];
    if (ND == 14) {
        ds[6] = fma(st.mm, s[6], fma(st.ml[0], plv[0], fma(st.ml[1], plv[1], st.ml[2] * plv[2])));
        ds[13] = fma(st.lmm, s[6], fma(st.lml[0], plv[0], fma(st.lml[1], plv[1], st.lml[2] * plv[2])));
    }
}

}  // namespace lto
