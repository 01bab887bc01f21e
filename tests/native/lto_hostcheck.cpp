// TEST-ONLY host build of the kernels' arithmetic (lto_math.cuh / lto_prop_generic.cuh are
// __host__ __device__).  Lets the CPU test-suite validate the hand-derived variational
// equations against the oracle's dual numbers without a GPU.  Never part of liblto_b200.so.
#include "../../lowthrustopt_b200/csrc/lto_prop_generic.cuh"
#include "../../lowthrustopt_b200/csrc/lto_hc_math.cuh"
#include <cstring>
using namespace lto;

static EPConst make_ep(const double* dp) {  // MU, DU, TU, Isp
    EPConst c; c.mu = dp[0]; c.m1 = 1.0 - dp[0]; c.kthr = dp[2] * dp[2] / dp[1] / 1e3; c.cmdot = dp[2] / (dp[3] * 9.81);
    c.default_mass = 1000.0; return c;
}
static SCConst make_sc(const double* ip) {  // MU DU TU thrustLimit mass td p rho Isp
    SCConst c; c.mu = ip[0]; c.m1 = 1.0 - ip[0]; c.kthr = ip[2] * ip[2] / ip[1] / 1e3; c.thrustLimit = ip[3]; c.mass = ip[4];
    c.omega = ip[5]; c.p = ip[6]; c.rho = ip[7]; c.cm = ip[2] / (c.kthr * ip[8] * 9.81); return c;
}

// The half-column kernel (lto_indirect_hc.cu) replayed on the host, thread by thread: the state in the second-order variables
// z = (r, v, lv, lvd), every STM column as the lane pair (half 0: (dr, dv); half 1: (dlv, dlv')) that swaps its stage positions,
// the joint / state-only controller, the Hairer initial step over the state.  Same functions as the device code (lto_hc_math.cuh).
template <int J>
static void hcs_state_stage(hcm::K3& Kr, hcm::K3& Kl, const double (&r)[3], const double (&v)[3], const double (&lv)[3], const double (&lvd)[3],
                            double h, double h2, const SCConst& c, const hcm::Law& lw, double (*U)[6], double (*W)[6], double (*G)[6]) {
    double R[3], V[3], M[3], N[3], kr[3], kl[3];
    hcm::stage_in<J>(Kr, r, v, h, h2, R, V);
    hcm::stage_in<J>(Kl, lv, lvd, h, h2, M, N);
    hcm::sc_eval2<true>(R, V, M, N, c.mu, c.m1, 2.0 * c.omega, c.p, lw, kr, kl, U[J], W[J], G[J]);
    for (int q = 0; q < 3; ++q) { Kr.k[J][q] = kr[q]; Kl.k[J][q] = kl[q]; }
}
template <int J>
static void hcs_col_stage(hcm::K3& Ka, hcm::K3& Kc, const double (&a)[3], const double (&ad)[3], const double (&cc)[3], const double (&cd)[3],
                          double h, double h2, double w2, const double (*U)[6], const double (*W)[6], const double (*G)[6]) {
    double Pa[3], Pad[3], Pc[3], Pcd[3], ka[3], kc[3];
    hcm::stage_in<J>(Ka, a, ad, h, h2, Pa, Pad);
    hcm::stage_in<J>(Kc, cc, cd, h, h2, Pc, Pcd);
    hcm::col_rhs(U[J], G[J], w2, Pa, Pad, Pc, ka);        // lane of half 0 receives the partner's position Pc
    hcm::col_rhs(U[J], W[J], w2, Pc, Pcd, Pa, kc);        // lane of half 1 receives Pa
    for (int q = 0; q < 3; ++q) { Ka.k[J][q] = ka[q]; Kc.k[J][q] = kc[q]; }
}
#define HCS_ALL(F, ...) F<0>(__VA_ARGS__); F<1>(__VA_ARGS__); F<2>(__VA_ARGS__); F<3>(__VA_ARGS__); F<4>(__VA_ARGS__); F<5>(__VA_ARGS__); F<6>(__VA_ARGS__); \
    F<7>(__VA_ARGS__); F<8>(__VA_ARGS__); F<9>(__VA_ARGS__); F<10>(__VA_ARGS__); F<11>(__VA_ARGS__); F<12>(__VA_ARGS__)

static double hcs_rms12(const double* e, const double* y, double atol, double rtol) {
    double s = 0.0;
    for (int i = 0; i < 12; ++i) { const double q = e[i] / fma(rtol, fabs(y[i]), atol); s = fma(q, q, s); }
    return sqrt(s / 12.0);
}
static void hcs_f(const double* x, const SCConst& c, const hcm::Law& lw, double* f) {       // right-hand side in the reference's variables
    double r[3], v[3], lv[3], lvd[3], kr[3], kl[3], U[6], W[6], G[6];
    const double w2 = 2.0 * c.omega;
    hcm::to_z(w2, x, r, v, lv, lvd);
    hcm::sc_eval2<false>(r, v, lv, lvd, c.mu, c.m1, w2, c.p, lw, kr, kl, U, W, G);
    double cn[3]; hcm::coriolis(w2, lvd, cn);
    for (int q = 0; q < 3; ++q) { f[q] = v[q]; f[3 + q] = kr[q]; f[6 + q] = -(kl[q] - cn[q]); f[9 + q] = lvd[q]; }   // lr' = -U lv = -(lv'' - C lv')
}


extern "C" {

// f and A = d f / d s (column-major) from the kernels' stage coefficients
int hc_sc_rhs_jac(int nd, const double* s, const double* ip, double* f, double* A) {
    SCConst c = make_sc(ip);
    if (nd == 12) { SCStage st; if (sc_stage<12>(s, c, c.thrustLimit, c.rho, f, st)) return -1;
        for (int j = 0; j < 12; ++j) { double e[12] = {0}; e[j] = 1.0; sc_col<12>(st, c.omega, e, A + 12 * j); } return 0; }
    if (nd == 14) { SCStage st; if (sc_stage<14>(s, c, c.thrustLimit, c.rho, f, st)) return -1;
        for (int j = 0; j < 14; ++j) { double e[14] = {0}; e[j] = 1.0; sc_col<14>(st, c.omega, e, A + 14 * j); } return 0; }
    return -2;
}
int hc_ep_rhs(int ns, const double* x, const double* u, double omega, const double* dp, double* f) {
    EPConst c = make_ep(dp);
    double un = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    EPStage st;
    if (ns == 6) ep_stage<6>(x, u, omega, -omega * un * c.cmdot, c, f, st); else ep_stage<7>(x, u, omega, -omega * un * c.cmdot, c, f, st);
    return 0;
}
int hc_ep_leg(int ns, int sens, const double* x0, const double* u, int backward, double t0, double t1, int mode, int nsteps,
              double tol, int err_norm, const double* dp, double* xend, double* S, double* maxErr, int* natt) {
    EPConst c = make_ep(dp); DirectCfg cfg{mode, nsteps, tol, err_norm, 100000};
    if (ns == 6) return sens ? ep_leg<6, true>(x0, u, backward, t0, t1, cfg, c, xend, S, maxErr, natt)
                             : ep_leg<6, false>(x0, u, backward, t0, t1, cfg, c, xend, S, maxErr, natt);
    return sens ? ep_leg<7, true>(x0, u, backward, t0, t1, cfg, c, xend, S, maxErr, natt)
                : ep_leg<7, false>(x0, u, backward, t0, t1, cfg, c, xend, S, maxErr, natt);
}
int hc_sc_seg(int nd, int sens, const double* x0, double t0, double t1, double atol, double rtol, int controller, int err_norm,
              const double* ip, double* xend, double* Phi, int* nacc, int* natt) {
    SCConst c = make_sc(ip); IndirectCfg cfg{atol, rtol, controller, err_norm, 100000};
    if (nd == 12) return sens ? sc_seg<12, true>(x0, t0, t1, cfg, c, c.thrustLimit, c.rho, xend, Phi, nacc, natt)
                              : sc_seg<12, false>(x0, t0, t1, cfg, c, c.thrustLimit, c.rho, xend, Phi, nacc, natt);
    return sens ? sc_seg<14, true>(x0, t0, t1, cfg, c, c.thrustLimit, c.rho, xend, Phi, nacc, natt)
                : sc_seg<14, false>(x0, t0, t1, cfg, c, c.thrustLimit, c.rho, xend, Phi, nacc, natt);
}

int hc_halfcol_seg12(const double* x0, double t0, double tf, double atol0, double rtol0, int joint, const double* ip, double* xend, double* Phi,
                     int* nacc, int* natt) {
    const SCConst c = make_sc(ip);
    const double w2 = 2.0 * c.omega;
    hcm::Law lw; lw.aL = c.thrustLimit * c.kthr / c.mass; lw.rho_inv = 1.0 / c.rho; lw.rq = lw.aL / (4.0 * c.rho);
    const double ts = joint ? 1.0 : state_tol_scale(c.p, c.rho);
    const double atol = atol0 * ts, rtol = rtol0 * ts;
    const double span = tf - t0;
    // Hairer-Norsett-Wanner initial step over the state components
    double f0[12], f1[12], y1[12], df[12];
    hcs_f(x0, c, lw, f0);
    const double d0 = hcs_rms12(x0, x0, atol, rtol), d1 = hcs_rms12(f0, x0, atol, rtol);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
    h0 = fmin(h0, span);
    for (int i = 0; i < 12; ++i) y1[i] = fma(h0, f0[i], x0[i]);
    hcs_f(y1, c, lw, f1);
    for (int i = 0; i < 12; ++i) df[i] = f1[i] - f0[i];
    const double d2 = hcs_rms12(df, x0, atol, rtol) / h0;
    const double dm = fmax(d1, d2);
    const double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : inv_eighth_root(dm / 0.01);
    double h = fmin(fmin(100.0 * h0, h1), span);
    double r[3], v[3], lv[3], lvd[3];
    hcm::to_z(w2, x0, r, v, lv, lvd);
    double ca[12][2][3], cad[12][2][3];                     // [column][half] (p, pd)
    for (int j = 0; j < 12; ++j) for (int hf = 0; hf < 2; ++hf) hcm::col_init(j, hf, w2, ca[j][hf], cad[j][hf]);
    double t = t0; int na = 0, nt = 0, status = 0; bool lastrej = false;
    const double inv_ne = joint ? 1.0 / 156.0 : 1.0 / 12.0;
    while (t < tf) {
        if (h < span * 1e-12) { status = LTO_ST_HMIN; break; }
        if (nt >= 100000) { status = LTO_ST_MAXSTEPS; break; }
        bool last = false;
        if (t + h >= tf) { h = tf - t; last = true; }
        ++nt;
        const double h2 = h * h;
        hcm::K3 Kr, Kl;
        double U[13][6], W[13][6], G[13][6];
        HCS_ALL(hcs_state_stage, Kr, Kl, r, v, lv, lvd, h, h2, c, lw, U, W, G);
        double rn[3], vn[3], lvn[3], lvdn[3];
        hcm::step_update(Kr, r, v, h, h2, rn, vn);
        hcm::step_update(Kl, lv, lvd, h, h2, lvn, lvdn);
        double s2 = joint ? hcm::state_err_sumsq<false>(Kr, Kl, w2, h, h2, r, v, lv, lvd, rn, vn, lvn, lvdn, atol, rtol)
                          : hcm::state_err_sumsq<true>(Kr, Kl, w2, h, h2, r, v, lv, lvd, rn, vn, lvn, lvdn, atol, rtol);
        double can[12][2][3], cadn[12][2][3];
        for (int j = 0; j < 12; ++j) {
            hcm::K3 Ka, Kc;
            HCS_ALL(hcs_col_stage, Ka, Kc, ca[j][0], cad[j][0], ca[j][1], cad[j][1], h, h2, w2, U, W, G);
            hcm::step_update(Ka, ca[j][0], cad[j][0], h, h2, can[j][0], cadn[j][0]);
            hcm::step_update(Kc, ca[j][1], cad[j][1], h, h2, can[j][1], cadn[j][1]);
            if (joint) {
                double ep[3], epd[3];
                hcm::step_error(Ka, h, h2, ep, epd);
                s2 += hcm::col_err_sumsq(0, w2, ca[j][0], cad[j][0], can[j][0], cadn[j][0], ep, epd, atol, rtol);
                hcm::step_error(Kc, h, h2, ep, epd);
                s2 += hcm::col_err_sumsq(1, w2, ca[j][1], cad[j][1], can[j][1], cadn[j][1], ep, epd, atol, rtol);
            }
        }
        const double u = s2 * inv_ne;                       // eest^2
        if (!(u == u)) { status = LTO_ST_NAN; break; }
        double q = (u == 0.0) ? 5.0 : 0.9 * pow(u, -1.0 / 16.0);
        q = fmin(5.0, fmax(0.2, q));
        if (u <= 1.0) {
            ++na;
            for (int k = 0; k < 3; ++k) { r[k] = rn[k]; v[k] = vn[k]; lv[k] = lvn[k]; lvd[k] = lvdn[k]; }
            memcpy(ca, can, sizeof ca); memcpy(cad, cadn, sizeof cad);
            if (last) { t = tf; break; }
            t += h;
            if (lastrej) q = fmin(q, 1.0);
            lastrej = false;
        } else { lastrej = true; q = fmin(q, 1.0); }
        h *= q;
    }
    double lr[3]; hcm::lr_of(w2, lv, lvd, lr);
    for (int q = 0; q < 3; ++q) { xend[q] = r[q]; xend[3 + q] = v[q]; xend[6 + q] = lr[q]; xend[9 + q] = lv[q]; }
    for (int j = 0; j < 12; ++j) {                          // column-major: Phi[j * 12 + row]
        double o[6];
        hcm::col_out(0, w2, ca[j][0], cad[j][0], o); for (int i = 0; i < 6; ++i) Phi[j * 12 + i] = o[i];
        hcm::col_out(1, w2, ca[j][1], cad[j][1], o); for (int i = 0; i < 6; ++i) Phi[j * 12 + 6 + i] = o[i];
    }
    *nacc = na; *natt = nt;
    return status;
}
}
