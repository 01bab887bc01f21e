// TEST-ONLY host build of the kernels' arithmetic (lto_math.cuh / lto_prop_generic.cuh are
// __host__ __device__).  Lets the CPU test-suite validate the hand-derived variational
// equations against the oracle's dual numbers without a GPU.  Never part of liblto_b200.so.
#include "../../lowthrustopt_b200/csrc/lto_prop_generic.cuh"
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
}
