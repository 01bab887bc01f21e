/* capi_smoke.c -- include/lto_b200.h used from plain C (what a cgo / ccall / ctypes binding sees): compiles as C99, links against
 * liblto_b200.so, and either reports "no device" (CPU-only machine: there is no fallback) or runs one small direct and one small
 * indirect call plus the Newton update and prints their results.
 *   gcc -std=c99 -Wall -Werror -I include tests/native/capi_smoke.c -L lowthrustopt_b200 -llto_b200 -Wl,-rpath,$PWD/lowthrustopt_b200 -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "lto_b200.h"

int main(void) {
    lto_direct_params dp;
    lto_indirect_params ip;
    lto_handle* h = NULL;
    int rc;
    if (lto_version() != LTO_B200_VERSION) { printf("version mismatch\n"); return 2; }
    lto_direct_params_default(&dp);
    lto_indirect_params_default(&ip);
    if (dp.g0 != 9.81 || dp.mode != LTO_FIXED || ip.reltol != 1e-13 || ip.err_norm != LTO_NORM_STATE_SENS) { printf("bad defaults\n"); return 2; }
    rc = lto_init(0, &h);
    if (rc == LTO_ERR_NODEVICE) { printf("nodevice: %s\n", lto_last_error(NULL)); return 0; }
    if (rc != LTO_SUCCESS) { printf("lto_init failed: %s\n", lto_last_error(NULL)); return 3; }
    {
        /* one direct segment: L2-orbit-like state, small thrust (multiShoot_CRTBP_direct.jl:66-143) */
        double Xa[7] = {1.12, 0.0, 0.02, 0.0, 0.18, 0.0, 1000.0}, Xb[7] = {1.119, 0.028, 0.0199, -0.012, 0.179, -0.002, 999.99};
        double ua[3] = {0.01, -0.02, 0.005}, ub[3] = {0.015, -0.01, 0.0}, ta = 0.0, tb = 0.158601;
        double defect[7], errors[1], jac[7 * 20];
        int32_t status[1];
        rc = lto_direct_defect_jac(h, &dp, 1, 7, 10, Xa, Xb, ua, ub, &ta, &tb, defect, errors, status, jac);
        if (rc) { printf("direct failed: %s\n", lto_last_error(h)); return 4; }
        printf("direct: status %d defect[0] %.12e J[0,0] %.12e\n", (int)status[0], defect[0], jac[0]);
        if (status[0] != LTO_ST_OK || !(fabs(jac[0] - 1.0) < 0.1)) return 5;
    }
    {
        /* a 3-node indirect trajectory: STM blocks, then the Newton update straight from them (multiShoot_CRTBP_indirect.jl:93-183) */
        double XC[3 * 12], t[3] = {0.0, 0.02, 0.04}, defect[2 * 12], phi[2 * 144], upd[3 * 12];
        int32_t status[2], nst[4], st1[1];
        int i, k;
        for (i = 0; i < 3; ++i) {
            double base[12] = {1.12, 0.0, 0.02, 0.0, 0.18, 0.0, 0.01, -0.02, 0.03, 0.02, 0.01, -0.01};
            for (k = 0; k < 12; ++k) XC[i * 12 + k] = base[k];
            XC[i * 12 + 1] += 0.0036 * i;                      /* roughly along the velocity */
        }
        ip.p = 2.0; ip.thrustLimit = 10.0;
        rc = lto_indirect_defect_jac_traj(h, &ip, 1, 3, 12, XC, t, NULL, NULL, defect, status, nst, phi);
        if (rc) { printf("indirect failed: %s\n", lto_last_error(h)); return 6; }
        rc = lto_indirect_newton(h, 1, 3, 0, phi, defect, upd, st1);
        if (rc) { printf("newton failed: %s\n", lto_last_error(h)); return 7; }
        printf("indirect: status %d %d steps %d Phi[0,0] %.12e | newton status %d upd[node 1][0] %.6e pinned %.1e\n", (int)status[0], (int)status[1],
               (int)nst[1], phi[0], (int)st1[0], upd[12], fabs(upd[0]) + fabs(upd[24]));
        if (status[0] != LTO_ST_OK || st1[0] != LTO_ST_OK || upd[0] != 0.0 || upd[24] != 0.0) return 8;
    }
    printf("launches %lld\n", (long long)lto_kernel_launches(h));
    lto_destroy(h);
    return 0;
}
