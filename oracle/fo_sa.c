/* fo_sa.c -- Spalart-Allmaras eddy viscosity and source term (TEST INFRASTRUCTURE, see fasts_oracle.h). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "fo_internal.h"
#include "fasts_oracle.h"

/* FastS/Compute/SA/vispalart.for (ides <= 1) with mulam.for, SA/chi.for, SA/xmut.for */
void fo_vispalart(const int *pi, const double *pr, const int ind[6], const double *rop, double *xmut)
{
    const int nd = pi[FO_NDIMDX];
    const double cmus1 = pr[FO_CS], temp01 = 1. / pr[FO_TEMP0], coesut = pr[FO_XMUL0] * (1. + cmus1 * temp01);
    FO_LOOP(ind) {
        const int l = fo_inddm(pi, i, j, k);
        const double amulam = coesut * sqrt(rop[l + 4 * nd] * temp01) / (1. + cmus1 / rop[l + 4 * nd]);
        const double anulam = rop[l] / amulam;
        const double anutild = fmax(rop[l + 5 * nd], 1.e-27);
        const double chi = fmax(anutild * anulam, 1.e-7);
        const double fvv1 = 1. / (1. + FO_SA_CV1 / (chi * chi * chi));
        const double amut = rop[l] * anutild * fvv1;
        const double xmuprov = fmax(amut, 0.);
        xmut[l] = amulam + fmin(xmuprov, pr[FO_RATIOM] * amulam);
    }
}

/* FastS/Compute/SA/spsource_SA.for:6 with sourceSA_grad_{3d,3dhomo,3dcart}.for, sourceSA_prod_dest.for,
   sourceSA_LU.for (SA_LOW_RE = SA_ROT_CORR = 0) */
void fo_spsource_SA(const int *pi, const double *pr, int nitcfg, const int ind[6], const double *xmut,
                    const double *rop, double *coe, const double *ti, const double *tj, const double *tk,
                    const double *vol, const double *dlng, double *drodm)
{
    (void)xmut;
    const int nd = pi[FO_NDIMDX], nm = pi[FO_NDIMDX_MTR], typ = pi[FO_ITYPZONE];
    const int incj = pi[FO_NIJK], inck = pi[FO_NIJK] * pi[FO_NIJK + 1];
    const int mi = pi[FO_NIJK_MTR], mj = pi[FO_NIJK_MTR + 1], mk = pi[FO_NIJK_MTR + 2];
    const double cmus1 = pr[FO_CS], temp01 = 1. / pr[FO_TEMP0], coesut = pr[FO_XMUL0] * (1. + cmus1 * temp01);
    const double c1 = FO_SA_CB2 / FO_SA_SIGMA;
    const double cw1 = (FO_SA_CB1 / FO_SA_CKARM / FO_SA_CKARM) + (1. + FO_SA_CB2) / FO_SA_SIGMA;
    const int store = pi[FO_ITYPCP] <= 1 && nitcfg <= 1;
    FO_LOOP(ind) {
        const int l = fo_inddm(pi, i, j, k), lt = typ == 2 ? 0 : fo_indmtr(pi, i, j, k);
        const int l1 = l + 1, l2 = l - 1, l3 = l + incj, l4 = l - incj, l5 = l + inck, l6 = l - inck;
        double g[4][3], h[3];
        for (int c = 0; c < 4; c++) {
            const double *q = rop + (c < 3 ? c + 1 : 5) * nd;
            const double u1 = q[l] + q[l2], u2 = q[l] + q[l3], u3 = q[l] + q[l1], u4 = q[l] + q[l4], u5 = q[l] + q[l5], u6 = q[l] + q[l6];
            if (typ == 0)
                for (int e = 0; e < 3; e++)
                    g[c][e] = (u3 * ti[lt + mi + e * nm] - u1 * ti[lt + e * nm] + u2 * tj[lt + mj + e * nm] - u4 * tj[lt + e * nm]
                             + u5 * tk[lt + mk + e * nm] - u6 * tk[lt + e * nm]);
            else if (typ == 1) {
                for (int e = 0; e < 2; e++)
                    g[c][e] = (u3 * ti[lt + mi + e * nm] - u1 * ti[lt + e * nm] + u2 * tj[lt + mj + e * nm] - u4 * tj[lt + e * nm]);
                g[c][2] = (u5 - u6) * tk[lt];
            } else { g[c][0] = (u3 - u1) * ti[0]; g[c][1] = (u2 - u4) * tj[0]; g[c][2] = (u5 - u6) * tk[0]; }
        }
        {
            const double *q = rop + 5 * nd;
            const double q0 = q[l] * rop[l];
            const double u1 = q0 + q[l2] * rop[l2], u2 = q0 + q[l3] * rop[l3], u3 = q0 + q[l1] * rop[l1],
                         u4 = q0 + q[l4] * rop[l4], u5 = q0 + q[l5] * rop[l5], u6 = q0 + q[l6] * rop[l6];
            if (typ == 0)
                for (int e = 0; e < 3; e++)
                    h[e] = g[3][e] * (u3 * ti[lt + mi + e * nm] - u1 * ti[lt + e * nm] + u2 * tj[lt + mj + e * nm] - u4 * tj[lt + e * nm]
                                    + u5 * tk[lt + mk + e * nm] - u6 * tk[lt + e * nm]);
            else if (typ == 1) {
                for (int e = 0; e < 2; e++)
                    h[e] = g[3][e] * (u3 * ti[lt + mi + e * nm] - u1 * ti[lt + e * nm] + u2 * tj[lt + mj + e * nm] - u4 * tj[lt + e * nm]);
                h[2] = g[3][2] * (u5 - u6) * tk[lt];
            } else { h[0] = g[3][0] * (u3 - u1) * ti[0]; h[1] = g[3][1] * (u2 - u4) * tj[0]; h[2] = g[3][2] * (u5 - u6) * tk[0]; }
        }
        const double xvol = 0.5 / vol[lt];
        const double rotx = g[2][1] - g[1][2], roty = g[0][2] - g[2][0], rotz = g[1][0] - g[0][1];
        const double rot = sqrt(rotx * rotx + roty * roty + rotz * rotz) * xvol;
        const double anvisc = c1 * (h[0] + h[1] + h[2]) * xvol * xvol;
        const double amulam = coesut * sqrt(rop[l + 4 * nd] * temp01) / (1. + cmus1 / rop[l + 4 * nd]);
        const double anulam = rop[l] / amulam;
        const double anutild = fmax(rop[l + 5 * nd], 1.e-27);
        const double chi = fmax(anutild * anulam, 1.e-7);
        const double f1 = 1. / (1. + FO_SA_CV1 / (chi * chi * chi));
        const double f2 = 1. - chi / (1. + chi * f1);
        const double dist = fmax(dlng[l], 1.e-27);
        double stild = rot + (anutild * f2) / (FO_SA_CKARM * FO_SA_CKARM * dist * dist);
        const double prod = rop[l] * FO_SA_CB1 * stild * anutild;
        stild = fmax(stild, 0.00000000000000000001);
        double r = anutild / (stild * FO_SA_CKARM * FO_SA_CKARM * dist * dist);
        r = fmin(r, 10.);
        const double r5 = r * r * r * r * r - 1.;
        const double gg = r * (1. + FO_SA_CW2 * r5);
        const double g2 = gg * gg, g6 = g2 * g2 * g2;          /* s**6 by repeated multiplication */
        const double fwg = gg * pow((1. + FO_SA_CW3) / (g6 + FO_SA_CW3), 1. / 6.);
        const double destruc = rop[l] * cw1 * fwg * (anutild / dist) * (anutild / dist);
        const double tsource = (prod + anvisc - destruc);
        if (store) {
            const double tsourceb = prod - destruc;
            double amutild = anutild * rop[l];
            amutild = fmax(amutild, 0.0000000000000001);
            const double tsourcenu = vol[lt] * tsourceb / amutild;
            coe[l + 5 * nd] = coe[l + 4 * nd] - fmin(tsourcenu, 0.) * coe[l];
        }
        drodm[l + 5 * nd] = drodm[l + 5 * nd] + vol[lt] * tsource;
    }
}
