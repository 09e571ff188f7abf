/*
 * fasts_oracle.c -- CPU restatement of FastS's residual-and-update hot path (part 1:
 * metrics, explicit chain, flux balance, ghost cells, driver).
 *
 * TEST INFRASTRUCTURE ONLY -- see fasts_oracle.h.  PARITY UNPINNED (no reference fixtures
 * exist for this path; DESIGN.md section "Oracle").
 *
 * Conventions: loops use the reference's 1-based (i,j,k) interior numbering; the address
 * helpers return 0-based offsets.  Paths in comments are relative to /root/reference/Fast.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "fo_param.h"
#include "fasts_oracle.h"
#include "fo_internal.h"

int fo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------ */
/* Metrics: FastC/Metric/cp_tijk.for:76-195 (3dfull), :196-290 (3dhomo), :291+ (3dcart)  */
/* Loop range = every cell incl. ghosts (FastC/Metric/indice_omp1.h), plus the imax+1,   */
/* jmax+1, kmax+1 face planes (li/lj/lk complements).                                    */
/* ------------------------------------------------------------------------------------ */
static void tijk_cell(const int *pi, const double *x, const double *y, const double *z,
                      double *ti, double *tj, double *tk, int i, int j, int k,
                      int do_i, int do_j, int do_k)
{
    const int typ = pi[FO_ITYPZONE];
    const int nm = pi[FO_NDIMDX_MTR];
    const int inci = 1, incj = pi[FO_NIJK_XYZ], inck = pi[FO_NIJK_XYZ] * pi[FO_NIJK_XYZ + 1];
    const int lt = (typ == 2) ? 0 : fo_indmtr(pi, i, j, k);
    const int l0 = fo_indcg(pi, i, j, k);
    const int l1 = l0 + incj, l2 = l0 + inck, l3 = l1 + inck;
    const int l4 = l0 + inci, l5 = l1 + inci, l6 = l2 + inci;
    if (do_i) {
        double tix10 = (y[l1] - y[l0]) * (z[l3] - z[l0]) - (z[l1] - z[l0]) * (y[l3] - y[l0]);
        double tiy10 = (z[l1] - z[l0]) * (x[l3] - x[l0]) - (x[l1] - x[l0]) * (z[l3] - z[l0]);
        double tiz10 = (x[l1] - x[l0]) * (y[l3] - y[l0]) - (y[l1] - y[l0]) * (x[l3] - x[l0]);
        double tix11 = (y[l2] - y[l3]) * (z[l0] - z[l2]) - (z[l2] - z[l3]) * (y[l0] - y[l2]);
        double tiy11 = (z[l2] - z[l3]) * (x[l0] - x[l2]) - (x[l2] - x[l3]) * (z[l0] - z[l2]);
        double tiz11 = (x[l2] - x[l3]) * (y[l0] - y[l2]) - (y[l2] - y[l3]) * (x[l0] - x[l2]);
        ti[lt] = .5 * (tix10 + tix11);
        if (typ != 2) ti[lt + nm] = .5 * (tiy10 + tiy11);
        if (typ == 0) ti[lt + 2 * nm] = .5 * (tiz10 + tiz11);
    }
    if (do_j) {
        double tjx10 = (y[l0] - y[l4]) * (z[l2] - z[l0]) - (z[l0] - z[l4]) * (y[l2] - y[l0]);
        double tjy10 = (z[l0] - z[l4]) * (x[l2] - x[l0]) - (x[l0] - x[l4]) * (z[l2] - z[l0]);
        double tjz10 = (x[l0] - x[l4]) * (y[l2] - y[l0]) - (y[l0] - y[l4]) * (x[l2] - x[l0]);
        double tjx11 = (y[l6] - y[l2]) * (z[l4] - z[l6]) - (z[l6] - z[l2]) * (y[l4] - y[l6]);
        double tjy11 = (z[l6] - z[l2]) * (x[l4] - x[l6]) - (x[l6] - x[l2]) * (z[l4] - z[l6]);
        double tjz11 = (x[l6] - x[l2]) * (y[l4] - y[l6]) - (y[l6] - y[l2]) * (x[l4] - x[l6]);
        if (typ == 2) tj[lt] = .5 * (tjy10 + tjy11);
        else {
            tj[lt] = .5 * (tjx10 + tjx11);
            tj[lt + nm] = .5 * (tjy10 + tjy11);
            if (typ == 0) tj[lt + 2 * nm] = .5 * (tjz10 + tjz11);
        }
    }
    if (do_k) {
        double tkx10 = (y[l4] - y[l0]) * (z[l5] - z[l4]) - (z[l4] - z[l0]) * (y[l5] - y[l4]);
        double tky10 = (z[l4] - z[l0]) * (x[l5] - x[l4]) - (x[l4] - x[l0]) * (z[l5] - z[l4]);
        double tkz10 = (x[l4] - x[l0]) * (y[l5] - y[l4]) - (y[l4] - y[l0]) * (x[l5] - x[l4]);
        double tkx11 = (y[l1] - y[l5]) * (z[l0] - z[l1]) - (z[l1] - z[l5]) * (y[l0] - y[l1]);
        double tky11 = (z[l1] - z[l5]) * (x[l0] - x[l1]) - (x[l1] - x[l5]) * (z[l0] - z[l1]);
        double tkz11 = (x[l1] - x[l5]) * (y[l0] - y[l1]) - (y[l1] - y[l5]) * (x[l0] - x[l1]);
        if (typ == 0) {
            tk[lt] = .5 * (tkx10 + tkx11);
            tk[lt + nm] = .5 * (tky10 + tky11);
            tk[lt + 2 * nm] = .5 * (tkz10 + tkz11);
        } else tk[lt] = .5 * (tkz10 + tkz11);
    }
}

void fo_cp_tijk(const int *pi, const double *x, const double *y, const double *z,
                double *ti, double *tj, double *tk)
{
    const int typ = pi[FO_ITYPZONE];
    const int nf = pi[FO_NIJK + 3], nfk = pi[FO_NIJK + 4];
    const int iv = pi[FO_IJKV], jv = pi[FO_IJKV + 1], kv = pi[FO_IJKV + 2];
    if (typ == 2) { /* one cell, no complement (init_metric: ind_mtr = 1) */
        tijk_cell(pi, x, y, z, ti, tj, tk, 1, 1, 1, 1, 1, 1);
        return;
    }
    int k0 = 1 - nfk, k1 = kv + nfk;
    if (typ == 1) { k0 = 1; k1 = 1; }
    for (int k = k0; k <= k1; k++)
        for (int j = 1 - nf; j <= jv + nf; j++)
            for (int i = 1 - nf; i <= iv + nf; i++)
                tijk_cell(pi, x, y, z, ti, tj, tk, i, j, k, 1, 1, 1);
    /* complements: li, lj always true for a whole zone; lk only when the k range reaches kv+1 */
    for (int k = k0; k <= k1; k++)
        for (int j = 1 - nf; j <= jv + nf; j++)
            tijk_cell(pi, x, y, z, ti, tj, tk, iv + nf + 1, j, k, 1, 0, 0);
    for (int k = k0; k <= k1; k++)
        for (int i = 1 - nf; i <= iv + nf; i++)
            tijk_cell(pi, x, y, z, ti, tj, tk, i, jv + nf + 1, k, 0, 1, 0);
    if (typ == 0)
        for (int j = 1 - nf; j <= jv + nf; j++)
            for (int i = 1 - nf; i <= iv + nf; i++)
                tijk_cell(pi, x, y, z, ti, tj, tk, i, j, kv + nfk + 1, 0, 0, 1);
}

/* FastC/Metric/cp_vol.for:74-114 (3dfull), :115-147 (3dhomo), :148+ (3dcart) */
void fo_cp_vol(const int *pi, const double *x, const double *y, const double *z,
               const double *ti, const double *tj, const double *tk, double *vol)
{
    const int typ = pi[FO_ITYPZONE];
    const int nm = pi[FO_NDIMDX_MTR];
    const int nf = pi[FO_NIJK + 3], nfk = pi[FO_NIJK + 4];
    const int iv = pi[FO_IJKV], jv = pi[FO_IJKV + 1], kv = pi[FO_IJKV + 2];
    const int inci = 1, incj = pi[FO_NIJK_XYZ], inck = pi[FO_NIJK_XYZ] * pi[FO_NIJK_XYZ + 1];
    const int mi = pi[FO_NIJK_MTR], mj = pi[FO_NIJK_MTR + 1], mk = pi[FO_NIJK_MTR + 2];
    int i0 = 1 - nf, i1 = iv + nf, j0 = 1 - nf, j1 = jv + nf, k0 = 1 - nfk, k1 = kv + nfk;
    if (typ == 1) { k0 = 1; k1 = 1; }
    if (typ == 2) { i0 = i1 = j0 = j1 = k0 = k1 = 1; }
    for (int k = k0; k <= k1; k++)
        for (int j = j0; j <= j1; j++)
            for (int i = i0; i <= i1; i++) {
                const int lt = (typ == 2) ? 0 : fo_indmtr(pi, i, j, k);
                const int l0 = fo_indcg(pi, i, j, k);
                const int l1 = l0 + incj, l2 = l0 + inck, l3 = l1 + inck;
                const int l4 = l0 + inci, l5 = l1 + inci, l6 = l2 + inci, l7 = l3 + inci;
                const int lip = lt + mi, ljp = lt + mj, lkp = lt + mk;
                double ax, ay, az;
                if (typ == 0) {
                    ax = .25 * ((x[l4] + x[l5] + x[l7] + x[l6]) * ti[lip] - (x[l0] + x[l1] + x[l3] + x[l2]) * ti[lt]
                              + (x[l1] + x[l3] + x[l5] + x[l7]) * tj[ljp] - (x[l0] + x[l2] + x[l4] + x[l6]) * tj[lt]
                              + (x[l2] + x[l3] + x[l6] + x[l7]) * tk[lkp] - (x[l0] + x[l1] + x[l4] + x[l5]) * tk[lt]);
                    ay = .25 * ((y[l4] + y[l5] + y[l7] + y[l6]) * ti[lip + nm] - (y[l0] + y[l1] + y[l3] + y[l2]) * ti[lt + nm]
                              + (y[l1] + y[l3] + y[l5] + y[l7]) * tj[ljp + nm] - (y[l0] + y[l2] + y[l4] + y[l6]) * tj[lt + nm]
                              + (y[l2] + y[l3] + y[l6] + y[l7]) * tk[lkp + nm] - (y[l0] + y[l1] + y[l4] + y[l5]) * tk[lt + nm]);
                    az = .25 * ((z[l4] + z[l5] + z[l7] + z[l6]) * ti[lip + 2 * nm] - (z[l0] + z[l1] + z[l3] + z[l2]) * ti[lt + 2 * nm]
                              + (z[l1] + z[l3] + z[l5] + z[l7]) * tj[ljp + 2 * nm] - (z[l0] + z[l2] + z[l4] + z[l6]) * tj[lt + 2 * nm]
                              + (z[l2] + z[l3] + z[l6] + z[l7]) * tk[lkp + 2 * nm] - (z[l0] + z[l1] + z[l4] + z[l5]) * tk[lt + 2 * nm]);
                } else if (typ == 1) {
                    ax = .25 * ((x[l4] + x[l5] + x[l7] + x[l6]) * ti[lip] - (x[l0] + x[l1] + x[l3] + x[l2]) * ti[lt]
                              + (x[l1] + x[l3] + x[l5] + x[l7]) * tj[ljp] - (x[l0] + x[l2] + x[l4] + x[l6]) * tj[lt]);
                    ay = .25 * ((y[l4] + y[l5] + y[l7] + y[l6]) * ti[lip + nm] - (y[l0] + y[l1] + y[l3] + y[l2]) * ti[lt + nm]
                              + (y[l1] + y[l3] + y[l5] + y[l7]) * tj[ljp + nm] - (y[l0] + y[l2] + y[l4] + y[l6]) * tj[lt + nm]);
                    az = .25 * ((z[l2] + z[l3] + z[l6] + z[l7]) * tk[lkp] - (z[l0] + z[l1] + z[l4] + z[l5]) * tk[lt]);
                } else {
                    ax = .25 * ((x[l4] + x[l5] + x[l7] + x[l6]) * ti[lip] - (x[l0] + x[l1] + x[l3] + x[l2]) * ti[lt]);
                    ay = .25 * ((y[l1] + y[l3] + y[l5] + y[l7]) * tj[ljp] - (y[l0] + y[l2] + y[l4] + y[l6]) * tj[lt]);
                    az = .25 * ((z[l2] + z[l3] + z[l6] + z[l7]) * tk[lkp] - (z[l0] + z[l1] + z[l4] + z[l5]) * tk[lt]);
                }
                vol[lt] = fmax((ax + ay + az) / 3., 1.e-30);
            }
}

/* ------------------------------------------------------------------------------------ */
/* FastS/Compute/invist.for: Sutherland viscosity                                         */
/* ------------------------------------------------------------------------------------ */
void fo_invist(const int *pi, const double *pr, const int ind[6], const double *rop, double *xmut)
{
    const double cmus1 = pr[FO_CS], temp01 = 1. / pr[FO_TEMP0];
    const double coesut = pr[FO_XMUL0] * (1. + cmus1 * temp01) * sqrt(temp01);
    const int v5 = 4 * pi[FO_NDIMDX];
    FO_LOOP(ind) {
        const int l = fo_inddm(pi, i, j, k);
        const double t1 = rop[l + v5];
        xmut[l] = coesut * sqrt(t1) * t1 / (t1 + cmus1);
    }
}

/* FastS/Compute/tstb3_global.for:38-70 (non chimera branch) */
void fo_tstb3_global(const int *pi, const double *pr, const int ind[6], double *coe, const double *vol)
{
    if (pi[FO_ITYPZONE] != 2) {
        FO_LOOP(ind) {
            const int l = fo_inddm(pi, i, j, k), lt = fo_indmtr(pi, i, j, k);
            coe[l] = pr[FO_DTC] / vol[lt];
        }
    } else {
        const double volinv = pr[FO_DTC] / vol[0];
        FO_LOOP(ind) coe[fo_inddm(pi, i, j, k)] = volinv;
    }
}

/* FastS/Compute/cpcfl3.for:6-800, fixed meshes (LALE = 0): viscous models :276-440, Euler :635-790.  Per interior cell
   cfloc = dt / min(vis, conv) (Euler: dt / conv) with conv = 2 vol surf / (|u| + c), vis = 2 rho vol^2 surf^2 (gamma/Pr) / mu,
   surf = 1 / max_d(|t_d(l)| + |t_d(l+d)|); cfl = [max, min, sum] (the caller divides the sum by the cell count,
   computePT.cpp:645). */
void fo_cpcfl3(const int *pi, const double *pr, const int ind[6], double cfl[3], const double *rop, const double *xmut,
               const double *ti, const double *tj, const double *tk, const double *vol)
{
    const int nd = pi[FO_NDIMDX], nm = pi[FO_NDIMDX_MTR], typ = pi[FO_ITYPZONE];
    const int inci = pi[FO_NIJK_MTR], incj = pi[FO_NIJK_MTR + 1], inck = pi[FO_NIJK_MTR + 2];
    const double rgp = pr[FO_CVINF] * (pr[FO_GAMMA] - 1.);
    const double gam1 = pr[FO_GAMMA] / pr[FO_PRANDT], gam2 = pr[FO_GAMMA] * rgp;
    const double dt = pr[FO_DTC];                       /* EXPLOC = 0 */
    const int visc = pi[FO_IFLOW] >= 2;
    double surf_cart = 0.;
    if (typ == 2) {
        const double ddi = 2. * sqrt(ti[0] * ti[0]), ddj = 2. * sqrt(tj[0] * tj[0]), ddk = 2. * sqrt(tk[0] * tk[0]);
        surf_cart = 1. / fmax(fmax(ddi, ddj), ddk);
    }
    FO_LOOP(ind) {
        const int l = fo_inddm(pi, i, j, k), lt = typ == 2 ? 0 : fo_indmtr(pi, i, j, k);
        double surf;
        if (typ == 0) {
            double ddi = sqrt(ti[lt + inci] * ti[lt + inci] + ti[lt + inci + nm] * ti[lt + inci + nm] + ti[lt + inci + 2 * nm] * ti[lt + inci + 2 * nm]);
            ddi = ddi + sqrt(ti[lt] * ti[lt] + ti[lt + nm] * ti[lt + nm] + ti[lt + 2 * nm] * ti[lt + 2 * nm]);
            double ddj = sqrt(tj[lt + incj] * tj[lt + incj] + tj[lt + incj + nm] * tj[lt + incj + nm] + tj[lt + incj + 2 * nm] * tj[lt + incj + 2 * nm]);
            ddj = ddj + sqrt(tj[lt] * tj[lt] + tj[lt + nm] * tj[lt + nm] + tj[lt + 2 * nm] * tj[lt + 2 * nm]);
            double ddk = sqrt(tk[lt + inck] * tk[lt + inck] + tk[lt + inck + nm] * tk[lt + inck + nm] + tk[lt + inck + 2 * nm] * tk[lt + inck + 2 * nm]);
            ddk = ddk + sqrt(tk[lt] * tk[lt] + tk[lt + nm] * tk[lt + nm] + tk[lt + 2 * nm] * tk[lt + 2 * nm]);
            surf = 1. / fmax(fmax(ddi, ddj), ddk);
        } else if (typ == 1) {
            double ddi = sqrt(ti[lt + inci] * ti[lt + inci] + ti[lt + inci + nm] * ti[lt + inci + nm]);
            ddi = ddi + sqrt(ti[lt] * ti[lt] + ti[lt + nm] * ti[lt + nm]);
            double ddj = sqrt(tj[lt + incj] * tj[lt + incj] + tj[lt + incj + nm] * tj[lt + incj + nm]);
            ddj = ddj + sqrt(tj[lt] * tj[lt] + tj[lt + nm] * tj[lt + nm]);
            double ddk = sqrt(tk[lt + inck] * tk[lt + inck]);
            ddk = ddk + sqrt(tk[lt] * tk[lt]);
            surf = 1. / fmax(fmax(ddi, ddj), ddk);
        } else if (typ == 2) surf = surf_cart;
        else {      /* 2-D */
            double ddi = sqrt(ti[lt + inci] * ti[lt + inci] + ti[lt + inci + nm] * ti[lt + inci + nm]);
            ddi = ddi + sqrt(ti[lt] * ti[lt] + ti[lt + nm] * ti[lt + nm]);
            double ddj = sqrt(tj[lt + incj] * tj[lt + incj] + tj[lt + incj + nm] * tj[lt + incj + nm]);
            ddj = ddj + sqrt(tj[lt] * tj[lt] + tj[lt + nm] * tj[lt + nm]);
            surf = 1. / fmax(ddi, ddj);
        }
        const double u = rop[l + nd], v = rop[l + 2 * nd], w = typ == 3 ? 0. : rop[l + 3 * nd], r = rop[l];
        const double c = sqrt(gam2 * rop[l + 4 * nd]);
        const double sp = sqrt(u * u + v * v + w * w);
        const double xinvspc = 1. / (sp + c);
        const double conv = 2. * vol[lt] * surf * xinvspc;
        double comp = conv;
        if (visc) {
            const double xinvkt = gam1 / xmut[l];
            const double vis = 2. * r * vol[lt] * vol[lt] * surf * surf * xinvkt;
            comp = fmin(vis, conv);
        }
        const double cfloc = dt / comp;
        cfl[0] = fmax(cfl[0], cfloc); cfl[1] = fmin(cfl[1], cfloc); cfl[2] = cfl[2] + cfloc;
    }
}

/* FastS/Compute/init_rhs.for:31-129 */
void fo_init_rhs(const int *pi, int nitcfg, const int ind[6], double *drodm)
{
    const double cxsi1 = -0.6830127018922193, cxsi2 = -4. / 3.;
    const int b = pi[FO_RK], neq = pi[FO_NEQ], nd = pi[FO_NDIMDX];
    double coefH = 0.;
    if (b == 3) { if (nitcfg == 2) coefH = cxsi1; else if (nitcfg == 3) coefH = cxsi2; }
    else if (b == 2) { if (nitcfg == 2) coefH = -1.; }
    if (pi[FO_ITYPCP] <= 1 || nitcfg == 1) {
        FO_LOOP(ind) { const int l = fo_inddm(pi, i, j, k); for (int e = 0; e < neq; e++) drodm[l + e * nd] = 0.; }
    } else {
        FO_LOOP(ind) { const int l = fo_inddm(pi, i, j, k); for (int e = 0; e < neq; e++) drodm[l + e * nd] = drodm[l + e * nd] * coefH; }
    }
}

/* FastS/Compute/SENSOR/wiggle2001.for:6 with bodies wiggle.for + wiggle_{i,j,k}.for.
   Fortran sign(1.,x) follows the sign bit of x (gfortran), hence copysign. */
static double wiggle_face(const double *rop, int nd, int l, int inc)
{
    const int nm = l - inc, nm2 = l - 2 * inc, np = l + inc;
    double test = 1e30;
    for (int e = 0; e < 5; e++) {
        const double *q = rop + e * nd;
        const double qm = (q[nm] - q[nm2]) * (q[l] - q[nm]);
        const double qp = (q[l] - q[nm]) * (q[np] - q[l]);
        const double f = copysign(1., qm) + copysign(1., qp);
        if (f < test) test = f;
    }
    return (test < -1e-12) ? 1. : 0.;
}

void fo_wiggle2001(const int *pi, const double *pr, const int ind[6], const double *rop, double *wig)
{
    (void)pr;
    const int nd = pi[FO_NDIMDX];
    const int inci = 1, incj = pi[FO_NIJK], inck = pi[FO_NIJK] * pi[FO_NIJK + 1];
    const int icorr = ind[1] == pi[FO_IJKV], jcorr = ind[3] == pi[FO_IJKV + 1], kcorr = ind[5] == pi[FO_IJKV + 2];
    FO_LOOP(ind) {
        const int l = fo_inddm(pi, i, j, k);
        wig[l + 2 * nd] = wiggle_face(rop, nd, l, inck);
        wig[l + nd] = wiggle_face(rop, nd, l, incj);
        wig[l] = wiggle_face(rop, nd, l, inci);
    }
    if (icorr)
        for (int k = ind[4]; k <= ind[5]; k++) for (int j = ind[2]; j <= ind[3]; j++) {
            const int l = fo_inddm(pi, ind[1] + 1, j, k); wig[l] = wiggle_face(rop, nd, l, inci); }
    if (jcorr)
        for (int k = ind[4]; k <= ind[5]; k++) for (int i = ind[0]; i <= ind[1]; i++) {
            const int l = fo_inddm(pi, i, ind[3] + 1, k); wig[l + nd] = wiggle_face(rop, nd, l, incj); }
    if (kcorr)
        for (int j = ind[2]; j <= ind[3]; j++) for (int i = ind[0]; i <= ind[1]; i++) {
            const int l = fo_inddm(pi, i, j, ind[5] + 1); wig[l + 2 * nd] = wiggle_face(rop, nd, l, inck); }
}

/* ------------------------------------------------------------------------------------ */
/* Flux balance: FastS/Compute/template_FluxAndBalance.for:5-272                          */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    const int *pi; const double *rop, *wig, *xmut, *vol;
    const double *t[3];           /* ti, tj, tk                                   */
    int inc[3], incm[3];          /* cell strides, metric strides                  */
    int nd, nm, typ, iflow, kflu, slope, neq;
    double gam, rgp, gam1, gam3, gam4, cmus1, temp01, coesut, sigma_1;
    double c1, c2, c3, opt0, c4, c5, c6, cvisq, wig_cte, psiroe, tmin_1;
    double tc0[3], volinv0;       /* 3dcart constants (template:112-119)           */
} fctx;

static void fctx_init(fctx *c, const int *pi, const double *pr, const double *rop, const double *wig,
                      const double *ti, const double *tj, const double *tk, const double *vol,
                      const double *xmut)
{
    c->pi = pi; c->rop = rop; c->wig = wig; c->xmut = xmut; c->vol = vol;
    c->t[0] = ti; c->t[1] = tj; c->t[2] = tk;
    c->inc[0] = 1; c->inc[1] = pi[FO_NIJK]; c->inc[2] = pi[FO_NIJK] * pi[FO_NIJK + 1];
    c->incm[0] = pi[FO_NIJK_MTR]; c->incm[1] = pi[FO_NIJK_MTR + 1]; c->incm[2] = pi[FO_NIJK_MTR + 2];
    c->nd = pi[FO_NDIMDX]; c->nm = pi[FO_NDIMDX_MTR]; c->typ = pi[FO_ITYPZONE];
    c->iflow = pi[FO_IFLOW]; c->kflu = pi[FO_KFLUDOM]; c->slope = pi[FO_SLOPE]; c->neq = pi[FO_NEQ];
    /* template:121-153 */
    c->gam = pr[FO_GAMMA];
    c->rgp = pr[FO_CVINF] * (c->gam - 1.);
    c->gam1 = c->gam / (c->gam - 1.);
    c->gam3 = c->gam1 / pr[FO_PRANDT] * c->rgp;
    c->gam4 = c->gam1 / pr[FO_PRANDT_TUR] * c->rgp;
    c->cmus1 = pr[FO_CS]; c->temp01 = 1. / pr[FO_TEMP0];
    c->coesut = pr[FO_XMUL0] * (1. + c->cmus1 * c->temp01);
    c->sigma_1 = 1. / FO_SA_SIGMA;
    const double roref = pr[FO_ROINF], uref = pr[FO_VINF];
    c->wig_cte = pr[FO_WIG_AMP]; c->psiroe = pr[FO_PSIROE];
    c->tmin_1 = 100. / pr[FO_TINF];
    c->c1 = 0.02 * uref; c->c2 = 0.02 / (uref * roref); c->c3 = -2.;
    c->opt0 = (double)pi[FO_SENSORTYPE];
    c->c6 = 1. / 6.; c->c4 = 5. * c->c6; c->c5 = 2. * c->c6; c->c6 = -1. * c->c6;
    c->cvisq = 1. / 3;
    c->tc0[0] = ti[0]; c->tc0[1] = tj[0]; c->tc0[2] = (pi[FO_NEQ_K] > 0) ? tk[0] : 0.;
    c->volinv0 = 0.5 / vol[0];
}

static inline double avmin(double c, double r)
{   /* template:108: sign(1.,c)*max(0.,min(abs(c),sign(1.,c)*r)) */
    const double s = copysign(1., c);
    return s * fmax(0., fmin(fabs(c), s * r));
}

/* laminar part of the viscous face flux for the three mesh families; adds into flu[1..4].
   Returns the face gradient of nutilde-independent quantities needed by the SA flux via g[]. */
static void visc_face(const fctx *c, int d0, int l, int lt, const double tc[3], double *flu,
                      double *xmutvol_out, double *volinv_out)
{
    const double *rop = c->rop; const int nd = c->nd;
    const int d1 = (d0 + 1) % 3, d2 = (d0 + 2) % 3;
    const int i0 = c->inc[d0], i1 = c->inc[d1], i2 = c->inc[d2];
    const int il = l - i0, ir = l;
    double G[5][3] = {{0}};  /* G[c][e]: d(u,v,w,T[,nutilde])/d(x,y,z) scaled (face gradient * 2 vol) */
    const int nv = c->iflow == 3 ? 5 : 4;   /* SA: nutilde gradient with the same metrics (SA/fluvisq_SA_*.for) */
    double volinv;
    if (c->typ == 2) {
        /* fluvisq_3dcart_{i,j,k}.for */
        volinv = c->volinv0;
        for (int cc = 0; cc < nv; cc++) {
            const double *q = rop + (cc + 1) * nd;
            const double f1 = q[il], f2 = q[ir], fv = f1 + f2;
            G[cc][d0] = (f2 - f1) * c->tc0[d0];
            if (cc >= 3) continue;
            for (int t = 0; t < 2; t++) {
                const int e = t ? d2 : d1, ie = t ? i2 : i1;
                if (!(cc == d0 || cc == e)) continue;
                const double fa = fv + q[il - ie] + q[ir - ie];
                const double fb = fv + q[il + ie] + q[ir + ie];
                G[cc][e] = (fb - fa) * (.25 * c->tc0[e]);
            }
        }
    } else if (c->typ == 0) {
        /* fluvisq_3dfull_{i,j,k}.for:34-130, the three files being cyclic permutations */
        const int m0 = c->incm[d0], m1 = c->incm[d1], m2 = c->incm[d2];
        const double *T0 = c->t[d0], *T1 = c->t[d1], *T2 = c->t[d2];
        double tn[3], tn1[3], ta[3], ta1[3], tb[3], tb1[3];
        for (int e = 0; e < 3; e++) {
            const int o = e * c->nm;
            tn[e] = .5 * (T0[lt + o] + T0[lt - m0 + o]);
            tn1[e] = .5 * (T0[lt + o] + T0[lt + m0 + o]);
            ta[e] = .125 * (T1[lt + o] + T1[lt - m0 + o]);
            ta1[e] = .125 * (T1[lt + m1 + o] + T1[lt + m1 - m0 + o]);
            tb[e] = .125 * (T2[lt + o] + T2[lt - m0 + o]);
            tb1[e] = .125 * (T2[lt + m2 + o] + T2[lt + m2 - m0 + o]);
        }
        for (int cc = 0; cc < nv; cc++) {
            const double *q = rop + (cc + 1) * nd;
            const double f1 = q[il], f2 = q[ir];
            const double f3 = q[ir] + q[il] + q[il - i1] + q[ir - i1];
            const double f4 = q[ir] + q[il] + q[il + i1] + q[ir + i1];
            const double f5 = q[ir] + q[il] + q[il - i2] + q[ir - i2];
            const double f6 = q[ir] + q[il] + q[il + i2] + q[ir + i2];
            for (int e = 0; e < 3; e++)
                G[cc][e] = (f2 * tn1[e] - f1 * tn[e]) + (f4 * ta1[e] - f3 * ta[e]) + (f6 * tb1[e] - f5 * tb[e]);
        }
        volinv = 1. / (c->vol[lt] + c->vol[lt - m0]);
    } else {
        /* fluvisq_3dhomo_{i,j,k}.for */
        const int nm = c->nm;
        if (d0 < 2) {
            const int dp = 1 - d0, ip = c->inc[dp], ik = c->inc[2];
            const int m0 = c->incm[d0], mp = c->incm[dp];
            const double *T0 = c->t[d0], *Tp = c->t[dp], *Tk = c->t[2];
            double tn[2], tn1[2], tp[2], tp1[2];
            for (int e = 0; e < 2; e++) {
                const int o = e * nm;
                tn[e] = .5 * (T0[lt + o] + T0[lt - m0 + o]);
                tn1[e] = .5 * (T0[lt + o] + T0[lt + m0 + o]);
                tp[e] = .125 * (Tp[lt + o] + Tp[lt - m0 + o]);
                tp1[e] = .125 * (Tp[lt + mp + o] + Tp[lt + mp - m0 + o]);
            }
            const double tkz = .125 * (Tk[lt] + Tk[lt - m0]);
            for (int cc = 0; cc < nv; cc++) {
                const double *q = rop + (cc + 1) * nd;
                const double f1 = q[il], f2 = q[ir];
                const double fa = q[ir] + q[il] + q[il - ip] + q[ir - ip];
                const double fb = q[ir] + q[il] + q[il + ip] + q[ir + ip];
                for (int e = 0; e < 2; e++)
                    G[cc][e] = (f2 * tn1[e] - f1 * tn[e]) + (fb * tp1[e] - fa * tp[e]);
                if (cc < 3) {
                    const double fk0 = q[ir] + q[il] + q[il - ik] + q[ir - ik];
                    const double fk1 = q[ir] + q[il] + q[il + ik] + q[ir + ik];
                    G[cc][2] = (fk1 - fk0) * tkz;
                }
            }
            volinv = 1. / (c->vol[lt] + c->vol[lt - m0]);
        } else {
            const int ii = c->inc[0], ij = c->inc[1];
            const int mi = c->incm[0], mj = c->incm[1];
            const double tiz = c->t[2][lt];
            const double tjx = 0.25 * c->t[0][lt], tjy = 0.25 * c->t[0][lt + nm];
            const double tjx1 = 0.25 * c->t[0][lt + mi], tjy1 = 0.25 * c->t[0][lt + mi + nm];
            const double tkx = 0.25 * c->t[1][lt], tky = 0.25 * c->t[1][lt + nm];
            const double tkx1 = 0.25 * c->t[1][lt + mj], tky1 = 0.25 * c->t[1][lt + mj + nm];
            for (int cc = 0; cc < nv; cc++) {
                const double *q = rop + (cc + 1) * nd;
                const double f1 = q[il], f2 = q[ir];
                G[cc][2] = (f2 - f1) * tiz;
                if (cc >= 3) continue;
                const double f3 = q[ir] + q[il] + q[il - ii] + q[ir - ii];
                const double f4 = q[ir] + q[il] + q[il + ii] + q[ir + ii];
                const double f5 = q[ir] + q[il] + q[il - ij] + q[ir - ij];
                const double f6 = q[ir] + q[il] + q[il + ij] + q[ir + ij];
                G[cc][0] = ((f4 * tjx1 - f3 * tjx) + (f6 * tkx1 - f5 * tkx));
                G[cc][1] = ((f4 * tjy1 - f3 * tjy) + (f6 * tky1 - f5 * tky));
            }
            volinv = 1. / (c->vol[lt] + c->vol[lt]);   /* lt002 = lt (fluvisq_3dhomo_k.for:8) */
        }
    }
    const double div = (G[0][0] + G[1][1] + G[2][2]) * c->cvisq;
    const double f1 = G[0][0] - div, f2 = G[0][1] + G[1][0], f3 = G[0][2] + G[2][0];
    const double f4 = G[1][1] - div, f5 = G[1][2] + G[2][1], f6 = G[2][2] - div;
    /* mut_interface.for / mut_prandtltb_interface.for (SA: generate_flu.py:216-227) */
    double xmutvol, xktvol;
    if (c->iflow == 3) {
        double xmulam = c->coesut * sqrt(rop[il + 4 * nd] * c->temp01) / (1. + c->cmus1 / rop[il + 4 * nd]);
        xmulam = xmulam + c->coesut * sqrt(rop[ir + 4 * nd] * c->temp01) / (1. + c->cmus1 / rop[ir + 4 * nd]);
        const double xmutot = c->xmut[ir] + c->xmut[il];
        const double xmutur = xmutot - xmulam;
        xktvol = (xmulam * c->gam3 + xmutur * c->gam4) * volinv;
        xmutvol = xmutot * volinv;
    } else {
        xmutvol = (c->xmut[ir] + c->xmut[il]) * volinv;
        xktvol = xmutvol * c->gam3;
    }
    const double ur = rop[ir + nd] + rop[il + nd], vr = rop[ir + 2 * nd] + rop[il + 2 * nd],
                 wr = rop[ir + 3 * nd] + rop[il + 3 * nd];
    double fv, fv5;
    if (c->typ == 0) {
        fv = -(2. * f1 * tc[0] + f2 * tc[1] + f3 * tc[2]) * xmutvol; fv5 = fv * ur; flu[1] += fv;
        fv = -(f2 * tc[0] + 2. * f4 * tc[1] + f5 * tc[2]) * xmutvol; fv5 = fv5 + fv * vr; flu[2] += fv;
        fv = -(f3 * tc[0] + f5 * tc[1] + 2. * f6 * tc[2]) * xmutvol; fv5 = fv5 + fv * wr; flu[3] += fv;
        fv = fv5 * 0.5 - (G[3][0] * tc[0] + G[3][1] * tc[1] + G[3][2] * tc[2]) * xktvol; flu[4] += fv;
    } else if (c->typ == 1 && d0 < 2) {
        fv = -(2. * f1 * tc[0] + f2 * tc[1]) * xmutvol; fv5 = fv * ur; flu[1] += fv;
        fv = -(f2 * tc[0] + 2. * f4 * tc[1]) * xmutvol; fv5 = fv5 + fv * vr; flu[2] += fv;
        fv = -(f3 * tc[0] + f5 * tc[1]) * xmutvol; fv5 = fv5 + fv * wr; flu[3] += fv;
        fv = fv5 * 0.5 - (G[3][0] * tc[0] + G[3][1] * tc[1]) * xktvol; flu[4] += fv;
    } else {
        /* single non-zero normal component tc[d0] (3dcart all faces, 3dhomo k face) */
        const double t0 = tc[d0];
        const double fu = (d0 == 0) ? 2. * f1 : (d0 == 1 ? f2 : f3);
        const double fvv = (d0 == 0) ? f2 : (d0 == 1 ? 2. * f4 : f5);
        const double fw = (d0 == 0) ? f3 : (d0 == 1 ? f5 : 2. * f6);
        fv = -(fu * t0) * xmutvol; fv5 = fv * ur; flu[1] += fv;
        fv = -(fvv * t0) * xmutvol; fv5 = fv5 + fv * vr; flu[2] += fv;
        fv = -(fw * t0) * xmutvol; fv5 = fv5 + fv * wr; flu[3] += fv;
        fv = fv5 * 0.5 - (G[3][d0] * t0) * xktvol; flu[4] += fv;
    }
    if (c->iflow == 3) {
        /* SA/mut_interface.for + SA/fluvisq_SA_<mesh>_<dir>.for / assemble_tauij_3dfull.for */
        double xmulam = c->coesut * sqrt(rop[il + 4 * nd] * c->temp01) / (1. + c->cmus1 / rop[il + 4 * nd]);
        xmulam = xmulam + c->coesut * sqrt(rop[ir + 4 * nd] * c->temp01) / (1. + c->cmus1 / rop[ir + 4 * nd]);
        double xsa = xmulam + rop[il + 5 * nd] * rop[il] + rop[ir + 5 * nd] * rop[ir];
        xsa = xsa * volinv * c->sigma_1;
        if (c->typ == 0) flu[5] = flu[5] - xsa * (G[4][0] * tc[0] + G[4][1] * tc[1] + G[4][2] * tc[2]);
        else if (c->typ == 1 && d0 < 2) flu[5] = flu[5] - xsa * (G[4][0] * tc[0] + G[4][1] * tc[1]);
        else flu[5] = flu[5] - xsa * G[4][d0] * tc[d0];
    }
    *xmutvol_out = xmutvol; *volinv_out = volinv;
}

/* Flux through the lower face of cell l in direction d0 (0 i, 1 j, 2 k). */
static void face_flux(const fctx *c, int d0, int l, int lt, double *flu)
{
    const double *rop = c->rop; const int nd = c->nd; const int inc = c->inc[d0];
    /* Normale/normale_<mesh>_<dir>.for */
    double tc[3] = {0., 0., 0.}, s;
    if (c->typ == 0) {
        const double *T = c->t[d0];
        tc[0] = T[lt]; tc[1] = T[lt + c->nm]; tc[2] = T[lt + 2 * c->nm];
        s = sqrt(tc[0] * tc[0] + tc[1] * tc[1] + tc[2] * tc[2]);
    } else if (c->typ == 1) {
        if (d0 < 2) { const double *T = c->t[d0]; tc[0] = T[lt]; tc[1] = T[lt + c->nm]; s = sqrt(tc[0] * tc[0] + tc[1] * tc[1]); }
        else { tc[2] = c->t[2][lt]; s = fabs(tc[2]); }
    } else { tc[d0] = c->tc0[d0]; s = fabs(tc[d0]); }

    const int nm = l - inc, nm2 = l - 2 * inc, np = l + inc;
    double qm[6], qp[6];   /* qm: right state, qp: left state */
    const int nvar = 5;
    for (int e = 0; e < nvar; e++) {
        const double *q = rop + e * nd;
        if (c->slope == 2) {            /* Slope/o3_slope_var.for */
            qm[e] = c->c4 * q[l] + c->c5 * q[nm] + c->c6 * q[np];
            qp[e] = c->c4 * q[nm] + c->c6 * q[nm2] + c->c5 * q[l];
        } else if (c->slope == 3) {     /* Slope/minmod_slope_var.for */
            const double roff = .5 * fmax(0., fmin(rop[l + 4 * nd] * c->tmin_1 - .5, 1.));
            const double delm = q[l] - q[nm], delp = q[np] - q[l], delq = q[nm] - q[nm2];
            const double slp = roff * avmin(delm, delp), slq = roff * avmin(delq, delm);
            qm[e] = q[l] - slp; qp[e] = q[nm] + slq;
        } else {                        /* Slope/o1_slope_var.for */
            qm[e] = q[l]; qp[e] = q[nm];
        }
    }
    const double qp1 = qp[0], qp2 = qp[1], qp3 = qp[2], qp4 = qp[3], qp5 = qp[4];
    const double qm1 = qm[0], qm2 = qm[1], qm3 = qm[2], qm4 = qm[3], qm5 = qm[4];
    const double rgp = c->rgp, gam1 = c->gam1;
    double u_conv = 0.;   /* convective velocity re-used by the SA face flux (fluFaceSA.for) */
    double qn1, qn2;
    if (c->kflu == 5) {
        /* etat_roe_GD.for + Vit_ent/fludiffer_3dfull_i.for (3dcart/3dhomo = zero components) */
        const double r1 = qp1, p1 = r1 * qp5 * rgp;
        const double h1 = gam1 * qp5 * rgp + .5 * (qp2 * qp2 + qp3 * qp3 + qp4 * qp4);
        const double r2 = qm1, p2 = r2 * qm5 * rgp;
        const double h2 = gam1 * qm5 * rgp + .5 * (qm2 * qm2 + qm3 * qm3 + qm4 * qm4);
        const double sq1 = sqrt(r1), sq2 = sqrt(r2);
        const double r = sq1 * sq2, r_1 = 1. / (sq1 + sq2);
        const double u = (sq1 * qp2 + sq2 * qm2) * r_1, v = (sq1 * qp3 + sq2 * qm3) * r_1,
                     w = (sq1 * qp4 + sq2 * qm4) * r_1, h = (sq1 * h1 + sq2 * h2) * r_1;
        const double q = .5 * (u * u + v * v + w * w);
        const double cc = sqrt((c->gam - 1.) * fabs(h - q));
        qn1 = qp2 * tc[0] + qp3 * tc[1] + qp4 * tc[2];
        qn2 = qm2 * tc[0] + qm3 * tc[1] + qm4 * tc[2];
        const double du = (qm2 - qp2) * s, dv = (qm3 - qp3) * s, dw = (qm4 - qp4) * s;
        const double dp = p2 - p1, dqn = qn2 - qn1;
        const double s_1 = 1.0 / fmax(s, 1.E-30);
        const double nx = tc[0] * s_1, ny = tc[1] * s_1, nz = tc[2] * s_1;
        const double qn = u * nx + v * ny + w * nz;
        double f1 = qn;
        const double f4 = c->psiroe * (fabs(f1) + cc);
        double f2 = 0.5 + copysign(0.5, fabs(f1) - f4);
        double f3 = (f2 - 1.0) * (f1 * f1 / f4 + f4) * 0.5 - f2 * fabs(f1);
        f1 = (r2 - r1 - dp / (cc * cc)) * s;
        flu[0] = f3 * (f1);
        flu[1] = f3 * (f1 * u + r * (du - nx * dqn));
        flu[2] = f3 * (f1 * v + r * (dv - ny * dqn));
        flu[3] = f3 * (f1 * w + r * (dw - nz * dqn));
        flu[4] = f3 * (f1 * q + r * (u * du + v * dv + w * dw - qn * dqn));
        f1 = qn + cc;
        f2 = 0.5 + copysign(0.5, fabs(f1) - f4);
        f3 = (f2 - 1.) * (f1 * f1 / f4 + f4) * 0.5 - f2 * fabs(f1);
        f1 = (dp + r * cc * dqn * s_1) / (2. * cc * cc) * s * f3;
        flu[0] += f1; flu[1] += f1 * (u + nx * cc); flu[2] += f1 * (v + ny * cc);
        flu[3] += f1 * (w + nz * cc); flu[4] += f1 * (h + qn * cc);
        f1 = qn - cc;
        f2 = 0.5 + copysign(0.5, fabs(f1) - f4);
        f3 = (f2 - 1.) * (f1 * f1 / f4 + f4) * 0.5 - f2 * fabs(f1);
        f1 = (dp - r * cc * dqn * s_1) / (2. * cc * cc) * f3 * s;
        flu[0] += f1; flu[1] += f1 * (u - nx * cc); flu[2] += f1 * (v - ny * cc);
        flu[3] += f1 * (w - nz * cc); flu[4] += f1 * (h - qn * cc);
        const double rou1 = r1 * qn1, rou2 = r2 * qn2, p1p2 = p1 + p2;
        flu[0] = (flu[0] + rou1 + rou2) * 0.5;
        flu[1] = (flu[1] + rou1 * qp2 + rou2 * qm2 + tc[0] * p1p2) * 0.5;
        flu[2] = (flu[2] + rou1 * qp3 + rou2 * qm3 + tc[1] * p1p2) * 0.5;
        flu[3] = (flu[3] + rou1 * qp4 + rou2 * qm4 + tc[2] * p1p2) * 0.5;
        flu[4] = (flu[4] + rou1 * h1 + rou2 * h2) * 0.5;
    } else {
        /* etat_GD.for */
        const double r1 = qp1, rou1 = r1 * qp2, rov1 = r1 * qp3, row1 = r1 * qp4, p1 = r1 * qp5 * rgp;
        const double h1 = gam1 * p1 + .5 * (rou1 * qp2 + rov1 * qp3 + row1 * qp4);
        const double r2 = qm1, rou2 = r2 * qm2, rov2 = r2 * qm3, row2 = r2 * qm4, p2 = r2 * qm5 * rgp;
        const double h2 = gam1 * p2 + .5 * (rou2 * qm2 + rov2 * qm3 + row2 * qm4);
        qn1 = qp2 * tc[0] + qp3 * tc[1] + qp4 * tc[2];
        qn2 = qm2 * tc[0] + qm3 * tc[1] + qm4 * tc[2];
        double u, tdu;
        if (c->kflu == 1) {   /* AUSM/fluFaceEuler.for:38-57 */
            const double cs = rgp * c->gam * rop[l + 4 * nd];
            const double son = sqrt(qn1 * qn1 / cs);
            const double tam = c->c3 * son + s;
            const double tam1 = fmax(0., tam) * c->c2;
            u = 0.25 * (qn1 + qn2) - tam1 * (p2 - p1);
            tdu = fmax(fabs(u), c->c1 * s) * c->wig_cte;
        } else {              /* SENSOR/fluFaceEuler.for:46-55 */
            const double wg = c->wig[l + d0 * nd];
            u = 0.25 * (qn1 + qn2) - c->c2 * s * (p2 - p1) * (c->opt0 * wg + 1. - c->opt0);
            tdu = fmax(fabs(u), c->c1 * s) * wg;
        }
        const double p1p2 = (p1 + p2) * 0.5;
        /* Vit_ent/fluvector_*.for */
        flu[0] = u * (r1 + r2) - tdu * (r2 - r1);
        flu[1] = u * (rou1 + rou2) - tdu * (rou2 - rou1) + tc[0] * p1p2;
        flu[2] = u * (rov1 + rov2) - tdu * (rov2 - rov1) + tc[1] * p1p2;
        flu[3] = u * (row1 + row2) - tdu * (row2 - row1) + tc[2] * p1p2;
        flu[4] = u * (h1 + h2) - tdu * (h2 - h1);
        u_conv = u; (void)tdu;
        if (c->neq == 6) {
            /* AUSM|SENSOR/fluFaceSA.for: nutilde reconstructed with the scheme's slope, times the MUSCL densities */
            const double *q = rop + 5 * nd; double qm6, qp6;
            if (c->slope == 2) { qm6 = c->c4 * q[l] + c->c5 * q[nm] + c->c6 * q[np]; qp6 = c->c4 * q[nm] + c->c6 * q[nm2] + c->c5 * q[l]; }
            else if (c->slope == 3) {
                const double roff = .5 * fmax(0., fmin(rop[l + 4 * nd] * c->tmin_1 - .5, 1.));
                const double delm = q[l] - q[nm], delp = q[np] - q[l], delq = q[nm] - q[nm2];
                qm6 = q[l] - roff * avmin(delm, delp); qp6 = q[nm] + roff * avmin(delq, delm);
            } else { qm6 = q[l]; qp6 = q[nm]; }
            qm6 = qm6 * qm1; qp6 = qp6 * qp1;
            flu[5] = u * (qm6 + qp6) - tdu * (qm6 - qp6);
        }
    }
    if (c->kflu == 5 && c->neq == 6) {
        /* ROE/fluFaceSA.for:2-34: first-order states, Roe-averaged normal velocity as dissipation speed */
        double p2 = rop[nm + nd], p3 = rop[nm + 2 * nd], p4 = rop[nm + 3 * nd];
        const double m2 = rop[l + nd], m3 = rop[l + 2 * nd], m4 = rop[l + 3 * nd];
        double n1 = p2 * tc[0] + p3 * tc[1] + p4 * tc[2];
        const double n2 = m2 * tc[0] + m3 * tc[1] + m4 * tc[2];
        double rp1 = rop[nm], rm1 = rop[l];
        const double qm6 = rop[l + 5 * nd] * rm1 * n2, qp6 = rop[nm + 5 * nd] * rp1 * n1;
        rp1 = sqrt(rp1); rm1 = sqrt(rm1);
        const double r_1 = 1. / (rp1 + rm1);
        p2 = r_1 * (rp1 * p2 + rm1 * m2); p3 = r_1 * (rp1 * p3 + rm1 * m3); p4 = r_1 * (rp1 * p4 + rm1 * m4);
        n1 = p2 * tc[0] + p3 * tc[1] + p4 * tc[2];
        const double s_1 = 1.0 / fmax(s, 1.E-30);
        const double tdu = fmax(c->c1 / s_1, fabs(n1));
        flu[5] = 0.5 * (qm6 + qp6 - tdu * (rop[l + 5 * nd] * rop[l] - rop[nm + 5 * nd] * rop[nm]));
    }
    (void)u_conv;
    if (c->iflow >= 2) {
        double xmutvol, volinv;
        visc_face(c, d0, l, lt, tc, flu, &xmutvol, &volinv);
    }
}

void fo_flux_balance(const int *pi, const double *pr, const int ind[6], int mask,
                     const double *rop, double *drodm, const double *wig,
                     const double *ti, const double *tj, const double *tk,
                     const double *vol, const double *xmut)
{
    if (ind[0] > ind[1] || ind[2] > ind[3] || ind[4] > ind[5]) return;
    fctx c; fctx_init(&c, pi, pr, rop, wig, ti, tj, tk, vol, xmut);
    const int neq = c.neq, nd = c.nd;
    double flu[6];
    const int is2d = (c.typ == 3);
    for (int k = ind[4]; k <= ind[5]; k++) {
        for (int j = ind[2]; j <= ind[3]; j++) {
            if (!is2d)
                for (int i = ind[0]; i <= ind[1]; i++) {
                    const int l = fo_inddm(pi, i, j, k), lt = c.typ == 2 ? 0 : fo_indmtr(pi, i, j, k), l0 = l - c.inc[2];
                    face_flux(&c, 2, l, lt, flu);
                    const int skip = (mask & 1) && k == ind[4];
                    for (int e = 0; e < neq; e++) { drodm[l + e * nd] += flu[e]; if (!skip) drodm[l0 + e * nd] -= flu[e]; }
                }
            for (int i = ind[0]; i <= ind[1]; i++) {
                const int l = fo_inddm(pi, i, j, k), lt = c.typ == 2 ? 0 : fo_indmtr(pi, i, j, k), l0 = l - c.inc[1];
                face_flux(&c, 1, l, lt, flu);
                for (int e = 0; e < neq; e++) { drodm[l + e * nd] += flu[e]; drodm[l0 + e * nd] -= flu[e]; }
            }
            for (int i = ind[0]; i <= ind[1]; i++) {
                const int l = fo_inddm(pi, i, j, k), lt = c.typ == 2 ? 0 : fo_indmtr(pi, i, j, k), l0 = l - c.inc[0];
                face_flux(&c, 0, l, lt, flu);
                for (int e = 0; e < neq; e++) { drodm[l + e * nd] += flu[e]; drodm[l0 + e * nd] -= flu[e]; }
            }
            {   /* icorr: missing i flux at imax+1 (template:222-233) */
                const int i = ind[1] + 1;
                const int l = fo_inddm(pi, i, j, k), lt = c.typ == 2 ? 0 : fo_indmtr(pi, i, j, k), ls = l - c.inc[0];
                face_flux(&c, 0, l, lt, flu);
                for (int e = 0; e < neq; e++) { drodm[ls + e * nd] -= flu[e]; drodm[l + e * nd] += flu[e]; }
            }
        }
        {   /* jcorr (template:236-250) */
            const int j = ind[3] + 1;
            for (int i = ind[0]; i <= ind[1]; i++) {
                const int l = fo_inddm(pi, i, j, k), lt = c.typ == 2 ? 0 : fo_indmtr(pi, i, j, k), ls = l - c.inc[1];
                face_flux(&c, 1, l, lt, flu);
                for (int e = 0; e < neq; e++) { drodm[ls + e * nd] -= flu[e]; drodm[l + e * nd] += flu[e]; }
            }
        }
    }
    if (!is2d) {   /* kcorr (template:255-270) */
        const int k = ind[5] + 1;
        for (int j = ind[2]; j <= ind[3]; j++)
            for (int i = ind[0]; i <= ind[1]; i++) {
                const int l = fo_inddm(pi, i, j, k), lt = c.typ == 2 ? 0 : fo_indmtr(pi, i, j, k), ls = l - c.inc[2];
                face_flux(&c, 2, l, lt, flu);
                for (int e = 0; e < neq; e++) { drodm[ls + e * nd] -= flu[e]; if (!(mask & 2)) drodm[l + e * nd] += flu[e]; }
            }
    }
}

/* ------------------------------------------------------------------------------------ */
/* Wall flux correction of the Roe scheme. FastS/BC/correct_flux.for:6-190 (lapply = 4),  */
/* FastS/Compute/template_correction.for:252-300 + assemble_drodm_corrflu.for (the interior */
/* Euler flux of the wall face is taken back out of the balance), FastS/BC/                 */
/* CL_correction_flu.for -> bflwall.for:96-131 / bflwallslip.for:104-190 (wall pressure     */
/* flux put in).  Non-ALE (qen = 0), npass = 0, whole-zone sub-domain.                      */
/* ------------------------------------------------------------------------------------ */
static int bc_window(const int *pi, int idir, const int *fen, const int ind_sdm[6], int out[6]);

void fo_correct_flux(const int *pi, const double *pr, const double *rop, double *drodm, const double *wig,
                     const double *ti, const double *tj, const double *tk, const double *vol, const double *xmut)
{
    if (pi[FO_KFLUDOM] != 5) return;
    fctx c; fctx_init(&c, pi, pr, rop, wig, ti, tj, tk, vol, xmut);
    c.iflow = 1;                                   /* iflow = 1: only the Euler flux is corrected (correct_flux.for:136) */
    const int nd = c.nd, neq = pi[FO_NEQ];
    const int pt_bcs = pi[FO_PT_BC], nb_bc = pi[pt_bcs];
    const int ind_sdm[6] = {1, pi[FO_IJKV], 1, pi[FO_IJKV + 1], 1, pi[FO_IJKV + 2]};
    const double gam2 = pr[FO_GAMMA] / (pr[FO_GAMMA] - 1.);
    for (int ndf = 0; ndf < nb_bc; ndf++) {
        const int pt_bc = pi[pt_bcs + 1 + ndf];
        const int ty = pi[pt_bc + FO_BC_TYPE], idir = pi[pt_bc + FO_BC_IDIR];
        const int wall = (ty == 4 || ty == 6 || ty == 7 || ty == 12), slip = (ty == 3);
        if (!wall && !slip) continue;
        int w[6];
        if (bc_window(pi, idir, pi + pt_bc + FO_BC_FEN, ind_sdm, w)) continue;
        const int d = (idir - 1) / 2, inc = c.inc[d];
        if (idir % 2 == 1) { w[2 * d] = 1; w[2 * d + 1] = 1; } else w[2 * d + 1] = w[2 * d];   /* correct_flux.for:116-131 */
        const double sens_c = (idir % 2 == 0) ? 1. : -1.;      /* template_correction.for:196-201 */
        const double sens_b = (idir % 2 == 0) ? -1. : 1.;      /* bflwall.for:81-86 */
        const int shift = (idir % 2 == 0) ? inc : 0;
        for (int k = w[4]; k <= w[5]; k++) for (int j = w[2]; j <= w[3]; j++) for (int i = w[0]; i <= w[1]; i++) {
            const int l = fo_inddm(pi, i, j, k), lt = c.typ == 2 ? 0 : fo_indmtr(pi, i, j, k), l0 = l - shift, ia = l - inc;
            double flu[6];
            face_flux(&c, d, l, lt, flu);
            for (int e = 0; e < 5; e++) drodm[l0 + e * nd] += flu[e] * sens_c;
            double tc[3] = {0., 0., 0.};
            if (c.typ == 0) { const double *T = c.t[d]; tc[0] = T[lt]; tc[1] = T[lt + c.nm]; tc[2] = T[lt + 2 * c.nm]; }
            else if (c.typ == 1) { if (d < 2) { const double *T = c.t[d]; tc[0] = T[lt]; tc[1] = T[lt + c.nm]; } else tc[2] = c.t[2][lt]; }
            else tc[d] = c.tc0[d];
            double r = 0.5 * (rop[l] + rop[ia]), u = 0.5 * (rop[l + nd] + rop[ia + nd]);
            double v = 0.5 * (rop[l + 2 * nd] + rop[ia + 2 * nd]), ww = 0.5 * (rop[l + 3 * nd] + rop[ia + 3 * nd]);
            const double u_int = tc[0] * u + tc[1] * v + tc[2] * ww;
            double f[6] = {0., 0., 0., 0., 0., 0.};
            if (wall) {
                const double p = 0.5 * (rop[l + 4 * nd] * rop[l] + rop[ia + 4 * nd] * rop[ia]) * c.rgp;
                f[1] = tc[0] * p + u_int * u * r; f[2] = tc[1] * p + u_int * v * r; f[3] = tc[2] * p + u_int * ww * r;
            } else {
                if (neq == 6) { r = rop[ia]; u = rop[ia + nd]; v = rop[ia + 2 * nd]; ww = rop[ia + 3 * nd]; }
                double p = rop[ia + 4 * nd] * rop[ia] * c.rgp;
                const double h = gam2 * p + r * (u * u + v * v + ww * ww) * 0.5;
                p = 0.5 * (rop[l + 4 * nd] * rop[l] + rop[ia + 4 * nd] * rop[ia]) * c.rgp;
                f[0] = u_int * r;
                f[1] = tc[0] * p + u_int * r * u; f[2] = tc[1] * p + u_int * r * v; f[3] = tc[2] * p + u_int * r * ww;
                f[4] = u_int * h;
                if (neq == 6) f[5] = u_int * r * rop[ia + 5 * nd];
            }
            for (int e = 0; e < neq; e++) drodm[l0 + e * nd] += f[e] * sens_b;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* FastS/Compute/core3ark3.for:22-25,65-86,108-154,203-236 (EXPLOC=0)                     */
/* ------------------------------------------------------------------------------------ */
void fo_core3ark3(const int *pi, const double *pr, int nitcfg, const int ind[6],
                  double *rop_1, const double *rop, const double *drodm, const double *coe)
{
    const double cgamma1 = 0.5, cgamma2 = 0.9106836025229591, cgamma3 = 0.3660254037844387;
    const double cv = pr[FO_CVINF], cvinv = 1. / cv, cvinv2 = 0.5 * cvinv;
    const int nd = pi[FO_NDIMDX], b = pi[FO_RK], neq = pi[FO_NEQ];
    double coefW = 1.;
    if (b == 2) coefW = (nitcfg == 1) ? 1. : 0.5;
    else if (b == 3) coefW = (nitcfg == 1) ? cgamma1 : (nitcfg == 2 ? cgamma2 : cgamma3);
    const double cmus1 = pr[FO_CS], temp01 = 1. / pr[FO_TEMP0];
    const double coesut = pr[FO_XMUL0] * (1. + cmus1 * temp01);
    const int twod = pi[FO_ITYPZONE] == 3;
    FO_LOOP(ind) {
        const int l = fo_inddm(pi, i, j, k);
        double r_1 = coefW * coe[l];
        const double tab1 = drodm[l] * r_1, tab2 = drodm[l + nd] * r_1, tab3 = drodm[l + 2 * nd] * r_1;
        const double tab4 = drodm[l + 3 * nd] * r_1, tab5 = drodm[l + 4 * nd] * r_1;
        const double ro_old = rop[l], u_old = rop[l + nd], v_old = rop[l + 2 * nd];
        const double w_old = rop[l + 3 * nd], t_old = rop[l + 4 * nd];
        rop_1[l] = rop[l] + tab1;
        r_1 = 1. / rop_1[l];
        rop_1[l + nd] = (ro_old * u_old + tab2) * r_1;
        rop_1[l + 2 * nd] = (ro_old * v_old + tab3) * r_1;
        double roe_old;
        if (!twod) {
            rop_1[l + 3 * nd] = (ro_old * w_old + tab4) * r_1;
            roe_old = ro_old * (cv * t_old + 0.5 * (u_old * u_old + v_old * v_old + w_old * w_old));
            rop_1[l + 4 * nd] = (roe_old + tab5) * r_1 * cvinv
                - cvinv2 * (rop_1[l + nd] * rop_1[l + nd] + rop_1[l + 2 * nd] * rop_1[l + 2 * nd] + rop_1[l + 3 * nd] * rop_1[l + 3 * nd]);
        } else {
            rop_1[l + 3 * nd] = 0.;
            roe_old = ro_old * (cv * t_old + 0.5 * (u_old * u_old + v_old * v_old));
            rop_1[l + 4 * nd] = (roe_old + tab5) * r_1 * cvinv
                - cvinv2 * (rop_1[l + nd] * rop_1[l + nd] + rop_1[l + 2 * nd] * rop_1[l + 2 * nd]);
        }
        if (neq == 6) {
            const double tab6 = drodm[l + 5 * nd] * (coefW * coe[l]);
            const double nu_old = rop[l + 5 * nd];
            double anulam = coesut * sqrt(rop_1[l + 4 * nd] * temp01) / (1. + cmus1 / rop_1[l + 4 * nd]);
            anulam = anulam * r_1 * pr[FO_RATIOM];
            double t = (ro_old * nu_old + tab6) * r_1;
            t = fmin(t, anulam);
            rop_1[l + 5 * nd] = fmax(t, 0.);
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* Ghost cells. FastS/BC/BCzone.cpp:60-865 with npass=0 (gsdr3.cpp:115), whole-zone        */
/* sub-domain; window arithmetic from FastC/BC/indice_cl_sdm.for.                          */
/* ------------------------------------------------------------------------------------ */
static int bc_window(const int *pi, int idir, const int *fen, const int ind_sdm[6], int out[6])
{
    const int nf = pi[FO_NIJK + 3], nfk = pi[FO_NIJK + 4];
    if (!(ind_sdm[0] <= fen[1] && ind_sdm[1] >= fen[0] && ind_sdm[2] <= fen[3] && ind_sdm[3] >= fen[2]
          && ind_sdm[4] <= fen[5] && ind_sdm[5] >= fen[4])) return 1;
    for (int d = 0; d < 3; d++) {
        out[2 * d] = fen[2 * d] > ind_sdm[2 * d] ? fen[2 * d] : ind_sdm[2 * d];
        out[2 * d + 1] = fen[2 * d + 1] < ind_sdm[2 * d + 1] ? fen[2 * d + 1] : ind_sdm[2 * d + 1];
    }
    const int d = (idir - 1) / 2, fic = d == 2 ? nfk : nf;
    if (idir % 2 == 1) { out[2 * d + 1] = out[2 * d] - 1; out[2 * d] = out[2 * d] - fic; }
    else { out[2 * d + 1] = out[2 * d] + fic; out[2 * d] = out[2 * d] + 1; }
    return 0;
}

void fo_bczone(const int *pi, const double *pr, int lrhs, double *rop, const double *ti,
               const double *tj, const double *tk, int ishift_lu[6])
{
    const int pt_bcs = pi[FO_PT_BC];
    const int nb_bc = pi[pt_bcs];
    const int nd = pi[FO_NDIMDX], neq = pi[FO_NEQ];
    const int ijkv[3] = {pi[FO_IJKV], pi[FO_IJKV + 1], pi[FO_IJKV + 2]};
    const int ind_sdm[6] = {1, ijkv[0], 1, ijkv[1], 1, ijkv[2]};
    int lskip[6] = {1, 1, 1, 1, 1, 1};
    int ificmax[6];
    const int lum = pi[FO_LU_MATCH] != 0;
    for (int p = 0; p < 4; p++) ificmax[p] = pi[FO_NIJK + 3] * lum;
    for (int p = 4; p < 6; p++) ificmax[p] = pi[FO_NIJK + 4] * lum;
    const double *tm[3] = {ti, tj, tk};

    if (lrhs == 1) {   /* BCzone.cpp:107-184: which faces are implicit-ed, sweep box extension */
        for (int ndf = 0; ndf < nb_bc; ndf++) {
            const int pt_bc = pi[pt_bcs + 1 + ndf];
            const int ipara = pi[pt_bc + FO_BC_IDIR] - 1, ty = pi[pt_bc + FO_BC_TYPE];
            if ((ty >= 3 && ty <= 7) || ty == 12 || ty == 18 || ty == 30) {
                lskip[ipara] = 0;
                if ((ty >= 3 && ty <= 7) || ty == 12) ificmax[ipara] = 1;
            }
        }
        for (int p = 0; p < 6; p++)
            ishift_lu[p] = ind_sdm[p] + ((p % 2) ? 1 : -1) * (1 - lskip[p]) * ificmax[p];
        /* lu_match==1: ishift vs ind_dm are identical for a whole zone => no extra shift (BCzone.cpp:145-184) */
        const int idirmax = pi[FO_ITYPZONE] == 3 ? 4 : 6;
        for (int idir = 0; idir < idirmax; idir++) {   /* default extrapolation of the rhs (BCzone.cpp:190-256) */
            if (lskip[idir]) continue;
            int fen[6] = {1, ijkv[0], 1, ijkv[1], 1, ijkv[2]};
            const int d = idir / 2;
            if (idir % 2 == 0) { fen[2 * d] = 1; fen[2 * d + 1] = 1; } else { fen[2 * d] = fen[2 * d + 1]; }
            int w[6];
            if (bc_window(pi, idir + 1, fen, ind_sdm, w)) continue;
            fo_bc_extrapolate(pi, idir + 1, lrhs, w, rop);
        }
    }
    for (int ndf = 0; ndf < nb_bc; ndf++) {
        const int pt_bc = pi[pt_bcs + 1 + ndf];
        const int idir = pi[pt_bc + FO_BC_IDIR];
        int ty = pi[pt_bc + FO_BC_TYPE];
        const int ipara = idir - 1;
        if (!(lskip[ipara] == 0 || lrhs == 0)) continue;
        int w[6];
        if (bc_window(pi, idir, pi + pt_bc + FO_BC_FEN, ind_sdm, w)) continue;
        if (lrhs == 1 && (ty == 0 || ty == 1 || ty == 2 || ty == 10 || ty == 11 || ty == 14 || ty == 16 || ty == 17
                          || ty == 20 || ty == 21)) ty = 0;   /* BCzone.cpp:330-332 */
        const double *data = NULL;
        if (pi[pt_bc + FO_BC_NBDATA] != 0) data = pr + pi[pt_bcs + 1 + ndf + nb_bc];
        (void)data;
        const int d = (idir - 1) / 2;
        switch (ty) {
        case 0:  fo_bc_extrapolate(pi, idir, lrhs, w, rop); break;
        case 1:  fo_bc_farfield(pi, pr, idir, lrhs, w, rop, tm[d]); break;
        case 3:  fo_bc_wall_inviscid(pi, pr, idir, lrhs, w, rop, tm[d]); break;
        case 4: case 6: fo_bc_wall_viscous(pi, pr, idir, lrhs, w, rop); break;
        case 11: fo_bc_periodic(pi, idir, lrhs, w, rop); break;
        case 22: break;   /* BCOverlap: filled by the transfers */
        default:
            fprintf(stderr, "fo_bczone: BC type %d outside the hot-path scope\n", ty);
            abort();
        }
    }
    (void)nd; (void)neq;
}

/* FastS/BC/correct_coins.for:41-58 and following (whole zone => every edge and corner) */
void fo_correct_coins(const int *pi, const int ind_in[6], double *rop)
{
    const int nd = pi[FO_NDIMDX], neq = pi[FO_NEQ];
    const int nf = pi[FO_NIJK + 3], nfk = pi[FO_NIJK + 4];
    const int iv = pi[FO_IJKV], jv = pi[FO_IJKV + 1], kv = pi[FO_IJKV + 2];
    int ind[6];
    ind[0] = ind_in[0] > 1 ? ind_in[0] : 1; ind[2] = ind_in[2] > 1 ? ind_in[2] : 1; ind[4] = ind_in[4] > 1 ? ind_in[4] : 1;
    ind[1] = ind_in[1] < iv ? ind_in[1] : iv; ind[3] = ind_in[3] < jv ? ind_in[3] : jv; ind[5] = ind_in[5] < kv ? ind_in[5] : kv;
    const int itest[6] = {1, iv, 1, jv, 1, kv};
#define MINI(a, b) ((a) < (b) ? (a) : (b))
    for (int idir = 0; idir < 2; idir++) {   /* faces imin / imax: edges with j and k */
        const int is = idir == 0 ? -1 : 1;
        if (ind[idir] != itest[idir]) continue;
        const int I0 = ind[idir];
        for (int jside = 0; jside < 2; jside++) {
            if (ind[2 + jside] != itest[2 + jside]) continue;
            const int J0 = ind[2 + jside], js = jside == 0 ? -1 : 1;
            for (int ne = 0; ne < neq; ne++) for (int k = ind[4]; k <= ind[5]; k++)
                for (int j = 1; j <= nf; j++) for (int i = 1; i <= nf; i++) {
                    const int i1 = MINI(nf, i + 1), j1 = MINI(nf, j + 1);
                    const int l1 = fo_inddm(pi, I0 + i * is, J0 + j * js, k);
                    const int l11 = fo_inddm(pi, I0 + i1 * is, J0, k);
                    const int l12 = fo_inddm(pi, I0, J0 + j1 * js, k);
                    rop[l1 + ne * nd] = (rop[l11 + ne * nd] + rop[l12 + ne * nd]) * 0.5;
                }
        }
        for (int kside = 0; kside < 2; kside++) {
            if (ind[4 + kside] != itest[4 + kside]) continue;
            const int K0 = ind[4 + kside], ks = kside == 0 ? -1 : 1;
            for (int ne = 0; ne < neq; ne++) for (int k = 1; k <= nfk; k++)
                for (int j = ind[2]; j <= ind[3]; j++) for (int i = 1; i <= nf; i++) {
                    const int k1 = MINI(nfk, k + 1), i1 = MINI(nf, i + 1);
                    const int l1 = fo_inddm(pi, I0 + i * is, j, K0 + k * ks);
                    const int l11 = fo_inddm(pi, I0, j, K0 + k1 * ks);
                    const int l12 = fo_inddm(pi, I0 + i1 * is, j, K0);
                    rop[l1 + ne * nd] = (rop[l11 + ne * nd] + rop[l12 + ne * nd]) * 0.5;
                }
        }
    }
    for (int idir = 2; idir < 4; idir++) {   /* faces jmin / jmax: edges with k */
        const int is = idir == 2 ? -1 : 1;
        if (ind[idir] != itest[idir]) continue;
        const int J0 = ind[idir];
        for (int kside = 0; kside < 2; kside++) {
            if (ind[4 + kside] != itest[4 + kside]) continue;
            const int K0 = ind[4 + kside], ks = kside == 0 ? -1 : 1;
            for (int ne = 0; ne < neq; ne++) for (int k = 1; k <= nfk; k++)
                for (int j = 1; j <= nf; j++) for (int i = ind[0]; i <= ind[1]; i++) {
                    const int k1 = MINI(nfk, k + 1), j1 = MINI(nf, j + 1);
                    const int l1 = fo_inddm(pi, i, J0 + j * is, K0 + k * ks);
                    const int l11 = fo_inddm(pi, i, J0, K0 + k1 * ks);
                    const int l12 = fo_inddm(pi, i, J0 + j1 * is, K0);
                    rop[l1 + ne * nd] = (rop[l11 + ne * nd] + rop[l12 + ne * nd]) * 0.5;
                }
        }
    }
    const double c3 = 1. / 3.;
    for (int corner = 0; corner < 8; corner++) {
        const int si = corner & 1, sj = (corner >> 1) & 1, sk = (corner >> 2) & 1;
        const int i0 = ind[si], j0 = ind[2 + sj], k0 = ind[4 + sk];
        if (i0 != itest[si] || j0 != itest[2 + sj] || k0 != itest[4 + sk]) continue;
        const int is = si ? 1 : -1, js = sj ? 1 : -1, ks = sk ? 1 : -1;
        for (int ne = 0; ne < neq; ne++) for (int k = 1; k <= nfk; k++)
            for (int j = 1; j <= nf; j++) for (int i = 1; i <= nf; i++) {
                const int i1 = MINI(nf, i + 1), j1 = MINI(nf, j + 1), k1 = MINI(nfk, k + 1);
                const int l = fo_inddm(pi, i0 + i * is, j0 + j * js, k0 + k * ks);
                const int l11 = fo_inddm(pi, i0, j0, k0 + k1 * ks);
                const int l12 = fo_inddm(pi, i0 + i1 * is, j0, k0);
                const int l13 = fo_inddm(pi, i0, j0 + j1 * js, k0);
                rop[l + ne * nd] = (rop[l11 + ne * nd] + rop[l12 + ne * nd] + rop[l13 + ne * nd]) * c3;
            }
    }
#undef MINI
}
