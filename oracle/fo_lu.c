/*
 * fo_lu.c -- implicit chain of the CPU oracle: LU spectral radii / diagonal (tstb3c, tstb3_impli), Newton
 * residual assembly (core3as2), residual norms (cprdu3s1), LU-SGS sweeps + Newton update (invlu_l / invlu_u and the
 * lu_ and mjr_ bodies), match-join transfers.  TEST INFRASTRUCTURE, see fasts_oracle.h.
 */
#include <math.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "fo_internal.h"
#include "fasts_oracle.h"

/* ---------------------------------------------------------------------------------------------------------
 * FastS/Compute/tstb3c.for:77-300 (global dt: coe(:,1) already = dt/vol) and tstb3_impli.for:60-330 with
 * dtloc.for (local dt: coe(:,1) = CFL*min(dt_conv, dt_visc)); fixed mesh, E_SCALAR_COMPUTER branch.
 * ------------------------------------------------------------------------------------------------------- */
static void coe_kernel(const int *pi, const double *pr, int first_it, const int ind[6], const double *rop,
                       double *coe, const double *xmut, const double *ti, const double *tj, const double *tk,
                       const double *vol, int local)
{
    const int nd = pi[FO_NDIMDX], nm = pi[FO_NDIMDX_MTR], typ = pi[FO_ITYPZONE];
    const int mi = pi[FO_NIJK_MTR], mj = pi[FO_NIJK_MTR + 1], mk = pi[FO_NIJK_MTR + 2];
    const double cut0x = local ? 1e-30 : 1e-20;
    const double gamma = pr[FO_GAMMA], rgp = pr[FO_CVINF] * (gamma - 1.), prandt = pr[FO_PRANDT];
    double xid = 1.5; if (first_it == 0) xid = 1.;
    if (local) xid = 1.;
    double gam1 = gamma / prandt;
    if (pi[FO_IFLOW] == 1) gam1 = local ? cut0x : 0.;
    const double gam1_1 = local ? 1. / gam1 : 0.;
    const double gam2 = gamma * rgp;
    double tcxi0 = 0, tcyj0 = 0, tczk0 = 0, tc2i0 = 0, tc2j0 = 0, tc2k0 = 0;
    if (typ == 2) { tcxi0 = ti[0]; tcyj0 = tj[0]; tczk0 = tk[0]; tc2i0 = sqrt(tcxi0 * tcxi0); tc2j0 = sqrt(tcyj0 * tcyj0); tc2k0 = sqrt(tczk0 * tczk0); }
    FO_LOOP(ind) {
        const int l = fo_inddm(pi, i, j, k);
        const int lt = typ == 2 ? 0 : fo_indmtr(pi, i, j, k);
        const int li = lt + mi, lj = lt + mj, lk = lt + mk;
        const double r = rop[l], detj = fmax(vol[lt], cut0x);
        const double u = rop[l + nd], v = rop[l + 2 * nd], w = rop[l + 3 * nd];
        const double c = sqrt(gam2 * rop[l + 4 * nd]);
        const double vmu = gam1 * (xmut ? xmut[l] : 0.);
        const double volu = detj * r;
        const double xmvis = vmu / volu;
        double tc2i, tc2j, tc2k, qni, qnj, qnk;
        if (typ == 0) {
            const double tcxi = .5 * (ti[lt] + ti[li]), tcyi = .5 * (ti[lt + nm] + ti[li + nm]), tczi = .5 * (ti[lt + 2 * nm] + ti[li + 2 * nm]);
            tc2i = .5 * (sqrt(ti[lt] * ti[lt] + ti[lt + nm] * ti[lt + nm] + ti[lt + 2 * nm] * ti[lt + 2 * nm])
                       + sqrt(ti[li] * ti[li] + ti[li + nm] * ti[li + nm] + ti[li + 2 * nm] * ti[li + 2 * nm]));
            qni = tcxi * u + tcyi * v + tczi * w;
            const double tcxj = .5 * (tj[lt] + tj[lj]), tcyj = .5 * (tj[lt + nm] + tj[lj + nm]), tczj = .5 * (tj[lt + 2 * nm] + tj[lj + 2 * nm]);
            tc2j = .5 * (sqrt(tj[lt] * tj[lt] + tj[lt + nm] * tj[lt + nm] + tj[lt + 2 * nm] * tj[lt + 2 * nm])
                       + sqrt(tj[lj] * tj[lj] + tj[lj + nm] * tj[lj + nm] + tj[lj + 2 * nm] * tj[lj + 2 * nm]));
            qnj = tcxj * u + tcyj * v + tczj * w;
            const double tcxk = .5 * (tk[lt] + tk[lk]), tcyk = .5 * (tk[lt + nm] + tk[lk + nm]), tczk = .5 * (tk[lt + 2 * nm] + tk[lk + 2 * nm]);
            tc2k = .5 * (sqrt(tk[lt] * tk[lt] + tk[lt + nm] * tk[lt + nm] + tk[lt + 2 * nm] * tk[lt + 2 * nm])
                       + sqrt(tk[lk] * tk[lk] + tk[lk + nm] * tk[lk + nm] + tk[lk + 2 * nm] * tk[lk + 2 * nm]));
            qnk = tcxk * u + tcyk * v + tczk * w;
        } else if (typ == 1) {
            const double tcxi = .5 * (ti[lt] + ti[li]), tcyi = .5 * (ti[lt + nm] + ti[li + nm]);
            tc2i = .5 * (sqrt(ti[lt] * ti[lt] + ti[lt + nm] * ti[lt + nm]) + sqrt(ti[li] * ti[li] + ti[li + nm] * ti[li + nm]));
            qni = tcxi * u + tcyi * v;
            const double tcxj = .5 * (tj[lt] + tj[lj]), tcyj = .5 * (tj[lt + nm] + tj[lj + nm]);
            tc2j = .5 * (sqrt(tj[lt] * tj[lt] + tj[lt + nm] * tj[lt + nm]) + sqrt(tj[lj] * tj[lj] + tj[lj + nm] * tj[lj + nm]));
            qnj = tcxj * u + tcyj * v;
            const double tczk = tk[lt];
            tc2k = sqrt(tk[lt] * tk[lt]);
            qnk = tczk * w;
        } else {
            tc2i = tc2i0; tc2j = tc2j0; tc2k = tc2k0;
            qni = tcxi0 * u; qnj = tcyj0 * v; qnk = tczk0 * w;
        }
        const double vmi = fabs(qni) + c * tc2i, vmj = fabs(qnj) + c * tc2j, vmk = fabs(qnk) + c * tc2k;
        coe[l + nd] = vmi * 2. + xmvis * tc2i * tc2i * 4.;
        coe[l + 2 * nd] = vmj * 2. + xmvis * tc2j * tc2j * 4.;
        coe[l + 3 * nd] = vmk * 2. + xmvis * tc2k * tc2k * 4.;
        if (local) {   /* dtloc.for */
            const double d = 1. / (tc2i * tc2i + tc2j * tc2j + tc2k * tc2k);
            const double dtvis = 0.5 * detj * d * gam1_1 / (xmut ? xmut[l] : 0.) * r;
            const double sp = sqrt(u * u + v * v + w * w) + c;
            const double dtconv = sqrt(d) / sp;
            coe[l] = pr[FO_CFL] * fmin(dtconv, dtvis);
        }
        coe[l + 4 * nd] = xid + coe[l] * (coe[l + nd] + coe[l + 2 * nd] + coe[l + 3 * nd]) * 0.5;
    }
}

void fo_tstb3c(const int *pi, const double *pr, int first_it, const int ind[6], const double *rop, double *coe,
               const double *xmut, const double *ti, const double *tj, const double *tk, const double *vol)
{ coe_kernel(pi, pr, first_it, ind, rop, coe, pi[FO_IFLOW] == 1 ? NULL : xmut, ti, tj, tk, vol, 0); }

void fo_tstb3_impli(const int *pi, const double *pr, int first_it, const int ind[6], const double *rop, double *coe,
                    const double *xmut, const double *ti, const double *tj, const double *tk, const double *vol)
{
    if (pi[FO_IFLOW] == 1) { fprintf(stderr, "oracle: local time step needs a viscous model (dtloc.for divides by xmut)\n"); abort(); }
    coe_kernel(pi, pr, first_it, ind, rop, coe, xmut, ti, tj, tk, vol, 1);
}

void fo_tstb3_expli(const int *pi, const double *pr, const int ind[6], const double *rop, double *coe,
                    const double *xmut, const double *ti, const double *tj, const double *tk, const double *vol)
{ (void)pi; (void)pr; (void)ind; (void)rop; (void)coe; (void)xmut; (void)ti; (void)tj; (void)tk; (void)vol;
  fprintf(stderr, "oracle: explicit local time step (tstb3_expli) is out of scope\n"); abort(); }

/* FastS/Compute/core3as2.for:37-195 */
void fo_core3as2(const int *pi, const double *pr, int nitcfg, int first_it, const int ind[6], const double *rop,
                 const double *rop_n, const double *rop_n1, double *drodm, const double *coe)
{
    const int nd = pi[FO_NDIMDX], neq = pi[FO_NEQ];
    const double cv = pr[FO_CVINF];
    double c1, c2, c3;
    if (first_it == 0 || pi[FO_DTLOC] == 1) { c1 = -1.; c2 = 1.; c3 = 0.; } else { c1 = -1.5; c2 = 2.; c3 = -0.5; }
    if (nitcfg == 1) {
        const double c4 = c1 + c2;
        if (pi[FO_DTLOC] == 1) {
            FO_LOOP(ind) { const int l = fo_inddm(pi, i, j, k); for (int e = 0; e < neq; e++) drodm[l + e * nd] = coe[l] * drodm[l + e * nd]; }
            return;
        }
        FO_LOOP(ind) {
            const int l = fo_inddm(pi, i, j, k);
            drodm[l] = coe[l] * drodm[l] + c4 * rop[l] + c3 * rop_n1[l];
            for (int e = 1; e < 4; e++)
                drodm[l + e * nd] = coe[l] * drodm[l + e * nd] + c4 * rop[l + e * nd] * rop[l] + c3 * rop_n1[l + e * nd] * rop_n1[l];
            const double roe = rop[l] * (cv * rop[l + 4 * nd] + 0.5 * (rop[l + nd] * rop[l + nd] + rop[l + 2 * nd] * rop[l + 2 * nd] + rop[l + 3 * nd] * rop[l + 3 * nd]));
            const double roe_n1 = rop_n1[l] * (cv * rop_n1[l + 4 * nd] + 0.5 * (rop_n1[l + nd] * rop_n1[l + nd] + rop_n1[l + 2 * nd] * rop_n1[l + 2 * nd] + rop_n1[l + 3 * nd] * rop_n1[l + 3 * nd]));
            drodm[l + 4 * nd] = coe[l] * drodm[l + 4 * nd] + c4 * roe + c3 * roe_n1;
            if (neq == 6) drodm[l + 5 * nd] = coe[l] * drodm[l + 5 * nd] + c4 * rop[l + 5 * nd] * rop[l] + c3 * rop_n1[l + 5 * nd] * rop_n1[l];
        }
    } else {
        FO_LOOP(ind) {
            const int l = fo_inddm(pi, i, j, k);
            drodm[l] = coe[l] * drodm[l] + c1 * rop[l] + c2 * rop_n[l] + c3 * rop_n1[l];
            for (int e = 1; e < 4; e++)
                drodm[l + e * nd] = coe[l] * drodm[l + e * nd] + c1 * rop[l + e * nd] * rop[l] + c2 * rop_n[l + e * nd] * rop_n[l]
                                  + c3 * rop_n1[l + e * nd] * rop_n1[l];
            const double roe = rop[l] * (cv * rop[l + 4 * nd] + 0.5 * (rop[l + nd] * rop[l + nd] + rop[l + 2 * nd] * rop[l + 2 * nd] + rop[l + 3 * nd] * rop[l + 3 * nd]));
            const double roe_n = rop_n[l] * (cv * rop_n[l + 4 * nd] + 0.5 * (rop_n[l + nd] * rop_n[l + nd] + rop_n[l + 2 * nd] * rop_n[l + 2 * nd] + rop_n[l + 3 * nd] * rop_n[l + 3 * nd]));
            const double roe_n1 = rop_n1[l] * (cv * rop_n1[l + 4 * nd] + 0.5 * (rop_n1[l + nd] * rop_n1[l + nd] + rop_n1[l + 2 * nd] * rop_n1[l + 2 * nd] + rop_n1[l + 3 * nd] * rop_n1[l + 3 * nd]));
            drodm[l + 4 * nd] = coe[l] * drodm[l + 4 * nd] + c1 * roe + c2 * roe_n + c3 * roe_n1;
            if (neq == 6)
                drodm[l + 5 * nd] = coe[l] * drodm[l + 5 * nd] + c1 * rop[l + 5 * nd] * rop[l] + c2 * rop_n[l + 5 * nd] * rop_n[l]
                                  + c3 * rop_n1[l + 5 * nd] * rop_n1[l];
        }
    }
}

/* FastS/Compute/cprdu3s1.for + cprdu3s1_incl.for:100-140, one residual block = the whole zone:
   rdm(nitcfg, ne, 1) = sum drodm^2, rdm(nitcfg, ne, 2) = max |drodm|  (Fortran order (nssiter, neq, 2)) */
void fo_cprdu3s1(const int *pi, int nitcfg, const int ind[6], const double *drodm, double *rdm)
{
    const int nd = pi[FO_NDIMDX], neq = pi[FO_NEQ], nss = pi[FO_NSSITER];
    for (int e = 0; e < neq; e++) {
        double s = 0., m = 0.;
        FO_LOOP(ind) { const double x = drodm[fo_inddm(pi, i, j, k) + e * nd]; const double xp = fabs(x); if (xp > m) m = xp; s += x * x; }
        rdm[(nitcfg - 1) + e * nss] = s;
        rdm[(nitcfg - 1) + e * nss + neq * nss] = m;
    }
}

/* ---------------------------------------------------------------------------------------------------------
 * LU-SGS.  B_d(l1) dW(l1): LU/lu_{i,j,k}_{3dfull,3dhomogene,3dcart}[_SA].for restated once with the cell-centred
 * mean normal tc of the NEIGHBOUR cell l1 (zero components for 3dhomo / 3dcart), signe = +0.5 (L) / -0.5 (U).
 * ------------------------------------------------------------------------------------------------------- */
typedef struct {
    const int *pi; const double *rop, *coe, *t[3];
    int inc[3], incm[3], nd, nm, typ, neq;
    double gam1, gamm1, gam2, cp, tc0[3];
} luctx;

static void lu_term(const luctx *c, int d, int l1, int lta, int ltb, double signe, const double *dw, double b[6])
{
    const int nd = c->nd, nm = c->nm;
    double tcx = 0., tcy = 0., tcz = 0.;
    const double *T = c->t[d];
    if (c->typ == 0) { tcx = 0.5 * (T[ltb] + T[lta]); tcy = 0.5 * (T[ltb + nm] + T[lta + nm]); tcz = 0.5 * (T[ltb + 2 * nm] + T[lta + 2 * nm]); }
    else if (c->typ == 1) { if (d < 2) { tcx = 0.5 * (T[ltb] + T[lta]); tcy = 0.5 * (T[ltb + nm] + T[lta + nm]); } else tcz = T[ltb]; }
    else { if (d == 0) tcx = c->tc0[0]; else if (d == 1) tcy = c->tc0[1]; else tcz = c->tc0[2]; }
    const double *rop = c->rop;
    const double u = rop[l1 + nd], v = rop[l1 + 2 * nd], w = rop[l1 + 3 * nd], t = rop[l1 + 4 * nd];
    const double q2 = .5 * (u * u + v * v + w * w), h = c->cp * t + q2, ph2 = c->gamm1 * q2;
    const double qn = tcx * u + tcy * v + tcz * w;
    const double cd = c->coe[l1 + (d + 1) * nd] * signe;
    const double gamm1 = c->gamm1, gam2 = c->gam2, gam1 = c->gam1;
    const double b11 = cd, b12 = tcx, b13 = tcy, b14 = tcz;
    const double b21 = (tcx * ph2 - u * qn), b22 = (qn - tcx * gam2 * u) + cd, b23 = (tcy * u - tcx * gamm1 * v), b24 = (tcz * u - tcx * gamm1 * w), b25 = tcx * gamm1;
    const double b31 = (tcy * ph2 - v * qn), b32 = (tcx * v - tcy * gamm1 * u), b33 = (qn - tcy * gam2 * v) + cd, b34 = (tcz * v - tcy * gamm1 * w), b35 = tcy * gamm1;
    const double b41 = (tcz * ph2 - w * qn), b42 = (tcx * w - tcz * gamm1 * u), b43 = (tcy * w - tcz * gamm1 * v), b44 = (qn - tcz * gam2 * w) + cd, b45 = tcz * gamm1;
    const double b51 = qn * (ph2 - h), b52 = (tcx * h - gamm1 * u * qn), b53 = (tcy * h - gamm1 * v * qn), b54 = (tcz * h - gamm1 * w * qn), b55 = (gam1 * qn) + cd;
    const double d1 = dw[l1], d2 = dw[l1 + nd], d3 = dw[l1 + 2 * nd], d4 = dw[l1 + 3 * nd], d5 = dw[l1 + 4 * nd];
    b[0] = b11 * d1 + b12 * d2 + b13 * d3 + b14 * d4;
    b[1] = b21 * d1 + b22 * d2 + b23 * d3 + b24 * d4 + b25 * d5;
    b[2] = b31 * d1 + b32 * d2 + b33 * d3 + b34 * d4 + b35 * d5;
    b[3] = b41 * d1 + b42 * d2 + b43 * d3 + b44 * d4 + b45 * d5;
    b[4] = b51 * d1 + b52 * d2 + b53 * d3 + b54 * d4 + b55 * d5;
    if (c->neq == 6) b[5] = (qn + cd) * dw[l1 + 5 * nd];
}

/* One cell of the lower sweep (invlu_l.for:79-215): normals 0.5*(t(lt)+t(lt-m)) of the cell l-d.  Shared by the serial sweep
   and the team sweep (fo_invlu_team): a cell's arithmetic only depends on its three lower neighbours being final. */
static inline void lu_lower_cell(const luctx *c, const int box[6], int overlap, double *drodm, int i, int j, int k)
{
    const int *pi = c->pi; const double *coe = c->coe;
    const int nd = c->nd, neq = c->neq;
    double b[6];
    const int l = fo_inddm(pi, i, j, k), lt = c->typ == 2 ? 0 : fo_indmtr(pi, i, j, k);
    const double xal = coe[l] * 0.5;
    int order[3], n = 0;
    const int hi = i > box[0], hj = j > box[2], hk = k > box[4];
    if (overlap) { if (hi) order[n++] = 0; if (hj) order[n++] = 1; if (hk) order[n++] = 2; }
    else if (c->typ == 2 && hj && hk) { order[n++] = 1; order[n++] = 2; if (hi) order[n++] = 0; }
    else { if (hk) order[n++] = 2; if (hj) order[n++] = 1; if (hi) order[n++] = 0; }
    for (int q = 0; q < n; q++) {
        const int d = order[q];
        lu_term(c, d, l - c->inc[d], lt - c->incm[d], lt, 0.5, drodm, b);
        for (int e = 0; e < neq; e++) drodm[l + e * nd] = drodm[l + e * nd] + b[e] * xal;
    }
    const double diag = 1. / coe[l + 4 * nd];
    for (int e = 0; e < 5; e++) drodm[l + e * nd] = drodm[l + e * nd] * diag;
    if (neq == 6) drodm[l + 5 * nd] = drodm[l + 5 * nd] / coe[l + 5 * nd];
}

/* One cell of the upper sweep (invlu_u.for:60-215): normals 0.5*(t(lt+2m)+t(lt+m)) of the cell l+d. */
static inline void lu_upper_cell(const luctx *c, const int box[6], int overlap, double *drodm, int i, int j, int k)
{
    const int *pi = c->pi; const double *coe = c->coe;
    const int nd = c->nd, neq = c->neq;
    double b[6];
    const int l = fo_inddm(pi, i, j, k), lt = c->typ == 2 ? 0 : fo_indmtr(pi, i, j, k);
    const double xal = coe[l] * -0.5;
    const double dg = coe[l + 4 * nd];
    for (int e = 0; e < 5; e++) drodm[l + e * nd] = drodm[l + e * nd] * dg;
    if (neq == 6) drodm[l + 5 * nd] = drodm[l + 5 * nd] * coe[l + 5 * nd];
    int order[3], n = 0;
    const int hi = i < box[1], hj = j < box[3], hk = k < box[5];
    if (overlap) { if (hi) order[n++] = 0; if (hj) order[n++] = 1; if (hk) order[n++] = 2; }
    else { if (hk) order[n++] = 2; if (hj) order[n++] = 1; if (hi) order[n++] = 0; }
    for (int q = 0; q < n; q++) {
        const int d = order[q];
        lu_term(c, d, l + c->inc[d], lt + c->incm[d], c->typ == 1 && d == 2 ? lt : lt + 2 * c->incm[d], -0.5, drodm, b);
        for (int e = 0; e < neq; e++) drodm[l + e * nd] = drodm[l + e * nd] + b[e] * xal;
    }
    const double diag = 1. / coe[l + 4 * nd];
    for (int e = 0; e < 5; e++) drodm[l + e * nd] = drodm[l + e * nd] * diag;
    if (neq == 6) drodm[l + 5 * nd] = drodm[l + 5 * nd] / coe[l + 5 * nd];
}

/* LU/mjr_newton.for / mjr_newton_SA.for */
static void mjr_newton(const int *pi, const double *pr, int l, const double *rop, double *rop_1, const double *dw)
{
    const int nd = pi[FO_NDIMDX], neq = pi[FO_NEQ];
    const double cvinv = 1. / pr[FO_CVINF], cvinv2 = 0.5 * cvinv;
    const double ro_old = rop[l], u_old = rop[l + nd], v_old = rop[l + 2 * nd], w_old = rop[l + 3 * nd], t_old = rop[l + 4 * nd];
    const double nu_old = neq == 6 ? rop[l + 5 * nd] : 0.;
    rop_1[l] = rop[l] + dw[l];
    const double r_1 = 1. / rop_1[l];
    rop_1[l + nd] = (ro_old * u_old + dw[l + nd]) * r_1;
    rop_1[l + 2 * nd] = (ro_old * v_old + dw[l + 2 * nd]) * r_1;
    rop_1[l + 3 * nd] = (ro_old * w_old + dw[l + 3 * nd]) * r_1;
    const double roe_old = ro_old * (pr[FO_CVINF] * t_old + 0.5 * (u_old * u_old + v_old * v_old + w_old * w_old));
    rop_1[l + 4 * nd] = (roe_old + dw[l + 4 * nd]) * r_1 * cvinv
        - cvinv2 * (rop_1[l + nd] * rop_1[l + nd] + rop_1[l + 2 * nd] * rop_1[l + 2 * nd] + rop_1[l + 3 * nd] * rop_1[l + 3 * nd]);
    if (neq == 6) {
        const double cmus1 = pr[FO_CS], temp01 = 1. / pr[FO_TEMP0], coesut = pr[FO_XMUL0] * (1. + cmus1 * temp01);
        double anulam = coesut * sqrt(rop_1[l + 4 * nd] * temp01) / (1. + cmus1 / rop_1[l + 4 * nd]);
        anulam = anulam * r_1 * pr[FO_RATIOM];
        double t = (ro_old * nu_old + dw[l + 5 * nd]) * r_1;
        t = fmin(t, anulam);
        rop_1[l + 5 * nd] = fmax(t, 0.);
    }
}

/* LU/invlu.for:163-240 (nb_relax >= 2, fixed mesh): invlussor1_l / invlussor_l / invlussor_u[_SA].for, lussor_initssor.for,
   mjr_drodm[_IO]_ssor.for, mjr_ssor.for -- symmetric Gauss-Seidel relaxations of (D + L + U) dW = rhs on the sweep box:
     lower, relaxation m:  dW*_l = (rhs_l + U-sum of relaxation m-1 at l + sum over the lower neighbours B_d(l-d) dW*(l-d) xal) / coe5
                           (invlussor_l.for: drodm_out enters as rhs + ssor, mjr_ssor.for of the previous upper sweep; the first
                           relaxation invlussor1_l.for starts from the rhs alone, lu_dinv_IO.for / mjr_drodm_IO_ssor.for);
                           the L-sum of the cell is kept in ssor (mjr_drodm_ssor.for) and one plane later the cell becomes
                           rhs + L-sum (mjr_ssor.for at l - inck, invlussor1_l.for:158-160)
     upper, relaxation m:  dW_l  = (rhs_l + L-sum_l + sum over the upper neighbours B_d(l+d) dW(l+d) xal) / coe5, the U-sum is
                           kept in ssor for the next lower sweep (invlussor_u.for:111-228, 245-272)
   Terms are added in i, j, k order (the include order of those files).  The Newton update runs plane by plane inside
   the LAST upper sweep (lussor_end = 1, invlussor_u.for:155-170, 232-244) over the interior ind_loop_sdm, so with rop
   aliased to rop_out the k term of plane k reads the updated state of plane k+1 exactly as in invlu_u.for.  The ssor
   array of the reference also spans the ghost rows of an OpenMP sub-domain; for the whole-zone box restated here those
   entries are never read.  The first relaxation equals LU-SGS: (D+U) dW = rhs + L-sum = D dW*. */
static void fo_invlussor(const int *pi, const double *pr, const int box[6], luctx *c, double *rop_out, double *drodm, int nrelax)
{
    const int nd = c->nd, neq = c->neq;
    const double *coe = c->coe;
    const int ideb = box[0], ifin = box[1], jdeb = box[2], jfin = box[3], kdeb = box[4], kfin = box[5];
    double *rhs = (double *)malloc(sizeof(double) * (size_t)nd * neq);
    double *low = (double *)calloc((size_t)nd * neq, sizeof(double));
    double *us = (double *)calloc((size_t)nd * neq, sizeof(double));     /* lussor_initssor.for: ssor = 0 */
    memcpy(rhs, drodm, sizeof(double) * (size_t)nd * neq);
    const int nb[6] = {1, pi[FO_IJKV], 1, pi[FO_IJKV + 1], 1, pi[FO_IJKV + 2]};
    double b[6];
    for (int m = 1; m <= nrelax; m++) {
        for (int k = kdeb; k <= kfin; k++) for (int j = jdeb; j <= jfin; j++) for (int i = ideb; i <= ifin; i++) {
            const int l = fo_inddm(pi, i, j, k), lt = c->typ == 2 ? 0 : fo_indmtr(pi, i, j, k);
            const double xal = coe[l] * 0.5;
            const int has[3] = {i > ideb, j > jdeb, k > kdeb};
            double acc[6], s[6] = {0., 0., 0., 0., 0., 0.};
            for (int e = 0; e < neq; e++) acc[e] = rhs[l + e * nd] + us[l + e * nd];
            for (int d = 0; d < 3; d++) {
                if (!has[d]) continue;
                lu_term(c, d, l - c->inc[d], lt - c->incm[d], lt, 0.5, drodm, b);
                for (int e = 0; e < neq; e++) { acc[e] = acc[e] + b[e] * xal; s[e] = s[e] + b[e] * xal; }
            }
            const double diag = 1. / coe[l + 4 * nd];
            for (int e = 0; e < 5; e++) drodm[l + e * nd] = acc[e] * diag;
            if (neq == 6) drodm[l + 5 * nd] = acc[5] / coe[l + 5 * nd];
            for (int e = 0; e < neq; e++) low[l + e * nd] = s[e];
        }
        for (int k = kfin; k >= kdeb; k--) {
            for (int j = jfin; j >= jdeb; j--) for (int i = ifin; i >= ideb; i--) {
                const int l = fo_inddm(pi, i, j, k), lt = c->typ == 2 ? 0 : fo_indmtr(pi, i, j, k);
                const double xal = coe[l] * -0.5;
                const int has[3] = {i < ifin, j < jfin, k < kfin};
                double acc[6], s[6] = {0., 0., 0., 0., 0., 0.};
                for (int e = 0; e < neq; e++) acc[e] = rhs[l + e * nd] + low[l + e * nd];
                for (int d = 0; d < 3; d++) {
                    if (!has[d]) continue;
                    lu_term(c, d, l + c->inc[d], lt + c->incm[d], c->typ == 1 && d == 2 ? lt : lt + 2 * c->incm[d], -0.5, drodm, b);
                    for (int e = 0; e < neq; e++) { acc[e] = acc[e] + b[e] * xal; s[e] = s[e] + b[e] * xal; }
                }
                const double diag = 1. / coe[l + 4 * nd];
                for (int e = 0; e < 5; e++) drodm[l + e * nd] = acc[e] * diag;
                if (neq == 6) drodm[l + 5 * nd] = acc[5] / coe[l + 5 * nd];
                for (int e = 0; e < neq; e++) us[l + e * nd] = s[e];
            }
            if (m == nrelax && k >= nb[4] && k <= nb[5])
                for (int j = nb[2]; j <= nb[3]; j++) for (int i = nb[0]; i <= nb[1]; i++)
                    mjr_newton(pi, pr, fo_inddm(pi, i, j, k), c->rop, rop_out, drodm);
        }
    }
    free(rhs); free(low); free(us);
}

/* LU/invlu.for:6-344 (nb_relax = 1, fixed mesh): invlu_l[_SA] then invlu_u[_SA]; lu_match = 1 selects the "overlap"
   variants whose only effect for a whole-zone box is the i,j,k accumulation order and a Newton update restricted to
   the interior (invlu_overlap_u.for).  In place: drodm holds rhs -> dW* -> dW. */
static void lu_ctx_init(luctx *c, const int *pi, const double *pr, const double *rop, const double *coe,
                        const double *ti, const double *tj, const double *tk)
{
    c->pi = pi; c->rop = rop; c->coe = coe; c->t[0] = ti; c->t[1] = tj; c->t[2] = tk;
    c->inc[0] = 1; c->inc[1] = pi[FO_NIJK]; c->inc[2] = pi[FO_NIJK] * pi[FO_NIJK + 1];
    c->incm[0] = pi[FO_NIJK_MTR]; c->incm[1] = pi[FO_NIJK_MTR + 1]; c->incm[2] = pi[FO_NIJK_MTR + 2];
    c->nd = pi[FO_NDIMDX]; c->nm = pi[FO_NDIMDX_MTR]; c->typ = pi[FO_ITYPZONE]; c->neq = pi[FO_NEQ];
    c->gam1 = pr[FO_GAMMA]; c->gamm1 = c->gam1 - 1.; c->cp = pr[FO_GAMMA] * pr[FO_CVINF]; c->gam2 = c->gamm1 - 1.;
    c->tc0[0] = ti[0]; c->tc0[1] = tj[0]; c->tc0[2] = tk[0];
}

static void fo_invlu_impl(const int *pi, const double *pr, const int box[6], const double *rop, double *rop_out, double *drodm,
                          const double *coe, const double *ti, const double *tj, const double *tk, int nrelax_forced);
void fo_invlu(const int *pi, const double *pr, const int box[6], const double *rop, double *rop_out, double *drodm,
              const double *coe, const double *ti, const double *tj, const double *tk)
{
    fo_invlu_impl(pi, pr, box, rop, rop_out, drodm, coe, ti, tj, tk, 0);
}
/* test hook: the relaxation sweeps with a given number of relaxations (1 must reproduce the nb_relax = 1 path) */
void fo_invlu_relax(const int *pi, const double *pr, const int box[6], const double *rop, double *rop_out, double *drodm,
                    const double *coe, const double *ti, const double *tj, const double *tk, int nrelax)
{
    fo_invlu_impl(pi, pr, box, rop, rop_out, drodm, coe, ti, tj, tk, nrelax);
}
static void fo_invlu_impl(const int *pi, const double *pr, const int box[6], const double *rop, double *rop_out, double *drodm,
                          const double *coe, const double *ti, const double *tj, const double *tk, int nrelax_forced)
{
    if (box[1] < box[0]) return;
    luctx c;
    lu_ctx_init(&c, pi, pr, rop, coe, ti, tj, tk);
    const int overlap = pi[FO_LU_MATCH] == 1;
    if (pi[FO_NB_RELAX] >= 2 || nrelax_forced > 0) { fo_invlussor(pi, pr, box, &c, rop_out, drodm, nrelax_forced > 0 ? nrelax_forced : pi[FO_NB_RELAX]); return; }
    /* SA + lu_match (invlu.for:32-49): invlu_overlap_{l,u}_SA.for are the 6-equation copies of the overlap variants
       (same includes with the _SA bodies).  extrap_coe.for:24-25,39,52,66 only acts on a face where the sweep box reaches
       the second ghost row (ind_loop = 1 - nfic), which for a whole-zone box needs a BC of type 18 / 30 (BCzone.cpp:117-127);
       with the BC families restated here the box stops at row 0, so extrap_coe is a no-op. */
    const int ideb = box[0], ifin = box[1], jdeb = box[2], jfin = box[3], kdeb = box[4], kfin = box[5];
    /* ---- lower sweep (invlu_l.for:79-215): normals 0.5*(t(lt)+t(lt-m)) of the cell l-d */
    for (int k = kdeb; k <= kfin; k++) for (int j = jdeb; j <= jfin; j++) for (int i = ideb; i <= ifin; i++)
        lu_lower_cell(&c, box, overlap, drodm, i, j, k);
    /* ---- upper sweep (invlu_u.for:60-215): normals 0.5*(t(lt+2m)+t(lt+m)) of the cell l+d, Newton update plane
       by plane (so the k term of plane k reads the already updated state of plane k+1 when rop aliases rop_out) */
    int nb[6];
    for (int q = 0; q < 6; q++) nb[q] = box[q];
    if (overlap) { nb[0] = 1; nb[1] = pi[FO_IJKV]; nb[2] = 1; nb[3] = pi[FO_IJKV + 1]; nb[4] = 1; nb[5] = pi[FO_IJKV + 2]; }
    for (int k = kfin; k >= kdeb; k--) {
        for (int j = jfin; j >= jdeb; j--) for (int i = ifin; i >= ideb; i--)
            lu_upper_cell(&c, box, overlap, drodm, i, j, k);
        if (k >= nb[4] && k <= nb[5])
            for (int j = nb[2]; j <= nb[3]; j++) for (int i = nb[0]; i <= nb[1]; i++)
                mjr_newton(pi, pr, fo_inddm(pi, i, j, k), rop, rop_out, drodm);
    }
}

/* The same two sweeps by a TEAM of threads, bit for bit (timed CPU baseline only: bench.py switches it on through
   fo_set_lu_team; the parity tests keep the serial sweep above).  Thread `rank` of `nper` owns a slab of j rows; a plane is cut
   into chunks of FO_LU_CHUNK cells along i and the team runs a (slab, chunk) wavefront: in the lower sweep rank r takes chunk
   g = (plane, chunk) once rank r-1 has finished the same chunk (its last row feeds the first row of r), in the upper sweep once
   rank r+1 has.  Inside a chunk the rows of the slab and the cells of a row are visited in sweep order, so every cell finds its
   three neighbours final: the values are those of the serial sweep whatever the interleaving.  The Newton update of a plane
   (mjr_newton, plane by plane as in invlu_u.for) is applied by each rank to its own slab after rank r-1 has finished sweeping
   that plane -- the in-plane neighbours (i+1, j+1) of a cell must still be read un-updated, the k neighbour (own slab) updated.
   prog: 2 * nper counters of finished chunks (lower, upper), zeroed by the caller, one set per zone. */
#define FO_LU_CHUNK 32
static inline void lu_wait(const long *p, long v)
{   /* spin, then give the core away: with more threads than cores the thread waited for may not be running */
    int spins = 0;
    while (__atomic_load_n(p, __ATOMIC_ACQUIRE) < v) {
        if (++spins < 2000) __builtin_ia32_pause(); else { sched_yield(); spins = 0; }
    }
}
void fo_invlu_team(const int *pi, const double *pr, const int box[6], const double *rop, double *rop_out, double *drodm,
                   const double *coe, const double *ti, const double *tj, const double *tk, int rank, int nper, long *prog)
{
    if (box[1] < box[0]) return;
    if (pi[FO_NB_RELAX] >= 2) {      /* relaxations: serial, by the first thread of the team */
        if (rank == 0) fo_invlu_impl(pi, pr, box, rop, rop_out, drodm, coe, ti, tj, tk, 0);
        return;
    }
    luctx c;
    lu_ctx_init(&c, pi, pr, rop, coe, ti, tj, tk);
    const int overlap = pi[FO_LU_MATCH] == 1;
    const int ideb = box[0], ifin = box[1], jdeb = box[2], jfin = box[3], kdeb = box[4], kfin = box[5];
    const int nj = jfin - jdeb + 1, ni = ifin - ideb + 1, nk = kfin - kdeb + 1;
    const int j0 = jdeb + (int)((long)nj * rank / nper), j1 = jdeb + (int)((long)nj * (rank + 1) / nper) - 1;   /* may be empty */
    const int nc = (ni + FO_LU_CHUNK - 1) / FO_LU_CHUNK;
    long *plo = prog, *pup = prog + nper;
    for (int kk = 0; kk < nk; kk++) {
        const int k = kdeb + kk;
        for (int cc = 0; cc < nc; cc++) {
            const long g = (long)kk * nc + cc;
            if (rank > 0) lu_wait(plo + rank - 1, g + 1);
            const int ia = ideb + cc * FO_LU_CHUNK, ib = ia + FO_LU_CHUNK - 1 < ifin ? ia + FO_LU_CHUNK - 1 : ifin;
            for (int j = j0; j <= j1; j++) for (int i = ia; i <= ib; i++) lu_lower_cell(&c, box, overlap, drodm, i, j, k);
            __atomic_store_n(plo + rank, g + 1, __ATOMIC_RELEASE);
        }
    }
    int nb[6];
    for (int q = 0; q < 6; q++) nb[q] = box[q];
    if (overlap) { nb[0] = 1; nb[1] = pi[FO_IJKV]; nb[2] = 1; nb[3] = pi[FO_IJKV + 1]; nb[4] = 1; nb[5] = pi[FO_IJKV + 2]; }
    for (int kk = 0; kk < nk; kk++) {
        const int k = kfin - kk;
        for (int cc = 0; cc < nc; cc++) {
            const long g = (long)kk * nc + cc;
            if (rank < nper - 1) lu_wait(pup + rank + 1, g + 1);
            const int ib = ifin - cc * FO_LU_CHUNK, ia = ib - FO_LU_CHUNK + 1 > ideb ? ib - FO_LU_CHUNK + 1 : ideb;
            for (int j = j1; j >= j0; j--) for (int i = ib; i >= ia; i--) lu_upper_cell(&c, box, overlap, drodm, i, j, k);
            __atomic_store_n(pup + rank, g + 1, __ATOMIC_RELEASE);
        }
        if (rank > 0) lu_wait(pup + rank - 1, (long)(kk + 1) * nc);     /* the slab below has read this plane's old state */
        if (k >= nb[4] && k <= nb[5]) {
            const int ja = j0 > nb[2] ? j0 : nb[2], jb = j1 < nb[3] ? j1 : nb[3];
            for (int j = ja; j <= jb; j++) for (int i = nb[0]; i <= nb[1]; i++)
                mjr_newton(pi, pr, fo_inddm(pi, i, j, k), rop, rop_out, drodm);
        }
    }
}

/* Match joins restricted to one donor / coefficient 1 (C/Com/setInterpTransfersFast.cpp:440-955 for interpolation
   type 0/1, ST/tgvMpi.py:136-139).  Table layout = csrc/fb_joins.cu (kind 0 joins only on the CPU side). */
void fo_transfers(fo_zone *zones, int nz, const int *ji, const double *prtc, int which, int nitcfg)
{
    (void)prtc; (void)nz;
    const int nj = ji[0];
    const int *idx = ji + 2 + 7 * nj;
    for (int q = 0; q < nj; q++) {
        const int *s = ji + 2 + 7 * q;
        if (s[0] != 0) continue;
        fo_zone *zd = &zones[s[1]], *zr = &zones[s[2]];
        const int neq = zd->pi[FO_NEQ], ndd = zd->pi[FO_NDIMDX], ndr = zr->pi[FO_NDIMDX];
        const int implicit = zd->pi[FO_ITYPCP] <= 1;
        const double *src; double *dst;
        if (which == 1) { src = zd->drodm; dst = zr->drodm; }
        else if (nitcfg == 1 || implicit) { src = zd->ro_p1; dst = zr->ro_p1; }
        else if (nitcfg % 2 == 0) { src = zd->ro; dst = zr->ro; }
        else { src = zd->ro_p1; dst = zr->ro_p1; }
        for (int n = 0; n < s[3]; n++)
            for (int e = 0; e < neq; e++) dst[idx[s[5] + n] + e * ndr] = src[idx[s[4] + n] + e * ndd];
    }
}

/* test hook: max-norm of rhs - (D + L + U) dW over the box, with the operators of the sweeps above (tests/test_oracle_pins.py) */
double fo_lu_residual(const int *pi, const double *pr, const int box[6], const double *rop, const double *rhs, const double *dw,
                      const double *coe, const double *ti, const double *tj, const double *tk)
{
    luctx c;
    lu_ctx_init(&c, pi, pr, rop, coe, ti, tj, tk);
    const int nd = c.nd, neq = c.neq;
    double worst = 0., b[6];
    for (int k = box[4]; k <= box[5]; k++) for (int j = box[2]; j <= box[3]; j++) for (int i = box[0]; i <= box[1]; i++) {
        const int l = fo_inddm(pi, i, j, k), lt = c.typ == 2 ? 0 : fo_indmtr(pi, i, j, k);
        const int lo[3] = {i > box[0], j > box[2], k > box[4]}, hi[3] = {i < box[1], j < box[3], k < box[5]};
        double acc[6];
        for (int e = 0; e < neq; e++) acc[e] = rhs[l + e * nd];
        for (int d = 0; d < 3; d++) {
            if (lo[d]) { lu_term(&c, d, l - c.inc[d], lt - c.incm[d], lt, 0.5, dw, b); for (int e = 0; e < neq; e++) acc[e] += b[e] * coe[l] * 0.5; }
            if (hi[d]) { lu_term(&c, d, l + c.inc[d], lt + c.incm[d], c.typ == 1 && d == 2 ? lt : lt + 2 * c.incm[d], -0.5, dw, b);
                         for (int e = 0; e < neq; e++) acc[e] += b[e] * coe[l] * -0.5; }
        }
        for (int e = 0; e < neq; e++) {
            const double r = fabs(acc[e] - dw[l + e * nd] * coe[l + (e < 5 ? 4 : 5) * nd]);
            if (r > worst) worst = r;
        }
    }
    (void)pr;
    return worst;
}
