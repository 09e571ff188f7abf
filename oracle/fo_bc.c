/*
 * fo_bc.c -- ghost-cell fills of the in-scope BC families (TEST INFRASTRUCTURE, see fasts_oracle.h).
 *
 * All `bvbs_*` routines are generated from FastS/BC/template_bc.for: for each (j,k) of the window
 * the first ghost row (closest to the interior) gets the "first rank" body and deeper rows the
 * "next rank" body; fixed mesh => c_ale = 0 (FastC/Compute/shape_tab_mtr.for).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "fo_internal.h"

typedef struct { int d, low, ir, iref, gface, g0, g1, first; int a0, a1, b0, b1; int da, db; } bcgeo;

static bcgeo geo(int idir, const int w[6])
{
    bcgeo g; g.d = (idir - 1) / 2; g.low = idir % 2;
    g.g0 = w[2 * g.d]; g.g1 = w[2 * g.d + 1];
    if (g.low) { g.first = g.g1; g.ir = g.g1 + 1; g.iref = 2 * g.g1 + 1; g.gface = g.g1 + 1; }
    else       { g.first = g.g0; g.ir = g.g0 - 1; g.iref = 2 * g.g0 - 1; g.gface = g.g0; }
    g.da = (g.d + 1) % 3; g.db = (g.d + 2) % 3;
    g.a0 = w[2 * g.da]; g.a1 = w[2 * g.da + 1]; g.b0 = w[2 * g.db]; g.b1 = w[2 * g.db + 1];
    return g;
}
static inline int addr(const int *pi, int d, int n, int da, int a, int db, int b)
{
    int ijk[3]; ijk[d] = n; ijk[da] = a; ijk[db] = b;
    return fo_inddm(pi, ijk[0], ijk[1], ijk[2]);
}
static inline int addrm(const int *pi, int d, int n, int da, int a, int db, int b)
{
    int ijk[3]; ijk[d] = n; ijk[da] = a; ijk[db] = b;
    return fo_indmtr(pi, ijk[0], ijk[1], ijk[2]);
}
/* unit-less face normal as read by the bvbs routines (shape_tab_mtr.for + e.g. BCWallInviscid.for:1-3) */
static void bc_normal(const int *pi, int idir, int lmtr, const double *tijk, double tc[3])
{
    const int d = (idir - 1) / 2;
    const int neq_mtr = d == 2 ? pi[FO_NEQ_K] : pi[FO_NEQ_IJ];
    const int nm = pi[FO_NDIMDX_MTR];
    tc[0] = tc[1] = tc[2] = 0.;
    if (neq_mtr == 3) { tc[0] = tijk[lmtr]; tc[1] = tijk[lmtr + nm]; tc[2] = tijk[lmtr + 2 * nm]; }
    else if (neq_mtr == 2) { tc[0] = tijk[lmtr]; tc[1] = tijk[lmtr + nm]; }
    else if (neq_mtr == 1) tc[d] = tijk[lmtr];
}

/* FastS/BC/bvbs_extrapolate.for + BCExtrapolate.for / BCExtrapolateSA.for */
void fo_bc_extrapolate(const int *pi, int idir, int lrhs, const int w[6], double *rop)
{
    const bcgeo g = geo(idir, w);
    const int nd = pi[FO_NDIMDX], neq = pi[FO_NEQ];
    const double c1 = lrhs == 1 ? 0. : 1., vmax = -1.e30;
    for (int b = g.b0; b <= g.b1; b++) for (int a = g.a0; a <= g.a1; a++) {
        const int l1 = addr(pi, g.d, g.ir, g.da, a, g.db, b);
        for (int n = g.g0; n <= g.g1; n++) {
            const int l = addr(pi, g.d, n, g.da, a, g.db, b);
            for (int e = 0; e < neq; e++) rop[l + e * nd] = fmax(vmax, rop[l1 + e * nd]) * c1;
        }
    }
}

/* FastS/BC/bvbs_periodique.for + BCPeriodic.for: copy from l + IJKV*inc */
void fo_bc_periodic(const int *pi, int idir, int lrhs, const int w[6], double *rop)
{
    const bcgeo g = geo(idir, w);
    const int nd = pi[FO_NDIMDX], neq = pi[FO_NEQ];
    const double c1 = lrhs == 1 ? 0. : 1.;
    const int shift = g.low ? pi[FO_IJKV + g.d] : -pi[FO_IJKV + g.d];
    for (int b = g.b0; b <= g.b1; b++) for (int a = g.a0; a <= g.a1; a++)
        for (int n = g.g0; n <= g.g1; n++) {
            const int l = addr(pi, g.d, n, g.da, a, g.db, b);
            const int l1 = addr(pi, g.d, n + shift, g.da, a, g.db, b);
            for (int e = 0; e < neq; e++) rop[l + e * nd] = rop[l1 + e * nd] * c1;
        }
}

/* FastS/BC/bvbs_wall_inviscid.for + BCWallInviscid.for:1-49 (+ BCWallInviscidSA_i.for) */
void fo_bc_wall_inviscid(const int *pi, const double *pr, int idir, int lrhs, const int w[6], double *rop,
                         const double *tijk)
{
    (void)pr; (void)lrhs;
    const bcgeo g = geo(idir, w);
    const int nd = pi[FO_NDIMDX], neq = pi[FO_NEQ];
    for (int b = g.b0; b <= g.b1; b++) for (int a = g.a0; a <= g.a1; a++) {
        const int lmtr = addrm(pi, g.d, g.gface, g.da, a, g.db, b);
        double tc[3]; bc_normal(pi, idir, lmtr, tijk, tc);
        double surf = sqrt(tc[0] * tc[0] + tc[1] * tc[1] + tc[2] * tc[2]);
        surf = fmax(surf, 1e-30);
        const double s_1 = 1. / surf;
        const double tcx = tc[0] * s_1, tcy = tc[1] * s_1, tcz = tc[2] * s_1;
        for (int n = g.g0; n <= g.g1; n++) {
            const int l = addr(pi, g.d, n, g.da, a, g.db, b);
            const int ldjr = addr(pi, g.d, g.iref - n, g.da, a, g.db, b);
            const double u = rop[ldjr + nd], v = rop[ldjr + 2 * nd], ww = rop[ldjr + 3 * nd];
            const double qn = (u * tcx + v * tcy + ww * tcz);
            const double ut = u - qn * tcx, vt = v - qn * tcy, wt = ww - qn * tcz;
            rop[l] = rop[ldjr];
            rop[l + nd] = 2. * ut - u;
            rop[l + 2 * nd] = 2. * vt - v;
            rop[l + 3 * nd] = 2. * wt - ww;
            rop[l + 4 * nd] = rop[ldjr + 4 * nd];
            if (neq == 6) rop[l + 5 * nd] = rop[ldjr + 5 * nd];
        }
    }
}

/* FastS/BC/bvbs_wall_viscous_adia.for + BCWallViscous.for (+ BCWallViscousSA_i.for: nutilde' = -nutilde) */
void fo_bc_wall_viscous(const int *pi, const double *pr, int idir, int lrhs, const int w[6], double *rop)
{
    (void)pr; (void)lrhs;
    const bcgeo g = geo(idir, w);
    const int nd = pi[FO_NDIMDX], neq = pi[FO_NEQ];
    for (int b = g.b0; b <= g.b1; b++) for (int a = g.a0; a <= g.a1; a++)
        for (int n = g.g0; n <= g.g1; n++) {
            const int l = addr(pi, g.d, n, g.da, a, g.db, b);
            const int ldjr = addr(pi, g.d, g.iref - n, g.da, a, g.db, b);
            const double u = rop[ldjr + nd], v = rop[ldjr + 2 * nd], ww = rop[ldjr + 3 * nd];
            rop[l] = rop[ldjr];
            rop[l + nd] = 2. * 0. - u;
            rop[l + 2 * nd] = 2. * 0. - v;
            rop[l + 3 * nd] = 2. * 0. - ww;
            rop[l + 4 * nd] = rop[ldjr + 4 * nd];
            if (neq == 6) rop[l + 5 * nd] = -rop[ldjr + 5 * nd];
        }
}

/* FastS/BC/bvbs_farfield.for: first rank = characteristic state (FastS/BC/non_reflection.for,
   BCFarfield_firstrank[_SA].for), deeper ranks = BC_nextrank[SA].for.  State = DEFAULT_STATE
   (BCzone.cpp:26-40, nbdata == 0).  Never called with lrhs=1 (BCzone.cpp:330-332). */
void fo_bc_farfield(const int *pi, const double *pr, int idir, int lrhs, const int w[6], double *rop,
                    const double *tijk)
{
    (void)lrhs;
    const bcgeo g = geo(idir, w);
    const int nd = pi[FO_NDIMDX], neq = pi[FO_NEQ];
    const double snorm = (idir % 2 == 0) ? 1. : -1.;
    const double gam1 = pr[FO_GAMMA] - 1., gam1_1 = 1. / gam1;
    const double cv = pr[FO_CVINF], cvinv = 1. / cv;
    double st[6];
    st[0] = pr[FO_ROINF]; st[1] = pr[FO_VXINF] * pr[FO_ROINF]; st[2] = pr[FO_VYINF] * pr[FO_ROINF];
    st[3] = pr[FO_VZINF] * pr[FO_ROINF];
    st[4] = pr[FO_PINF] / (pr[FO_GAMMA] - 1.) + 0.5 * (pr[FO_VXINF] * st[1] + pr[FO_VYINF] * st[2] + pr[FO_VZINF] * st[3]);
    st[5] = pr[FO_RONUTILDEINF];
    const double roe = st[0], rue = st[1], rve = st[2], rwe = st[3], ete = st[4];
    const double nue = neq == 6 ? st[5] / roe : 0.;
    /* MUSCL weights handed over by BCzone.cpp:334-338 */
    const double c4 = 5. / 6., c5 = 2. / 6., c6 = -1. / 6.;
    const double c0 = 1. / c6, c1 = -(c4 + c5) * c0, c2 = -c6 * c0, c3 = (2. - c5 - c4) * c0;
    int incv[3] = {1, pi[FO_NIJK], pi[FO_NIJK] * pi[FO_NIJK + 1]};
    const int inci = g.low ? incv[g.d] : -incv[g.d];
    for (int b = g.b0; b <= g.b1; b++) for (int a = g.a0; a <= g.a1; a++) {
        const int lmtr = addrm(pi, g.d, g.gface, g.da, a, g.db, b);
        double tc[3]; bc_normal(pi, idir, lmtr, tijk, tc);
        const double tcx = tc[0] * snorm, tcy = tc[1] * snorm, tcz = tc[2] * snorm;
        {   /* first rank */
            const int l = addr(pi, g.d, g.first, g.da, a, g.db, b);
            const int l1 = addr(pi, g.d, g.ir, g.da, a, g.db, b);
            const double sn = sqrt(tcx * tcx + tcy * tcy + tcz * tcz);
            const double roi = rop[l1], rui = rop[l1 + nd] * roi, rvi = rop[l1 + 2 * nd] * roi, rwi = rop[l1 + 3 * nd] * roi;
            const double eti = rop[l1 + 4 * nd] * pr[FO_CVINF] * roi + .5 * (rui * rui + rvi * rvi + rwi * rwi) / roi;
            const double qen = 0.;
            const double qref1 = .5 * (roe + roi), qref2 = .5 * (rue + rui), qref3 = .5 * (rve + rvi),
                         qref4 = .5 * (rwe + rwi), qref5 = .5 * (ete + eti);
            const double r = qref1, ri = 1. / r;
            double u = qref2 * ri, v = qref3 * ri, ww = qref4 * ri;
            const double p = gam1 * (qref5 - .5 * r * (u * u + v * v + ww * ww));
            const double c = sqrt(pr[FO_GAMMA] * p * ri), ci = 1. / c, s_1 = 1. / sn;
            const double qn = tcx * u + tcy * v + tcz * ww - qen;
            double tnx, tny;
            if (sn == 0.) { tnx = 1.; tny = 0.; } else { tnx = tcx / sn; tny = tcy / sn; }
            const double eigv1 = qn, eigv4 = qn + sn * c, eigv5 = qn - sn * c;
            double ext[5], in[5];
            for (int side = 0; side < 2; side++) {
                const double q1 = side ? roi : roe, q2 = side ? rui : rue, q3 = side ? rvi : rve,
                             q4 = side ? rwi : rwe, q5 = side ? eti : ete;
                const double s1 = q1, s2 = (q2 - u * q1) * ri, s3 = (q3 - v * q1) * ri, s4 = (q4 - ww * q1) * ri;
                const double s5 = gam1 * (.5 * (u * u + v * v + ww * ww) * q1 - u * q2 - v * q3 - ww * q4 + q5);
                double *o = side ? in : ext;
                o[0] = s1 - ci * ci * s5;
                o[1] = tcz * s_1 * (tnx * s2 + tny * s3) - s4;
                o[2] = -tny * s2 + tnx * s3;
                o[3] = .5 * s_1 * (tcx * s2 + tcy * s3 + tcz * s4) + .5 * ri * ci * s5;
                o[4] = -.5 * s_1 * (tcx * s2 + tcy * s3 + tcz * s4) + .5 * ri * ci * s5;
            }
            const double rvar1 = eigv1 < 0. ? ext[0] : in[0], rvar2 = eigv1 < 0. ? ext[1] : in[1],
                         rvar3 = eigv1 < 0. ? ext[2] : in[2], rvar4 = eigv4 < 0. ? ext[3] : in[3],
                         rvar5 = eigv5 < 0. ? ext[4] : in[4];
            const double svar1 = rvar1 + r * ci * (rvar4 + rvar5);
            const double svar2 = tcz * s_1 * tnx * rvar2 - tny * rvar3 + tcx * s_1 * (rvar4 - rvar5);
            const double svar3 = tcz * s_1 * tny * rvar2 + tnx * rvar3 + tcy * s_1 * (rvar4 - rvar5);
            const double svar4 = -rvar2 + tcz * s_1 * (rvar4 - rvar5);
            const double svar5 = r * c * (rvar4 + rvar5);
            const double qvar5 = .5 * (u * u + v * v + ww * ww) * svar1 + r * (u * svar2 + v * svar3 + ww * svar4) + gam1_1 * svar5;
            const double roinv = 1. / svar1;
            u = u + r * svar2 * roinv; v = v + r * svar3 * roinv; ww = ww + r * svar4 * roinv;
            rop[l] = svar1; rop[l + nd] = u; rop[l + 2 * nd] = v; rop[l + 3 * nd] = ww;
            rop[l + 4 * nd] = (qvar5 * roinv - .5 * (u * u + v * v + ww * ww)) * cvinv;
            if (neq == 6) rop[l + 5 * nd] = eigv1 < 0. ? nue : rop[l1 + 5 * nd];
        }
        /* next ranks, walking away from the interior */
        const int step = g.low ? -1 : 1;
        for (int n = g.first + step; n >= g.g0 && n <= g.g1; n += step) {
            const int l = addr(pi, g.d, n, g.da, a, g.db, b);
            const int l0 = l + inci, l1 = l + 2 * inci, l2 = l + 3 * inci;
            rop[l] = rop[l0] * c3 + rop[l1] * c1 + rop[l2] * c2;
            const int bad = neq == 6 ? (rop[l] <= 0.3 * rop[l0]) : (fabs(rop[l] / rop[l0] - 1.) >= 0.7);
            for (int e = 1; e < neq; e++)
                rop[l + e * nd] = bad ? rop[l0 + e * nd] : rop[l0 + e * nd] * c3 + rop[l1 + e * nd] * c1 + rop[l2 + e * nd] * c2;
            if (bad) rop[l] = rop[l0];
        }
    }
}
