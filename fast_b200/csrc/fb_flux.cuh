// fb_flux.cuh -- face flux device functions (MUSCL + AUSM+(P) / SENSOR / Roe inviscid, face-gradient viscous).
//
// Reference semantics: S/Compute/template_FluxAndBalance.for:108-190 (constants), {AUSM,SENSOR,ROE}/fluFaceEuler.for,
// Slope/*_slope_var.for, etat_GD.for, etat_roe_GD.for, Vit_ent/{qn,fluvector,fludiffer}_*, fluvisq_{3dfull,3dhomo,
// 3dcart}_{i,j,k}.for, mut_interface.for.  The reference generates 96 routines per scheme; here one template
// <MESH, FLOW, SCHEME, SLOPE, DIR> is instantiated by the launcher.  Everything is FP64; a face is described by the
// cell l on its upper side (flux through the LOWER face of cell l in direction DIR).
#pragma once
#include "fb_common.cuh"
#include "fb_view.cuh"

// MESH: 0 3dfull, 1 3dhomo, 2 3dcart.  FLOW: 1 euler, 2 laminar, 3 SA.  SCHEME: 1 ausmpred, 2 senseur, 5 roe.
// SLOPE: 1 o1, 2 o3, 3 minmod.

__device__ __forceinline__ double avmin(double c, double r)
{
    const double s = copysign(1., c);
    return s * fmax(0., fmin(fabs(c), s * r));
}

template <int MESH, int DIR, class M>
__device__ __forceinline__ void face_normal(const ZoneDev &z, const M &mt, int lt, double tc[3], double &s)
{
    tc[0] = tc[1] = tc[2] = 0.;
    if (MESH == 0) {
        const double *T = DIR == 0 ? z.ti : (DIR == 1 ? z.tj : z.tk);
        tc[0] = __ldg(T + lt); tc[1] = __ldg(T + lt + z.nm); tc[2] = __ldg(T + lt + 2 * z.nm);
        s = sqrt(tc[0] * tc[0] + tc[1] * tc[1] + tc[2] * tc[2]);
    } else if (MESH == 1) {
        if (DIR < 2) {
            tc[0] = DIR == 0 ? mt.ti(0, lt) : mt.tj(0, lt); tc[1] = DIR == 0 ? mt.ti(1, lt) : mt.tj(1, lt);
            s = sqrt(tc[0] * tc[0] + tc[1] * tc[1]);
        } else { tc[2] = mt.tk(lt); s = fabs(tc[2]); }
    } else { tc[DIR] = z.tc0[DIR]; s = fabs(tc[DIR]); }
}

template <int SLOPE>
__device__ __forceinline__ void muscl(const ZoneDev &z, double qm2_, double qm1_, double q0, double qp1_, double T0,
                                      double &qL, double &qR)
{   // q(l-2), q(l-1), q(l), q(l+1) -> left (qp in the reference) and right (qm) states of the lower face of l
    if (SLOPE == 2) {
        const double c6 = 1. / 6., c4 = 5. * c6, c5 = 2. * c6, c6m = -1. * c6;
        qR = c4 * q0 + c5 * qm1_ + c6m * qp1_;
        qL = c4 * qm1_ + c6m * qm2_ + c5 * q0;
    } else if (SLOPE == 3) {
        const double roff = .5 * fmax(0., fmin(T0 * z.tmin_1 - .5, 1.));
        const double delm = q0 - qm1_, delp = qp1_ - q0, delq = qm1_ - qm2_;
        qR = q0 - roff * avmin(delm, delp);
        qL = qm1_ + roff * avmin(delq, delm);
    } else { qR = q0; qL = qm1_; }
}

// laminar viscous flux of one face, added into flu[1..4]; returns volinv (needed by the SA diffusion flux)
template <int MESH, int FLOW, int DIR, class V, class M>
__device__ __forceinline__ double visc_face(const ZoneDev &z, const V &v, const M &mt, int l, int lt, const double tc[3], double *flu)
{
    constexpr int D1 = (DIR + 1) % 3, D2 = (DIR + 2) % 3;
    const int il = v.template dn<DIR>(l), ir = l;
    // the eight cross neighbours of the face (il, ir) in the two tangential directions
    const int ilm1 = v.template dn<D1>(il), irm1 = v.template dn<D1>(ir), ilp1 = v.template up<D1>(il), irp1 = v.template up<D1>(ir);
    const int ilm2 = v.template dn<D2>(il), irm2 = v.template dn<D2>(ir), ilp2 = v.template up<D2>(il), irp2 = v.template up<D2>(ir);
    constexpr int NV = FLOW == 3 ? 5 : 4;   // SA: nutilde gradient with the same averaged metrics (SA/fluvisq_SA_*.for)
    double G[NV][3];
#pragma unroll
    for (int c = 0; c < NV; c++) { G[c][0] = G[c][1] = G[c][2] = 0.; }
    double volinv;
    if (MESH == 2) {
        volinv = 0.5 / z.vol0;
#pragma unroll
        for (int c = 0; c < NV; c++) {
            const int e = c + 1;
            const double f1 = v.q(e, il), f2 = v.q(e, ir), fv = f1 + f2;
            G[c][DIR] = (f2 - f1) * z.tc0[DIR];
            if (c >= 3) continue;
            if (c == DIR || c == D1) {
                const double fa = fv + v.q(e, ilm1) + v.q(e, irm1), fb = fv + v.q(e, ilp1) + v.q(e, irp1);
                G[c][D1] = (fb - fa) * (.25 * z.tc0[D1]);
            }
            if (c == DIR || c == D2) {
                const double fa = fv + v.q(e, ilm2) + v.q(e, irm2), fb = fv + v.q(e, ilp2) + v.q(e, irp2);
                G[c][D2] = (fb - fa) * (.25 * z.tc0[D2]);
            }
        }
    } else if (MESH == 0) {
        const int mm[3] = {z.mi, z.mj, z.mk};
        const int m0 = mm[DIR], m1 = mm[D1], m2 = mm[D2];
        const double *T0 = DIR == 0 ? z.ti : (DIR == 1 ? z.tj : z.tk);
        const double *T1 = D1 == 0 ? z.ti : (D1 == 1 ? z.tj : z.tk);
        const double *T2 = D2 == 0 ? z.ti : (D2 == 1 ? z.tj : z.tk);
        double tn[3], tn1[3], ta[3], ta1[3], tb[3], tb1[3];
#pragma unroll
        for (int e = 0; e < 3; e++) {
            const int o = e * z.nm;
            const double t00 = __ldg(T0 + lt + o);
            tn[e] = .5 * (t00 + __ldg(T0 + lt - m0 + o));
            tn1[e] = .5 * (t00 + __ldg(T0 + lt + m0 + o));
            ta[e] = .125 * (__ldg(T1 + lt + o) + __ldg(T1 + lt - m0 + o));
            ta1[e] = .125 * (__ldg(T1 + lt + m1 + o) + __ldg(T1 + lt + m1 - m0 + o));
            tb[e] = .125 * (__ldg(T2 + lt + o) + __ldg(T2 + lt - m0 + o));
            tb1[e] = .125 * (__ldg(T2 + lt + m2 + o) + __ldg(T2 + lt + m2 - m0 + o));
        }
#pragma unroll
        for (int c = 0; c < NV; c++) {
            const int e_ = c + 1;
            const double f1 = v.q(e_, il), f2 = v.q(e_, ir);
            const double f3 = f2 + f1 + v.q(e_, ilm1) + v.q(e_, irm1);
            const double f4 = f2 + f1 + v.q(e_, ilp1) + v.q(e_, irp1);
            const double f5 = f2 + f1 + v.q(e_, ilm2) + v.q(e_, irm2);
            const double f6 = f2 + f1 + v.q(e_, ilp2) + v.q(e_, irp2);
#pragma unroll
            for (int e = 0; e < 3; e++)
                G[c][e] = (f2 * tn1[e] - f1 * tn[e]) + (f4 * ta1[e] - f3 * ta[e]) + (f6 * tb1[e] - f5 * tb[e]);
        }
        volinv = 1. / (__ldg(z.vol + lt) + __ldg(z.vol + lt - m0));
    } else {   // 3dhomo
        if (DIR < 2) {
            constexpr int DP = 1 - (DIR < 2 ? DIR : 0);
            // DP is D1 for DIR = 0 and D2 for DIR = 1; k is the other tangential direction
            const int ilmp = DIR == 0 ? ilm1 : ilm2, irmp = DIR == 0 ? irm1 : irm2, ilpp = DIR == 0 ? ilp1 : ilp2, irpp = DIR == 0 ? irp1 : irp2;
            const int ilmk = DIR == 0 ? ilm2 : ilm1, irmk = DIR == 0 ? irm2 : irm1, ilpk = DIR == 0 ? ilp2 : ilp1, irpk = DIR == 0 ? irp2 : irp1;
            // metric neighbours: one cell back / forward in the face direction, one forward in the in-plane tangential one
            const int ltm = mt.template dn<DIR>(lt), ltp = mt.template up<DIR>(lt), ltq = mt.template up<DP>(lt), ltqm = mt.template dn<DIR>(ltq);
            double tn[2], tn1[2], tp[2], tp1[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const double t00 = DIR == 0 ? mt.ti(e, lt) : mt.tj(e, lt);
                tn[e] = .5 * (t00 + (DIR == 0 ? mt.ti(e, ltm) : mt.tj(e, ltm)));
                tn1[e] = .5 * (t00 + (DIR == 0 ? mt.ti(e, ltp) : mt.tj(e, ltp)));
                tp[e] = .125 * ((DP == 0 ? mt.ti(e, lt) : mt.tj(e, lt)) + (DP == 0 ? mt.ti(e, ltm) : mt.tj(e, ltm)));
                tp1[e] = .125 * ((DP == 0 ? mt.ti(e, ltq) : mt.tj(e, ltq)) + (DP == 0 ? mt.ti(e, ltqm) : mt.tj(e, ltqm)));
            }
            const double tkz = .125 * (mt.tk(lt) + mt.tk(ltm));
#pragma unroll
            for (int c = 0; c < NV; c++) {
                const int e_ = c + 1;
                const double f1 = v.q(e_, il), f2 = v.q(e_, ir);
                const double fa = f2 + f1 + v.q(e_, ilmp) + v.q(e_, irmp), fb = f2 + f1 + v.q(e_, ilpp) + v.q(e_, irpp);
#pragma unroll
                for (int e = 0; e < 2; e++) G[c][e] = (f2 * tn1[e] - f1 * tn[e]) + (fb * tp1[e] - fa * tp[e]);
                if (c < 3) {
                    const double fk0 = f2 + f1 + v.q(e_, ilmk) + v.q(e_, irmk), fk1 = f2 + f1 + v.q(e_, ilpk) + v.q(e_, irpk);
                    G[c][2] = (fk1 - fk0) * tkz;
                }
            }
            volinv = 1. / (mt.vol(lt) + mt.vol(ltm));
        } else {
            const double tiz = mt.tk(lt);
            const int lti = mt.template up<0>(lt), ltj = mt.template up<1>(lt);
            const double tjx = 0.25 * mt.ti(0, lt), tjy = 0.25 * mt.ti(1, lt);
            const double tjx1 = 0.25 * mt.ti(0, lti), tjy1 = 0.25 * mt.ti(1, lti);
            const double tkx = 0.25 * mt.tj(0, lt), tky = 0.25 * mt.tj(1, lt);
            const double tkx1 = 0.25 * mt.tj(0, ltj), tky1 = 0.25 * mt.tj(1, ltj);
#pragma unroll
            for (int c = 0; c < NV; c++) {
                const int e_ = c + 1;
                const double f1 = v.q(e_, il), f2 = v.q(e_, ir);
                G[c][2] = (f2 - f1) * tiz;
                if (c >= 3) continue;
                const double f3 = f2 + f1 + v.q(e_, ilm1) + v.q(e_, irm1), f4 = f2 + f1 + v.q(e_, ilp1) + v.q(e_, irp1);
                const double f5 = f2 + f1 + v.q(e_, ilm2) + v.q(e_, irm2), f6 = f2 + f1 + v.q(e_, ilp2) + v.q(e_, irp2);
                G[c][0] = ((f4 * tjx1 - f3 * tjx) + (f6 * tkx1 - f5 * tkx));
                G[c][1] = ((f4 * tjy1 - f3 * tjy) + (f6 * tky1 - f5 * tky));
            }
            const double vv = mt.vol(lt);
            volinv = 1. / (vv + vv);
        }
    }
    const double div = (G[0][0] + G[1][1] + G[2][2]) * (1. / 3);
    const double f1 = G[0][0] - div, f2 = G[0][1] + G[1][0], f3 = G[0][2] + G[2][0];
    const double f4 = G[1][1] - div, f5 = G[1][2] + G[2][1], f6 = G[2][2] - div;
    double xmutvol, xktvol;
    const double Tl = v.q(4, il), Tr = v.q(4, ir);
    if (FLOW == 3) {
        // SA/mut_interface.for: Sutherland at both cells, one division each (mulam, fb_source.cuh)
        double xmulam = z.coesut * sqrt(Tl * z.temp01) * Tl / (Tl + z.cmus1);
        xmulam = xmulam + z.coesut * sqrt(Tr * z.temp01) * Tr / (Tr + z.cmus1);
        const double xmutot = v.mu(ir) + v.mu(il);
        const double xmutur = xmutot - xmulam;
        xktvol = (xmulam * z.gam3 + xmutur * z.gam4) * volinv;
        xmutvol = xmutot * volinv;
    } else {
        xmutvol = (v.mu(ir) + v.mu(il)) * volinv;
        xktvol = xmutvol * z.gam3;
    }
    const double ur = v.q(1, ir) + v.q(1, il), vr = v.q(2, ir) + v.q(2, il), wr = v.q(3, ir) + v.q(3, il);
    double fv, fv5;
    if (MESH == 0) {
        fv = -(2. * f1 * tc[0] + f2 * tc[1] + f3 * tc[2]) * xmutvol; fv5 = fv * ur; flu[1] += fv;
        fv = -(f2 * tc[0] + 2. * f4 * tc[1] + f5 * tc[2]) * xmutvol; fv5 = fv5 + fv * vr; flu[2] += fv;
        fv = -(f3 * tc[0] + f5 * tc[1] + 2. * f6 * tc[2]) * xmutvol; fv5 = fv5 + fv * wr; flu[3] += fv;
        fv = fv5 * 0.5 - (G[3][0] * tc[0] + G[3][1] * tc[1] + G[3][2] * tc[2]) * xktvol; flu[4] += fv;
    } else if (MESH == 1 && DIR < 2) {
        fv = -(2. * f1 * tc[0] + f2 * tc[1]) * xmutvol; fv5 = fv * ur; flu[1] += fv;
        fv = -(f2 * tc[0] + 2. * f4 * tc[1]) * xmutvol; fv5 = fv5 + fv * vr; flu[2] += fv;
        fv = -(f3 * tc[0] + f5 * tc[1]) * xmutvol; fv5 = fv5 + fv * wr; flu[3] += fv;
        fv = fv5 * 0.5 - (G[3][0] * tc[0] + G[3][1] * tc[1]) * xktvol; flu[4] += fv;
    } else {
        const double t0 = tc[DIR];
        const double fu = (DIR == 0) ? 2. * f1 : (DIR == 1 ? f2 : f3);
        const double fvv = (DIR == 0) ? f2 : (DIR == 1 ? 2. * f4 : f5);
        const double fw = (DIR == 0) ? f3 : (DIR == 1 ? f5 : 2. * f6);
        fv = -(fu * t0) * xmutvol; fv5 = fv * ur; flu[1] += fv;
        fv = -(fvv * t0) * xmutvol; fv5 = fv5 + fv * vr; flu[2] += fv;
        fv = -(fw * t0) * xmutvol; fv5 = fv5 + fv * wr; flu[3] += fv;
        fv = fv5 * 0.5 - (G[3][DIR] * t0) * xktvol; flu[4] += fv;
    }
    (void)f6;
    if (FLOW == 3) {
        // SA/mut_interface.for: Sutherland at both cells, one division each (mulam, fb_source.cuh)
        double xmulam = z.coesut * sqrt(Tl * z.temp01) * Tl / (Tl + z.cmus1);
        xmulam = xmulam + z.coesut * sqrt(Tr * z.temp01) * Tr / (Tr + z.cmus1);
        double xsa = xmulam + v.q(5, il) * v.q(0, il) + v.q(5, ir) * v.q(0, ir);
        xsa = xsa * volinv * z.sigma_1;
        constexpr int C = NV - 1;
        if (MESH == 0) flu[5] = flu[5] - xsa * (G[C][0] * tc[0] + G[C][1] * tc[1] + G[C][2] * tc[2]);
        else if (MESH == 1 && DIR < 2) flu[5] = flu[5] - xsa * (G[C][0] * tc[0] + G[C][1] * tc[1]);
        else flu[5] = flu[5] - xsa * G[C][DIR] * tc[DIR];
    }
    return volinv;
}

#include "fb_sa.cuh"

// l: cell index in the view v; wigb: SENSOR switch byte of the cell (bit d = lower face in direction d; only read
// for SCHEME == 2, loaded once per cell by the caller, early, so that its latency is hidden); lt: metric index
// wcomp (SCHEME == 2, first sub-iteration): the switch of this face is evaluated here from the MUSCL stencil that is
// already in registers (SENSOR/wiggle2001.for + wiggle.for: 1 when some primitive variable has two consecutive sign
// changes of its differences across the face) and OR-ed into wigb; the caller stores the byte for the later stages.
// mt / lt: metric accessor of 3dhomo zones and the cell's index in it (MESH 0 / 2 read z.ti.. / constants with the global lt)
template <int MESH, int FLOW, int SCHEME, int SLOPE, int DIR, class V, class M>
__device__ __forceinline__ void face_flux(const ZoneDev &z, const V &vw, const M &mt, int l, unsigned &wigb, bool wcomp, int lt, double *flu)
{
    constexpr int NEQ = FLOW == 3 ? 6 : 5;
    double tc[3], s;
    face_normal<MESH, DIR>(z, mt, lt, tc, s);
    const int nm = vw.template dn<DIR>(l), np = vw.template up<DIR>(l);
    const double T0 = vw.q(4, l);
    double qL[NEQ], qR[NEQ];
    // wiggle.for: test = min over the variables of sign(1,qm) + sign(1,qp), switch on when test < -1e-12, i.e. when for
    // some variable BOTH products carry the sign bit (sign() copies the sign bit, -0 included): an AND of sign bits
    int wsign = 0;
#pragma unroll
    for (int e = 0; e < 5; e++) {
        const double s0 = vw.template qm2<DIR>(e, l), s1 = vw.q(e, nm), s2 = vw.q(e, l), s3 = vw.q(e, np);
        muscl<SLOPE>(z, s0, s1, s2, s3, T0, qL[e], qR[e]);
        if (SCHEME == 2 && wcomp) {
            const double dc = s2 - s1;
            const double qm = (s1 - s0) * dc, qp = dc * (s3 - s2);
            wsign |= __double2hiint(qm) & __double2hiint(qp);
        }
    }
    if (SCHEME == 2 && wcomp) wigb |= ((unsigned)wsign >> 31) << DIR;
    const double rgp = z.rgp, gam1 = z.gam1;
    double qn1 = qL[1] * tc[0] + qL[2] * tc[1] + qL[3] * tc[2];
    double qn2 = qR[1] * tc[0] + qR[2] * tc[1] + qR[3] * tc[2];
    if (SCHEME == 5) {
        const double r1 = qL[0], p1 = r1 * qL[4] * rgp;
        const double h1 = gam1 * qL[4] * rgp + .5 * (qL[1] * qL[1] + qL[2] * qL[2] + qL[3] * qL[3]);
        const double r2 = qR[0], p2 = r2 * qR[4] * rgp;
        const double h2 = gam1 * qR[4] * rgp + .5 * (qR[1] * qR[1] + qR[2] * qR[2] + qR[3] * qR[3]);
        const double sq1 = sqrt(r1), sq2 = sqrt(r2);
        const double r = sq1 * sq2, r_1 = 1. / (sq1 + sq2);
        const double u = (sq1 * qL[1] + sq2 * qR[1]) * r_1, v = (sq1 * qL[2] + sq2 * qR[2]) * r_1,
                     w = (sq1 * qL[3] + sq2 * qR[3]) * r_1, h = (sq1 * h1 + sq2 * h2) * r_1;
        const double q = .5 * (u * u + v * v + w * w);
        const double cc = sqrt((z.gam - 1.) * fabs(h - q));
        const double du = (qR[1] - qL[1]) * s, dv = (qR[2] - qL[2]) * s, dw = (qR[3] - qL[3]) * s;
        const double dp = p2 - p1, dqn = qn2 - qn1;
        const double s_1 = 1.0 / fmax(s, 1.E-30);
        const double nx = tc[0] * s_1, ny = tc[1] * s_1, nz = tc[2] * s_1;
        const double qn = u * nx + v * ny + w * nz;
        double f1 = qn;
        const double f4 = z.psiroe * (fabs(f1) + cc);
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
        flu[1] = (flu[1] + rou1 * qL[1] + rou2 * qR[1] + tc[0] * p1p2) * 0.5;
        flu[2] = (flu[2] + rou1 * qL[2] + rou2 * qR[2] + tc[1] * p1p2) * 0.5;
        flu[3] = (flu[3] + rou1 * qL[3] + rou2 * qR[3] + tc[2] * p1p2) * 0.5;
        flu[4] = (flu[4] + rou1 * h1 + rou2 * h2) * 0.5;
        if (FLOW == 3) flu[5] = roe_sa_flux(z, vw, l, nm, tc, s_1);
    } else {
        const double r1 = qL[0], rou1 = r1 * qL[1], rov1 = r1 * qL[2], row1 = r1 * qL[3], p1 = r1 * qL[4] * rgp;
        const double h1 = gam1 * p1 + .5 * (rou1 * qL[1] + rov1 * qL[2] + row1 * qL[3]);
        const double r2 = qR[0], rou2 = r2 * qR[1], rov2 = r2 * qR[2], row2 = r2 * qR[3], p2 = r2 * qR[4] * rgp;
        const double h2 = gam1 * p2 + .5 * (rou2 * qR[1] + rov2 * qR[2] + row2 * qR[3]);
        double u, tdu;
        if (SCHEME == 1) {
            const double cs = rgp * z.gam * T0;
            const double son = fabs(qn1) * rsqrt(cs);      // = sqrt(qn1^2 / cs) of fluFaceEuler.for without the division
            const double tam = -2. * son + s;
            const double tam1 = fmax(0., tam) * z.c2;
            u = 0.25 * (qn1 + qn2) - tam1 * (p2 - p1);
            tdu = fmax(fabs(u), z.c1 * s) * z.wig_cte;
        } else {
            const double wg = (double)((wigb >> DIR) & 1);
            u = 0.25 * (qn1 + qn2) - z.c2 * s * (p2 - p1) * (z.opt0 * wg + 1. - z.opt0);
            tdu = fmax(fabs(u), z.c1 * s) * wg;
        }
        const double p1p2 = (p1 + p2) * 0.5;
        flu[0] = u * (r1 + r2) - tdu * (r2 - r1);
        flu[1] = u * (rou1 + rou2) - tdu * (rou2 - rou1) + tc[0] * p1p2;
        flu[2] = u * (rov1 + rov2) - tdu * (rov2 - rov1) + tc[1] * p1p2;
        flu[3] = u * (row1 + row2) - tdu * (row2 - row1) + tc[2] * p1p2;
        flu[4] = u * (h1 + h2) - tdu * (h2 - h1);
        if (FLOW == 3) flu[5] = ausm_sa_flux<SLOPE, DIR>(z, vw, l, nm, np, T0, u, tdu, r1, r2);
    }
    if (FLOW >= 2) visc_face<MESH, FLOW, DIR>(z, vw, mt, l, lt, tc, flu);
}
