// fb_source.cuh -- per-cell source / assembly pieces fused into k_rhs: SA eddy viscosity (SA/vispalart.for),
// SA source (SA/spsource_SA.for + sourceSA_grad_*.for + sourceSA_prod_dest.for + sourceSA_LU.for) and the
// Newton residual assembly (core3as2.for).
#pragma once
#include "fb_common.cuh"
#include "fb_view.cuh"

__device__ __forceinline__ double mulam(const ZoneDev &z, double T)
{   // S/Compute/mulam.for: coesut sqrt(T/T0) / (1 + Cs/T), written with one division (= coesut sqrt(T/T0) T / (T + Cs))
    return z.coesut * sqrt(T * z.temp01) * T / (T + z.cmus1);
}

// SA/vispalart.for: mu + min(max(rho nutilde fv1, 0), ratiom mu)
__device__ __forceinline__ double vispalart_cell(const ZoneDev &z, const double *__restrict__ rop, int l)
{
    const int nd = z.nd;
    const double amulam = mulam(z, rop[l + 4 * nd]);
    const double anulam = rop[l] / amulam;
    const double anutild = fmax(rop[l + 5 * nd], 1.e-27);
    const double chi = fmax(anutild * anulam, 1.e-7);
    const double fvv1 = 1. / (1. + SA_CV1 / (chi * chi * chi));
    const double amut = rop[l] * anutild * fvv1;
    const double xmuprov = fmax(amut, 0.);
    return amulam + fmin(xmuprov, z.ratiom * amulam);
}

// SA source of cell l: returns vol*tsource; when STORE the LU diagonal coe(:,6) is updated (nitcfg<=1, implicit)
// l: cell index in the view, lg: global cell index (wall distance, coe store)
template <int MESH, class V, class M>
__device__ __forceinline__ double sa_source_cell(const ZoneDev &z, const V &v, const M &mt, int l, int lg, int lt, bool store, double coe1,
                                                 double dist_in = -1.)
{
    const int nd = z.nd, nm = z.nm;
    const int l1 = v.template up<0>(l), l2 = v.template dn<0>(l), l3 = v.template up<1>(l), l4 = v.template dn<1>(l),
              l5 = v.template up<2>(l), l6 = v.template dn<2>(l);
    const double c1 = SA_CB2 / SA_SIGMA;
    const double cw1 = (SA_CB1 / SA_CKARM / SA_CKARM) + (1. + SA_CB2) / SA_SIGMA;
    double g[4][3];      // u,v,w,nutilde gradients (x 2 vol)
    double h[3];         // rho*nutilde gradient (x 2 vol)
    double vol;
    if (MESH == 2) {
        vol = z.vol0;
        const double tix = z.tc0[0], tjy = z.tc0[1], tkz = z.tc0[2];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int e_ = c < 3 ? c + 1 : 5;
            const double q0 = v.q(e_, l);
            const double u1 = q0 + v.q(e_, l2), u2 = q0 + v.q(e_, l3), u3 = q0 + v.q(e_, l1), u4 = q0 + v.q(e_, l4), u5 = q0 + v.q(e_, l5), u6 = q0 + v.q(e_, l6);
            g[c][0] = (u3 - u1) * tix; g[c][1] = (u2 - u4) * tjy; g[c][2] = (u5 - u6) * tkz;
        }
        const double q0 = v.q(5, l) * v.q(0, l);
        const double u1 = q0 + v.q(5, l2) * v.q(0, l2), u2 = q0 + v.q(5, l3) * v.q(0, l3), u3 = q0 + v.q(5, l1) * v.q(0, l1),
                     u4 = q0 + v.q(5, l4) * v.q(0, l4), u5 = q0 + v.q(5, l5) * v.q(0, l5), u6 = q0 + v.q(5, l6) * v.q(0, l6);
        h[0] = g[3][0] * (u3 - u1) * tix; h[1] = g[3][1] * (u2 - u4) * tjy; h[2] = g[3][2] * (u5 - u6) * tkz;
    } else if (MESH == 1) {
        vol = mt.vol(lt);
        const int lti = mt.template up<0>(lt), ltj = mt.template up<1>(lt);
        const double tix = mt.ti(0, lt), tiy = mt.ti(1, lt), tix1 = mt.ti(0, lti), tiy1 = mt.ti(1, lti);
        const double tjx = mt.tj(0, lt), tjy = mt.tj(1, lt), tjx1 = mt.tj(0, ltj), tjy1 = mt.tj(1, ltj);
        const double tkz = mt.tk(lt);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int e_ = c < 3 ? c + 1 : 5;
            const double q0 = v.q(e_, l);
            const double u1 = q0 + v.q(e_, l2), u2 = q0 + v.q(e_, l3), u3 = q0 + v.q(e_, l1), u4 = q0 + v.q(e_, l4), u5 = q0 + v.q(e_, l5), u6 = q0 + v.q(e_, l6);
            g[c][0] = (u3 * tix1 - u1 * tix + u2 * tjx1 - u4 * tjx);
            g[c][1] = (u3 * tiy1 - u1 * tiy + u2 * tjy1 - u4 * tjy);
            g[c][2] = (u5 - u6) * tkz;
        }
        const double q0 = v.q(5, l) * v.q(0, l);
        const double u1 = q0 + v.q(5, l2) * v.q(0, l2), u2 = q0 + v.q(5, l3) * v.q(0, l3), u3 = q0 + v.q(5, l1) * v.q(0, l1),
                     u4 = q0 + v.q(5, l4) * v.q(0, l4), u5 = q0 + v.q(5, l5) * v.q(0, l5), u6 = q0 + v.q(5, l6) * v.q(0, l6);
        h[0] = g[3][0] * (u3 * tix1 - u1 * tix + u2 * tjx1 - u4 * tjx);
        h[1] = g[3][1] * (u3 * tiy1 - u1 * tiy + u2 * tjy1 - u4 * tjy);
        h[2] = g[3][2] * (u5 - u6) * tkz;
    } else {
        vol = __ldg(z.vol + lt);
        double ti0[3], ti1[3], tj0[3], tj1[3], tk0[3], tk1[3];
#pragma unroll
        for (int e = 0; e < 3; e++) {
            ti0[e] = __ldg(z.ti + lt + e * nm); ti1[e] = __ldg(z.ti + lt + z.mi + e * nm);
            tj0[e] = __ldg(z.tj + lt + e * nm); tj1[e] = __ldg(z.tj + lt + z.mj + e * nm);
            tk0[e] = __ldg(z.tk + lt + e * nm); tk1[e] = __ldg(z.tk + lt + z.mk + e * nm);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int e_ = c < 3 ? c + 1 : 5;
            const double q0 = v.q(e_, l);
            const double u1 = q0 + v.q(e_, l2), u2 = q0 + v.q(e_, l3), u3 = q0 + v.q(e_, l1), u4 = q0 + v.q(e_, l4), u5 = q0 + v.q(e_, l5), u6 = q0 + v.q(e_, l6);
#pragma unroll
            for (int e = 0; e < 3; e++)
                g[c][e] = (u3 * ti1[e] - u1 * ti0[e] + u2 * tj1[e] - u4 * tj0[e] + u5 * tk1[e] - u6 * tk0[e]);
        }
        const double q0 = v.q(5, l) * v.q(0, l);
        const double u1 = q0 + v.q(5, l2) * v.q(0, l2), u2 = q0 + v.q(5, l3) * v.q(0, l3), u3 = q0 + v.q(5, l1) * v.q(0, l1),
                     u4 = q0 + v.q(5, l4) * v.q(0, l4), u5 = q0 + v.q(5, l5) * v.q(0, l5), u6 = q0 + v.q(5, l6) * v.q(0, l6);
#pragma unroll
        for (int e = 0; e < 3; e++)
            h[e] = g[3][e] * (u3 * ti1[e] - u1 * ti0[e] + u2 * tj1[e] - u4 * tj0[e] + u5 * tk1[e] - u6 * tk0[e]);
    }
    const double xvol = 0.5 / vol;
    const double rotx = g[2][1] - g[1][2], roty = g[0][2] - g[2][0], rotz = g[1][0] - g[0][1];
    const double rot = sqrt(rotx * rotx + roty * roty + rotz * rotz) * xvol;
    const double anvisc = c1 * (h[0] + h[1] + h[2]) * xvol * xvol;
    // sourceSA_prod_dest.for
    const double rho = v.q(0, l);
    const double amulam = mulam(z, v.q(4, l));
    const double anulam = rho / amulam;
    const double anutild = fmax(v.q(5, l), 1.e-27);
    const double chi = fmax(anutild * anulam, 1.e-7);
    // (the same algebra as sourceSA_prod_dest.for with the divisions shared: fv1 = chi^3 / (chi^3 + cv1), one reciprocal of
    // (kappa d)^2 for S~, r and the destruction term, x^(1/6) as cbrt(sqrt(x)) -- the FP64 pipe is what bounds this kernel, and
    // pow() alone costs more instructions than the rest of the source; differences at round-off level)
    const double chi3 = chi * chi * chi;
    const double f1 = chi3 / (chi3 + SA_CV1);
    const double f2 = 1. - chi / (1. + chi * f1);
    // dist_in: the wall distance when the caller has loaded it ahead of time (k_rhs3); < 0: read here
    const double dist = fmax(dist_in >= 0. ? dist_in : z.xmut[lg + nd], 1.e-27);
    const double ikd2 = 1. / (SA_CKARM * SA_CKARM * dist * dist);
    double stild = rot + (anutild * f2) * ikd2;
    const double prod = rho * SA_CB1 * stild * anutild;
    stild = fmax(stild, 0.00000000000000000001);
    double r = anutild * ikd2 / stild;
    r = fmin(r, 10.);
    const double r5 = r * r * r * r * r - 1.;
    const double gg = r * (1. + SA_CW2 * r5);
    const double g2 = gg * gg, g6 = g2 * g2 * g2;
    const double fwg = gg * cbrt(sqrt((1. + SA_CW3) / (g6 + SA_CW3)));
    const double destruc = rho * cw1 * fwg * (anutild * anutild) * (ikd2 * (SA_CKARM * SA_CKARM));
    const double tsource = (prod + anvisc - destruc);
    if (store) {   // sourceSA_LU.for
        const double tsourceb = prod - destruc;
        double amutild = anutild * rho;
        amutild = fmax(amutild, 0.0000000000000001);
        const double tsourcenu = vol * tsourceb / amutild;
        z.coe[lg + 5 * nd] = FB_LD(z.coe + lg + 4 * nd) - fmin(tsourcenu, 0.) * coe1;
    }
    return vol * tsource;
}

template <int MESH, bool IMPL, class V, class M>
__device__ __forceinline__ void sa_source(const ZoneDev &z, const V &v, const M &mt, int l, int lg, int lt, int nitcfg, double coe1, double *acc,
                                          double dist_in = -1.)
{
    acc[5] += sa_source_cell<MESH>(z, v, mt, l, lg, lt, IMPL && nitcfg <= 1, coe1, dist_in);
}

// core3as2.for:37-195: Newton residual from the flux balance and the BDF time terms (conservative variables)
// r0 (= rop_ssiter) is read through the view v at cell lv; rn / rm at the global cell l
template <int NEQ, class V>
__device__ __forceinline__ void implicit_assemble(const ZoneDev &z, const V &v, int lv, const double *r0,
                                                  const double *__restrict__ rn, const double *__restrict__ rm, int l,
                                                  double coe1, double c1, double c2, double c3, double *acc, const double *npre = nullptr)
{   // npre: W(n) of the cell when the caller has loaded it ahead of time (k_rhs3, only looked at when rn != r0)
    const int nd = z.nd; const double cv = z.cv;
    const bool first = (rn == r0);   // nitcfg == 1: rop_ssiter is ro itself => c4 = c1 + c2
    if (first && z.dtloc == 1) {
#pragma unroll
        for (int e = 0; e < NEQ; e++) acc[e] = coe1 * acc[e];
        return;
    }
    const double a0 = v.q(0, lv), a1 = v.q(1, lv), a2 = v.q(2, lv), a3 = v.q(3, lv), a4 = v.q(4, lv);
    // level n-1 only enters with BDF2 (c3 != 0): first-order / local time stepping skips its loads (0 * W(n-1) = 0)
    double m0 = 0., m1 = 0., m2 = 0., m3 = 0., m4 = 0., m5 = 0.;
    if (c3 != 0.) {
        m0 = FB_LD(rm + l); m1 = FB_LD(rm + l + nd); m2 = FB_LD(rm + l + 2 * nd); m3 = FB_LD(rm + l + 3 * nd); m4 = FB_LD(rm + l + 4 * nd);
        if (NEQ == 6) m5 = FB_LD(rm + l + 5 * nd);
    }
    const double roe = a0 * (cv * a4 + 0.5 * (a1 * a1 + a2 * a2 + a3 * a3));
    const double roe_n1 = m0 * (cv * m4 + 0.5 * (m1 * m1 + m2 * m2 + m3 * m3));
    if (first) {
        const double c4 = c1 + c2;
        acc[0] = coe1 * acc[0] + c4 * a0 + c3 * m0;
        acc[1] = coe1 * acc[1] + c4 * a1 * a0 + c3 * m1 * m0;
        acc[2] = coe1 * acc[2] + c4 * a2 * a0 + c3 * m2 * m0;
        acc[3] = coe1 * acc[3] + c4 * a3 * a0 + c3 * m3 * m0;
        acc[4] = coe1 * acc[4] + c4 * roe + c3 * roe_n1;
        if (NEQ == 6) acc[NEQ - 1] = coe1 * acc[NEQ - 1] + c4 * v.q(5, lv) * a0 + c3 * m5 * m0;
    } else {
        double n0, n1, n2, n3, n4, n5 = 0.;
        if (npre) { n0 = npre[0]; n1 = npre[1]; n2 = npre[2]; n3 = npre[3]; n4 = npre[4]; if (NEQ == 6) n5 = npre[5]; }
        else {
            n0 = FB_LD(rn + l); n1 = FB_LD(rn + l + nd); n2 = FB_LD(rn + l + 2 * nd); n3 = FB_LD(rn + l + 3 * nd); n4 = FB_LD(rn + l + 4 * nd);
            if (NEQ == 6) n5 = FB_LD(rn + l + 5 * nd);
        }
        const double roe_n = n0 * (cv * n4 + 0.5 * (n1 * n1 + n2 * n2 + n3 * n3));
        acc[0] = coe1 * acc[0] + c1 * a0 + c2 * n0 + c3 * m0;
        acc[1] = coe1 * acc[1] + c1 * a1 * a0 + c2 * n1 * n0 + c3 * m1 * m0;
        acc[2] = coe1 * acc[2] + c1 * a2 * a0 + c2 * n2 * n0 + c3 * m2 * m0;
        acc[3] = coe1 * acc[3] + c1 * a3 * a0 + c2 * n3 * n0 + c3 * m3 * m0;
        acc[4] = coe1 * acc[4] + c1 * roe + c2 * roe_n + c3 * roe_n1;
        if (NEQ == 6)
            acc[NEQ - 1] = coe1 * acc[NEQ - 1] + c1 * v.q(5, lv) * a0 + c2 * n5 * n0 + c3 * m5 * m0;
    }
}
