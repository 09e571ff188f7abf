// fb_lu_prod.cuh -- Jacobian-times-increment product shared by the LU sweep kernels (fb_lu.cu, fb_implicit.cu)
#pragma once

// B_d(l) . dW(l) with the cell-centred mean normal (tcx,tcy,tcz) of cell l (lu_{i,j,k}_*.for; signe folded into cd).
// The reference spells out the 5x5 flux Jacobian; the product is evaluated here in its factored form
//   theta = n.dm - qn drho,  dp = (g-1)(q2 drho - q.dm + dE),  A dW = [n.dm, qn dm + q theta + n dp, qn (dE + dp) + h theta]
// (identical algebra, ~4x fewer FP64 operations, differences at round-off level).
template <int NEQ>
__device__ __forceinline__ void lu_prod(double tcx, double tcy, double tcz, double u, double v, double w, double t,
                                        double cd, double gam, double cp, const double *d, double *b)
{
    const double gamm1 = gam - 1.;
    const double q2 = .5 * (u * u + v * v + w * w), h = cp * t + q2;
    const double qn = tcx * u + tcy * v + tcz * w;
    const double ndm = tcx * d[1] + tcy * d[2] + tcz * d[3];
    const double theta = ndm - qn * d[0];
    const double dp = gamm1 * ((q2 * d[0] - (u * d[1] + v * d[2] + w * d[3])) + d[4]);
    b[0] = ndm + cd * d[0];
    b[1] = (qn * d[1] + u * theta) + (tcx * dp + cd * d[1]);
    b[2] = (qn * d[2] + v * theta) + (tcy * dp + cd * d[2]);
    b[3] = (qn * d[3] + w * theta) + (tcz * dp + cd * d[3]);
    b[4] = (qn * (d[4] + dp) + h * theta) + cd * d[4];
    if (NEQ == 6) b[NEQ - 1] = (qn + cd) * d[NEQ - 1];
}


// 1/x for the sweep's critical path: MUFU seed + two Newton steps (relative error below 2 ulp), 5 dependent operations
// instead of the ~12-deep IEEE division sequence with its slow-path branch.  The operands here (rho + drho, the LU
// diagonals coe5 / coe6 >= 1, T + Cs) are positive and far from the denormal range.
__device__ __forceinline__ double lu_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.);
    y = fma(y, e, y);
    e = fma(-x, y, 1.);
    return fma(y, e, y);
}

// ---- full-Jacobian form of the same product, read straight from the zone arrays (hyperplane kernel k_lu of fb_implicit.cu,
// U-sum of the nb_relax >= 2 relaxations in fb_lu.cu)
// B_d(l1) * dW(l1) with the cell-centred mean normal of the neighbour cell l1 (lu_{i,j,k}_*.for)
template <int MESH, int NEQ>
__device__ __forceinline__ void lu_term(const ZoneDev &z, const double *__restrict__ rop, int d, int l1, int lta, int ltb,
                                        double signe, bool updated, double *b)
{
    const int nd = z.nd, nm = z.nm;
    double tcx = 0., tcy = 0., tcz = 0.;
    const double *T = d == 0 ? z.ti : (d == 1 ? z.tj : z.tk);
    if (MESH == 0) { tcx = 0.5 * (T[ltb] + T[lta]); tcy = 0.5 * (T[ltb + nm] + T[lta + nm]); tcz = 0.5 * (T[ltb + 2 * nm] + T[lta + 2 * nm]); }
    else if (MESH == 1) { if (d < 2) { tcx = 0.5 * (T[ltb] + T[lta]); tcy = 0.5 * (T[ltb + nm] + T[lta + nm]); } else tcz = T[ltb]; }
    else { if (d == 0) tcx = z.tc0[0]; else if (d == 1) tcy = z.tc0[1]; else tcz = z.tc0[2]; }
    const double *dw = z.drodm;
    const double d1 = dw[l1], d2 = dw[l1 + nd], d3 = dw[l1 + 2 * nd], d4 = dw[l1 + 3 * nd], d5 = dw[l1 + 4 * nd];
    double u = rop[l1 + nd], v = rop[l1 + 2 * nd], w = rop[l1 + 3 * nd], t = rop[l1 + 4 * nd];
    if (updated) {   // mjr_newton.for applied on the fly to the k+1 neighbour
        const double ro_old = rop[l1];
        const double cvinv = 1. / z.cv, cvinv2 = 0.5 * cvinv;
        const double r_1 = 1. / (ro_old + d1);
        const double un = (ro_old * u + d2) * r_1, vn = (ro_old * v + d3) * r_1, wn = (ro_old * w + d4) * r_1;
        const double roe_old = ro_old * (z.cv * t + 0.5 * (u * u + v * v + w * w));
        t = (roe_old + d5) * r_1 * cvinv - cvinv2 * (un * un + vn * vn + wn * wn);
        u = un; v = vn; w = wn;
    }
    const double gam1 = z.gam, gamm1 = gam1 - 1., gam2 = gamm1 - 1., cp = z.gam * z.cv;
    const double q2 = .5 * (u * u + v * v + w * w), h = cp * t + q2, ph2 = gamm1 * q2;
    const double qn = tcx * u + tcy * v + tcz * w;
    const double cd = z.coe[l1 + (d + 1) * nd] * signe;
    const double b11 = cd, b12 = tcx, b13 = tcy, b14 = tcz;
    const double b21 = (tcx * ph2 - u * qn), b22 = (qn - tcx * gam2 * u) + cd, b23 = (tcy * u - tcx * gamm1 * v), b24 = (tcz * u - tcx * gamm1 * w), b25 = tcx * gamm1;
    const double b31 = (tcy * ph2 - v * qn), b32 = (tcx * v - tcy * gamm1 * u), b33 = (qn - tcy * gam2 * v) + cd, b34 = (tcz * v - tcy * gamm1 * w), b35 = tcy * gamm1;
    const double b41 = (tcz * ph2 - w * qn), b42 = (tcx * w - tcz * gamm1 * u), b43 = (tcy * w - tcz * gamm1 * v), b44 = (qn - tcz * gam2 * w) + cd, b45 = tcz * gamm1;
    const double b51 = qn * (ph2 - h), b52 = (tcx * h - gamm1 * u * qn), b53 = (tcy * h - gamm1 * v * qn), b54 = (tcz * h - gamm1 * w * qn), b55 = (gam1 * qn) + cd;
    b[0] = b11 * d1 + b12 * d2 + b13 * d3 + b14 * d4;
    b[1] = b21 * d1 + b22 * d2 + b23 * d3 + b24 * d4 + b25 * d5;
    b[2] = b31 * d1 + b32 * d2 + b33 * d3 + b34 * d4 + b35 * d5;
    b[3] = b41 * d1 + b42 * d2 + b43 * d3 + b44 * d4 + b45 * d5;
    b[4] = b51 * d1 + b52 * d2 + b53 * d3 + b54 * d4 + b55 * d5;
    if (NEQ == 6) b[NEQ - 1] = (qn + cd) * dw[l1 + 5 * nd];
}

