// fb_lu_prod.cuh -- Jacobian-times-increment product shared by the LU sweep kernels (fb_lu.cu, fb_lu2.cu)
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
