// fb_sa.cuh -- Spalart-Allmaras convective face fluxes (reference: S/Compute/{AUSM,SENSOR,ROE}/fluFaceSA.for).
#pragma once

// AUSM / SENSOR: nutilde reconstructed with the scheme's slope times the MUSCL densities, same u / tdu as the mean flow
template <int SLOPE>
__device__ __forceinline__ double ausm_sa_flux(const ZoneDev &z, const double *__restrict__ rop, int l, int inc,
                                               double T0, double u, double tdu, double r1, double r2)
{
    const double *q = rop + 5 * z.nd;
    double qL, qR;
    muscl<SLOPE>(z, q[l - 2 * inc], q[l - inc], q[l], q[l + inc], T0, qL, qR);
    const double qm6 = qR * r2, qp6 = qL * r1;
    return u * (qm6 + qp6) - tdu * (qm6 - qp6);
}

// Roe: first-order states, Roe-averaged normal velocity as dissipation speed (ROE/fluFaceSA.for:2-34)
__device__ __forceinline__ double roe_sa_flux(const ZoneDev &z, const double *__restrict__ rop, int l, int inc,
                                              const double tc[3], double s_1)
{
    const int nd = z.nd, nm = l - inc;
    double p2 = rop[nm + nd], p3 = rop[nm + 2 * nd], p4 = rop[nm + 3 * nd];
    const double m2 = rop[l + nd], m3 = rop[l + 2 * nd], m4 = rop[l + 3 * nd];
    double n1 = p2 * tc[0] + p3 * tc[1] + p4 * tc[2];
    const double n2 = m2 * tc[0] + m3 * tc[1] + m4 * tc[2];
    double rp1 = rop[nm], rm1 = rop[l];
    const double nul = rop[l + 5 * nd], num = rop[nm + 5 * nd];
    const double qm6 = nul * rm1 * n2, qp6 = num * rp1 * n1;
    const double d6 = nul * rm1 - num * rp1;
    rp1 = sqrt(rp1); rm1 = sqrt(rm1);
    const double r_1 = 1. / (rp1 + rm1);
    p2 = r_1 * (rp1 * p2 + rm1 * m2); p3 = r_1 * (rp1 * p3 + rm1 * m3); p4 = r_1 * (rp1 * p4 + rm1 * m4);
    n1 = p2 * tc[0] + p3 * tc[1] + p4 * tc[2];
    const double tdu = fmax(z.c1 / s_1, fabs(n1));
    return 0.5 * (qm6 + qp6 - tdu * d6);
}
