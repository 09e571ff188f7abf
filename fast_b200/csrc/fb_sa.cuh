// fb_sa.cuh -- Spalart-Allmaras convective face fluxes (reference: S/Compute/{AUSM,SENSOR,ROE}/fluFaceSA.for).
#pragma once

// AUSM / SENSOR: nutilde reconstructed with the scheme's slope times the MUSCL densities, same u / tdu as the mean flow
template <int SLOPE, int DIR, class V>
__device__ __forceinline__ double ausm_sa_flux(const ZoneDev &z, const V &v, int l, int nm, int np,
                                               double T0, double u, double tdu, double r1, double r2)
{
    double qL, qR;
    muscl<SLOPE>(z, v.template qm2<DIR>(5, l), v.q(5, nm), v.q(5, l), v.q(5, np), T0, qL, qR);
    const double qm6 = qR * r2, qp6 = qL * r1;
    return u * (qm6 + qp6) - tdu * (qm6 - qp6);
}

// Roe: first-order states, Roe-averaged normal velocity as dissipation speed (ROE/fluFaceSA.for:2-34)
template <class V>
__device__ __forceinline__ double roe_sa_flux(const ZoneDev &z, const V &v, int l, int nm, const double tc[3], double s_1)
{
    double p2 = v.q(1, nm), p3 = v.q(2, nm), p4 = v.q(3, nm);
    const double m2 = v.q(1, l), m3 = v.q(2, l), m4 = v.q(3, l);
    double n1 = p2 * tc[0] + p3 * tc[1] + p4 * tc[2];
    const double n2 = m2 * tc[0] + m3 * tc[1] + m4 * tc[2];
    double rp1 = v.q(0, nm), rm1 = v.q(0, l);
    const double nul = v.q(5, l), num = v.q(5, nm);
    const double qm6 = nul * rm1 * n2, qp6 = num * rp1 * n1;
    const double d6 = nul * rm1 - num * rp1;
    rp1 = sqrt(rp1); rm1 = sqrt(rm1);
    const double r_1 = 1. / (rp1 + rm1);
    p2 = r_1 * (rp1 * p2 + rm1 * m2); p3 = r_1 * (rp1 * p3 + rm1 * m3); p4 = r_1 * (rp1 * p4 + rm1 * m4);
    n1 = p2 * tc[0] + p3 * tc[1] + p4 * tc[2];
    const double tdu = fmax(z.c1 / s_1, fabs(n1));
    return 0.5 * (qm6 + qp6 - tdu * d6);
}
