/* fo_sa.c -- Spalart-Allmaras pieces (TEST INFRASTRUCTURE, see fasts_oracle.h). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "fo_internal.h"
#include "fasts_oracle.h"

double fo_roe_sa_flux(const double *rop, int nd, int l, int inc, double qn1, double qn2, double r1, double r2,
                      int slope, double c4, double c5, double c6)
{ (void)rop;(void)nd;(void)l;(void)inc;(void)qn1;(void)qn2;(void)r1;(void)r2;(void)slope;(void)c4;(void)c5;(void)c6;
  fprintf(stderr, "fo_roe_sa_flux: not restated yet\n"); abort(); }
double fo_sa_visc_flux(const double *rop, const double *xmut, const double *const t[3], const int inc[3],
                       const int incm[3], int nd, int nm, int typ, const double tc0[3], int d0, int l, int lt,
                       const double tc[3], double volinv, double sigma_1)
{ (void)rop;(void)xmut;(void)t;(void)inc;(void)incm;(void)nd;(void)nm;(void)typ;(void)tc0;(void)d0;(void)l;(void)lt;(void)tc;(void)volinv;(void)sigma_1;
  fprintf(stderr, "fo_sa_visc_flux: not restated yet\n"); abort(); }
