/* fo_internal.h -- address helpers shared by the oracle translation units (TEST INFRASTRUCTURE). */
#ifndef FO_INTERNAL_H
#define FO_INTERNAL_H
#include "fo_param.h"

/* FastS/formule_param.h: inddm, 0-based */
static inline int fo_inddm(const int *pi, int i, int j, int k)
{
    return (i + pi[FO_NIJK + 3] - 1) + (j + pi[FO_NIJK + 3] - 1) * pi[FO_NIJK]
         + (k + pi[FO_NIJK + 4] - 1) * pi[FO_NIJK] * pi[FO_NIJK + 1];
}
/* FastS/formule_mtr_param.h: indmtr, 0-based */
static inline int fo_indmtr(const int *pi, int i, int j, int k)
{
    return (i + pi[FO_NIJK_MTR + 3] - 1) * pi[FO_NIJK_MTR] + (j + pi[FO_NIJK_MTR + 3] - 1) * pi[FO_NIJK_MTR + 1]
         + (k + pi[FO_NIJK_MTR + 4] - 1) * pi[FO_NIJK_MTR + 2];
}
/* FastC/formule_xyz_param.h: indcg (node address), 0-based */
static inline int fo_indcg(const int *pi, int i, int j, int k)
{
    return (i + pi[FO_NIJK_XYZ + 3] - 1) + (j + pi[FO_NIJK_XYZ + 3] - 1) * pi[FO_NIJK_XYZ]
         + (k + pi[FO_NIJK_XYZ + 4] - 1) * pi[FO_NIJK_XYZ] * pi[FO_NIJK_XYZ + 1];
}

#define FO_LOOP(ind) \
    for (int k = (ind)[4]; k <= (ind)[5]; k++) \
    for (int j = (ind)[2]; j <= (ind)[3]; j++) \
    for (int i = (ind)[0]; i <= (ind)[1]; i++)

/* fo_bc.c */
void fo_bc_extrapolate(const int *pi, int idir, int lrhs, const int w[6], double *rop);
void fo_bc_periodic(const int *pi, int idir, int lrhs, const int w[6], double *rop);
void fo_bc_wall_inviscid(const int *pi, const double *pr, int idir, int lrhs, const int w[6], double *rop, const double *tijk);
void fo_bc_wall_viscous(const int *pi, const double *pr, int idir, int lrhs, const int w[6], double *rop);
void fo_bc_farfield(const int *pi, const double *pr, int idir, int lrhs, const int w[6], double *rop, const double *tijk);

#endif
