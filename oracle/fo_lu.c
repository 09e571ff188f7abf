/* fo_lu.c -- implicit chain (TEST INFRASTRUCTURE, see fasts_oracle.h); filled in below. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "fo_internal.h"
#include "fasts_oracle.h"
#define NOTYET(name) { fprintf(stderr, #name ": not restated yet\n"); abort(); }
void fo_vispalart(const int *pi, const double *pr, const int ind[6], const double *rop, double *xmut) NOTYET(fo_vispalart)
void fo_tstb3c(const int *pi, const double *pr, int first_it, const int ind[6], const double *rop, double *coe, const double *xmut, const double *ti, const double *tj, const double *tk, const double *vol) NOTYET(fo_tstb3c)
void fo_tstb3_impli(const int *pi, const double *pr, int first_it, const int ind[6], const double *rop, double *coe, const double *xmut, const double *ti, const double *tj, const double *tk, const double *vol) NOTYET(fo_tstb3_impli)
void fo_tstb3_expli(const int *pi, const double *pr, const int ind[6], const double *rop, double *coe, const double *xmut, const double *ti, const double *tj, const double *tk, const double *vol) NOTYET(fo_tstb3_expli)
void fo_spsource_SA(const int *pi, const double *pr, int nitcfg, const int ind[6], const double *xmut, const double *rop, double *coe, const double *ti, const double *tj, const double *tk, const double *vol, const double *dist, double *drodm) NOTYET(fo_spsource_SA)
void fo_core3as2(const int *pi, const double *pr, int nitcfg, int first_it, const int ind[6], const double *rop_1, const double *rop_n, const double *rop_n1, double *drodm, const double *coe) NOTYET(fo_core3as2)
void fo_cprdu3s1(const int *pi, int nitcfg, const int ind[6], const double *drodm, double *rdm_blk) NOTYET(fo_cprdu3s1)
void fo_invlu(const int *pi, const double *pr, const int ind_lu[6], const double *rop, double *rop_out, double *drodm, const double *coe, const double *ti, const double *tj, const double *tk) NOTYET(fo_invlu)
void fo_transfers(fo_zone *zones, int nz, const int *pitc, const double *prtc, int which, int nitcfg) NOTYET(fo_transfers)
