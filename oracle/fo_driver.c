/*
 * fo_driver.c -- one sub-iteration for every zone (FastS/Compute/gsdr3.cpp:57-484, rhs.cpp, lhs.cpp,
 * bcs.cpp, transfert_multiblock.cpp) and the NIT x nssiter loop (computePT.cpp:534-679).
 * TEST INFRASTRUCTURE, see fasts_oracle.h.  One task per zone (souszones_list with one sub-zone).
 * nthreads > 1 only changes HOW the work is cut, never a bit of the result (tests/test_oracle_pins.py): k slabs for the rhs; for the
 * LU one zone per thread, or -- fo_set_lu_team(1), the timed CPU baseline with more threads than zones -- teams of threads per zone
 * (fo_invlu_team).  The small parity tests run one thread, the full-size GPU parity tests and the baseline all host threads.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "fo_internal.h"
#include "fasts_oracle.h"

static void ranges(const int *pi, int sdm[6], int grad[6], int coe[6], int ssa[6])
{   /* FastC/HPC_LAYER/indice_boucle_ssdom.for, one thread / one cache block */
    const int nf = pi[FO_NIJK + 3], nfk = pi[FO_NIJK + 4];
    const int inck = nfk == 0 ? 0 : 1;
    for (int d = 0; d < 3; d++) {
        const int n = pi[FO_IJKV + d], f = d == 2 ? nfk : nf, g = d == 2 ? inck : 1;
        sdm[2 * d] = 1; sdm[2 * d + 1] = n;
        grad[2 * d] = 1 - g; grad[2 * d + 1] = n + g;
        coe[2 * d] = 1 - f; coe[2 * d + 1] = n + f;
        ssa[2 * d] = 1 - g; ssa[2 * d + 1] = n + g;
    }
}

/* cut [lo,hi] of direction k (j in 2D) into `n` slabs, slab `t` */
static void slab(const int full[6], int n, int t, int out[6], int dir)
{
    memcpy(out, full, 6 * sizeof(int));
    const int lo = full[2 * dir], hi = full[2 * dir + 1], len = hi - lo + 1;
    const int a = lo + (int)((long)len * t / n), b = lo + (int)((long)len * (t + 1) / n) - 1;
    out[2 * dir] = a; out[2 * dir + 1] = b;
}

/* FastS/Compute/navier_stokes_struct.for:216-496 for one zone */
static void rhs_zone(fo_zone *z, int nitcfg, int nssiter, int nitrun, int first_it, int lssiter_verif,
                     double *rop_ssiter, int nthreads)
{
    (void)nssiter; (void)nitrun;
    const int *pi = z->pi; const double *pr = z->pr;
    int sdm[6], grad[6], coe[6], ssa[6];
    ranges(pi, sdm, grad, coe, ssa);
    const int implicit = pi[FO_ITYPCP] <= 1;
    const int nd = pi[FO_NDIMDX];
    const int sdir = pi[FO_NIJK + 4] == 0 ? 1 : 2;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > sdm[2 * sdir + 1]) nthreads = sdm[2 * sdir + 1];

#pragma omp parallel num_threads(nthreads) if (nthreads > 1)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        const int t = 0, nt = 1;
#endif
        int s_sdm[6], s_grad[6], s_coe[6], s_ssa[6];
        slab(sdm, nt, t, s_sdm, sdir); slab(grad, nt, t, s_grad, sdir);
        slab(coe, nt, t, s_coe, sdir); slab(ssa, nt, t, s_ssa, sdir);
        if (nitcfg == 1) {   /* :216-301 */
            if (pi[FO_IFLOW] == 2) fo_invist(pi, pr, implicit ? s_coe : s_grad, rop_ssiter, z->xmut);
            else if (pi[FO_IFLOW] == 3) fo_vispalart(pi, pr, s_coe, rop_ssiter, z->xmut);
#pragma omp barrier
            /* cptst3.for:38-121 */
            if (pi[FO_DTLOC] == 1) {
                if (implicit) fo_tstb3_impli(pi, pr, first_it, s_coe, rop_ssiter, z->coe, z->xmut, z->ti, z->tj, z->tk, z->vol);
                else fo_tstb3_expli(pi, pr, s_sdm, rop_ssiter, z->coe, z->xmut, z->ti, z->tj, z->tk, z->vol);
            } else if (implicit) {
                fo_tstb3_global(pi, pr, s_coe, z->coe, z->vol);
                fo_tstb3c(pi, pr, first_it, s_coe, rop_ssiter, z->coe, z->xmut, z->ti, z->tj, z->tk, z->vol);
            } else fo_tstb3_global(pi, pr, s_sdm, z->coe, z->vol);
            /* cptst3.for:113-120: CFL statistics at the first sub-iteration of a check iteration; initial values and the final
               division by the cell count: computePT.cpp:519-522, 619-645 */
            if (lssiter_verif != 0) {
#pragma omp barrier
#pragma omp single
                {
                    z->cfl[0] = 0.; z->cfl[1] = 100000000.; z->cfl[2] = 0.;
                    fo_cpcfl3(pi, pr, sdm, z->cfl, rop_ssiter, z->xmut, z->ti, z->tj, z->tk, z->vol);
                    z->cfl[2] = z->cfl[2] / ((double)sdm[1] * (double)sdm[3] * (double)sdm[5]);
                }
            }
        }
#pragma omp barrier
        /* src_term.for: init_rhs, SA source, sensor */
        fo_init_rhs(pi, nitcfg, s_sdm, z->drodm);
        /* the slabs of ind_ssa (interior +- 1) are not those of ind_sdm: a plane may be zeroed by one thread and receive its
           source from another */
#pragma omp barrier
        if (pi[FO_IFLOW] == 3 && pi[FO_ILES] == 0)
            fo_spsource_SA(pi, pr, nitcfg, s_ssa, z->xmut, rop_ssiter, z->coe, z->ti, z->tj, z->tk, z->vol,
                           z->xmut + nd, z->drodm);
        if (pi[FO_KFLUDOM] == 2 && nitcfg == 1) fo_wiggle2001(pi, pr, s_sdm, rop_ssiter, z->wig);
#pragma omp barrier
        {
            int mask = 0;
            if (nt > 1) { if (t > 0) mask |= 1; if (t < nt - 1) mask |= 2; }
            if (sdir == 1 && nt > 1) { fprintf(stderr, "oracle: slab threading needs a 3-D zone\n"); abort(); }
            fo_flux_balance(pi, pr, s_sdm, mask, rop_ssiter, z->drodm, z->wig, z->ti, z->tj, z->tk, z->vol, z->xmut);
        }
#pragma omp barrier
        /* navier_stokes_struct.for:376-383: wall flux correction (Roe) between the balance and the update */
#pragma omp single
        fo_correct_flux(pi, pr, rop_ssiter, z->drodm, z->wig, z->ti, z->tj, z->tk, z->vol, z->xmut);
        if (implicit) fo_core3as2(pi, pr, nitcfg, first_it, s_sdm, rop_ssiter, z->ro, z->ro_m1, z->drodm, z->coe);
        else if (nitcfg % 2 == 1) fo_core3ark3(pi, pr, nitcfg, s_sdm, z->ro_p1, z->ro, z->drodm, z->coe);
        else fo_core3ark3(pi, pr, nitcfg, s_sdm, z->ro, z->ro_p1, z->drodm, z->coe);
    }
}

static int g_lu_team = 0;
void fo_set_lu_team(int on) { g_lu_team = on != 0; }
int fo_get_lu_team(void) { return g_lu_team; }

void fo_gsdr3(fo_zone *zones, int nz, int nitcfg, int nssiter, int nitrun, int first_it,
              int lssiter_verif, int lexit_lu, const int *pitc, const double *prtc, int nthreads)
{
    const int implicit = zones[0].pi[FO_ITYPCP] <= 1;
    const int kimpli = implicit;
    for (int n = 0; n < nz; n++) {
        fo_zone *z = &zones[n];
        double *ssiter;
        if (nitcfg == 1) ssiter = z->ro;
        else if (implicit) ssiter = z->ro_p1;
        else ssiter = (nitcfg % 2 == 0) ? z->ro_p1 : z->ro;
        rhs_zone(z, nitcfg, nssiter, nitrun, first_it, lssiter_verif, ssiter, nthreads);
    }
    if (kimpli) {
        /* lhs.cpp:1-15: lu_match => the ghost rows of drodm come from the neighbours */
        if (zones[0].pi[FO_LU_MATCH] == 1 && pitc != NULL) fo_transfers(zones, nz, pitc, prtc, 1, nitcfg);
        int *shift = (int *)malloc(6 * nz * sizeof(int));
        for (int n = 0; n < nz; n++) {   /* lhs.cpp:57-67 */
            fo_zone *z = &zones[n];
            fo_bczone(z->pi, z->pr, 1, z->drodm, z->ti, z->tj, z->tk, shift + 6 * n);
            int lcorner = 1;
            {   /* lcorner=0 as soon as one wall-type window exists (BCzone.cpp:118-131) */
                const int pt_bcs = z->pi[FO_PT_BC], nb = z->pi[pt_bcs];
                for (int b = 0; b < nb; b++) {
                    const int ty = z->pi[z->pi[pt_bcs + 1 + b] + FO_BC_TYPE];
                    if ((ty >= 3 && ty <= 7) || ty == 12 || ty == 18 || ty == 30) lcorner = 0;
                }
            }
            if (lcorner == 0) fo_correct_coins(z->pi, shift + 6 * n, z->drodm);
        }
        if (lssiter_verif == 1)
            for (int n = 0; n < nz; n++) {
                fo_zone *z = &zones[n];
                const int ind[6] = {1, z->pi[FO_IJKV], 1, z->pi[FO_IJKV + 1], 1, z->pi[FO_IJKV + 2]};
                if (z->rdm) fo_cprdu3s1(z->pi, nitcfg, ind, z->drodm, z->rdm);
            }
        if (lexit_lu == 0) {
            /* threads: one zone per thread; with more threads than zones (and fo_set_lu_team(1): the timed baseline) the zones are
               swept by teams of nper threads each (fo_invlu_team, same bits) */
            const int nteam = nz < nthreads ? nz : nthreads, nper = nteam > 0 ? nthreads / nteam : 1;
            int teams_ran = 0;
            if (g_lu_team && nper > 1) {
                long *prog = (long *)calloc((size_t)nz * 2 * nper, sizeof(long));
                int ok = 1;
#pragma omp parallel num_threads(nteam * nper)
                {
                    if (omp_get_num_threads() != nteam * nper) {      /* fewer threads than asked for: a team would wait forever */
#pragma omp single
                        ok = 0;
                    } else {
                        const int tid = omp_get_thread_num(), team = tid / nper, rank = tid % nper;
                        for (int n = team; n < nz; n += nteam) {
                            fo_zone *z = &zones[n];
                            fo_invlu_team(z->pi, z->pr, shift + 6 * n, z->ro_p1 == NULL ? NULL : (nitcfg == 1 ? z->ro : z->ro_p1),
                                          z->ro_p1, z->drodm, z->coe, z->ti, z->tj, z->tk, rank, nper, prog + (size_t)n * 2 * nper);
                        }
                    }
                }
                free(prog);
                teams_ran = ok;
            }
            if (!teams_ran) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) if (nthreads > 1)
                for (int n = 0; n < nz; n++) {
                    fo_zone *z = &zones[n];
                    fo_invlu(z->pi, z->pr, shift + 6 * n, z->ro_p1 == NULL ? NULL : (nitcfg == 1 ? z->ro : z->ro_p1),
                             z->ro_p1, z->drodm, z->coe, z->ti, z->tj, z->tk);
                }
            }
        }
        free(shift);
    }
    if (lexit_lu == 0) {
        if (pitc != NULL) fo_transfers(zones, nz, pitc, prtc, 0, nitcfg);
        for (int n = 0; n < nz; n++) {   /* bcs.cpp:32-40 */
            fo_zone *z = &zones[n];
            double *cl;
            if (nitcfg == 1 || implicit) cl = z->ro_p1; else cl = (nitcfg % 2 == 0) ? z->ro : z->ro_p1;
            int dummy[6];
            fo_bczone(z->pi, z->pr, 0, cl, z->ti, z->tj, z->tk, dummy);
            const int ind[6] = {1, z->pi[FO_IJKV], 1, z->pi[FO_IJKV + 1], 1, z->pi[FO_IJKV + 2]};
            fo_correct_coins(z->pi, ind, cl);
        }
    }
}

void fo_compute(fo_zone *zones, int nz, int nitrun, int nit, int nssiter, int modulo_verif,
                int *first_it, const int *pitc, const double *prtc, int nthreads)
{
    const int itypcp = zones[0].pi[FO_ITYPCP];
    for (int it = 0; it < nit; it++) {
        const int nitrun_loc = nitrun * nit + it;
        for (int nstep = 1; nstep <= nssiter; nstep++) {
            /* souszones_list.cpp:36-83 */
            int lssiter_verif = 0, lexit_lu = 0;
            if (nitrun_loc % modulo_verif == 0) { lssiter_verif = 1; if (nstep == nssiter && itypcp != 2) lexit_lu = 1; }
            if (lssiter_verif == 0 && nstep == nssiter && itypcp == 1) continue;   /* computePT.cpp:582 */
            fo_gsdr3(zones, nz, nstep, nssiter, nitrun_loc, *first_it, lssiter_verif, lexit_lu, pitc, prtc, nthreads);
        }
        *first_it = 1;
        /* computePT.cpp:675-679 swaps local pointers only; the persistent rotation is FastS._compute's switchPointers__(zones,
           case, nitmax%2+2) (S/PyTree.py:226-228), a no-op (order 2, C/PyTree.py:2299) for explicit schemes with an even number
           of stages: RK2's last stage has written `ro` itself (gsdr3.cpp:162-173) */
        if (itypcp == 2 && nssiter % 2 == 0) continue;
        for (int n = 0; n < nz; n++) {
            double *sv = zones[n].ro_m1; zones[n].ro_m1 = zones[n].ro; zones[n].ro = zones[n].ro_p1; zones[n].ro_p1 = sv;
        }
    }
}

/* warmup's first ghost fill (S/PyTree.py:536, _fillGhostcells): transfers then BCs on one time level */
void fo_fill_ghost(fo_zone *zones, int nz, int level, const int *pitc, const double *prtc)
{
    if (pitc != NULL) {
        /* fo_transfers works on rop_CL = ro_p1 at nitcfg 1: alias the requested level there */
        double **sv = (double **)malloc(nz * sizeof(double *));
        for (int n = 0; n < nz; n++) { sv[n] = zones[n].ro_p1; zones[n].ro_p1 = level == 0 ? zones[n].ro : (level == 1 ? zones[n].ro_m1 : zones[n].ro_p1); }
        fo_transfers(zones, nz, pitc, prtc, 0, 1);
        for (int n = 0; n < nz; n++) zones[n].ro_p1 = sv[n];
        free(sv);
    }
    for (int n = 0; n < nz; n++) {
        fo_zone *z = &zones[n];
        double *a = level == 0 ? z->ro : (level == 1 ? z->ro_m1 : z->ro_p1);
        int dummy[6];
        fo_bczone(z->pi, z->pr, 0, a, z->ti, z->tj, z->tk, dummy);
        const int ind[6] = {1, z->pi[FO_IJKV], 1, z->pi[FO_IJKV + 1], 1, z->pi[FO_IJKV + 2]};
        fo_correct_coins(z->pi, ind, a);
    }
}
