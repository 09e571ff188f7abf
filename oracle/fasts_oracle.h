/*
 * fasts_oracle.h -- CPU restatement of FastS's residual-and-update hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (fast_b200/) may link,
 * import or call this file; it exists so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs can check and time the reference's
 * algorithm.  PARITY UNPINNED: the reference ships no golden vectors for this path
 * and its Fortran cannot be compiled here (no Fortran compiler, no Cassiopee), so
 * this restatement is pinned only by in-source known answers and analytic invariants
 * (see tests/test_oracle_*.py and DESIGN.md).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/Fast).  Real literals follow the reference build's r8 convention
 * (ELSAPROD=x86_r8, FastS/setup.scons:74): all reals are FP64.
 */
#ifndef FASTS_ORACLE_H
#define FASTS_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fo_zone {
    int    *pi;                 /* Parameter_int  (FastS/param_solver.h:80-216)   */
    double *pr;                 /* Parameter_real (FastS/param_solver.h:228-312)  */
    double *ro, *ro_m1, *ro_p1; /* rop(NDIMDX,NEQ): Density / _M1 / _P1 blocks     */
    double *xmut;               /* [mu+mut | dist] (FastS/Compute/src_term.for:96) */
    double *x, *y, *z;          /* node coordinates incl. ghost nodes              */
    double *ti, *tj, *tk, *vol; /* metrics (FastS/Metric/allocate_metric.cpp:466)  */
    double *drodm;              /* rhs work array  NDIMDX*NEQ                      */
    double *drodm2;             /* LU output (ssor==NULL => in place)              */
    double *coe;                /* NDIMDX*NEQ_COE                                  */
    double *wig;                /* NDIMDX*3                                        */
    double *rdm;                /* residual norms (nssiter,neq,2,nblk)             */
    double  cfl[3];             /* max, min, sum                                   */
} fo_zone;

/* metrics: FastC/Metric/cp_tijk.for:6, cp_vol.for:6 (whole zone incl. ghost cells) */
void fo_cp_tijk(const int *pi, const double *x, const double *y, const double *z,
                double *ti, double *tj, double *tk);
void fo_cp_vol(const int *pi, const double *x, const double *y, const double *z,
               const double *ti, const double *tj, const double *tk, double *vol);

/* pieces of navier_stokes_struct.for:216-496, callable one by one */
void fo_invist(const int *pi, const double *pr, const int ind[6], const double *rop, double *xmut);
void fo_vispalart(const int *pi, const double *pr, const int ind[6], const double *rop, double *xmut);
void fo_tstb3_global(const int *pi, const double *pr, const int ind[6], double *coe, const double *vol);
void fo_tstb3c(const int *pi, const double *pr, int first_it, const int ind[6], const double *rop,
               double *coe, const double *xmut, const double *ti, const double *tj,
               const double *tk, const double *vol);
void fo_tstb3_impli(const int *pi, const double *pr, int first_it, const int ind[6], const double *rop,
               double *coe, const double *xmut, const double *ti, const double *tj,
               const double *tk, const double *vol);
void fo_tstb3_expli(const int *pi, const double *pr, const int ind[6], const double *rop,
               double *coe, const double *xmut, const double *ti, const double *tj,
               const double *tk, const double *vol);
void fo_cpcfl3(const int *pi, const double *pr, const int ind[6], double cfl[3], const double *rop, const double *xmut,
               const double *ti, const double *tj, const double *tk, const double *vol);
void fo_init_rhs(const int *pi, int nitcfg, const int ind[6], double *drodm);
void fo_wiggle2001(const int *pi, const double *pr, const int ind[6], const double *rop, double *wig);
void fo_spsource_SA(const int *pi, const double *pr, int nitcfg, const int ind[6], const double *xmut,
                    const double *rop, double *coe, const double *ti, const double *tj,
                    const double *tk, const double *vol, const double *dist, double *drodm);
/* mask bit0: do not write the k-lower neighbour on the first k plane, bit1: do not write the
   ghost cell of the k-upper complement face (only the OpenMP slab decomposition sets them) */
void fo_flux_balance(const int *pi, const double *pr, const int ind[6], int mask,
                     const double *rop, double *drodm, const double *wig,
                     const double *ti, const double *tj, const double *tk,
                     const double *vol, const double *xmut);
void fo_correct_flux(const int *pi, const double *pr, const double *rop, double *drodm, const double *wig,
                     const double *ti, const double *tj, const double *tk, const double *vol, const double *xmut);
void fo_core3ark3(const int *pi, const double *pr, int nitcfg, const int ind[6],
                  double *rop_1, const double *rop, const double *drodm, const double *coe);
void fo_core3as2(const int *pi, const double *pr, int nitcfg, int first_it, const int ind[6],
                 const double *rop_1, const double *rop_n, const double *rop_n1,
                 double *drodm, const double *coe);
void fo_cprdu3s1(const int *pi, int nitcfg, const int ind[6], const double *drodm, double *rdm_blk);
/* max |rhs - (D + L + U) dW| over the sweep box, with the operators of fo_invlu (pin of the nb_relax >= 2 relaxations) */
double fo_lu_residual(const int *pi, const double *pr, const int box[6], const double *rop, const double *rhs, const double *dw,
                      const double *coe, const double *ti, const double *tj, const double *tk);
void fo_invlu_relax(const int *pi, const double *pr, const int box[6], const double *rop, double *rop_out, double *drodm,
                    const double *coe, const double *ti, const double *tj, const double *tk, int nrelax);
void fo_invlu(const int *pi, const double *pr, const int ind_lu[6], const double *rop,
              double *rop_out, double *drodm, const double *coe,
              const double *ti, const double *tj, const double *tk);

/* the same sweeps by a team of `nper` threads (every thread of the team calls it with its rank; prog: 2 * nper zeroed counters);
   bit-identical to fo_invlu, used by the timed CPU baseline only (fo_set_lu_team) */
void fo_invlu_team(const int *pi, const double *pr, const int box[6], const double *rop, double *rop_out, double *drodm,
                   const double *coe, const double *ti, const double *tj, const double *tk, int rank, int nper, long *prog);
/* 1: fo_gsdr3 spreads the threads left over after one per zone over the zones' sweeps (fo_invlu_team); 0 (default): one zone per thread */
void fo_set_lu_team(int on);
int fo_get_lu_team(void);

/* ghost cells: FastS/BC/BCzone.cpp:60, correct_coins.for */
void fo_bczone(const int *pi, const double *pr, int lrhs, double *rop, const double *ti,
               const double *tj, const double *tk, int ishift_lu[6]);
void fo_correct_coins(const int *pi, const int ind[6], double *rop);

/* match joins: FastC/Com/setInterpTransfersFast.cpp:440-955 restricted to 1-donor joins */
void fo_transfers(fo_zone *zones, int nz, const int *pitc, const double *prtc, int which /*0 rop_CL, 1 drodm*/,
                  int nitcfg);

/* one sub-iteration of every zone (gsdr3.cpp:57-484) and the NIT x nssiter driver (computePT.cpp:534-679) */
void fo_gsdr3(fo_zone *zones, int nz, int nitcfg, int nssiter, int nitrun, int first_it,
              int lssiter_verif, int lexit_lu, const int *pitc, const double *prtc, int nthreads);
void fo_compute(fo_zone *zones, int nz, int nitrun, int nit, int nssiter, int modulo_verif,
                int *first_it, const int *pitc, const double *prtc, int nthreads);

void fo_fill_ghost(fo_zone *zones, int nz, int level, const int *pitc, const double *prtc);
int fo_max_threads(void);
#ifdef __cplusplus
}
#endif
#endif
