/*
 * fo_param.h -- slot numbers of the reference's flat per-zone tables.
 *
 * TEST INFRASTRUCTURE (oracle).  The numeric values are the reference's wire
 * format for `Parameter_int` / `Parameter_real` (Fast/FastS/FastS/param_solver.h:80-312);
 * a drop-in has to use the same slots, so they are restated here under an FO_ prefix.
 */
#ifndef FO_PARAM_H
#define FO_PARAM_H

/* Parameter_int (param_solver.h:80-216) */
#define FO_NIJK        0   /* ni,nj,nk (cells incl. ghosts), nfic_ij, nfic_k */
#define FO_NIJK_MTR    5   /* metric strides i,j,k + ghost counts            */
#define FO_NIJK_XYZ   10   /* node dims + ghost counts                       */
#define FO_IJKV       20   /* interior cells                                 */
#define FO_MXSSDOM_LU 24
#define FO_ITYPZONE   25   /* 0 3dfull, 1 3dhomo, 2 3dcart, 3 2d             */
#define FO_IFLOW      27   /* 1 euler, 2 laminar, 3 SA                       */
#define FO_ILES       28
#define FO_ITYPCP     29   /* <=1 implicit, 2 explicit                       */
#define FO_SIZE_SSDOM 30
#define FO_KFLUDOM    33   /* 1 ausmpred, 2 senseur, 5 roe                   */
#define FO_LALE       34
#define FO_NEQ        36
#define FO_NEQ_IJ     37
#define FO_NEQ_K      38
#define FO_NEQ_COE    39
#define FO_NDIMDX     41
#define FO_NDIMDX_XYZ 42
#define FO_NDIMDX_MTR 43
#define FO_DTLOC      46
#define FO_SA_INT     47
#define FO_RK         52
#define FO_LEVEL      53
#define FO_EXPLOC     54
#define FO_NSSITER    58
#define FO_SLOPE      64   /* 1 o1, 2 o3, 3 minmod                           */
#define FO_SHIFTVAR   66
#define FO_PT_BC      70
#define FO_NB_RELAX   71
#define FO_IMPLICITSOLVER 74
#define FO_LU_MATCH   75
#define FO_SENSORTYPE 85
#define FO_SA_DIST   132
#define FO_PARAM_INT_SIZE 137

/* BC record (param_solver.h:220-223) */
#define FO_BC_TYPE   0
#define FO_BC_IDIR   1
#define FO_BC_FEN    2
#define FO_BC_NBDATA 8

/* Parameter_real (param_solver.h:228-312) */
#define FO_DTC      0
#define FO_GAMMA    1
#define FO_CVINF    2
#define FO_ROINF    3
#define FO_PINF     4
#define FO_VINF     5
#define FO_TINF     6
#define FO_MINF     7
#define FO_RONUTILDEINF 8
#define FO_REYNOLDS 9
#define FO_PRANDT  10
#define FO_XMUL0   11
#define FO_TEMP0   12
#define FO_CS      13
#define FO_EPSI_NEWTON 14
#define FO_CFL     15
#define FO_SA_REAL 16
#define FO_RATIOM  18   /* SA_REAL + SA_RATIOM - 1 */
#define FO_VXINF   19
#define FO_VYINF   20
#define FO_VZINF   21
#define FO_TEMPS   22
#define FO_PSIROE  34
#define FO_PRANDT_TUR 38
#define FO_WIG_AMP 72
#define FO_PARAM_REAL_SIZE 73

/* SA constants (param_solver.h:315-327) */
#define FO_SA_CKARM 0.41
#define FO_SA_CB1   0.1355
#define FO_SA_CB2   0.622
#define FO_SA_CV1   357.911
#define FO_SA_CV2   0.7
#define FO_SA_CV3   0.9
#define FO_SA_CW2   0.3
#define FO_SA_CW3   64.
#define FO_SA_SIGMA (2./3.)

#endif
