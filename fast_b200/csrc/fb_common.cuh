// fb_common.cuh -- device-visible zone descriptor and address helpers (sm_100a, FP64).
//
// Data layout in HBM (per zone, all FP64, same as the reference so host buffers copy 1:1):
//   rop[level]  : NDIMDX*NEQ   variable-major SoA, primitive (rho,u,v,w,T[,nutilde])  (S/formule_param.h)
//   drodm       : NDIMDX*NEQ   residual / Newton increment
//   coe         : NDIMDX*NEQ_COE (implicit only; explicit keeps dt/vol in registers)
//   xmut        : NDIMDX*(1|2) [mu(+mut) | wall distance]
//   ti,tj,tk,vol: NDIMDX_MTR*(3|2|1) face normals x area, cell volume (node-sized strides, NIJK_MTR)
//   wigb        : NDIMDX bytes, bit d = SENSOR switch of the lower face in direction d (reference: 3 doubles)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Parameter_int / Parameter_real slots (reference wire format, S/param_solver.h:80-312)
enum {
    PI_NIJK = 0, PI_NIJK_MTR = 5, PI_NIJK_XYZ = 10, PI_IJKV = 20, PI_ITYPZONE = 25, PI_IFLOW = 27, PI_ILES = 28,
    PI_ITYPCP = 29, PI_KFLUDOM = 33, PI_LALE = 34, PI_NEQ = 36, PI_NEQ_IJ = 37, PI_NEQ_K = 38, PI_NEQ_COE = 39,
    PI_NDIMDX = 41, PI_NDIMDX_XYZ = 42, PI_NDIMDX_MTR = 43, PI_DTLOC = 46, PI_SA_INT = 47, PI_RK = 52,
    PI_EXPLOC = 54, PI_NSSITER = 58, PI_SLOPE = 64, PI_SHIFTVAR = 66, PI_PT_BC = 70, PI_NB_RELAX = 71,
    PI_IMPLICITSOLVER = 74, PI_LU_MATCH = 75, PI_SENSORTYPE = 85, PI_SIZE = 137
};
enum {
    PR_DTC = 0, PR_GAMMA = 1, PR_CVINF = 2, PR_ROINF = 3, PR_PINF = 4, PR_VINF = 5, PR_TINF = 6, PR_RONUTILDEINF = 8,
    PR_PRANDT = 10, PR_XMUL0 = 11, PR_TEMP0 = 12, PR_CS = 13, PR_CFL = 15, PR_RATIOM = 18, PR_VXINF = 19,
    PR_VYINF = 20, PR_VZINF = 21, PR_PSIROE = 34, PR_PRANDT_TUR = 38, PR_WIG_AMP = 72, PR_SIZE = 73
};
enum { BC_TYPE = 0, BC_IDIR = 1, BC_FEN = 2, BC_NBDATA = 8 };

#define SA_CKARM 0.41
#define SA_CB1 0.1355
#define SA_CB2 0.622
#define SA_CV1 357.911
#define SA_CV2 0.7
#define SA_CV3 0.9
#define SA_CW2 0.3
#define SA_CW3 64.
#define SA_SIGMA (2. / 3.)

// wall faces whose flux is corrected when the zone runs the Roe scheme (S/BC/correct_flux.for, lapply = 4)
#define FB_MAXCORR 8
struct CorrWin { int idir, slip; int w[6]; };   // face-cell range: the cell whose LOWER face is the wall face

struct ZoneDev {
    int ni, nj, nk, nf, nfk, iv, jv, kv, nd;
    int incj, inck;
    int mi, mj, mk, mf, mfk, nm;
    int xi, xj, xf, xfk, ndx;
    int typ, iflow, kflu, slope, neq, itypcp, neq_coe, rk, dtloc, sensortype;
    const double *x, *y, *z;
    double *ti, *tj, *tk, *vol;
    double *xmut, *drodm, *coe;
    unsigned char *wigb;
    double gam, rgp, gam1, gam3, gam4, cmus1, temp01, coesut, coesut_vis, sigma_1;
    double c1, c2, opt0, wig_cte, psiroe, tmin_1, cv, dtc, ratiom, cfl, roinf, xmul0, prandt;
    double tc0[3], vol0;
    double vinf[3], pinf, ronutildeinf;   // DEFAULT_STATE of the BCs (S/BC/BCzone.cpp:26-40)
    int ncorr; CorrWin corr[FB_MAXCORR];
};

// Loads of arrays that OTHER kernels rewrite between two launches of the reader (time levels, coe, drodm) go to L2 (ld.global.cg):
// with the per-zone kernels of a phase spread over several streams, L1-cached copies of such data were observed to survive
// from one sub-iteration into the next (profiles/r02_race.md) -- the hardware keeps the L1s of the SMs coherent only between
// dependent grids of ONE stream.  Staged data (TMA, cp.async) already bypasses L1.
#define FB_LD(p) __ldcg(p)

__device__ __forceinline__ int inddm(const ZoneDev &z, int i, int j, int k)
{
    return (i + z.nf - 1) + (j + z.nf - 1) * z.incj + (k + z.nfk - 1) * z.inck;
}
__device__ __forceinline__ int indmtr(const ZoneDev &z, int i, int j, int k)
{
    return (i + z.mf - 1) * z.mi + (j + z.mf - 1) * z.mj + (k + z.mfk - 1) * z.mk;
}
__device__ __forceinline__ int indcg(const ZoneDev &z, int i, int j, int k)
{
    return (i + z.xf - 1) + (j + z.xf - 1) * z.xi + (k + z.xfk - 1) * z.xi * z.xj;
}
