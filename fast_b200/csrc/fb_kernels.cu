// fb_kernels.cu -- metric precompute, viscosity, SENSOR switch, fused rhs kernels, ghost-cell kernels (sm_100a).
//
// Kernel map (reference function -> kernel):
//   cp_tijk_ + cp_vol_ (C/Metric/cp_tijk.for, cp_vol.for)                   -> k_tijk, k_vol
//   invist / vispalart (S/Compute/invist.for, SA/vispalart.for)              -> k_visc
//   wiggle2001 (S/Compute/SENSOR/wiggle2001.for)                              -> k_sensor   (3 bits / cell, not 3 doubles)
//   init_rhs + flu*_select + core3ark3 | spsource_SA + core3as2              -> k_rhs<...> (ONE fused kernel per sub-iteration)
//   BCzone/bvbs_* (S/BC/BCzone.cpp, bvbs_*.for)                               -> k_bc       (all windows of a zone in one launch)
//   correct_coins (S/BC/correct_coins.for)                                    -> k_corners
// All kernels are HBM-bound streaming / stencil kernels: i is the fastest thread index so that every warp reads and
// writes 256 B contiguous segments of each SoA variable; the (j,k) stencil neighbours hit L1/L2.
#include <cstdio>
#include <cstdlib>
#include "fb_kernels.cuh"
#include "fb_flux.cuh"
#include "fb_source.cuh"
#include "fb_rhs_decl.cuh"

// ------------------------------------------------------------------------------------------------ metrics
__global__ void k_tijk(ZoneDev z, int i0, int i1, int j0, int j1, int k0, int k1)
{
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = j0 + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = k0 + blockIdx.z;
    if (i > i1 || j > j1 || k > k1) return;
    const int typ = z.typ, nm = z.nm;
    const double *x = z.x, *y = z.y, *zz = z.z;
    const int inci = 1, incj = z.xi, inck = z.xi * z.xj;
    const int lt = typ == 2 ? 0 : indmtr(z, i, j, k);
    const int l0 = indcg(z, i, j, k);
    const int l1 = l0 + incj, l2 = l0 + inck, l3 = l1 + inck, l4 = l0 + inci, l5 = l1 + inci, l6 = l2 + inci;
    // upper-bound planes only carry the face that lies on them (complement loops of cp_tijk.for)
    const bool last_i = (i == i1 && typ != 2), last_j = (j == j1 && typ != 2), last_k = (k == k1 && typ == 0);
    const bool do_i = !last_j && !last_k, do_j = !last_i && !last_k, do_k = !last_i && !last_j;
    if (do_i) {
        const double tix10 = (y[l1] - y[l0]) * (zz[l3] - zz[l0]) - (zz[l1] - zz[l0]) * (y[l3] - y[l0]);
        const double tiy10 = (zz[l1] - zz[l0]) * (x[l3] - x[l0]) - (x[l1] - x[l0]) * (zz[l3] - zz[l0]);
        const double tiz10 = (x[l1] - x[l0]) * (y[l3] - y[l0]) - (y[l1] - y[l0]) * (x[l3] - x[l0]);
        const double tix11 = (y[l2] - y[l3]) * (zz[l0] - zz[l2]) - (zz[l2] - zz[l3]) * (y[l0] - y[l2]);
        const double tiy11 = (zz[l2] - zz[l3]) * (x[l0] - x[l2]) - (x[l2] - x[l3]) * (zz[l0] - zz[l2]);
        const double tiz11 = (x[l2] - x[l3]) * (y[l0] - y[l2]) - (y[l2] - y[l3]) * (x[l0] - x[l2]);
        z.ti[lt] = .5 * (tix10 + tix11);
        if (typ != 2) z.ti[lt + nm] = .5 * (tiy10 + tiy11);
        if (typ == 0) z.ti[lt + 2 * nm] = .5 * (tiz10 + tiz11);
    }
    if (do_j) {
        const double tjx10 = (y[l0] - y[l4]) * (zz[l2] - zz[l0]) - (zz[l0] - zz[l4]) * (y[l2] - y[l0]);
        const double tjy10 = (zz[l0] - zz[l4]) * (x[l2] - x[l0]) - (x[l0] - x[l4]) * (zz[l2] - zz[l0]);
        const double tjz10 = (x[l0] - x[l4]) * (y[l2] - y[l0]) - (y[l0] - y[l4]) * (x[l2] - x[l0]);
        const double tjx11 = (y[l6] - y[l2]) * (zz[l4] - zz[l6]) - (zz[l6] - zz[l2]) * (y[l4] - y[l6]);
        const double tjy11 = (zz[l6] - zz[l2]) * (x[l4] - x[l6]) - (x[l6] - x[l2]) * (zz[l4] - zz[l6]);
        const double tjz11 = (x[l6] - x[l2]) * (y[l4] - y[l6]) - (y[l6] - y[l2]) * (x[l4] - x[l6]);
        if (typ == 2) z.tj[lt] = .5 * (tjy10 + tjy11);
        else {
            z.tj[lt] = .5 * (tjx10 + tjx11);
            z.tj[lt + nm] = .5 * (tjy10 + tjy11);
            if (typ == 0) z.tj[lt + 2 * nm] = .5 * (tjz10 + tjz11);
        }
    }
    if (do_k && (typ != 1 || true)) {
        const double tkx10 = (y[l4] - y[l0]) * (zz[l5] - zz[l4]) - (zz[l4] - zz[l0]) * (y[l5] - y[l4]);
        const double tky10 = (zz[l4] - zz[l0]) * (x[l5] - x[l4]) - (x[l4] - x[l0]) * (zz[l5] - zz[l4]);
        const double tkz10 = (x[l4] - x[l0]) * (y[l5] - y[l4]) - (y[l4] - y[l0]) * (x[l5] - x[l4]);
        const double tkx11 = (y[l1] - y[l5]) * (zz[l0] - zz[l1]) - (zz[l1] - zz[l5]) * (y[l0] - y[l1]);
        const double tky11 = (zz[l1] - zz[l5]) * (x[l0] - x[l1]) - (x[l1] - x[l5]) * (zz[l0] - zz[l1]);
        const double tkz11 = (x[l1] - x[l5]) * (y[l0] - y[l1]) - (y[l1] - y[l5]) * (x[l0] - x[l1]);
        if (typ == 0) {
            z.tk[lt] = .5 * (tkx10 + tkx11);
            z.tk[lt + nm] = .5 * (tky10 + tky11);
            z.tk[lt + 2 * nm] = .5 * (tkz10 + tkz11);
        } else z.tk[lt] = .5 * (tkz10 + tkz11);
    }
}

__global__ void k_vol(ZoneDev z, int i0, int i1, int j0, int j1, int k0, int k1)
{
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = j0 + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = k0 + blockIdx.z;
    if (i > i1 || j > j1 || k > k1) return;
    const int typ = z.typ, nm = z.nm;
    const double *x = z.x, *y = z.y, *zz = z.z, *ti = z.ti, *tj = z.tj, *tk = z.tk;
    const int inci = 1, incj = z.xi, inck = z.xi * z.xj;
    const int lt = typ == 2 ? 0 : indmtr(z, i, j, k);
    const int l0 = indcg(z, i, j, k);
    const int l1 = l0 + incj, l2 = l0 + inck, l3 = l1 + inck, l4 = l0 + inci, l5 = l1 + inci, l6 = l2 + inci, l7 = l3 + inci;
    const int lip = lt + z.mi, ljp = lt + z.mj, lkp = lt + z.mk;
    double ax, ay, az;
    if (typ == 0) {
        ax = .25 * ((x[l4] + x[l5] + x[l7] + x[l6]) * ti[lip] - (x[l0] + x[l1] + x[l3] + x[l2]) * ti[lt]
                  + (x[l1] + x[l3] + x[l5] + x[l7]) * tj[ljp] - (x[l0] + x[l2] + x[l4] + x[l6]) * tj[lt]
                  + (x[l2] + x[l3] + x[l6] + x[l7]) * tk[lkp] - (x[l0] + x[l1] + x[l4] + x[l5]) * tk[lt]);
        ay = .25 * ((y[l4] + y[l5] + y[l7] + y[l6]) * ti[lip + nm] - (y[l0] + y[l1] + y[l3] + y[l2]) * ti[lt + nm]
                  + (y[l1] + y[l3] + y[l5] + y[l7]) * tj[ljp + nm] - (y[l0] + y[l2] + y[l4] + y[l6]) * tj[lt + nm]
                  + (y[l2] + y[l3] + y[l6] + y[l7]) * tk[lkp + nm] - (y[l0] + y[l1] + y[l4] + y[l5]) * tk[lt + nm]);
        az = .25 * ((zz[l4] + zz[l5] + zz[l7] + zz[l6]) * ti[lip + 2 * nm] - (zz[l0] + zz[l1] + zz[l3] + zz[l2]) * ti[lt + 2 * nm]
                  + (zz[l1] + zz[l3] + zz[l5] + zz[l7]) * tj[ljp + 2 * nm] - (zz[l0] + zz[l2] + zz[l4] + zz[l6]) * tj[lt + 2 * nm]
                  + (zz[l2] + zz[l3] + zz[l6] + zz[l7]) * tk[lkp + 2 * nm] - (zz[l0] + zz[l1] + zz[l4] + zz[l5]) * tk[lt + 2 * nm]);
    } else if (typ == 1) {
        ax = .25 * ((x[l4] + x[l5] + x[l7] + x[l6]) * ti[lip] - (x[l0] + x[l1] + x[l3] + x[l2]) * ti[lt]
                  + (x[l1] + x[l3] + x[l5] + x[l7]) * tj[ljp] - (x[l0] + x[l2] + x[l4] + x[l6]) * tj[lt]);
        ay = .25 * ((y[l4] + y[l5] + y[l7] + y[l6]) * ti[lip + nm] - (y[l0] + y[l1] + y[l3] + y[l2]) * ti[lt + nm]
                  + (y[l1] + y[l3] + y[l5] + y[l7]) * tj[ljp + nm] - (y[l0] + y[l2] + y[l4] + y[l6]) * tj[lt + nm]);
        az = .25 * ((zz[l2] + zz[l3] + zz[l6] + zz[l7]) * tk[lkp] - (zz[l0] + zz[l1] + zz[l4] + zz[l5]) * tk[lt]);
    } else {
        ax = .25 * ((x[l4] + x[l5] + x[l7] + x[l6]) * ti[lip] - (x[l0] + x[l1] + x[l3] + x[l2]) * ti[lt]);
        ay = .25 * ((y[l1] + y[l3] + y[l5] + y[l7]) * tj[ljp] - (y[l0] + y[l2] + y[l4] + y[l6]) * tj[lt]);
        az = .25 * ((zz[l2] + zz[l3] + zz[l6] + zz[l7]) * tk[lkp] - (zz[l0] + zz[l1] + zz[l4] + zz[l5]) * tk[lt]);
    }
    z.vol[lt] = fmax((ax + ay + az) / 3., 1.e-30);
}

static inline dim3 grid3(int ni, int nj, int nk, dim3 b) { return dim3((ni + b.x - 1) / b.x, (nj + b.y - 1) / b.y, nk); }

void launch_metric(const ZoneDev &z, cudaStream_t st)
{
    // every cell incl. ghosts (C/Metric/indice_omp1.h) + the imax+1 / jmax+1 / kmax+1 face planes
    int i0 = 1 - z.nf, i1 = z.iv + z.nf, j0 = 1 - z.nf, j1 = z.jv + z.nf, k0 = 1 - z.nfk, k1 = z.kv + z.nfk;
    if (z.typ == 1) { k0 = 1; k1 = 1; }
    if (z.typ == 2) { i0 = i1 = j0 = j1 = k0 = k1 = 1; }
    dim3 b(32, 4, 1);
    {
        int e1 = i1, f1 = j1, g1 = k1;
        if (z.typ != 2) { e1 += 1; f1 += 1; }
        if (z.typ == 0) g1 += 1;
        k_tijk<<<grid3(e1 - i0 + 1, f1 - j0 + 1, g1 - k0 + 1, b), b, 0, st>>>(z, i0, e1, j0, f1, k0, g1);
    }
    k_vol<<<grid3(i1 - i0 + 1, j1 - j0 + 1, k1 - k0 + 1, b), b, 0, st>>>(z, i0, i1, j0, j1, k0, k1);
}

// ------------------------------------------------------------------------------------------------ viscosity
// S/Compute/invist.for (laminar) / SA/vispalart.for (mu + mut) on interior +- ext cells
template <int FLOW>
__global__ void k_visc(ZoneDev z, const double *__restrict__ rop, int ext, int extk)
{
    const int i = 1 - ext + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 - ext + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = 1 - extk + blockIdx.z;
    if (i > z.iv + ext || j > z.jv + ext) return;
    const int l = inddm(z, i, j, k);
    const double t1 = rop[l + 4 * z.nd];
    if (FLOW == 2) z.xmut[l] = z.coesut_vis * sqrt(t1) * t1 / (t1 + z.cmus1);
    else z.xmut[l] = vispalart_cell(z, rop, l);
}

void launch_invist(const ZoneDev &z, const double *rop, int ext, cudaStream_t st)
{
    const int extk = z.nfk == 0 ? 0 : ext;
    dim3 b(64, 4, 1);
    dim3 g = grid3(z.iv + 2 * ext, z.jv + 2 * ext, z.kv + 2 * extk, b);
    if (z.iflow == 2) k_visc<2><<<g, b, 0, st>>>(z, rop, ext, extk);
    else if (z.iflow == 3) k_visc<3><<<g, b, 0, st>>>(z, rop, ext, extk);
}

// ------------------------------------------------------------------------------------------------ SENSOR switch
// S/Compute/SENSOR/wiggle2001.for + wiggle.for: bit d of wigb[l] = 1 when some primitive variable has two
// consecutive sign changes of its differences across the lower face of l in direction d.
__device__ __forceinline__ int wiggle_face(const double *__restrict__ rop, int nd, int l, int inc)
{
    double test = 1e30;
#pragma unroll
    for (int e = 0; e < 5; e++) {
        const double *q = rop + e * nd;
        const double a = q[l - 2 * inc], b = q[l - inc], c = q[l], d = q[l + inc];
        const double qm = (b - a) * (c - b), qp = (c - b) * (d - c);
        test = fmin(test, copysign(1., qm) + copysign(1., qp));
    }
    return test < -1e-12 ? 1 : 0;
}

__global__ void k_sensor(ZoneDev z, const double *__restrict__ rop)
{
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = 1 + blockIdx.z;
    if (i > z.iv + 1 || j > z.jv + 1) return;
    const int l = inddm(z, i, j, k);
    int b = 0;
    // faces of the complement planes exist only where the two other indices are interior
    if (j <= z.jv && k <= z.kv) b |= wiggle_face(rop, z.nd, l, 1);
    if (i <= z.iv && k <= z.kv) b |= wiggle_face(rop, z.nd, l, z.incj) << 1;
    if (i <= z.iv && j <= z.jv) b |= wiggle_face(rop, z.nd, l, z.inck) << 2;
    z.wigb[l] = (unsigned char)b;
}

bool fb_sensor_fused()
{   // default: stand-alone k_sensor at nstep 1.  FAST_B200_SENSOR=f evaluates wiggle2001 inside the first-sub-iteration rhs
    // kernel from the MUSCL stencils it already holds (parity-green, but measured slower on TGV 256^3: +0.55 ms on that
    // launch against 0.36 ms for k_sensor -- kept for the next profiling pass)
    static const bool f = [] { const char *e = getenv("FAST_B200_SENSOR"); return e && e[0] == 'f'; }();
    return f;
}

// k-marching version: every primitive-variable difference is formed ONCE.  The switch only needs signs:
//   qm = d(l-1) d(l),  qp = d(l) d(l+1),  d(l) = q(l) - q(l-1);   sign(1,qm) + sign(1,qp) = -2  <=>  both products carry
// the sign bit, and the sign bit of an IEEE product is the XOR of the factors' sign bits (zeros included).  A thread owns
// the column (i,j), keeps plane p-1 in registers, packs the 15 sign bits of its cell (5 variables x 3 directions) in one
// word, gets the i neighbours' words by warp shuffle, the j neighbours' through shared memory and the k neighbours' from its
// own history (the byte of plane p-1 is completed and stored when plane p has been visited).  15 loads per cell instead
// of 60; bit-identical to k_sensor.
#define SEN_BJ 8
#define SEN_KC 32
__global__ void __launch_bounds__(32 * SEN_BJ) k_sensor2(ZoneDev z, const double *__restrict__ rop)
{
    __shared__ unsigned sS[2][SEN_BJ][32];
    const int li = threadIdx.x, m = threadIdx.y;
    const int i = blockIdx.x * 30 + li, j = blockIdx.y * (SEN_BJ - 2) + m;       // columns i0-1 .. i0+30, rows j0-1 .. j0+6
    const int k0 = 1 + blockIdx.z * SEN_KC, k1 = min(k0 + SEN_KC - 1, z.kv + 1);
    const int nd = z.nd;
    // columns whose differences exist: i in [0, iv+2], j in [0, jv+2] (two ghost layers); others carry zeros
    const bool col = i <= z.iv + 2 && j <= z.jv + 2;
    const bool out = li >= 1 && li <= 30 && m >= 1 && m <= SEN_BJ - 2 && i <= z.iv + 1 && j <= z.jv + 1;
    double qp[5];                                  // plane p-1 of the own column
    const int lc = inddm(z, col ? i : 1, col ? j : 1, k0 - 2);
#pragma unroll
    for (int e = 0; e < 5; e++) qp[e] = rop[lc + e * nd];
    unsigned sk1 = 0, sk2 = 0, ij1 = 0;            // k sign bits of planes p-1, p-2; i/j face bits of plane p-1
    for (int p = k0 - 1; p <= k1 + 1; p++) {
        const int l = lc + (p - (k0 - 2)) * z.inck;
        unsigned S = 0;
        if (col) {
#pragma unroll
            for (int e = 0; e < 5; e++) {
                const double q = rop[l + e * nd];
                const unsigned bi_ = (unsigned)__double2hiint(q - rop[l - 1 + e * nd]) >> 31;
                const unsigned bj_ = (unsigned)__double2hiint(q - rop[l - z.incj + e * nd]) >> 31;
                const unsigned bk_ = (unsigned)__double2hiint(q - qp[e]) >> 31;
                S |= (bi_ << e) | (bj_ << (5 + e)) | (bk_ << (10 + e));
                qp[e] = q;
            }
        }
        sS[p & 1][m][li] = S;
        const unsigned Sl = __shfl_up_sync(0xffffffffu, S, 1), Sr = __shfl_down_sync(0xffffffffu, S, 1);
        __syncthreads();
        if (out) {
            const unsigned Sd = sS[p & 1][m - 1][li], Su = sS[p & 1][m + 1][li];
            const unsigned wi = (Sl ^ S) & (S ^ Sr) & 0x1fu;
            const unsigned wj = ((Sd ^ S) & (S ^ Su) >> 0) & (0x1fu << 5);
            const unsigned skc = S & (0x1fu << 10);
            // the byte of plane p-1: its i / j bits were formed one iteration ago, its k bit needs sk(p-2), sk(p-1), sk(p)
            if (p - 1 >= k0 && p - 1 <= k1) {
                const unsigned wk = (sk2 ^ sk1) & (sk1 ^ skc);
                const int kk = p - 1;
                unsigned b = ij1;
                if (!(j <= z.jv && kk <= z.kv)) b &= ~1u;          // faces of the complement planes exist only where the
                if (!(i <= z.iv && kk <= z.kv)) b &= ~2u;          // two other indices are interior (k_sensor)
                if (wk && i <= z.iv && j <= z.jv) b |= 4u;
                z.wigb[l - z.inck] = (unsigned char)b;
            }
            ij1 = (wi ? 1u : 0u) | (wj ? 2u : 0u);
            sk2 = sk1; sk1 = skc;
        }
    }
}

void launch_sensor(const ZoneDev &z, const double *rop, cudaStream_t st)
{
    static const bool v1 = [] { const char *e = getenv("FAST_B200_SENSOR"); return e && e[0] == '1'; }();   // A/B: first version
    if (v1 || z.nf < 2 || z.nfk < 2) {
        dim3 b(64, 4, 1);
        k_sensor<<<grid3(z.iv + 1, z.jv + 1, z.kv + 1, b), b, 0, st>>>(z, rop);
        return;
    }
    dim3 b(32, SEN_BJ, 1);
    dim3 g((z.iv + 1 + 29) / 30, (z.jv + 1 + SEN_BJ - 3) / (SEN_BJ - 2), (z.kv + 1 + SEN_KC - 1) / SEN_KC);
    k_sensor2<<<g, b, 0, st>>>(z, rop);
}

template <int MESH>
static void launch_rhs_f(const ZoneDev &z, RhsArgs &a, bool impl, cudaStream_t st)
{
    switch (z.iflow) {
    case 1: launch_rhs_k<MESH, 1>(z, a, impl, st); break;
    case 2: launch_rhs_k<MESH, 2>(z, a, impl, st); break;
    default: launch_rhs_k<MESH, 3>(z, a, impl, st); break;
    }
}

void launch_rhs(const ZoneDev &z, const double *rop, double *rop_out, const double *ro_n, const double *ro_n1,
                int nitcfg, int first_it, cudaStream_t st, int part, int faces, int pdl)
{
    RhsArgs a; a.part = z.itypcp <= 1 ? 0 : part; a.tile0 = a.gx = a.gy = a.gz = 0; a.faces = faces; a.pdl = pdl;
    a.rop = rop; a.rop_out = rop_out; a.ro_n = ro_n; a.ro_n1 = ro_n1; a.nitcfg = nitcfg; a.wfuse = fb_sensor_fused() ? 1 : 0;
    const bool impl = z.itypcp <= 1;
    a.coefH = 0.; a.coefW = 1.;
    if (z.rk == 3) {
        a.coefH = nitcfg == 2 ? -0.6830127018922193 : (nitcfg == 3 ? -4. / 3. : 0.);
        a.coefW = nitcfg == 1 ? 0.5 : (nitcfg == 2 ? 0.9106836025229591 : 0.3660254037844387);
    } else if (z.rk == 2) { a.coefH = nitcfg == 2 ? -1. : 0.; a.coefW = nitcfg == 1 ? 1. : 0.5; }
    // core3as2.for:37-46
    if (first_it == 0 || z.dtloc == 1) { a.c1 = -1.; a.c2 = 1.; a.c3 = 0.; }
    else { a.c1 = -1.5; a.c2 = 2.; a.c3 = -0.5; }
    switch (z.typ) {
    case 0: launch_rhs_f<0>(z, a, impl, st); break;
    case 1: launch_rhs_f<1>(z, a, impl, st); break;
    default: launch_rhs_f<2>(z, a, impl, st); break;
    }
}

#include "fb_bc.cuh"
