// fb_implicit.cu -- implicit chain: LU spectral radii / diagonal, SA diagonal on the ghost ring, residual norms,
// LU-SGS sweeps on i+j+k hyperplane wavefronts, Newton update.
//
// Reference -> kernel:
//   tstb3_global + tstb3c | tstb3_impli + dtloc (S/Compute/tstb3c.for:77-300, tstb3_impli.for, dtloc.for)   -> k_coe
//   sourceSA_LU on interior +-1 (S/Compute/SA/spsource_SA.for, ind_ssa)                                     -> k_sa_ring
//   cprdu3s1 (S/Compute/cprdu3s1.for, cprdu3s1_incl.for)                                                    -> k_resid
//   invlu_l[_SA] / invlu_u[_SA] (S/Compute/LU/invlu_l.for:79-215, invlu_u.for:60-215, lu_*.for, mjr_*.for)  -> k_lu<...>
//   mjr_newton[_SA] (S/Compute/LU/mjr_newton.for)                                                           -> k_newton
// The reference sweeps i fastest, then j, then k; a cell only depends on its three lower (upper) neighbours, so all
// cells of a hyperplane i+j+k = const are independent and the wavefront reproduces exactly the sequential result.
#include <cstdio>
#include <cstdlib>
#include "fb_implicit.cuh"
#include "fb_lu_prod.cuh"
#include "fb_source.cuh"

static inline dim3 grid3i(int ni, int nj, int nk, dim3 b) { return dim3((ni + b.x - 1) / b.x, (nj + b.y - 1) / b.y, nk); }

// ------------------------------------------------------------------------------------------------ coe
template <int MESH>
__global__ void k_coe(ZoneDev z, const double *__restrict__ rop, int first_it)
{
    const int i = 1 - z.nf + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 - z.nf + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = 1 - z.nfk + blockIdx.z;
    if (i > z.iv + z.nf || j > z.jv + z.nf) return;
    const int nd = z.nd, nm = z.nm;
    const int l = inddm(z, i, j, k), lt = MESH == 2 ? 0 : indmtr(z, i, j, k);
    const bool local = z.dtloc == 1;
    const double cut0x = local ? 1e-30 : 1e-20;
    const double prandt = z.prandt;
    double xid = 1.5; if (first_it == 0) xid = 1.;
    if (local) xid = 1.;
    double gam1 = z.gam / prandt;
    if (z.iflow == 1) gam1 = local ? cut0x : 0.;
    const double gam2 = z.gam * z.rgp;
    const double r = rop[l];
    const double vol = MESH == 2 ? z.vol0 : __ldg(z.vol + lt);
    const double detj = fmax(vol, cut0x);
    const double u = rop[l + nd], v = rop[l + 2 * nd], w = rop[l + 3 * nd];
    const double c = sqrt(gam2 * rop[l + 4 * nd]);
    const double xm = z.iflow == 1 ? 0. : z.xmut[l];
    const double vmu = gam1 * xm, volu = detj * r, xmvis = vmu / volu;
    double tc2i, tc2j, tc2k, qni, qnj, qnk;
    if (MESH == 0) {
        const int li = lt + z.mi, lj = lt + z.mj, lk = lt + z.mk;
        const double *ti = z.ti, *tj = z.tj, *tk = z.tk;
        const double a0 = ti[lt], a1 = ti[lt + nm], a2 = ti[lt + 2 * nm], b0 = ti[li], b1 = ti[li + nm], b2 = ti[li + 2 * nm];
        tc2i = .5 * (sqrt(a0 * a0 + a1 * a1 + a2 * a2) + sqrt(b0 * b0 + b1 * b1 + b2 * b2));
        qni = (.5 * (a0 + b0)) * u + (.5 * (a1 + b1)) * v + (.5 * (a2 + b2)) * w;
        const double c0 = tj[lt], c1 = tj[lt + nm], c2 = tj[lt + 2 * nm], d0 = tj[lj], d1 = tj[lj + nm], d2 = tj[lj + 2 * nm];
        tc2j = .5 * (sqrt(c0 * c0 + c1 * c1 + c2 * c2) + sqrt(d0 * d0 + d1 * d1 + d2 * d2));
        qnj = (.5 * (c0 + d0)) * u + (.5 * (c1 + d1)) * v + (.5 * (c2 + d2)) * w;
        const double e0 = tk[lt], e1 = tk[lt + nm], e2 = tk[lt + 2 * nm], f0 = tk[lk], f1 = tk[lk + nm], f2 = tk[lk + 2 * nm];
        tc2k = .5 * (sqrt(e0 * e0 + e1 * e1 + e2 * e2) + sqrt(f0 * f0 + f1 * f1 + f2 * f2));
        qnk = (.5 * (e0 + f0)) * u + (.5 * (e1 + f1)) * v + (.5 * (e2 + f2)) * w;
    } else if (MESH == 1) {
        const int li = lt + z.mi, lj = lt + z.mj;
        const double *ti = z.ti, *tj = z.tj;
        const double a0 = ti[lt], a1 = ti[lt + nm], b0 = ti[li], b1 = ti[li + nm];
        tc2i = .5 * (sqrt(a0 * a0 + a1 * a1) + sqrt(b0 * b0 + b1 * b1));
        qni = (.5 * (a0 + b0)) * u + (.5 * (a1 + b1)) * v;
        const double c0 = tj[lt], c1 = tj[lt + nm], d0 = tj[lj], d1 = tj[lj + nm];
        tc2j = .5 * (sqrt(c0 * c0 + c1 * c1) + sqrt(d0 * d0 + d1 * d1));
        qnj = (.5 * (c0 + d0)) * u + (.5 * (c1 + d1)) * v;
        const double tz = z.tk[lt];
        tc2k = sqrt(tz * tz); qnk = tz * w;
    } else {
        tc2i = sqrt(z.tc0[0] * z.tc0[0]); tc2j = sqrt(z.tc0[1] * z.tc0[1]); tc2k = sqrt(z.tc0[2] * z.tc0[2]);
        qni = z.tc0[0] * u; qnj = z.tc0[1] * v; qnk = z.tc0[2] * w;
    }
    const double vmi = fabs(qni) + c * tc2i, vmj = fabs(qnj) + c * tc2j, vmk = fabs(qnk) + c * tc2k;
    const double c2_ = vmi * 2. + xmvis * tc2i * tc2i * 4.;
    const double c3_ = vmj * 2. + xmvis * tc2j * tc2j * 4.;
    const double c4_ = vmk * 2. + xmvis * tc2k * tc2k * 4.;
    double coe1;
    if (local) {
        const double d = 1. / (tc2i * tc2i + tc2j * tc2j + tc2k * tc2k);
        const double gam1_1 = 1. / gam1;
        const double dtvis = 0.5 * detj * d * gam1_1 / xm * r;
        const double sp = sqrt(u * u + v * v + w * w) + c;
        const double dtconv = sqrt(d) / sp;
        coe1 = z.cfl * fmin(dtconv, dtvis);
    } else coe1 = z.dtc / vol;   // tstb3_global.for (no floor on vol there)
    z.coe[l] = coe1; z.coe[l + nd] = c2_; z.coe[l + 2 * nd] = c3_; z.coe[l + 3 * nd] = c4_;
    z.coe[l + 4 * nd] = xid + coe1 * (c2_ + c3_ + c4_) * 0.5;
}

void launch_coe(const ZoneDev &z, const double *rop, int first_it, cudaStream_t st)
{
    dim3 b(64, 4, 1);
    dim3 g = grid3i(z.iv + 2 * z.nf, z.jv + 2 * z.nf, z.kv + 2 * z.nfk, b);
    if (z.typ == 0) k_coe<0><<<g, b, 0, st>>>(z, rop, first_it);
    else if (z.typ == 1) k_coe<1><<<g, b, 0, st>>>(z, rop, first_it);
    else k_coe<2><<<g, b, 0, st>>>(z, rop, first_it);
}

// ------------------------------------------------------------------------------------------------ CFL statistics
// S/Compute/cpcfl3.for (fixed meshes: viscous :276-440, Euler :635-790), called from cptst3.for:113-120 at the first
// sub-iteration of a check iteration (lssiter_verif), explicit and implicit: per interior cell
//   cfloc = dt / min(vis, conv),  conv = 2 vol surf / (|u| + c),  vis = 2 rho vol^2 surf^2 (gamma/Pr) / mu,
//   surf = 1 / max_d(|t_d(l)| + |t_d(l+d)|)
// reduced to [max, min, sum] over the zone (computePT.cpp:519-522 initial values, :619-645 gather) with warp shuffles, one
// atomic per block and statistic.  cflv[3] counts cells whose state is not finite (the device side of stopAtNan).
__device__ __forceinline__ void atomic_max_pos(double *addr, double v)
{   // v >= 0: IEEE ordering of non-negative doubles == ordering of their bit patterns
    atomicMax((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ void atomic_min_pos(double *addr, double v)
{
    atomicMin((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
}
__global__ void k_cfl_init(double *cflv) { if (threadIdx.x == 0) { cflv[0] = 0.; cflv[1] = 100000000.; cflv[2] = 0.; cflv[3] = 0.; } }

template <int MESH>
__global__ void k_cfl(ZoneDev z, const double *__restrict__ rop, double *cflv)
{
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = 1 + blockIdx.z;
    const bool ok = i <= z.iv && j <= z.jv;
    double cmax = 0., cmin = 100000000., csum = 0., bad = 0.;
    if (ok) {
        const int nd = z.nd, nm = z.nm;
        const int l = inddm(z, i, j, k), lt = MESH == 2 ? 0 : indmtr(z, i, j, k);
        double surf, vol;
        if (MESH == 0) {
            const int li = lt + z.mi, lj = lt + z.mj, lk = lt + z.mk;
            const double *ti = z.ti, *tj = z.tj, *tk = z.tk;
            double ddi = sqrt(ti[li] * ti[li] + ti[li + nm] * ti[li + nm] + ti[li + 2 * nm] * ti[li + 2 * nm]);
            ddi = ddi + sqrt(ti[lt] * ti[lt] + ti[lt + nm] * ti[lt + nm] + ti[lt + 2 * nm] * ti[lt + 2 * nm]);
            double ddj = sqrt(tj[lj] * tj[lj] + tj[lj + nm] * tj[lj + nm] + tj[lj + 2 * nm] * tj[lj + 2 * nm]);
            ddj = ddj + sqrt(tj[lt] * tj[lt] + tj[lt + nm] * tj[lt + nm] + tj[lt + 2 * nm] * tj[lt + 2 * nm]);
            double ddk = sqrt(tk[lk] * tk[lk] + tk[lk + nm] * tk[lk + nm] + tk[lk + 2 * nm] * tk[lk + 2 * nm]);
            ddk = ddk + sqrt(tk[lt] * tk[lt] + tk[lt + nm] * tk[lt + nm] + tk[lt + 2 * nm] * tk[lt + 2 * nm]);
            surf = 1. / fmax(fmax(ddi, ddj), ddk); vol = z.vol[lt];
        } else if (MESH == 1) {
            const int li = lt + z.mi, lj = lt + z.mj, lk = lt + z.mk;
            const double *ti = z.ti, *tj = z.tj, *tk = z.tk;
            double ddi = sqrt(ti[li] * ti[li] + ti[li + nm] * ti[li + nm]);
            ddi = ddi + sqrt(ti[lt] * ti[lt] + ti[lt + nm] * ti[lt + nm]);
            double ddj = sqrt(tj[lj] * tj[lj] + tj[lj + nm] * tj[lj + nm]);
            ddj = ddj + sqrt(tj[lt] * tj[lt] + tj[lt + nm] * tj[lt + nm]);
            double ddk = sqrt(tk[lk] * tk[lk]);
            ddk = ddk + sqrt(tk[lt] * tk[lt]);
            surf = 1. / fmax(fmax(ddi, ddj), ddk); vol = z.vol[lt];
        } else {
            const double ddi = 2. * sqrt(z.tc0[0] * z.tc0[0]), ddj = 2. * sqrt(z.tc0[1] * z.tc0[1]), ddk = 2. * sqrt(z.tc0[2] * z.tc0[2]);
            surf = 1. / fmax(fmax(ddi, ddj), ddk); vol = z.vol0;
        }
        const double r = rop[l], u = rop[l + nd], v = rop[l + 2 * nd], w = rop[l + 3 * nd], t = rop[l + 4 * nd];
        const double c = sqrt(z.gam * z.rgp * t);
        const double sp = sqrt(u * u + v * v + w * w);
        const double xinvspc = 1. / (sp + c);
        const double conv = 2. * vol * surf * xinvspc;
        double comp = conv;
        if (z.iflow >= 2) {
            const double xinvkt = (z.gam / z.prandt) / z.xmut[l];
            const double vis = 2. * r * vol * vol * surf * surf * xinvkt;
            comp = fmin(vis, conv);
        }
        const double cfloc = z.dtc / comp;
        cmax = cfloc; cmin = cfloc; csum = cfloc;
        double fin = r + u + v + w + t;
        if (z.neq == 6) fin += rop[l + 5 * nd];
        if (!isfinite(fin) || !isfinite(cfloc)) { bad = 1.; cmax = 0.; cmin = 100000000.; csum = 0.; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o)); cmin = fmin(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
        csum += __shfl_xor_sync(0xffffffffu, csum, o); bad += __shfl_xor_sync(0xffffffffu, bad, o);
    }
    __shared__ double sh[8][4];
    const int t = threadIdx.y * blockDim.x + threadIdx.x, lane = t & 31, warp = t >> 5;
    if (lane == 0) { sh[warp][0] = cmax; sh[warp][1] = cmin; sh[warp][2] = csum; sh[warp][3] = bad; }
    __syncthreads();
    if (t == 0) {
        const int nw = (blockDim.x * blockDim.y + 31) >> 5;
        for (int q = 1; q < nw; q++) { cmax = fmax(cmax, sh[q][0]); cmin = fmin(cmin, sh[q][1]); csum += sh[q][2]; bad += sh[q][3]; }
        atomic_max_pos(cflv, cmax); atomic_min_pos(cflv + 1, cmin); atomicAdd(cflv + 2, csum);
        if (bad > 0.) atomicAdd(cflv + 3, bad);
    }
}

void launch_cfl(const ZoneDev &z, const double *rop, double *cflv, cudaStream_t st)
{
    k_cfl_init<<<1, 32, 0, st>>>(cflv);
    dim3 b(64, 4, 1);
    dim3 g = grid3i(z.iv, z.jv, z.kv, b);
    if (z.typ == 0) k_cfl<0><<<g, b, 0, st>>>(z, rop, cflv);
    else if (z.typ == 1) k_cfl<1><<<g, b, 0, st>>>(z, rop, cflv);
    else k_cfl<2><<<g, b, 0, st>>>(z, rop, cflv);
}

// ------------------------------------------------------------------------------------------------ SA ghost ring
// spsource_SA loops on ind_ssa = interior +- 1: the LU diagonal coe(:,6) is also needed on the first ghost row (the
// sweep box is extended across wall faces).  Interior cells are handled inside k_rhs; this kernel does the ring.
template <int MESH>
__global__ void k_sa_ring(ZoneDev z, const double *__restrict__ rop)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // 0..iv+1
    const int j = blockIdx.y * blockDim.y + threadIdx.y;          // 0..jv+1
    const int k = blockIdx.z;                                     // 0..kv+1
    if (i > z.iv + 1 || j > z.jv + 1) return;
    const bool inner = i >= 1 && i <= z.iv && j >= 1 && j <= z.jv && k >= 1 && k <= z.kv;
    if (inner) return;
    const int l = inddm(z, i, j, k), lt = MESH == 2 ? 0 : indmtr(z, i, j, k);
    sa_source_cell<MESH>(z, GView(z, rop), GMet1(z), l, l, lt, true, z.coe[l]);
}

void launch_sa_ring(const ZoneDev &z, const double *rop, cudaStream_t st)
{
    dim3 b(64, 4, 1);
    dim3 g = grid3i(z.iv + 2, z.jv + 2, z.kv + 2, b);
    if (z.typ == 0) k_sa_ring<0><<<g, b, 0, st>>>(z, rop);
    else if (z.typ == 1) k_sa_ring<1><<<g, b, 0, st>>>(z, rop);
    else k_sa_ring<2><<<g, b, 0, st>>>(z, rop);
}

// ------------------------------------------------------------------------------------------------ residual norms
__global__ void k_resid_zero(double *rdm, int slot, int nss, int neq)
{
    const int e = threadIdx.x;
    if (e < neq) { rdm[slot + e * nss] = 0.; rdm[slot + e * nss + neq * nss] = 0.; }
}

__global__ void k_resid(ZoneDev z, double *rdm, int slot, int nss, double *cflv)
{
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = 1 + blockIdx.z;
    const bool ok = i <= z.iv && j <= z.jv;
    const int l = ok ? inddm(z, i, j, k) : 0;
    __shared__ double ss[8][6], sm[8][6];
    const int lane = (threadIdx.y * blockDim.x + threadIdx.x) & 31, warp = (threadIdx.y * blockDim.x + threadIdx.x) >> 5;
    for (int e = 0; e < z.neq; e++) {
        const double x = ok ? z.drodm[l + e * z.nd] : 0.;
        double s = x * x, m = fabs(x);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o)); }
        if (lane == 0) { ss[warp][e] = s; sm[warp][e] = m; }
    }
    __syncthreads();
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    if (t < z.neq) {
        double s = 0., m = 0.;
        for (int q = 0; q < nw; q++) { s += ss[q][t]; m = fmax(m, sm[q][t]); }
        atomicAdd(&rdm[slot + t * nss], s);
        atomic_max_pos(&rdm[slot + t * nss + z.neq * nss], m);
        if (!isfinite(s)) atomicAdd(cflv + 3, 1.);      // device side of stopAtNan (S/PyTree.py:2070-2073)
    }
}

// ------------------------------------------------------------------------------------------------ LU-SGS
struct LuArgs {
    const double *rop;      // rop_ssiter
    int box[6];
    int plane;              // hyperplane index, 0-based from the sweep's starting corner
    int bi, bj, bk;
    int alias;              // upper sweep, rop_ssiter == rop_out: k neighbour already Newton-updated (invlu_u.for order)
    int order;              // 0: k,j,i (invlu_l/u)  1: i,j,k (overlap variants, lu_match)  2: 3dcart lower j,k,i
};

template <int MESH, int NEQ, bool UPPER>
__global__ void k_lu(ZoneDev z, LuArgs a)
{
    // one thread per (jj,kk) of the box; ii follows from the hyperplane index
    const int jj = blockIdx.x * blockDim.x + threadIdx.x, kk = blockIdx.y * blockDim.y + threadIdx.y;
    if (jj >= a.bj || kk >= a.bk) return;
    const int ii = a.plane - jj - kk;
    if (ii < 0 || ii >= a.bi) return;
    const int i = UPPER ? a.box[1] - ii : a.box[0] + ii;
    const int j = UPPER ? a.box[3] - jj : a.box[2] + jj;
    const int k = UPPER ? a.box[5] - kk : a.box[4] + kk;
    const int nd = z.nd;
    const int l = inddm(z, i, j, k), lt = MESH == 2 ? 0 : indmtr(z, i, j, k);
    double acc[NEQ], b[NEQ];
#pragma unroll
    for (int e = 0; e < NEQ; e++) acc[e] = z.drodm[l + e * nd];
    const double coe5 = z.coe[l + 4 * nd];
    double coe6 = 1.;
    if (NEQ == 6) coe6 = z.coe[l + 5 * nd];
    if (UPPER) {
#pragma unroll
        for (int e = 0; e < 5; e++) acc[e] = acc[e] * coe5;
        if (NEQ == 6) acc[NEQ - 1] = acc[NEQ - 1] * coe6;
    }
    const double signe = UPPER ? -0.5 : 0.5;
    const double xal = z.coe[l] * signe;
    const int sgn = UPPER ? 1 : -1;
    const int inc[3] = {1, z.incj, z.inck}, incm[3] = {z.mi, z.mj, z.mk};
    const bool has[3] = {ii > 0, jj > 0, kk > 0};
    int ord[3];
    if (a.order == 1) { ord[0] = 0; ord[1] = 1; ord[2] = 2; }
    else if (a.order == 2 && has[1] && has[2]) { ord[0] = 1; ord[1] = 2; ord[2] = 0; }
    else { ord[0] = 2; ord[1] = 1; ord[2] = 0; }
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const int d = ord[q];
        if (!has[d]) continue;
        const int l1 = l + sgn * inc[d];
        int lta, ltb;
        if (UPPER) { lta = lt + incm[d]; ltb = (MESH == 1 && d == 2) ? lt : lt + 2 * incm[d]; }
        else { lta = lt - incm[d]; ltb = lt; }
        lu_term<MESH, NEQ>(z, a.rop, d, l1, lta, ltb, signe, UPPER && a.alias && d == 2, b);
#pragma unroll
        for (int e = 0; e < NEQ; e++) acc[e] = acc[e] + b[e] * xal;
    }
    const double diag = 1. / coe5;
#pragma unroll
    for (int e = 0; e < 5; e++) z.drodm[l + e * nd] = acc[e] * diag;
    if (NEQ == 6) z.drodm[l + 5 * nd] = acc[NEQ - 1] / coe6;
}

// LU/mjr_newton.for / mjr_newton_SA.for over the Newton box
template <int NEQ>
__global__ void k_newton(ZoneDev z, const double *__restrict__ rop, double *__restrict__ out, int i0, int i1, int j0, int j1, int k0)
{
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = j0 + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = k0 + blockIdx.z;
    if (i > i1 || j > j1) return;
    const int nd = z.nd, l = inddm(z, i, j, k);
    const double *dw = z.drodm;
    const double cvinv = 1. / z.cv, cvinv2 = 0.5 * cvinv;
    const double ro_old = rop[l], u_old = rop[l + nd], v_old = rop[l + 2 * nd], w_old = rop[l + 3 * nd], t_old = rop[l + 4 * nd];
    double nu_old = 0.; if (NEQ == 6) nu_old = rop[l + 5 * nd];
    const double rn = ro_old + dw[l];
    const double r_1 = 1. / rn;
    const double un = (ro_old * u_old + dw[l + nd]) * r_1, vn = (ro_old * v_old + dw[l + 2 * nd]) * r_1, wn = (ro_old * w_old + dw[l + 3 * nd]) * r_1;
    const double roe_old = ro_old * (z.cv * t_old + 0.5 * (u_old * u_old + v_old * v_old + w_old * w_old));
    const double tn = (roe_old + dw[l + 4 * nd]) * r_1 * cvinv - cvinv2 * (un * un + vn * vn + wn * wn);
    out[l] = rn; out[l + nd] = un; out[l + 2 * nd] = vn; out[l + 3 * nd] = wn; out[l + 4 * nd] = tn;
    if (NEQ == 6) {
        double anulam = z.coesut * sqrt(tn * z.temp01) / (1. + z.cmus1 / tn);
        anulam = anulam * r_1 * z.ratiom;
        double t = (ro_old * nu_old + dw[l + 5 * nd]) * r_1;
        t = fmin(t, anulam);
        out[l + 5 * nd] = fmax(t, 0.);
    }
}


template <int MESH, int NEQ>
static int run_lu(const ZoneDev &z, const LuPlan &p, const double *rop, double *rop_out, int lu_match, cudaStream_t st)
{
    LuArgs a;
    a.rop = rop;
    for (int i = 0; i < 6; i++) a.box[i] = p.box[i];
    a.bi = p.box[1] - p.box[0] + 1; a.bj = p.box[3] - p.box[2] + 1; a.bk = p.box[5] - p.box[4] + 1;
    a.alias = (rop == rop_out) ? 1 : 0;
    dim3 b(32, 4, 1), g((a.bj + 31) / 32, (a.bk + 3) / 4, 1);
    int n = 0;
    a.order = lu_match ? 1 : (MESH == 2 ? 2 : 0);
    for (int h = 0; h < p.nplanes; h++) { a.plane = h; k_lu<MESH, NEQ, false><<<g, b, 0, st>>>(z, a); n++; }
    a.order = lu_match ? 1 : 0;
    for (int h = 0; h < p.nplanes; h++) { a.plane = h; k_lu<MESH, NEQ, true><<<g, b, 0, st>>>(z, a); n++; }
    int nb[6];
    for (int i = 0; i < 6; i++) nb[i] = p.box[i];
    if (lu_match) { nb[0] = 1; nb[1] = z.iv; nb[2] = 1; nb[3] = z.jv; nb[4] = 1; nb[5] = z.kv; }
    dim3 bn(64, 4, 1), gn((nb[1] - nb[0] + 64) / 64, (nb[3] - nb[2] + 4) / 4, nb[5] - nb[4] + 1);
    k_newton<NEQ><<<gn, bn, 0, st>>>(z, rop, rop_out, nb[0], nb[1], nb[2], nb[3], nb[4]);
    return n + 1;
}

int launch_lu(const ZoneDev &z, const LuPlan &p, const double *rop, double *rop_out, int lu_match, cudaStream_t st)
{
    // default: tile-pipelined wavefront (fb_lu.cu); FAST_B200_LU=plane keeps the one-launch-per-hyperplane kernels for A/B checks
    static const int use_plane = [] { const char *e = getenv("FAST_B200_LU"); return e && e[0] == 'p'; }();
    if (!use_plane || p.nrelax >= 2) return launch_lu_tiles(z, p, rop, rop_out, lu_match, st);   // relaxations: tile kernel only
    if (z.neq == 6) {
        if (z.typ == 0) return run_lu<0, 6>(z, p, rop, rop_out, lu_match, st);
        if (z.typ == 1) return run_lu<1, 6>(z, p, rop, rop_out, lu_match, st);
        return run_lu<2, 6>(z, p, rop, rop_out, lu_match, st);
    }
    if (z.typ == 0) return run_lu<0, 5>(z, p, rop, rop_out, lu_match, st);
    if (z.typ == 1) return run_lu<1, 5>(z, p, rop, rop_out, lu_match, st);
    return run_lu<2, 5>(z, p, rop, rop_out, lu_match, st);
}

void launch_resid(const ZoneDev &z, double *rdm, double *cflv, int nitcfg, int nss, cudaStream_t st)
{
    const int slot = nitcfg - 1;
    k_resid_zero<<<1, 32, 0, st>>>(rdm, slot, nss, z.neq);
    dim3 b(64, 4, 1);
    k_resid<<<grid3i(z.iv, z.jv, z.kv, b), b, 0, st>>>(z, rdm, slot, nss, cflv);
}
