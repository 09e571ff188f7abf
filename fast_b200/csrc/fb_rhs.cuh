// fb_rhs.cuh -- fused rhs kernel template; instantiated per (MESH, FLOW) in fb_rhs_m*f*.cu so that the 81 x 2
// variants compile in parallel.
#pragma once
#include <cstdlib>
#include "fb_kernels.cuh"
#include "fb_flux.cuh"
#include "fb_source.cuh"

static inline dim3 grid3r(int ni, int nj, int nk, dim3 b) { return dim3((ni + b.x - 1) / b.x, (nj + b.y - 1) / b.y, nk); }

// ------------------------------------------------------------------------------------------------ fused rhs
// One thread per interior cell, gather form: the six face fluxes of the cell are evaluated and summed in the
// reference's accumulation order (+Fk +Fj +Fi -Fi(i+1) -Fj(j+1) -Fk(k+1), template_FluxAndBalance.for:192-271),
// then the time-integration step is applied in registers so that drodm and rop_out are written once.
#include "fb_rhs_decl.cuh"

template <int MESH, int FLOW, int SCHEME, int SLOPE, bool IMPL>
__global__ void __launch_bounds__(128) k_rhs(ZoneDev z, RhsArgs a)
{
    constexpr int NEQ = FLOW == 3 ? 6 : 5;
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = 1 + blockIdx.z;
    if (i > z.iv || j > z.jv) return;
    const int nd = z.nd;
    const int l = inddm(z, i, j, k);
    const int lt = MESH == 2 ? 0 : indmtr(z, i, j, k);
    const double *__restrict__ rop = a.rop;
    double acc[NEQ], f[NEQ];
    // init_rhs.for:85-129
    if (IMPL || a.nitcfg == 1) {
#pragma unroll
        for (int e = 0; e < NEQ; e++) acc[e] = 0.;
    } else {
#pragma unroll
        for (int e = 0; e < NEQ; e++) acc[e] = z.drodm[l + e * nd] * a.coefH;
    }
    double coe1;   // dt/vol (tstb3_global.for) -- explicit: recomputed, implicit: from coe(:,1)
    if (IMPL) coe1 = z.coe[l];
    else coe1 = MESH == 2 ? z.dtc / z.vol0 : z.dtc / __ldg(z.vol + lt);
    if (FLOW == 3) sa_source<MESH, IMPL>(z, rop, l, lt, i, j, k, a.nitcfg, coe1, acc);
    // lower faces
    face_flux<MESH, FLOW, SCHEME, SLOPE, 2>(z, rop, l, lt, f);
#pragma unroll
    for (int e = 0; e < NEQ; e++) acc[e] += f[e];
    face_flux<MESH, FLOW, SCHEME, SLOPE, 1>(z, rop, l, lt, f);
#pragma unroll
    for (int e = 0; e < NEQ; e++) acc[e] += f[e];
    face_flux<MESH, FLOW, SCHEME, SLOPE, 0>(z, rop, l, lt, f);
#pragma unroll
    for (int e = 0; e < NEQ; e++) acc[e] += f[e];
    // upper faces = lower faces of the next cell
    face_flux<MESH, FLOW, SCHEME, SLOPE, 0>(z, rop, l + 1, MESH == 2 ? 0 : lt + z.mi, f);
#pragma unroll
    for (int e = 0; e < NEQ; e++) acc[e] -= f[e];
    face_flux<MESH, FLOW, SCHEME, SLOPE, 1>(z, rop, l + z.incj, MESH == 2 ? 0 : lt + z.mj, f);
#pragma unroll
    for (int e = 0; e < NEQ; e++) acc[e] -= f[e];
    face_flux<MESH, FLOW, SCHEME, SLOPE, 2>(z, rop, l + z.inck, MESH == 2 ? 0 : lt + z.mk, f);
#pragma unroll
    for (int e = 0; e < NEQ; e++) acc[e] -= f[e];

    const double cv = z.cv;
    if (!IMPL) {
        // core3ark3.for:108-154 (+203-236 for nutilde)
#pragma unroll
        for (int e = 0; e < NEQ; e++) z.drodm[l + e * nd] = acc[e];
        const double *__restrict__ ro = a.rop;   // the stage is updated from its own input (navier_stokes_struct.for:477-483)
        double r_1 = a.coefW * coe1;
        const double ro_old = ro[l], u_old = ro[l + nd], v_old = ro[l + 2 * nd], w_old = ro[l + 3 * nd], t_old = ro[l + 4 * nd];
        const double tab1 = acc[0] * r_1, tab2 = acc[1] * r_1, tab3 = acc[2] * r_1, tab4 = acc[3] * r_1, tab5 = acc[4] * r_1;
        const double cvinv = 1. / cv, cvinv2 = 0.5 * cvinv;
        const double rn = ro_old + tab1;
        const double r1n = 1. / rn;
        const double un = (ro_old * u_old + tab2) * r1n, vn = (ro_old * v_old + tab3) * r1n, wn = (ro_old * w_old + tab4) * r1n;
        const double roe_old = ro_old * (cv * t_old + 0.5 * (u_old * u_old + v_old * v_old + w_old * w_old));
        const double tn = (roe_old + tab5) * r1n * cvinv - cvinv2 * (un * un + vn * vn + wn * wn);
        double *__restrict__ out = a.rop_out;
        out[l] = rn; out[l + nd] = un; out[l + 2 * nd] = vn; out[l + 3 * nd] = wn; out[l + 4 * nd] = tn;
        if (FLOW == 3) {
            const double tab6 = acc[NEQ - 1] * r_1;
            double anulam = z.coesut * sqrt(tn * z.temp01) / (1. + z.cmus1 / tn);
            anulam = anulam * r1n * z.ratiom;
            double t = (ro_old * ro[l + 5 * nd] + tab6) * r1n;
            t = fmin(t, anulam);
            out[l + 5 * nd] = fmax(t, 0.);
        }
    } else {
        // core3as2.for:70-100: drodm = coe1*drodm + c1 W(rop_ssiter) + c2 W(ro) + c3 W(ro_M1)
        const double *__restrict__ r0 = a.rop, *__restrict__ rn = a.ro_n, *__restrict__ rm = a.ro_n1;
        implicit_assemble<NEQ>(z, r0, rn, rm, l, coe1, a.c1, a.c2, a.c3, acc);
#pragma unroll
        for (int e = 0; e < NEQ; e++) z.drodm[l + e * nd] = acc[e];
    }
}

// ------------------------------------------------------------------------------------------------ fused rhs, v2
// k-marching kernel: every face flux is evaluated ONCE and shared by its two cells.
//   CTA = 32 x (RJ+1) threads over a tile of 31 x RJ cells, marching over a chunk of k planes.
//   thread (li, m): lower i face and lower j face of cell (i0+li, j0+m, k) and its lower k face;
//     F_i(i+1) comes from lane li+1 (warp shuffle; lane 31 only owns a face), F_j(j+1) from row m+1 (shared memory,
//     double-buffered by plane parity: one __syncthreads per plane; row RJ only owns faces), F_k(k+1) is the lower k
//     face of the next plane: the cell of plane k-1 is finished (and written) when plane k has produced it.
//   The sum keeps the reference's order +F_k +F_j +F_i -F_i(i+1) -F_j(j+1) -F_k(k+1)
//   (template_FluxAndBalance.for:192-271), so the result is bitwise the one of the gather kernel above.
#define RHS2_RJ 8
#define RHS2_KC 32

template <int MESH, int FLOW, bool IMPL>
__device__ __forceinline__ void rhs_finish(const ZoneDev &z, const RhsArgs &a, int l, double coe1, double *acc)
{
    constexpr int NEQ = FLOW == 3 ? 6 : 5;
    const int nd = z.nd;
    const double cv = z.cv;
    if (!IMPL) {
        // core3ark3.for:108-154 (+203-236 for nutilde)
#pragma unroll
        for (int e = 0; e < NEQ; e++) z.drodm[l + e * nd] = acc[e];
        const double *__restrict__ ro = a.rop;
        double r_1 = a.coefW * coe1;
        const double ro_old = ro[l], u_old = ro[l + nd], v_old = ro[l + 2 * nd], w_old = ro[l + 3 * nd], t_old = ro[l + 4 * nd];
        const double tab1 = acc[0] * r_1, tab2 = acc[1] * r_1, tab3 = acc[2] * r_1, tab4 = acc[3] * r_1, tab5 = acc[4] * r_1;
        const double cvinv = 1. / cv, cvinv2 = 0.5 * cvinv;
        const double rn = ro_old + tab1;
        const double r1n = 1. / rn;
        const double un = (ro_old * u_old + tab2) * r1n, vn = (ro_old * v_old + tab3) * r1n, wn = (ro_old * w_old + tab4) * r1n;
        const double roe_old = ro_old * (cv * t_old + 0.5 * (u_old * u_old + v_old * v_old + w_old * w_old));
        const double tn = (roe_old + tab5) * r1n * cvinv - cvinv2 * (un * un + vn * vn + wn * wn);
        double *__restrict__ out = a.rop_out;
        out[l] = rn; out[l + nd] = un; out[l + 2 * nd] = vn; out[l + 3 * nd] = wn; out[l + 4 * nd] = tn;
        if (FLOW == 3) {
            const double tab6 = acc[NEQ - 1] * r_1;
            double anulam = z.coesut * sqrt(tn * z.temp01) / (1. + z.cmus1 / tn);
            anulam = anulam * r1n * z.ratiom;
            double t = (ro_old * ro[l + 5 * nd] + tab6) * r1n;
            t = fmin(t, anulam);
            out[l + 5 * nd] = fmax(t, 0.);
        }
    } else {
        // core3as2.for:70-100: drodm = coe1*drodm + c1 W(rop_ssiter) + c2 W(ro) + c3 W(ro_M1)
        implicit_assemble<NEQ>(z, a.rop, a.ro_n, a.ro_n1, l, coe1, a.c1, a.c2, a.c3, acc);
#pragma unroll
        for (int e = 0; e < NEQ; e++) z.drodm[l + e * nd] = acc[e];
    }
}

template <int MESH, int FLOW, int SCHEME, int SLOPE, bool IMPL>
__global__ void __launch_bounds__(32 * (RHS2_RJ + 1)) k_rhs2(ZoneDev z, RhsArgs a)
{
    constexpr int NEQ = FLOW == 3 ? 6 : 5, RJ = RHS2_RJ;
    __shared__ double sFj[2][RJ + 1][NEQ][32];
    const int li = threadIdx.x, m = threadIdx.y;
    const int i = 1 + blockIdx.x * 31 + li, j = 1 + blockIdx.y * RJ + m;
    const int k0 = 1 + blockIdx.z * RHS2_KC, k1 = min(k0 + RHS2_KC - 1, z.kv);
    const bool cell = li < 31 && m < RJ && i <= z.iv && j <= z.jv;      // owns a cell
    const bool fi = m < RJ && j <= z.jv && i <= z.iv + 1;               // owns a lower i face that somebody needs
    const bool fj = li < 31 && i <= z.iv && j <= z.jv + 1;              // owns a lower j face that somebody needs
    const int nd = z.nd;
    const double *__restrict__ rop = a.rop;
    double acc[NEQ], f[NEQ];
#pragma unroll
    for (int e = 0; e < NEQ; e++) acc[e] = 0.;
    double coe1p = 0.;        // dt/vol of the cell being finished (plane k-1)
    for (int k = k0; k <= k1 + 1; k++) {
        const int l = inddm(z, i, j, k);
        const int lt = MESH == 2 ? 0 : indmtr(z, i, j, k);
        const bool last = k == k1 + 1;
        // lower k face: finishes the cell of plane k-1, starts the cell of plane k
        if (cell) {
            face_flux<MESH, FLOW, SCHEME, SLOPE, 2>(z, rop, l, lt, f);
            if (k > k0) {
#pragma unroll
                for (int e = 0; e < NEQ; e++) acc[e] -= f[e];
                rhs_finish<MESH, FLOW, IMPL>(z, a, l - z.inck, coe1p, acc);
            }
        }
        if (last) break;      // uniform
        double coe1 = 0.;
        if (cell) {
            // init_rhs.for:85-129
            if (IMPL || a.nitcfg == 1) {
#pragma unroll
                for (int e = 0; e < NEQ; e++) acc[e] = 0.;
            } else {
#pragma unroll
                for (int e = 0; e < NEQ; e++) acc[e] = z.drodm[l + e * nd] * a.coefH;
            }
            if (IMPL) coe1 = z.coe[l];
            else coe1 = MESH == 2 ? z.dtc / z.vol0 : z.dtc / __ldg(z.vol + lt);
            if (FLOW == 3) sa_source<MESH, IMPL>(z, rop, l, lt, i, j, k, a.nitcfg, coe1, acc);
#pragma unroll
            for (int e = 0; e < NEQ; e++) acc[e] += f[e];
        }
        coe1p = coe1;
        // lower j face -> shared memory
        if (fj) {
            face_flux<MESH, FLOW, SCHEME, SLOPE, 1>(z, rop, l, lt, f);
#pragma unroll
            for (int e = 0; e < NEQ; e++) sFj[k & 1][m][e][li] = f[e];
            if (cell) {
#pragma unroll
                for (int e = 0; e < NEQ; e++) acc[e] += f[e];
            }
        }
        // lower i face; the upper one comes from the next lane
        if (fi) face_flux<MESH, FLOW, SCHEME, SLOPE, 0>(z, rop, l, lt, f);
        else {
#pragma unroll
            for (int e = 0; e < NEQ; e++) f[e] = 0.;
        }
#pragma unroll
        for (int e = 0; e < NEQ; e++) {
            const double up = __shfl_down_sync(0xffffffffu, f[e], 1);
            if (cell) acc[e] = (acc[e] + f[e]) - up;
        }
        __syncthreads();
        if (cell) {
#pragma unroll
            for (int e = 0; e < NEQ; e++) acc[e] -= sFj[k & 1][m + 1][e][li];
        }
    }
}

template <int MESH, int FLOW, int SCHEME, int SLOPE>
static void launch_rhs2_t(const ZoneDev &z, const RhsArgs &a, bool impl, cudaStream_t st)
{
    dim3 b(32, RHS2_RJ + 1, 1);
    dim3 g((z.iv + 30) / 31, (z.jv + RHS2_RJ - 1) / RHS2_RJ, (z.kv + RHS2_KC - 1) / RHS2_KC);
    if (impl) k_rhs2<MESH, FLOW, SCHEME, SLOPE, true><<<g, b, 0, st>>>(z, a);
    else k_rhs2<MESH, FLOW, SCHEME, SLOPE, false><<<g, b, 0, st>>>(z, a);
}

template <int MESH, int FLOW, int SCHEME, int SLOPE>
static void launch_rhs_t(const ZoneDev &z, const RhsArgs &a, bool impl, cudaStream_t st)
{
    static const bool gather = [] { const char *e = getenv("FAST_B200_RHS"); return e && e[0] == 'g'; }();
    if (!gather) { launch_rhs2_t<MESH, FLOW, SCHEME, SLOPE>(z, a, impl, st); return; }
    dim3 b(64, 2, 1);
    dim3 g = grid3r(z.iv, z.jv, z.kv, b);
    if (impl) k_rhs<MESH, FLOW, SCHEME, SLOPE, true><<<g, b, 0, st>>>(z, a);
    else k_rhs<MESH, FLOW, SCHEME, SLOPE, false><<<g, b, 0, st>>>(z, a);
}
template <int MESH, int FLOW, int SCHEME>
static void launch_rhs_s(const ZoneDev &z, const RhsArgs &a, bool impl, cudaStream_t st)
{
    switch (z.slope) {
    case 2: launch_rhs_t<MESH, FLOW, SCHEME, 2>(z, a, impl, st); break;
    case 3: launch_rhs_t<MESH, FLOW, SCHEME, 3>(z, a, impl, st); break;
    default: launch_rhs_t<MESH, FLOW, SCHEME, 1>(z, a, impl, st); break;
    }
}
template <int MESH, int FLOW>
void launch_rhs_k(const ZoneDev &z, const RhsArgs &a, bool impl, cudaStream_t st)
{
    switch (z.kflu) {
    case 1: launch_rhs_s<MESH, FLOW, 1>(z, a, impl, st); break;
    case 2: launch_rhs_s<MESH, FLOW, 2>(z, a, impl, st); break;
    default: launch_rhs_s<MESH, FLOW, 5>(z, a, impl, st); break;
    }
}
