// fb_view.cuh -- how the flux / source device functions reach the primitive state rop(NDIMDX,NEQ) and mu(+mut).
//
//   GView: straight from global memory (L1/L2 path; any zone size).  Cell index = the reference's l (S/formule_param.h).
//   SView: from the TMA-staged shared-memory window of the k-marching kernel k_rhs3 (fb_rhs.cuh): planes k-1, k, k+1
//          as [slot][var][row][col] boxes with compile-time row / variable strides, so every access is an LDS with an
//          immediate offset; plane k-2 of the thread's own column (only needed by the k-face MUSCL stencil) is carried
//          in registers.  Cell index = offset of (col,row) inside the slot of plane k.
// up<D>/dn<D> move one cell in direction D; in an SView the k moves are only valid from plane k.
#pragma once
#include "fb_common.cuh"

struct GView {
    const double *__restrict__ rop; const double *__restrict__ xmut;
    int nd, incj, inck;
    __device__ __forceinline__ GView(const ZoneDev &z, const double *r) : rop(r), xmut(z.xmut), nd(z.nd), incj(z.incj), inck(z.inck) {}
    __device__ __forceinline__ double q(int e, int l) const { return rop[l + e * nd]; }
    __device__ __forceinline__ double mu(int l) const { return xmut[l]; }
    template <int D> __device__ __forceinline__ int up(int l) const { return l + (D == 0 ? 1 : (D == 1 ? incj : inck)); }
    template <int D> __device__ __forceinline__ int dn(int l) const { return l - (D == 0 ? 1 : (D == 1 ? incj : inck)); }
    template <int D> __device__ __forceinline__ double qm2(int e, int l) const { return rop[l - 2 * (D == 0 ? 1 : (D == 1 ? incj : inck)) + e * nd]; }
};

template <int RS, int VS, int NEQ, int MUOFF>
struct SView {
    const double *s;          // shared-memory ring
    int kup, kdn;             // offset (doubles) from the slot of plane k to the slots of planes k+1 / k-1
    const double *km2;        // register array: own column at plane k-2
    __device__ __forceinline__ double q(int e, int l) const { return s[l + e * VS]; }
    __device__ __forceinline__ double mu(int l) const { return s[l + MUOFF]; }
    template <int D> __device__ __forceinline__ int up(int l) const { return l + (D == 0 ? 1 : (D == 1 ? RS : kup)); }
    template <int D> __device__ __forceinline__ int dn(int l) const { return l - (D == 0 ? 1 : (D == 1 ? RS : -kdn)); }
    template <int D> __device__ __forceinline__ double qm2(int e, int l) const { return D == 2 ? km2[e] : s[l - 2 * (D == 0 ? 1 : RS) + e * VS]; }
};

// 3dhomo metrics (ti, tj: two components; tk, vol: one; all independent of k).  GMet1 reads the node-strided plane
// arrays through L1; SMet1 is the copy of the tile's (36 x rows) neighbourhood that k_rhs3 keeps in shared memory for
// the whole k march (same in-plane index as the state window, immediate offsets).  lt = metric index of the cell.
struct GMet1 {
    const double *pti, *ptj, *ptk, *pvol; int nm, mi, mj;
    __device__ __forceinline__ GMet1(const ZoneDev &z) : pti(z.ti), ptj(z.tj), ptk(z.tk), pvol(z.vol), nm(z.nm), mi(z.mi), mj(z.mj) {}
    __device__ __forceinline__ double ti(int e, int lt) const { return __ldg(pti + lt + e * nm); }
    __device__ __forceinline__ double tj(int e, int lt) const { return __ldg(ptj + lt + e * nm); }
    __device__ __forceinline__ double tk(int lt) const { return __ldg(ptk + lt); }
    __device__ __forceinline__ double vol(int lt) const { return __ldg(pvol + lt); }
    template <int D> __device__ __forceinline__ int up(int lt) const { return lt + (D == 0 ? mi : mj); }
    template <int D> __device__ __forceinline__ int dn(int lt) const { return lt - (D == 0 ? mi : mj); }
};
template <int RS, int VS>
struct SMet1 {
    const double *s;          // [ti.x, ti.y, tj.x, tj.y, tk, vol][rows][RS]
    __device__ __forceinline__ double ti(int e, int lt) const { return s[lt + e * VS]; }
    __device__ __forceinline__ double tj(int e, int lt) const { return s[lt + (2 + e) * VS]; }
    __device__ __forceinline__ double tk(int lt) const { return s[lt + 4 * VS]; }
    __device__ __forceinline__ double vol(int lt) const { return s[lt + 5 * VS]; }
    template <int D> __device__ __forceinline__ int up(int lt) const { return lt + (D == 0 ? 1 : RS); }
    template <int D> __device__ __forceinline__ int dn(int lt) const { return lt - (D == 0 ? 1 : RS); }
};
