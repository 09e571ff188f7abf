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
