// fb_lu.cu -- LU-SGS lower / upper sweeps + Newton update as a tile-pipelined wavefront (sm_100a, FP64).
//
// Reference: S/Compute/LU/invlu.for:6-344 -> invlu_l[_SA].for:79-215 (forward sweep, i fastest then j then k),
// invlu_u[_SA].for:60-215 (multiply by D, backward sweep, Newton update plane by plane), bodies lu_{i,j,k}_*.for,
// mjr_newton[_SA].for.  A cell depends only on its three lower (upper) neighbours, so every ordering that respects
// (i-1,j,k), (i,j-1,k), (i,j,k-1) reproduces the sequential recurrence exactly.
//
// Mapping onto the B200 (this replaces one launch per i+j+k hyperplane):
//   * the (j,k) cross-section of the sweep box is cut into TJ x TK tiles of grid lines; ONE CTA owns one tile for the
//     whole length of the box in i.  Thread (jj,kk) owns line (jj,kk) and marches along i, skewed by jj+kk steps, so
//     that at step s the CTA works on the cells of its tile that lie on the hyperplane ii+jj+kk = s.
//   * a cell computes its own dW*, then the three products B_d(l).dW(l) (the reference evaluates them in the receiving
//     cell; the value is identical because B_d only depends on the donor cell) and hands them to (jj+1,kk), (jj,kk+1)
//     through shared memory and to itself (ii+1) in registers: ONE __syncthreads per step.
//   * the products that leave the tile go through self-cleaning mailboxes in global memory (one 48-byte entry per
//     boundary cell, empty = signalling-NaN pattern, strong relaxed.gpu 8-byte accesses, prefetched two steps ahead):
//     no flags, no fences, the consumer tile trails its producers by TJ (TK) steps.  Tiles take tickets in i+j+k
//     order of their (J,K) index, so a tile only ever waits for tiles that are already running or finished.
//   * every per-cell input (rhs, coe(1:6), primitive state, cell-centred mean normals) is staged through shared
//     memory by two service warps with 8-byte cp.async copies whose lanes run along i (coalesced 64-byte segments per
//     line), double-buffered in chunks of CI cells per line with the line's own skew; results (dW* / the Newton-
//     updated state) leave through the same ring, written back along i.  The compute warps only touch shared memory.
// HBM traffic per cell: lower reads rhs+coe+(u,v,w,T)+normals, writes dW*; upper reads dW*+coe+state+normals and
// writes ro_P1 directly (mjr_newton fused, dW itself is never stored).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <type_traits>
#include "fb_implicit.cuh"

#define LU_CI 8            // cells per line per staged chunk
#define LU_TJ 8
#define LU_MB 6            // doubles per mailbox entry
#define LU_EMPTY 0xFFF7DEADBEEF5A5AULL   // signalling NaN: never produced by arithmetic
#define LU_SPIN_MAX (1 << 21)
#define LU_NC 256          // compute threads: 4 lanes per line
#define LU_NT 320          // + 2 service warps

template <int MESH, int NEQ, bool UPPER>
struct LuCfg {
    static constexpr int TJ = LU_TJ;
    static constexpr int TK = (MESH == 0 && NEQ == 6) ? 6 : 8;
    static constexpr int L = TJ * TK;
    static constexpr int NC = NEQ == 6 ? 6 : 5;            // coe columns staged
    static constexpr int NS = UPPER ? NEQ : 4;             // state variables staged (upper: all, lower: u,v,w,T)
    static constexpr int S0 = UPPER ? 0 : 1;
    static constexpr int NT = MESH == 0 ? 9 : (MESH == 1 ? 5 : 0);
    static constexpr int A_DW = 0, A_COE = NEQ, A_ST = NEQ + NC, A_T = NEQ + NC + NS;
    static constexpr int A = NEQ + NC + NS + NT;
    static constexpr size_t SMEM = (size_t)(2 * A * L * LU_CI + 2 * 3 * L * LU_MB + 4 * 16 * LU_MB) * sizeof(double);
};

int lu_tile_tk(int mesh, int neq) { return (mesh == 0 && neq == 6) ? 6 : 8; }

struct LuTileArgs {
    const double *rop;      // rop_ssiter
    double *rop_out;        // ro_P1
    double *dw;             // drodm: rhs -> dW*
    const double *coe;
    const double *tcm;      // cell-centred mean normals (k_lunorm)
    long long tstride;      // doubles between two components of tcm
    double *ifj, *ifk;      // mailboxes
    const int *order;       // ticket -> J | K << 16
    int *ticket;
    int *err;
    int box[6], nbx[6];
    int bi, bj, bk, NJ, NK;
    int alias, order_mode;
    int plane;              // ni*nj (3dhomo: tcm is a plane array)
    unsigned long long *dbg;   // FAST_B200_LU_DEBUG: per tile {start, first step, end, smid} (globaltimer ns)
};
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %smid;" : "=r"(r)); return r; }

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// mailbox accesses: strong (volatile) 8-byte accesses served by L2; an entry validates itself word by word
__device__ __forceinline__ bool mb_empty(unsigned long long v) { return v == LU_EMPTY; }

// B_d(l) . dW(l) with the cell-centred mean normal (tcx,tcy,tcz) of cell l (lu_{i,j,k}_*.for; signe folded into cd).
// The reference spells out the 5x5 flux Jacobian; the product is evaluated here in its factored form
//   theta = n.dm - qn drho,  dp = (g-1)(q2 drho - q.dm + dE),  A dW = [n.dm, qn dm + q theta + n dp, qn (dE + dp) + h theta]
// (identical algebra, ~4x fewer FP64 operations, differences at round-off level).
template <int NEQ>
__device__ __forceinline__ void lu_prod(double tcx, double tcy, double tcz, double u, double v, double w, double t,
                                        double cd, double gam, double cp, const double *d, double *b)
{
    const double gamm1 = gam - 1.;
    const double q2 = .5 * (u * u + v * v + w * w), h = cp * t + q2;
    const double qn = tcx * u + tcy * v + tcz * w;
    const double ndm = tcx * d[1] + tcy * d[2] + tcz * d[3];
    const double theta = ndm - qn * d[0];
    const double dp = gamm1 * ((q2 * d[0] - (u * d[1] + v * d[2] + w * d[3])) + d[4]);
    b[0] = ndm + cd * d[0];
    b[1] = (qn * d[1] + u * theta) + (tcx * dp + cd * d[1]);
    b[2] = (qn * d[2] + v * theta) + (tcy * dp + cd * d[2]);
    b[3] = (qn * d[3] + w * theta) + (tcz * dp + cd * d[3]);
    b[4] = (qn * (d[4] + dp) + h * theta) + cd * d[4];
    if (NEQ == 6) b[NEQ - 1] = (qn + cd) * d[NEQ - 1];
}

template <int MESH, int NEQ, bool UPPER>
__global__ void __launch_bounds__(LU_NT, 1) k_lu_tile(ZoneDev z, LuTileArgs a)
{
    using C = LuCfg<MESH, NEQ, UPPER>;
    constexpr int TJ = C::TJ, TK = C::TK, L = C::L, A = C::A, CI = LU_CI, EX = LU_MB;
    extern __shared__ double smem[];
    double *ring = smem;                              // [2][A][L][CI]
    double *exb = smem + 2 * A * L * CI;              // [2][3][L][EX] products handed to (ii+1), (jj+1), (kk+1)
    unsigned long long *mbs = (unsigned long long *)(exb + 2 * 3 * L * EX);   // [4][16][EX] prefetched mailbox entries
    __shared__ int s_tile;
    const int tid = threadIdx.x;
    if (tid == 0) s_tile = a.order[atomicAdd(a.ticket, 1)];
    __syncthreads();
    const int J = s_tile & 0xffff, K = s_tile >> 16;
    if (a.dbg && tid == 0) { a.dbg[4 * (K * a.NJ + J)] = gtime(); a.dbg[4 * (K * a.NJ + J) + 3] = smid(); }
    const int bi = a.bi, bj = a.bj, bk = a.bk;
    const int nst = bi + TJ + TK - 2;
    const int nd = z.nd;
    const int sg = UPPER ? -1 : 1;
    // cell address of sweep coordinates (ii, jjg, kkg): l0 + sg * (ii + jjg * incj + kkg * inck)
    const int l0 = UPPER ? inddm(z, a.box[1], a.box[3], a.box[5]) : inddm(z, a.box[0], a.box[2], a.box[4]);

    if (tid >= LU_NC) {
        // ------------------------------------------------------------------------------------ service warps
        const int st = tid - LU_NC;
        const int q = st / CI, cell = st % CI;           // line slot (= kk) and cell inside the chunk
        const bool qok = q < TK;
        const int kk = q, kkg = K * TK + kk;
        // rule(t): the line of row kk with (t - sigma) % CI == 0 finished chunk m-1 at step t-1 (m = (t-sigma)/CI):
        // write chunk m-1 back, refill its buffer with chunk m+1.  Executed one iteration late (t = s-1) so that the
        // barrier of iteration s-1 orders it after the compute step that produced the data.
        auto rule = [&](int t) {
            if (!qok) return;
            const int jj = ((t - kk) % TJ + TJ) % TJ;          // TJ == CI
            const int jjg = J * TJ + jj;
            if (jjg >= bj || kkg >= bk) return;
            const int sigma = jj + kk;
            const int m = (t - sigma) / CI;                     // exact: (t - sigma) is a multiple of CI (may be < 0)
            const int lam = kk * TJ + jj;
            const int lrow = l0 + sg * (jjg * z.incj + kkg * z.inck);
            {   // write back chunk m-1
                const int c = m - 1, ii = c * CI + cell;
                if (c >= 0 && ii < bi) {
                    const int l = lrow + sg * ii;
                    const double *src = ring + ((size_t)((c & 1) * A + C::A_DW) * L + lam) * CI + cell;
                    if (!UPPER) {
#pragma unroll
                        for (int e = 0; e < NEQ; e++) a.dw[l + e * nd] = src[(size_t)e * L * CI];
                    } else {
                        const int i = a.box[1] - ii, j = a.box[3] - jjg, k = a.box[5] - kkg;
                        if (i >= a.nbx[0] && i <= a.nbx[1] && j >= a.nbx[2] && j <= a.nbx[3] && k >= a.nbx[4] && k <= a.nbx[5]) {
#pragma unroll
                            for (int e = 0; e < NEQ; e++) a.rop_out[l + e * nd] = src[(size_t)e * L * CI];
                        }
                    }
                }
            }
            {   // refill with chunk m+1
                const int c = m + 1, ii = c * CI + cell;
                if (c >= 0 && ii < bi) {
                    const int l = lrow + sg * ii;
                    double *dst = ring + ((size_t)((c & 1) * A) * L + lam) * CI + cell;
#pragma unroll
                    for (int e = 0; e < NEQ; e++) cp_async8(dst + (size_t)(C::A_DW + e) * L * CI, a.dw + l + e * nd);
#pragma unroll
                    for (int e = 0; e < C::NC; e++) cp_async8(dst + (size_t)(C::A_COE + e) * L * CI, a.coe + l + e * nd);
#pragma unroll
                    for (int e = 0; e < C::NS; e++) cp_async8(dst + (size_t)(C::A_ST + e) * L * CI, a.rop + l + (C::S0 + e) * nd);
                    if (MESH != 2) {
                        const int lt = MESH == 1 ? l % a.plane : l;
#pragma unroll
                        for (int e = 0; e < C::NT; e++) cp_async8(dst + (size_t)(C::A_T + e) * L * CI, a.tcm + lt + e * a.tstride);
                    }
                }
            }
        };
        // mailbox prefetch: thread st < 48 fetches 16 bytes of the entry that interface line il needs at step sp and
        // parks them in shared memory (slot sp & 3); the load is issued before the refill work of the iteration and
        // consumed after it, so its L2 latency overlaps with the cp.async issue
        const int il = st / 3, part = st % 3;
        const bool isj = il < 8;
        const int mjj = isj ? 0 : il - 8, mkk = isj ? il : 0;
        const int mjg = J * TJ + mjj, mkg = K * TK + mkk;
        const bool mvalid = st < 48 && mjg < bj && mkg < bk && mkk < TK && (isj ? J > 0 : K > 0);
        const unsigned long long *mrow = nullptr;
        if (mvalid) mrow = (const unsigned long long *)(isj ? a.ifj + ((size_t)((J - 1) * bk + mkg) * bi) * LU_MB
                                                              : a.ifk + ((size_t)((K - 1) * bj + mjg) * bi) * LU_MB) + 2 * part;
        const int msig = mjj + mkk;
        for (int t = -CI; t <= -2; t++) { rule(t); cp_async_commit(); }
        for (int sp = 0; sp < 2; sp++) {
            const int ii = sp - msig;
            if (mvalid && ii >= 0 && ii < bi) {
                const volatile unsigned long long *p = mrow + (size_t)ii * LU_MB;
                unsigned long long *dst = mbs + ((size_t)((sp & 3) * 16 + il) * EX) + 2 * part;
                dst[0] = p[0]; dst[1] = p[1];
            }
        }
        for (int s = 0; s < nst; s++) {
            const int sp = s + 2, ii = sp - msig;
            const bool doit = mvalid && ii >= 0 && ii < bi;
            unsigned long long v0 = LU_EMPTY, v1 = LU_EMPTY;
            if (doit) { const volatile unsigned long long *p = mrow + (size_t)ii * LU_MB; v0 = p[0]; v1 = p[1]; }
            rule(s - 1);
            cp_async_commit();
            if (doit) { unsigned long long *dst = mbs + ((size_t)((sp & 3) * 16 + il) * EX) + 2 * part; dst[0] = v0; dst[1] = v1; }
            cp_async_wait<CI - 1>();
            __syncthreads();
        }
        __syncthreads();
        for (int s = nst; s <= nst + CI; s++) rule(s - 1);
        cp_async_wait<0>();
        return;
    }

    // ---------------------------------------------------------------------------------------- compute warps
    // 4 lanes per line: lane r < 3 evaluates the product of direction r, all lanes redo the (cheap) dW* of the line, so
    // that the three Jacobian products run on different lanes of the same instruction stream and the 8 compute warps
    // keep two warps per scheduler on the FP64 pipe.
    const int lam = tid >> 2, r = tid & 3;
    const int jj = lam % TJ, kk = lam / TJ;
    const int jjg = J * TJ + jj, kkg = K * TK + kk;
    const bool lvalid = lam < L && jjg < bj && kkg < bk;
    const int sigma = jj + kk;
    const bool has_j = jjg > 0, has_k = kkg > 0;
    const bool mbj_in = lvalid && jj == 0 && J > 0, mbk_in = lvalid && kk == 0 && K > 0;
    const bool mbj_out = lvalid && r == 1 && jj == TJ - 1 && J + 1 < a.NJ, mbk_out = lvalid && r == 2 && kk == TK - 1 && K + 1 < a.NK;
    volatile unsigned long long *mbj_r = mbj_in ? (unsigned long long *)(a.ifj + ((size_t)((J - 1) * bk + kkg) * bi) * LU_MB) : nullptr;
    volatile unsigned long long *mbk_r = mbk_in ? (unsigned long long *)(a.ifk + ((size_t)((K - 1) * bj + jjg) * bi) * LU_MB) : nullptr;
    volatile unsigned long long *mb_w = nullptr;
    if (mbj_out) mb_w = (unsigned long long *)(a.ifj + ((size_t)(J * bk + kkg) * bi) * LU_MB);
    if (mbk_out) mb_w = (unsigned long long *)(a.ifk + ((size_t)(K * bj + jjg) * bi) * LU_MB);
    // Newton box test of the own line (upper only)
    const int jg = UPPER ? a.box[3] - jjg : a.box[2] + jjg, kg = UPPER ? a.box[5] - kkg : a.box[4] + kkg;
    const bool line_in_nb = jg >= a.nbx[2] && jg <= a.nbx[3] && kg >= a.nbx[4] && kg <= a.nbx[5];
    const double signe = UPPER ? -0.5 : 0.5;
    const double gam = z.gam, cp = z.gam * z.cv;
    // accumulation order of the reference (invlu_l.for / invlu_u.for: k,j,i; overlap variants: i,j,k; 3dcart lower: j,k,i
    // when both exist): which product sits at position 0, 1, 2 of the sum
    int ord0 = 2, ord1 = 1, ord2 = 0;
    if (a.order_mode == 1) { ord0 = 0; ord1 = 1; ord2 = 2; }
    else if (a.order_mode == 2 && has_j && has_k) { ord0 = 1; ord1 = 2; ord2 = 0; }
    const int rr = r < 3 ? r : 2;                      // lane 3 repeats the k product (not stored)

    // one incoming product: from shared memory (neighbour line of the tile), from the prefetched mailbox entry (tile
    // boundary; falls back to polling global memory while the producer tile has not delivered), or zero (box boundary)
    auto mailbox = [&](volatile unsigned long long *row, const unsigned long long *pre, int ii, double *bv) {
        unsigned long long v[NEQ];
        bool ok = true;
#pragma unroll
        for (int e = 0; e < NEQ; e++) { v[e] = pre[e]; ok = ok && !mb_empty(v[e]); }
        volatile unsigned long long *p = row + (size_t)ii * LU_MB;
        int spin = 0;
        while (!ok) {
            ok = true;
#pragma unroll
            for (int e = 0; e < NEQ; e++) { v[e] = p[e]; ok = ok && !mb_empty(v[e]); }
            if (++spin > LU_SPIN_MAX || (((spin & 1023) == 0) && *(volatile int *)a.err)) { *a.err = 1; break; }
        }
#pragma unroll
        for (int e = 0; e < NEQ; e++) bv[e] = __longlong_as_double((long long)v[e]);
        if (r == 0) {
#pragma unroll
            for (int e = 0; e < LU_MB; e++) p[e] = LU_EMPTY;      // self-cleaning: ready for the next sweep
        }
    };

    if (a.dbg && tid == 0) a.dbg[4 * (K * a.NJ + J) + 1] = gtime();
    for (int s = 0; s < nst; s++) {
        __syncthreads();
        const int ii = s - sigma;
        if (!(lvalid && ii >= 0 && ii < bi)) continue;
        const int c = ii / CI, slot = ii % CI;
        const double *R = ring + ((size_t)((c & 1) * A) * L + lam) * CI + slot;
#define RING(arr) R[(size_t)(arr) * L * CI]
        double acc[NEQ], bi_[EX], bj_[EX], bk_[EX];
#pragma unroll
        for (int e = 0; e < NEQ; e++) acc[e] = RING(C::A_DW + e);
        const double coe1 = RING(C::A_COE), coe5 = RING(C::A_COE + 4);
        double coe6 = 1.;
        if (NEQ == 6) coe6 = RING(C::A_COE + 5);
        // state, mean normal and spectral radius of the lane's direction
        double ro = 0., u, v, w, t, nut = 0.;
        if (UPPER) { ro = RING(C::A_ST); u = RING(C::A_ST + 1); v = RING(C::A_ST + 2); w = RING(C::A_ST + 3); t = RING(C::A_ST + 4); if (NEQ == 6) nut = RING(C::A_ST + 5); }
        else { u = RING(C::A_ST); v = RING(C::A_ST + 1); w = RING(C::A_ST + 2); t = RING(C::A_ST + 3); }
        double tcx = 0., tcy = 0., tcz = 0.;
        if (MESH == 0) { tcx = RING(C::A_T + 3 * rr); tcy = RING(C::A_T + 3 * rr + 1); tcz = RING(C::A_T + 3 * rr + 2); }
        else if (MESH == 1) {
            if (rr < 2) { tcx = RING(C::A_T + 2 * rr); tcy = RING(C::A_T + 2 * rr + 1); } else tcz = RING(C::A_T + 4);
        } else { if (rr == 0) tcx = z.tc0[0]; else if (rr == 1) tcy = z.tc0[1]; else tcz = z.tc0[2]; }
        const double cd = RING(C::A_COE + 1 + rr) * signe;
        // incoming products (zero where the neighbour is outside the sweep box: adding 0 leaves acc unchanged)
        const double *xin = exb + (size_t)((s + 1) & 1) * 3 * L * EX;
        {
            const double2 *p = (const double2 *)(xin + (size_t)lam * EX);
            const double2 x0 = p[0], x1 = p[1], x2 = p[2];
            const bool h = ii > 0;
            bi_[0] = h ? x0.x : 0.; bi_[1] = h ? x0.y : 0.; bi_[2] = h ? x1.x : 0.; bi_[3] = h ? x1.y : 0.; bi_[4] = h ? x2.x : 0.; bi_[5] = h ? x2.y : 0.;
        }
        if (jj > 0) {
            const double2 *p = (const double2 *)(xin + (size_t)(L + lam - 1) * EX);
            const double2 x0 = p[0], x1 = p[1], x2 = p[2];
            bj_[0] = x0.x; bj_[1] = x0.y; bj_[2] = x1.x; bj_[3] = x1.y; bj_[4] = x2.x; bj_[5] = x2.y;
        } else if (has_j) mailbox(mbj_r, mbs + (size_t)((s & 3) * 16 + kk) * EX, ii, bj_);
        else {
#pragma unroll
            for (int e = 0; e < EX; e++) bj_[e] = 0.;
        }
        if (kk > 0) {
            const double2 *p = (const double2 *)(xin + (size_t)(2 * L + lam - TJ) * EX);
            const double2 x0 = p[0], x1 = p[1], x2 = p[2];
            bk_[0] = x0.x; bk_[1] = x0.y; bk_[2] = x1.x; bk_[3] = x1.y; bk_[4] = x2.x; bk_[5] = x2.y;
        } else if (has_k) mailbox(mbk_r, mbs + (size_t)((s & 3) * 16 + 8 + jj) * EX, ii, bk_);
        else {
#pragma unroll
            for (int e = 0; e < EX; e++) bk_[e] = 0.;
        }
        if (UPPER) {
#pragma unroll
            for (int e = 0; e < 5; e++) acc[e] = acc[e] * coe5;
            if (NEQ == 6) acc[NEQ - 1] = acc[NEQ - 1] * coe6;
        }
        const double xal = coe1 * signe;
#pragma unroll
        for (int e = 0; e < NEQ; e++) {
            const double t0 = ord0 == 0 ? bi_[e] : (ord0 == 1 ? bj_[e] : bk_[e]);
            const double t1 = ord1 == 0 ? bi_[e] : (ord1 == 1 ? bj_[e] : bk_[e]);
            const double t2 = ord2 == 0 ? bi_[e] : (ord2 == 1 ? bj_[e] : bk_[e]);
            acc[e] = ((acc[e] + t0 * xal) + t1 * xal) + t2 * xal;
        }
        const double diag = 1. / coe5;
        double dwn[NEQ];
#pragma unroll
        for (int e = 0; e < 5; e++) dwn[e] = acc[e] * diag;
        if (NEQ == 6) dwn[NEQ - 1] = acc[NEQ - 1] / coe6;
        // Newton update of the own cell (mjr_newton[_SA].for); with rop aliased to rop_out the reference has already
        // updated plane k+1 when it sweeps plane k, i.e. the k product of a cell uses its UPDATED state
        if (UPPER) {
            const int ig = a.box[1] - ii;
            const bool in_nb = line_in_nb && ig >= a.nbx[0] && ig <= a.nbx[1];
            const double cvinv = 1. / z.cv, cvinv2 = 0.5 * cvinv;
            const double rn = ro + dwn[0];
            const double r_1 = 1. / rn;
            const double un = (ro * u + dwn[1]) * r_1, vn = (ro * v + dwn[2]) * r_1, wn = (ro * w + dwn[3]) * r_1;
            const double roe_old = ro * (z.cv * t + 0.5 * (u * u + v * v + w * w));
            const double tn = (roe_old + dwn[4]) * r_1 * cvinv - cvinv2 * (un * un + vn * vn + wn * wn);
            if (in_nb && r == 3) {
                double *W = const_cast<double *>(R);
                W[(size_t)(C::A_DW) * L * CI] = rn; W[(size_t)(C::A_DW + 1) * L * CI] = un; W[(size_t)(C::A_DW + 2) * L * CI] = vn;
                W[(size_t)(C::A_DW + 3) * L * CI] = wn; W[(size_t)(C::A_DW + 4) * L * CI] = tn;
                if (NEQ == 6) {
                    double anulam = z.coesut * sqrt(tn * z.temp01) / (1. + z.cmus1 / tn);
                    anulam = anulam * r_1 * z.ratiom;
                    double tt = (ro * nut + dwn[NEQ - 1]) * r_1;
                    tt = fmin(tt, anulam);
                    W[(size_t)(C::A_DW + 5) * L * CI] = fmax(tt, 0.);
                }
            }
            if (a.alias && in_nb && rr == 2) { u = un; v = vn; w = wn; t = tn; }
        } else if (r == 3) {
            double *W = const_cast<double *>(R);
#pragma unroll
            for (int e = 0; e < NEQ; e++) W[(size_t)(C::A_DW + e) * L * CI] = dwn[e];
        }
        // outgoing product of the lane's direction
        double b[EX];
        b[EX - 1] = 0.;
        lu_prod<NEQ>(tcx, tcy, tcz, u, v, w, t, cd, gam, cp, dwn, b);
        if (r < 3) {
            double2 *p = (double2 *)(exb + ((size_t)((s & 1) * 3 + r) * L + lam) * EX);
            p[0] = make_double2(b[0], b[1]); p[1] = make_double2(b[2], b[3]); p[2] = make_double2(b[4], b[5]);
            if (mb_w) {
                volatile unsigned long long *o = mb_w + (size_t)ii * LU_MB;
#pragma unroll
                for (int e = 0; e < LU_MB; e++) o[e] = (unsigned long long)__double_as_longlong(b[e]);
            }
        }
#undef RING
    }
    __syncthreads();
    if (a.dbg && tid == 0) a.dbg[4 * (K * a.NJ + J) + 2] = gtime();
}

// cell-centred mean normals 0.5*(t_d(l) + t_d(l + d)) of every cell (lu_{i,j,k}_*.for build them on the fly)
template <int MESH>
__global__ void k_lunorm(ZoneDev z, double *tcm, long long tstride)
{
    const int i = 1 - z.nf + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 - z.nf + blockIdx.y * blockDim.y + threadIdx.y;
    const int k = 1 - z.nfk + blockIdx.z;
    if (i > z.iv + z.nf || j > z.jv + z.nf) return;
    const int l = inddm(z, i, j, k), lt = indmtr(z, i, j, k), nm = z.nm;
    if (MESH == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            tcm[l + (0 + c) * tstride] = 0.5 * (z.ti[lt + z.mi + c * nm] + z.ti[lt + c * nm]);
            tcm[l + (3 + c) * tstride] = 0.5 * (z.tj[lt + z.mj + c * nm] + z.tj[lt + c * nm]);
            tcm[l + (6 + c) * tstride] = 0.5 * (z.tk[lt + z.mk + c * nm] + z.tk[lt + c * nm]);
        }
    } else {
        if (blockIdx.z != 0) return;
        const int lp = l % (z.ni * z.nj);
#pragma unroll
        for (int c = 0; c < 2; c++) {
            tcm[lp + (0 + c) * tstride] = 0.5 * (z.ti[lt + z.mi + c * nm] + z.ti[lt + c * nm]);
            tcm[lp + (2 + c) * tstride] = 0.5 * (z.tj[lt + z.mj + c * nm] + z.tj[lt + c * nm]);
        }
        tcm[lp + 4 * tstride] = z.tk[lt];
    }
}

__global__ void k_fill_u64(unsigned long long *p, size_t n, unsigned long long v)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// ------------------------------------------------------------------------------------------------ host side
int lu_plan_build(LuPlan &p, const ZoneDev &z, const int box[6])
{
    lu_plan_free(p);
    for (int i = 0; i < 6; i++) p.box[i] = box[i];
    const int bi = box[1] - box[0] + 1, bj = box[3] - box[2] + 1, bk = box[5] - box[4] + 1;
    p.nplanes = bi + bj + bk - 2;
    p.tk = lu_tile_tk(z.typ, z.neq);
    p.NJ = (bj + LU_TJ - 1) / LU_TJ; p.NK = (bk + p.tk - 1) / p.tk;
    std::vector<int> ord;
    for (int K = 0; K < p.NK; K++) for (int J = 0; J < p.NJ; J++) ord.push_back(J | (K << 16));
    std::stable_sort(ord.begin(), ord.end(), [](int a, int b) { return (a & 0xffff) + (a >> 16) < (b & 0xffff) + (b >> 16); });
    if (cudaMalloc(&p.d_order, ord.size() * sizeof(int)) != cudaSuccess) return 1;
    cudaMemcpy(p.d_order, ord.data(), ord.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (cudaMalloc(&p.d_ctl, 4 * sizeof(int)) != cudaSuccess) return 1;
    cudaMemset(p.d_ctl, 0, 4 * sizeof(int));
    p.n_ifj = (size_t)std::max(p.NJ - 1, 0) * bk * bi * LU_MB;
    p.n_ifk = (size_t)std::max(p.NK - 1, 0) * bj * bi * LU_MB;
    const size_t nmb = p.n_ifj + p.n_ifk;
    if (nmb) {
        if (cudaMalloc(&p.d_mb, nmb * sizeof(double)) != cudaSuccess) return 1;
        k_fill_u64<<<592, 256>>>((unsigned long long *)p.d_mb, nmb, LU_EMPTY);
    }
    if (z.typ != 2) {
        p.tstride = z.typ == 0 ? (long long)z.nd : (long long)z.ni * z.nj;
        const int nt = z.typ == 0 ? 9 : 5;
        if (cudaMalloc(&p.d_tcm, (size_t)nt * p.tstride * sizeof(double)) != cudaSuccess) return 1;
        cudaMemset(p.d_tcm, 0, (size_t)nt * p.tstride * sizeof(double));
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

void lu_plan_free(LuPlan &p)
{
    cudaFree(p.d_order); cudaFree(p.d_ctl); cudaFree(p.d_mb); cudaFree(p.d_tcm);
    p.d_order = nullptr; p.d_ctl = nullptr; p.d_mb = nullptr; p.d_tcm = nullptr;
}

void lu_plan_metrics(LuPlan &p, const ZoneDev &z, cudaStream_t st)
{
    if (z.typ == 2 || !p.d_tcm) return;
    dim3 b(64, 4, 1), g((z.iv + 2 * z.nf + 63) / 64, (z.jv + 2 * z.nf + 3) / 4, z.typ == 0 ? z.kv + 2 * z.nfk : 1);
    if (z.typ == 0) k_lunorm<0><<<g, b, 0, st>>>(z, p.d_tcm, p.tstride);
    else k_lunorm<1><<<g, b, 0, st>>>(z, p.d_tcm, p.tstride);
}

template <int MESH, int NEQ>
static int run_lu_tiles(const ZoneDev &z, const LuPlan &p, const double *rop, double *rop_out, int lu_match, cudaStream_t st)
{
    LuTileArgs a;
    a.rop = rop; a.rop_out = rop_out; a.dw = z.drodm; a.coe = z.coe; a.tcm = p.d_tcm; a.tstride = p.tstride;
    a.ifj = p.d_mb; a.ifk = p.d_mb ? p.d_mb + p.n_ifj : nullptr;
    a.order = p.d_order; a.err = p.d_ctl + 2;
    for (int i = 0; i < 6; i++) { a.box[i] = p.box[i]; a.nbx[i] = p.box[i]; }
    if (lu_match) { a.nbx[0] = 1; a.nbx[1] = z.iv; a.nbx[2] = 1; a.nbx[3] = z.jv; a.nbx[4] = 1; a.nbx[5] = z.kv; }
    a.bi = p.box[1] - p.box[0] + 1; a.bj = p.box[3] - p.box[2] + 1; a.bk = p.box[5] - p.box[4] + 1;
    a.NJ = p.NJ; a.NK = p.NK; a.alias = (rop == rop_out) ? 1 : 0; a.plane = z.ni * z.nj;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_lu_tile<MESH, NEQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LuCfg<MESH, NEQ, false>::SMEM);
        cudaFuncSetAttribute(k_lu_tile<MESH, NEQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LuCfg<MESH, NEQ, true>::SMEM);
        attr_done = true;
    }
    cudaMemsetAsync(p.d_ctl, 0, 2 * sizeof(int), st);
    const int ntiles = p.NJ * p.NK;
    static const bool dbg = getenv("FAST_B200_LU_DEBUG") != nullptr;
    static unsigned long long *d_dbg = nullptr; static int dbg_calls = 0;
    a.dbg = nullptr;
    if (dbg && dbg_calls < 3) { if (!d_dbg) cudaMalloc(&d_dbg, 2 * 4 * 65536 * sizeof(unsigned long long)); a.dbg = d_dbg; }
    a.ticket = p.d_ctl; a.order_mode = lu_match ? 1 : (MESH == 2 ? 2 : 0);
    k_lu_tile<MESH, NEQ, false><<<ntiles, LU_NT, LuCfg<MESH, NEQ, false>::SMEM, st>>>(z, a);
    a.ticket = p.d_ctl + 1; a.order_mode = lu_match ? 1 : 0;
    if (a.dbg) a.dbg += 4 * 65536;
    k_lu_tile<MESH, NEQ, true><<<ntiles, LU_NT, LuCfg<MESH, NEQ, true>::SMEM, st>>>(z, a);
    if (a.dbg) {
        cudaStreamSynchronize(st);
        std::vector<unsigned long long> h(2 * 4 * 65536);
        cudaMemcpy(h.data(), d_dbg, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        for (int sw = 0; sw < 2; sw++) {
            const unsigned long long *q = h.data() + sw * 4 * 65536; unsigned long long t0 = ~0ULL, t1 = 0;
            for (int t = 0; t < ntiles; t++) { t0 = std::min(t0, q[4 * t]); t1 = std::max(t1, q[4 * t + 2]); }
            fprintf(stderr, "LUDBG sweep %d call %d: NJ %d NK %d tiles %d bi %d total %.1f us\n", sw, dbg_calls, p.NJ, p.NK, ntiles, a.bi, (t1 - t0) * 1e-3);
            for (int t = 0; t < ntiles; t++) {
                const int J = t % p.NJ, K = t / p.NJ;
                if (dbg_calls == 1 && (J % 4 == 0) && (K % 2 == 0 || K == p.NK - 1))
                    fprintf(stderr, "LUDBG   tile J %2d K %2d sm %3llu start %8.1f first %8.1f end %8.1f dur %8.1f\n", J, K, q[4 * t + 3],
                            (q[4 * t] - t0) * 1e-3, (q[4 * t + 1] - t0) * 1e-3, (q[4 * t + 2] - t0) * 1e-3, (q[4 * t + 2] - q[4 * t]) * 1e-3);
            }
        }
        dbg_calls++;
    }
    return 3;
}

int launch_lu_tiles(const ZoneDev &z, const LuPlan &p, const double *rop, double *rop_out, int lu_match, cudaStream_t st)
{
    if (z.neq == 6) {
        if (z.typ == 0) return run_lu_tiles<0, 6>(z, p, rop, rop_out, lu_match, st);
        if (z.typ == 1) return run_lu_tiles<1, 6>(z, p, rop, rop_out, lu_match, st);
        return run_lu_tiles<2, 6>(z, p, rop, rop_out, lu_match, st);
    }
    if (z.typ == 0) return run_lu_tiles<0, 5>(z, p, rop, rop_out, lu_match, st);
    if (z.typ == 1) return run_lu_tiles<1, 5>(z, p, rop, rop_out, lu_match, st);
    return run_lu_tiles<2, 5>(z, p, rop, rop_out, lu_match, st);
}

int lu_plan_error(const LuPlan &p)
{
    int e = 0;
    if (p.d_ctl) cudaMemcpy(&e, p.d_ctl + 2, sizeof(int), cudaMemcpyDeviceToHost);
    return e;
}
