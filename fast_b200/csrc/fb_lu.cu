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
#include <map>
#include <cstring>
#include <cuda.h>
#include "fb_implicit.cuh"
#include "fb_kernels.cuh"

#define LU_CI 8            // cells per line per staged chunk
#define LU_TJ 8
#define LU_MB 6            // doubles per mailbox entry
#define LU_EMPTY 0xFFF7DEADBEEF5A5AULL   // signalling NaN: never produced by arithmetic
#define LU_SPIN_MAX (1 << 21)
#define LU_NC 256          // compute threads: 4 lanes per line
#define LU_NR 128          // refill threads (cp.async, lanes along i)
#define LU_NT 448          // + 64 write-back / mailbox threads
// per-role cycle counters of the FAST_B200_LU_DEBUG report (LUCLK lines) cost two clock reads per step and role: build
// with -DLU_CLK to get them (make VAR=clk EXTRA=-DLU_CLK); otherwise they read as zero
#ifdef LU_CLK
#define LU_CLOCK() clock64()
#else
#define LU_CLOCK() 0LL
#endif

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
    static constexpr size_t SMEM = (size_t)(2 * A * L * LU_CI + 2 * 3 * L * LU_MB + 6 * 16 * LU_MB) * sizeof(double);
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
    int newton;             // upper sweep: 1 = Newton update written to rop_out (mjr_newton fused), 0 = dW stored back (nb_relax >= 2)
    int plane;              // ni*nj (3dhomo: tcm is a plane array)
    const CUtensorMap *tm;     // k_lu_cell: skewed views of drodm, coe, tcm (device memory)
    const CUtensorMap *tm_rop; // ... and of rop_ssiter (u, v, w, T for the lower sweep, all variables for the upper one)
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

#include "fb_lu_prod.cuh"

// one tile (J | K << 16 = tile) of zone z for the whole length of the sweep box; z and a live in the constant bank
// (k_lu_tile: one zone per launch) or in shared memory (k_lu_tile_b: all zones of a solver in one launch)
template <int MESH, int NEQ, bool UPPER>
__device__ __forceinline__ void lu_tile_body(const ZoneDev &z, const LuTileArgs &a, const int tile, double *smem)
{
    using C = LuCfg<MESH, NEQ, UPPER>;
    constexpr int TJ = C::TJ, TK = C::TK, L = C::L, A = C::A, CI = LU_CI, EX = LU_MB;
    double *ring = smem;                              // [2][A][L][CI]
    double *exb = smem + 2 * A * L * CI;              // [2][L][3][EX] products handed to (ii+1), (jj+1), (kk+1)
    unsigned long long *mbs = (unsigned long long *)(exb + 2 * 3 * L * EX);   // [4][16][EX] prefetched mailbox words
    double *mbx = (double *)(mbs + 4 * 16 * EX);      // [2][16][EX] validated products of the neighbour tiles
    const int tid = threadIdx.x;
    // exchange buffers start at zero: products that are multiplied by a zero weight (box boundary) must be finite
    for (int i = tid; i < 2 * 3 * L * EX + 6 * 16 * EX; i += LU_NT) exb[i] = 0.;
    __syncthreads();
    const int J = tile & 0xffff, K = tile >> 16;
    if (a.dbg && tid == 0) { a.dbg[16 * (K * a.NJ + J)] = gtime(); a.dbg[16 * (K * a.NJ + J) + 3] = smid(); }
    const int bi = a.bi, bj = a.bj, bk = a.bk;
    const int nst = bi + TJ + TK - 2;
    const int nd = z.nd;
    const int sg = UPPER ? -1 : 1;
    // cell address of sweep coordinates (ii, jjg, kkg): l0 + sg * (ii + jjg * incj + kkg * inck)
    const int l0 = UPPER ? inddm(z, a.box[1], a.box[3], a.box[5]) : inddm(z, a.box[0], a.box[2], a.box[4]);

    if (tid >= LU_NC) {
        // ------------------------------------------------------------------------------------ service warps
        // Line bookkeeping shared by the roles.  rule(t): the line of row kk with (t - sigma) % CI == 0 finished chunk
        // m-1 at step t-1 (m = (t - sigma) / CI): chunk m-1 is written back (one iteration later, t = s-1, so that the
        // barrier of iteration s-1 orders it after the compute step that produced it) and its buffer refilled with
        // chunk m+1 (two iterations later, t = s-2, after the write-back has read it).
        const int st = tid - LU_NC;
        const int q = (st / CI) % 8, cell = st % CI;     // line slot (= kk) and cell inside the chunk
        const int kk = q, kkg = K * TK + kk;
        struct Ln { int lam, lrow, m, jjg; bool ok; };
        auto line_of = [&](int t) {
            Ln o;
            const int jj = ((t - kk) % TJ + TJ) % TJ;          // TJ == CI
            o.jjg = J * TJ + jj;
            o.ok = q < TK && o.jjg < bj && kkg < bk;
            o.m = (t - (jj + kk)) / CI;                         // exact: a multiple of CI (may be < 0)
            o.lam = kk * TJ + jj;
            o.lrow = l0 + sg * (o.jjg * z.incj + kkg * z.inck);
            return o;
        };
        if (st < LU_NR) {
            // ---- refill role: 128 threads, lanes along i, the staged arrays split in two halves
            const int half = st / 64;
            auto refill = [&](int t) {
                const Ln o = line_of(t);
                const int c = o.m + 1, ii = c * CI + cell;
                if (!(o.ok && c >= 0 && ii < bi)) return;
                const int l = o.lrow + sg * ii;
                double *dst = ring + ((size_t)((c & 1) * A) * L + o.lam) * CI + cell;
                if (half == 0) {
#pragma unroll
                    for (int e = 0; e < NEQ; e++) cp_async8(dst + (size_t)(C::A_DW + e) * L * CI, a.dw + l + e * nd);
#pragma unroll
                    for (int e = 0; e < C::NC; e++) cp_async8(dst + (size_t)(C::A_COE + e) * L * CI, a.coe + l + e * nd);
                } else {
#pragma unroll
                    for (int e = 0; e < C::NS; e++) cp_async8(dst + (size_t)(C::A_ST + e) * L * CI, a.rop + l + (C::S0 + e) * nd);
                    if (MESH != 2) {
                        const int lt = MESH == 1 ? l % a.plane : l;
#pragma unroll
                        for (int e = 0; e < C::NT; e++) cp_async8(dst + (size_t)(C::A_T + e) * L * CI, a.tcm + lt + e * a.tstride);
                    }
                }
            };
            for (int t = -CI; t <= -3; t++) { refill(t); cp_async_commit(); }
            long long tw = 0, tq = 0;
            for (int s = 0; s < nst; s++) {
                const long long t0 = LU_CLOCK();
                refill(s - 2);
                cp_async_commit();
                const long long t1 = LU_CLOCK();
                cp_async_wait<CI - 2>();
                const long long t2 = LU_CLOCK();
                tw += t1 - t0; tq += t2 - t1;
                __syncthreads();
            }
            if (a.dbg && st == 0) { a.dbg[16 * (K * a.NJ + J) + 9] = tw; a.dbg[16 * (K * a.NJ + J) + 10] = tq; }
            __syncthreads();
            cp_async_wait<0>();
            return;
        }
        // ---- write-back + mailbox role: 64 threads
        auto writeback = [&](int t) {   // upper sweep: Newton update mjr_newton[_SA].for applied here, off the compute warps
            const Ln o = line_of(t);
            const int c = o.m - 1, ii = c * CI + cell;
            if (!(o.ok && c >= 0 && ii < bi)) return;
            const int l = o.lrow + sg * ii;
            const double *src = ring + ((size_t)((c & 1) * A) * L + o.lam) * CI + cell;
#define SRC(arr) src[(size_t)(arr) * L * CI]
            if (!UPPER || !a.newton) {
#pragma unroll
                for (int e = 0; e < NEQ; e++) a.dw[l + e * nd] = SRC(C::A_DW + e);
            } else {
                const int i = a.box[1] - ii, j = a.box[3] - o.jjg, k = a.box[5] - kkg;
                if (i >= a.nbx[0] && i <= a.nbx[1] && j >= a.nbx[2] && j <= a.nbx[3] && k >= a.nbx[4] && k <= a.nbx[5]) {
                    const double ro = SRC(C::A_ST), u = SRC(C::A_ST + 1), v = SRC(C::A_ST + 2), w = SRC(C::A_ST + 3), t_ = SRC(C::A_ST + 4);
                    const double cvinv = 1. / z.cv, cvinv2 = 0.5 * cvinv;
                    const double rn = ro + SRC(C::A_DW);
                    const double r_1 = lu_rcp(rn);
                    const double un = (ro * u + SRC(C::A_DW + 1)) * r_1, vn = (ro * v + SRC(C::A_DW + 2)) * r_1, wn = (ro * w + SRC(C::A_DW + 3)) * r_1;
                    const double roe_old = ro * (z.cv * t_ + 0.5 * (u * u + v * v + w * w));
                    const double tn = (roe_old + SRC(C::A_DW + 4)) * r_1 * cvinv - cvinv2 * (un * un + vn * vn + wn * wn);
                    a.rop_out[l] = rn; a.rop_out[l + nd] = un; a.rop_out[l + 2 * nd] = vn; a.rop_out[l + 3 * nd] = wn; a.rop_out[l + 4 * nd] = tn;
                    if (NEQ == 6) {
                        double anulam = z.coesut * sqrt(tn * z.temp01) * tn * lu_rcp(tn + z.cmus1);   // = coesut sqrt(T/T0) / (1 + Cs/T)
                        anulam = anulam * r_1 * z.ratiom;
                        double tt = (ro * SRC(C::A_ST + NEQ - 1) + SRC(C::A_DW + NEQ - 1)) * r_1;
                        tt = fmin(tt, anulam);
                        a.rop_out[l + 5 * nd] = fmax(tt, 0.);
                    }
                }
            }
#undef SRC
        };
        // mailboxes: thread sw < 48 owns 16 bytes of the entry of interface line il (0..7: j interface, row kk = il;
        // 8..15: k interface, column jj = il-8).  Words needed at step X are loaded at iteration X-3 (registers carried
        // over the loop edge, no stall), parked in shared memory at X-2 and, before the barrier of step X, delivered
        // validated into mbx (polling global memory while the producer tile has not delivered); the entry is re-armed.
        // The compute warps never see a mailbox.
        const int sw = st - LU_NR;
        const int il = sw / 3, part = sw % 3;
        const bool isj = il < 8;
        const int mjj = isj ? 0 : il - 8, mkk = isj ? il : 0;
        const int mjg = J * TJ + mjj, mkg = K * TK + mkk;
        const bool mvalid = sw < 48 && mjg < bj && mkg < bk && mkk < TK && (isj ? J > 0 : K > 0);
        volatile unsigned long long *mrow = (volatile unsigned long long *)a.order;   // harmless readable dummy
        if (mvalid) mrow = (unsigned long long *)(isj ? a.ifj + ((size_t)((J - 1) * bk + mkg) * bi) * LU_MB
                                                        : a.ifk + ((size_t)((K - 1) * bj + mjg) * bi) * LU_MB) + 2 * part;
        const int msig = mjj + mkk;
        unsigned long long *mpre = mbs + (size_t)il * EX + 2 * part;        // + (step & 3) * 16 * EX
        double *mout = mbx + (size_t)il * EX + 2 * part;                     // + (step & 1) * 16 * EX
        auto entry = [&](int X) {    // clamped: always a readable address
            int ii = X - msig; ii = ii < 0 ? 0 : (ii >= bi ? bi - 1 : ii);
            return mrow + (mvalid ? (size_t)ii * LU_MB : 0);
        };
        auto needed = [&](int X) { const int ii = X - msig; return mvalid && ii >= 0 && ii < bi; };
        auto deliver = [&](int X) {
            if (!needed(X)) return;
            const unsigned long long *pre = mpre + (size_t)(X & 3) * 16 * EX;
            unsigned long long v0 = pre[0], v1 = pre[1];
            volatile unsigned long long *p = entry(X);
            int spin = 0;
            while (mb_empty(v0) || mb_empty(v1)) {
                v0 = p[0]; v1 = p[1];
                if (++spin > LU_SPIN_MAX || (((spin & 1023) == 0) && *(volatile int *)a.err)) { *a.err = 1; break; }
            }
            double *o = mout + (size_t)(X & 1) * 16 * EX;
            o[0] = __longlong_as_double((long long)v0); o[1] = __longlong_as_double((long long)v1);
            p[0] = LU_EMPTY; p[1] = LU_EMPTY;                                  // self-cleaning: ready for the next sweep
        };
        const bool mth = sw < 48;                                            // threads 48..63 only write back
        for (int X = 0; X < 2; X++) {
            volatile unsigned long long *p = entry(X);
            unsigned long long *d = mpre + (size_t)(X & 3) * 16 * EX;
            if (mth) { d[0] = p[0]; d[1] = p[1]; }
        }
        unsigned long long w0, w1;
        { volatile unsigned long long *p = entry(2); w0 = p[0]; w1 = p[1]; }
        long long twb = 0, tdl = 0, tpk = 0;
        for (int s = 0; s < nst; s++) {
            const long long t0 = LU_CLOCK();
            if (mth) { unsigned long long *d = mpre + (size_t)((s + 2) & 3) * 16 * EX; d[0] = w0; d[1] = w1; }
            { volatile unsigned long long *p = entry(s + 3); w0 = p[0]; w1 = p[1]; }
            const long long t1 = LU_CLOCK();
            writeback(s - 1);
            const long long t2 = LU_CLOCK();
            deliver(s);
            const long long t3 = LU_CLOCK();
            tpk += t1 - t0; twb += t2 - t1; tdl += t3 - t2;
            __syncthreads();
        }
        if (a.dbg && sw == 0) { a.dbg[16 * (K * a.NJ + J) + 11] = tpk; a.dbg[16 * (K * a.NJ + J) + 12] = twb; a.dbg[16 * (K * a.NJ + J) + 13] = tdl; }
        __syncthreads();
        for (int s = nst; s <= nst + CI; s++) writeback(s - 1);
        return;
    }

    // ---------------------------------------------------------------------------------------- compute warps
    // 4 lanes per line: lane r < 3 evaluates the product of direction r, all lanes redo the (cheap) dW* of the line, so
    // that the three Jacobian products run on different lanes of the same instruction stream and the 8 compute warps
    // keep two warps per scheduler on the FP64 pipe.  The step is branch-free: products from outside the sweep box are
    // zeros, products from neighbour tiles arrive validated in mbx (service warps).
    const int lam = tid >> 2, r = tid & 3;
    const int jj = lam % TJ, kk = lam / TJ;
    const int jjg = J * TJ + jj, kkg = K * TK + kk;
    const bool lvalid = lam < L && jjg < bj && kkg < bk;
    const int sigma = jj + kk;
    const bool has_j = jjg > 0, has_k = kkg > 0;
    volatile unsigned long long *mb_w = nullptr;
    if (lvalid && r == 1 && jj == TJ - 1 && J + 1 < a.NJ) mb_w = (unsigned long long *)(a.ifj + ((size_t)(J * bk + kkg) * bi) * LU_MB);
    if (lvalid && r == 2 && kk == TK - 1 && K + 1 < a.NK) mb_w = (unsigned long long *)(a.ifk + ((size_t)(K * bj + jjg) * bi) * LU_MB);
    // Newton box test of the own line (upper, aliased state only)
    const int jg = UPPER ? a.box[3] - jjg : a.box[2] + jjg, kg = UPPER ? a.box[5] - kkg : a.box[4] + kkg;
    const bool line_in_nb = jg >= a.nbx[2] && jg <= a.nbx[3] && kg >= a.nbx[4] && kg <= a.nbx[5];
    const double signe = UPPER ? -0.5 : 0.5;
    const double gam = z.gam, cp = z.gam * z.cv;
    const int rr = r < 3 ? r : 2;                      // lane 3 repeats the k product (not stored) and writes dW back
    // where the j / k products come from: neighbour line of the tile (exb) or neighbour tile (mbx); offsets in doubles
    // relative to the buffer of the step parity
    const int offj = jj > 0 ? ((lam - 1) * 3 + 1) * EX : -1, offk = kk > 0 ? ((lam - TJ) * 3 + 2) * EX : -1;
    const int mofj = kk * EX, mofk = (8 + jj) * EX;
    if (a.dbg && tid == 0) a.dbg[16 * (K * a.NJ + J) + 1] = gtime();
    long long tcomp = 0;
    for (int s = 0; s < nst; s++) {
        __syncthreads();
        if (a.dbg && tid == 0 && (s == 16 || s == 128 || s == 256 || s == 384)) a.dbg[16 * (K * a.NJ + J) + 4 + (s >> 7)] = gtime();
        const int ii = s - sigma;
        if (!(lvalid && ii >= 0 && ii < bi)) continue;
        const long long tc0 = LU_CLOCK();
        const int c = ii / CI, slot = ii % CI;
        double *R = ring + ((size_t)((c & 1) * A) * L + lam) * CI + slot;
#define RING(arr) R[(size_t)(arr) * L * CI]
        double acc[NEQ];
#pragma unroll
        for (int e = 0; e < NEQ; e++) acc[e] = RING(C::A_DW + e);
        const double coe1 = RING(C::A_COE), coe5 = RING(C::A_COE + 4);
        double coe6 = 1.;
        if (NEQ == 6) coe6 = RING(C::A_COE + 5);
        // state, mean normal and spectral radius of the lane's direction
        double ro = 0., u, v, w, t;
        if (UPPER) { ro = RING(C::A_ST); u = RING(C::A_ST + 1); v = RING(C::A_ST + 2); w = RING(C::A_ST + 3); t = RING(C::A_ST + 4); }
        else { u = RING(C::A_ST); v = RING(C::A_ST + 1); w = RING(C::A_ST + 2); t = RING(C::A_ST + 3); }
        // every lane loads all normals / radii (broadcast, conflict-free) and keeps those of its direction
        double tcx = 0., tcy = 0., tcz = 0.;
        if (MESH == 0) {
            const double a0 = RING(C::A_T), a1 = RING(C::A_T + 1), a2 = RING(C::A_T + 2), b0 = RING(C::A_T + 3), b1 = RING(C::A_T + 4),
                         b2 = RING(C::A_T + 5), c0 = RING(C::A_T + 6), c1 = RING(C::A_T + 7), c2 = RING(C::A_T + 8);
            tcx = rr == 0 ? a0 : (rr == 1 ? b0 : c0); tcy = rr == 0 ? a1 : (rr == 1 ? b1 : c1); tcz = rr == 0 ? a2 : (rr == 1 ? b2 : c2);
        } else if (MESH == 1) {
            const double a0 = RING(C::A_T), a1 = RING(C::A_T + 1), b0 = RING(C::A_T + 2), b1 = RING(C::A_T + 3), c2 = RING(C::A_T + 4);
            tcx = rr == 0 ? a0 : (rr == 1 ? b0 : 0.); tcy = rr == 0 ? a1 : (rr == 1 ? b1 : 0.); tcz = rr == 2 ? c2 : 0.;
        } else { tcx = rr == 0 ? z.tc0[0] : 0.; tcy = rr == 1 ? z.tc0[1] : 0.; tcz = rr == 2 ? z.tc0[2] : 0.; }
        const double cd0 = RING(C::A_COE + 1), cd1 = RING(C::A_COE + 2), cd2 = RING(C::A_COE + 3);
        const double cd = (rr == 0 ? cd0 : (rr == 1 ? cd1 : cd2)) * signe;
        // incoming products; the sum follows the reference's k, j, i order (invlu_l.for / invlu_u.for); the overlap and
        // 3dcart variants only permute the three additions (a last-bit effect, inside the stated 1e-10 tolerance)
        const double *xin = exb + (size_t)((s + 1) & 1) * 3 * L * EX;
        const double *xmb = mbx + (size_t)(s & 1) * 16 * EX;
        const double2 *pi_ = (const double2 *)(xin + (size_t)lam * 3 * EX);
        const double2 *pj_ = (const double2 *)(offj >= 0 ? xin + offj : xmb + mofj);
        const double2 *pk_ = (const double2 *)(offk >= 0 ? xin + offk : xmb + mofk);
        const double2 i0 = pi_[0], i1 = pi_[1], i2 = pi_[2], j0 = pj_[0], j1 = pj_[1], j2 = pj_[2], k0 = pk_[0], k1 = pk_[1], k2 = pk_[2];
        const double xal = coe1 * signe;
        const double xi = ii > 0 ? xal : 0., xj = has_j ? xal : 0., xk = has_k ? xal : 0.;
        if (UPPER) {
#pragma unroll
            for (int e = 0; e < 5; e++) acc[e] = acc[e] * coe5;
            if (NEQ == 6) acc[NEQ - 1] = acc[NEQ - 1] * coe6;
        }
        acc[0] = ((acc[0] + k0.x * xk) + j0.x * xj) + i0.x * xi;
        acc[1] = ((acc[1] + k0.y * xk) + j0.y * xj) + i0.y * xi;
        acc[2] = ((acc[2] + k1.x * xk) + j1.x * xj) + i1.x * xi;
        acc[3] = ((acc[3] + k1.y * xk) + j1.y * xj) + i1.y * xi;
        acc[4] = ((acc[4] + k2.x * xk) + j2.x * xj) + i2.x * xi;
        if (NEQ == 6) acc[NEQ - 1] = ((acc[NEQ - 1] + k2.y * xk) + j2.y * xj) + i2.y * xi;
        const double diag = lu_rcp(coe5);
        double dwn[NEQ];
#pragma unroll
        for (int e = 0; e < 5; e++) dwn[e] = acc[e] * diag;
        if (NEQ == 6) dwn[NEQ - 1] = acc[NEQ - 1] * lu_rcp(coe6);
        if (r == 3) {
#pragma unroll
            for (int e = 0; e < NEQ; e++) RING(C::A_DW + e) = dwn[e];
        }
        // with rop aliased to rop_out the reference has already Newton-updated plane k+1 when it sweeps plane k
        // (invlu_u.for), i.e. the k product of a cell uses its UPDATED state
        if (UPPER && a.alias) {
            const int ig = a.box[1] - ii;
            const bool in_nb = line_in_nb && ig >= a.nbx[0] && ig <= a.nbx[1] && rr == 2;
            const double cvinv = 1. / z.cv, cvinv2 = 0.5 * cvinv;
            const double r_1 = lu_rcp(ro + dwn[0]);
            const double un = (ro * u + dwn[1]) * r_1, vn = (ro * v + dwn[2]) * r_1, wn = (ro * w + dwn[3]) * r_1;
            const double roe_old = ro * (z.cv * t + 0.5 * (u * u + v * v + w * w));
            const double tn = (roe_old + dwn[4]) * r_1 * cvinv - cvinv2 * (un * un + vn * vn + wn * wn);
            u = in_nb ? un : u; v = in_nb ? vn : v; w = in_nb ? wn : w; t = in_nb ? tn : t;
        }
        // outgoing product of the lane's direction
        double b[EX];
        b[EX - 1] = 0.;
        lu_prod<NEQ>(tcx, tcy, tcz, u, v, w, t, cd, gam, cp, dwn, b);
        if (r < 3) {
            double2 *p = (double2 *)(exb + ((size_t)((s & 1) * L + lam) * 3 + r) * EX);
            p[0] = make_double2(b[0], b[1]); p[1] = make_double2(b[2], b[3]); p[2] = make_double2(b[4], b[5]);
            if (mb_w) {
                volatile unsigned long long *o = mb_w + (size_t)ii * LU_MB;
#pragma unroll
                for (int e = 0; e < LU_MB; e++) o[e] = (unsigned long long)__double_as_longlong(b[e]);
            }
        }
#undef RING
        tcomp += LU_CLOCK() - tc0;
    }
    __syncthreads();
    if (a.dbg && tid == 0) { a.dbg[16 * (K * a.NJ + J) + 2] = gtime(); a.dbg[16 * (K * a.NJ + J) + 8] = tcomp; }
    if (a.dbg && tid == 252) a.dbg[16 * (K * a.NJ + J) + 14] = tcomp;
}

template <int MESH, int NEQ, bool UPPER>
__global__ void __launch_bounds__(LU_NT, 1) k_lu_tile(ZoneDev z, LuTileArgs a)
{
    extern __shared__ double smem[];
    __shared__ int s_tile;
    if (threadIdx.x == 0) s_tile = a.order[atomicAdd(a.ticket, 1)];
    __syncthreads();
    lu_tile_body<MESH, NEQ, UPPER>(z, a, s_tile, smem);
}

// Multi-zone solvers (C4 / C5: several blocks per GPU): the sweeps of all zones in ONE launch.  A zone on its own leaves
// most SMs idle -- its tiles trail each other by ~12 steps per (J,K) hop, so a 128^3 block (256 tiles of 144 steps) spends
// most of a sweep ramping up and down -- and, with one CTA per SM, tiles of a second launch cannot become resident while
// the first launch's tiles sit waiting.  Here tickets run over the tiles of ALL zones ordered by J+K, zones interleaved,
// so that every zone's wavefront advances on its share of the SMs; the zone descriptor and the sweep arguments of the
// tile's zone are copied from a device table into shared memory.  Dependencies stay inside a zone (invlu_l.for:79-96).
struct LuBatchItem { ZoneDev z; LuTileArgs a; };
template <int MESH, int NEQ, bool UPPER>
__global__ void __launch_bounds__(LU_NT, 1) k_lu_tile_b(const LuBatchItem *items, const int2 *order, int *ticket)
{
    extern __shared__ double smem[];
    __shared__ int2 s_t;
    __shared__ __align__(16) LuBatchItem s_it;
    if (threadIdx.x == 0) s_t = order[atomicAdd(ticket, 1)];
    __syncthreads();
    static_assert(sizeof(LuBatchItem) % 8 == 0, "LuBatchItem is copied in 8-byte words");
    const unsigned long long *src = (const unsigned long long *)(items + s_t.x);
    for (int i = threadIdx.x; i < (int)(sizeof(LuBatchItem) / 8); i += LU_NT) ((unsigned long long *)&s_it)[i] = src[i];
    __syncthreads();
    lu_tile_body<MESH, NEQ, UPPER>(s_it.z, s_it.a, s_t.y, smem);
}

// ================================================================================================ k_lu_cell
// Thread-per-cell variant of the same tile-pipelined wavefront, staged by TMA (default whenever the zone's strides can be
// encoded: ni and NDIMDX even; k_lu_tile above stays the 8-byte cp.async path for the other zones).
// What changed against k_lu_tile, and why (profiles/r01_summary.md: the four-lanes-per-line step was a chain of 207 / 255
// instructions per warp at 5.5 - 8 cycles each, every lane redoing dW*, the Newton update evaluated by all four lanes AND
// again by the write-back role; profiles/r02_lu_cell_v1.md: a first thread-per-cell version fed by 16-byte cp.async spent
// 2200 cycles per step in its service warps, 23 scattered LDGSTS per line):
//   * ONE thread owns a line: it gathers the three incoming products, forms dW, evaluates the three outgoing products,
//     keeps B_i.dW in registers for its own next cell and, in the upper sweep, applies the Newton update once -- the updated
//     state feeds both the k product (aliased sweeps) and the write-back.  ~4x fewer instructions per cell.
//   * a tile is 8 x 16 lines (8 x 12 for 3dfull SA) and advances by SUPER-STEPS of two cells per line (one 16-byte granule):
//     256 cells per barrier and CTA instead of 64; the two cells of a thread are consecutive along i, so only B_i.dW of the
//     first feeds the second and the rest of their work overlaps.
//   * staging is ONE TMA box per staged array group and super-step.  The tensor maps describe each zone array through a
//     SKEWED view: element (c0, c1, c2, e) -> array[e][c0 + c1 (ni - 2) + c2 (ni nj - 2)], i.e. c0 = i + 2 j + 2 k.  In that
//     view the granules needed by the lines (jj, kk) of a tile at super-step S -- cells i = 2 (S - jj - kk) + {0, 1} of
//     rows that are ni apart -- form a dense (2, 8, TK, nvar) box at c0 = 2 S + const: the skew of the wavefront is folded
//     into the strides, nothing is fetched twice and no load instruction is issued by any thread (the four-lanes kernel
//     needs 23 cp.async per cell).  The view overlaps itself (stride of c1 < extent of c0), which cuTensorMapEncodeTiled
//     accepts; coordinates outside [0, ni + 2 nj + 2 nk) x [0, nj) x [0, nk) are zero-filled and the largest address of the
//     view is the last cell of the array (tools/probe/tma_skew_probe.cu checks all of this on the device).  Box starts must
//     be 16-byte aligned, so the pairing of cells into granules follows the parity of the first column of the sweep box.
//     The k-invariant normals of 3dhomo zones come through a rank-3 view (c0 = i + 2 j) into their own ring of 20 super-steps:
//     line kk needs at super-step S what line 0 needed at S - kk.
//   * three slots (LC_SLOTS_MAX) of (all arrays) x (lines) x (2 cells): compute works on slot S, the service threads write
//     the results of S-1 back (16-byte stores, masked per cell) and then re-arm that slot for S + slots - 1.
// Mailboxes (one 48-byte entry per boundary cell, two per super-step), tickets, error flag and the (J + K) tile order are those of
// k_lu_tile.
#ifndef LC_SLOTS_MAX
#define LC_SLOTS_MAX 3     // slots of the staging ring; boxes are asked for SLOTS - 1 super-steps ahead.  4 (where they fit in 227 KB:
                           // all but the 3dfull 5-equation zones) measured no faster (C5 598 v 600, plate 396 v 398 Mcell.it/s): the
                           // sweep does not wait for its boxes
#endif
#define LC_NSL 20          // super-steps kept in the 3dhomo normals ring (>= TK + slots)
#define LC_NSV 128         // service threads

template <int MESH, int NEQ, bool UPPER>
struct LcCfg {
    static constexpr int TJ = 8;
    static constexpr int TK = (MESH == 0 && NEQ == 6) ? 12 : 16;
    static constexpr int L = TJ * TK;
    static constexpr int NC = NEQ == 6 ? 6 : 5;
    static constexpr int NS = UPPER ? NEQ : 4;
    static constexpr int S0 = UPPER ? 0 : 1;
    static constexpr int NT = MESH == 0 ? 9 : 0;        // normals staged in the slot (3dfull); 3dhomo: separate ring
    static constexpr int A_DW = 0, A_COE = NEQ, A_ST = NEQ + NC, A_T = NEQ + NC + NS;
    static constexpr int A = NEQ + NC + NS + NT;
    static constexpr int NIF = TJ + TK;                 // interface lines: TK towards (J-1,K), TJ towards (J,K-1)
    static constexpr int NTHR = L + LC_NSV;
    static constexpr int AS = L * 2;                    // doubles per staged array and slot
    static constexpr int SLOT = A * AS;
    static constexpr int NRS = 5 * TJ * 2;              // doubles per super-step of the 3dhomo normals ring
    static constexpr int NRING = MESH == 1 ? LC_NSL * NRS : 0;
    static constexpr int EXB = 2 * 2 * 2 * L * LU_MB;   // [buffer][direction][cell][line][6]
    static constexpr int MBW = 2 * LU_MB;               // words per interface line and super-step
    static constexpr unsigned TX = (unsigned)((SLOT + (MESH == 1 ? NRS : 0)) * sizeof(double));
    static constexpr size_t smem_of(int slots) { return (size_t)(slots * SLOT + NRING + EXB + 6 * NIF * MBW) * sizeof(double) + 64; }
    static constexpr int SLOTS = smem_of(LC_SLOTS_MAX) <= 227 * 1024 ? LC_SLOTS_MAX : 3;
    static constexpr size_t SMEM = smem_of(SLOTS);
};
int lu_cell_tk(int mesh, int neq) { return (mesh == 0 && neq == 6) ? 12 : 16; }
long long g_lu_cell_launches = 0;   // sweeps run by the TMA-staged thread-per-cell kernel (fb_counter "lu_cell")

__device__ __forceinline__ void lc_mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void lc_mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lc_mbar_wait(unsigned long long *bar, unsigned parity)
{
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(b), "r"(parity) : "memory");
}
__device__ __forceinline__ void lc_tma4(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void lc_tma3(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <int MESH, int NEQ, bool UPPER>
__device__ __forceinline__ void lu_cell_body(const ZoneDev &z, const LuTileArgs &a, const int tile, double *smem)
{
    using C = LcCfg<MESH, NEQ, UPPER>;
    constexpr int TJ = C::TJ, TK = C::TK, L = C::L, EX = LU_MB, NIF = C::NIF, AS = C::AS, MBW = C::MBW, LC_SLOTS = C::SLOTS;
    double *ring = smem;                                   // [LC_SLOTS][A][L][2]
    double *nring = smem + LC_SLOTS * C::SLOT;             // 3dhomo: [LC_NSL][5][TJ][2]
    double *exb = nring + C::NRING;                        // [2][2][2][L][EX]: products handed to (jj+1) and (kk+1)
    unsigned long long *mbs = (unsigned long long *)(exb + C::EXB);   // [4][NIF][MBW] prefetched mailbox words
    double *mbx = (double *)(mbs + 4 * NIF * MBW);         // [2][NIF][MBW] validated products of the neighbour tiles
    unsigned long long *full = (unsigned long long *)(mbx + 2 * NIF * MBW);   // [LC_SLOTS] "slot has landed"
    const int tid = threadIdx.x;
    for (int i = tid; i < C::EXB + 6 * NIF * MBW; i += C::NTHR) exb[i] = 0.;
    if (tid == 0) {
        for (int q = 0; q < LC_SLOTS; q++) lc_mbar_init(full + q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int J = tile & 0xffff, K = tile >> 16;
    const int bi = a.bi, bj = a.bj, bk = a.bk;
    const int nd = z.nd;
    const int sg = UPPER ? -1 : 1;
    const int l0 = UPPER ? inddm(z, a.box[1], a.box[3], a.box[5]) : inddm(z, a.box[0], a.box[2], a.box[4]);
    // pairing of the cells of a line into 16-byte granules: super-step m of a line holds cells 2m - p and 2m - p + 1
    const int ir = (UPPER ? a.box[1] : a.box[0]) + z.nf - 1;            // array column of the first cell of the sweep
    const int p = UPPER ? ((ir + 1) & 1) : (ir & 1);
    const int nm = (bi - 1 + p) / 2 + 1;                                 // super-steps of a line
    const int nst = nm + TJ + TK - 2;

    if (tid >= L) {
        // ------------------------------------------------------------------------------------ service warps
        const int st = tid - L;
        const int lam = st, jj = lam % TJ, kk = lam / TJ;
        const int jjg = J * TJ + jj, kkg = K * TK + kk;
        const bool lok = lam < L && jjg < bj && kkg < bk;
        const int sigma = jj + kk;
        const int lam_s = UPPER ? L - 1 - lam : lam;                    // line index inside the TMA box
        const int lrow = l0 + sg * (jjg * z.incj + kkg * z.inck);
        const int jg = UPPER ? a.box[3] - jjg : a.box[2] + jjg, kg = UPPER ? a.box[5] - kkg : a.box[4] + kkg;
        const bool line_in_nb = jg >= a.nbx[2] && jg <= a.nbx[3] && kg >= a.nbx[4] && kg <= a.nbx[5];
        // box origin of the tile in the skewed views (see the header): c0 = cbase + 2 S (lower) / cbase - 2 S (upper)
        const int jr = (UPPER ? a.box[3] : a.box[2]) + z.nf - 1, kr = (UPPER ? a.box[5] : a.box[4]) + z.nfk - 1;
        const int c1 = UPPER ? jr - TJ * J - (TJ - 1) : jr + TJ * J;
        const int c2 = UPPER ? kr - TK * K - (TK - 1) : kr + TK * K;
        const int cn = UPPER ? ir + p - 1 + 2 * (jr - TJ * J) : ir - p + 2 * (jr + TJ * J);       // rank-3 view of the 3dhomo normals
        const int cb = cn + (UPPER ? 2 * (kr - TK * K) : 2 * (kr + TK * K));
        auto fetch = [&](int S) {       // one thread: every staged array of super-step S
            if (S >= nst) return;
            const int q = S % LC_SLOTS;
            double *dst = ring + (size_t)q * C::SLOT;
            const int c0 = cb + sg * 2 * S;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the slot was last written through the generic proxy
            lc_mbar_expect_tx(full + q, C::TX);
            lc_tma4(dst + (size_t)C::A_DW * AS, a.tm + 0, c0, c1, c2, 0, full + q);
            lc_tma4(dst + (size_t)C::A_COE * AS, a.tm + 1, c0, c1, c2, 0, full + q);
            lc_tma4(dst + (size_t)C::A_ST * AS, a.tm_rop + (UPPER ? 1 : 0), c0, c1, c2, C::S0, full + q);
            if (MESH == 0) lc_tma4(dst + (size_t)C::A_T * AS, a.tm + 2, c0, c1, c2, 0, full + q);
            if (MESH == 1) lc_tma3(nring + (size_t)(S % LC_NSL) * C::NRS, a.tm + 2, cn + sg * 2 * S, c1, 0, full + q);
        };
        // results of super-step S: one 16-byte store per array when both cells of the granule are to be written
        auto writeback = [&](int S) {
            const int m = S - sigma;
            if (!(lok && m >= 0 && m < nm)) return;
            const int iiA = 2 * m - p, iiB = iiA + 1;
            const bool nwt = UPPER && a.newton;
            bool vA = iiA >= 0, vB = iiB < bi;
            if (nwt) {
                const int ig = a.box[1] - iiA;
                vA = vA && line_in_nb && ig >= a.nbx[0] && ig <= a.nbx[1];
                vB = vB && line_in_nb && ig - 1 >= a.nbx[0] && ig - 1 <= a.nbx[1];
            }
            if (!vA && !vB) return;
            const double *src = ring + (size_t)(S % LC_SLOTS) * C::SLOT + (size_t)lam_s * 2;
            double *out = nwt ? a.rop_out : a.dw;
            const int arr0 = nwt ? C::A_ST : C::A_DW;
            const int g0 = UPPER ? lrow - iiB : lrow + iiA;              // even cell address: start of the granule
            if (vA && vB) {
#pragma unroll
                for (int e = 0; e < NEQ; e++)
                    *(double2 *)(out + g0 + (size_t)e * nd) = *(const double2 *)(src + (size_t)(arr0 + e) * AS);
            } else {
                const int pos = vA ? (UPPER ? 1 : 0) : (UPPER ? 0 : 1);
#pragma unroll
                for (int e = 0; e < NEQ; e++) out[g0 + pos + (size_t)e * nd] = src[(size_t)(arr0 + e) * AS + pos];
            }
        };
        // -- mailboxes: thread st < 3 NIF owns 16 bytes of the two entries (cells A and B) of interface line il (0..TK-1: j
        // interface, row kk = il; TK..: k interface, column jj = il - TK) at every super-step.  Words needed at super-step X
        // are loaded at iteration X-3 (registers carried over the loop edge), parked in shared memory at X-2 and delivered
        // validated into mbx at X-1 (polling global memory while the producer tile has not delivered); the entries are re-armed.
        const int il = st / 3, part = st % 3;
        const bool isj = il < TK;
        const int mjj = isj ? 0 : il - TK, mkk = isj ? il : 0;
        const int mjg = J * TJ + mjj, mkg = K * TK + mkk;
        const bool mth = st < 3 * NIF;
        const bool mvalid = mth && mjg < bj && mkg < bk && (isj ? J > 0 : K > 0);
        volatile unsigned long long *mrow = (volatile unsigned long long *)a.order;   // harmless readable dummy
        if (mvalid) mrow = (unsigned long long *)(isj ? a.ifj + ((size_t)((J - 1) * bk + mkg) * bi) * LU_MB
                                                        : a.ifk + ((size_t)((K - 1) * bj + mjg) * bi) * LU_MB) + 2 * part;
        const int msig = mjj + mkk;
        unsigned long long *mpre = mbs + (size_t)(mth ? il : 0) * MBW + 2 * part;     // + (X & 3) * NIF * MBW; cell B: + EX
        double *mout = mbx + (size_t)(mth ? il : 0) * MBW + 2 * part;                  // + (X & 1) * NIF * MBW
        auto cell_of = [&](int X, int c) { return 2 * (X - msig) - p + c; };
        auto entry = [&](int ii) {    // clamped: always a readable address
            ii = ii < 0 ? 0 : (ii >= bi ? bi - 1 : ii);
            return mrow + (mvalid ? (size_t)ii * LU_MB : 0);
        };
        auto deliver = [&](int X) {
            if (!mvalid) return;
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int ii = cell_of(X, c);
                if (ii < 0 || ii >= bi) continue;
                const unsigned long long *pre = mpre + (size_t)(X & 3) * NIF * MBW + c * EX;
                unsigned long long v0 = pre[0], v1 = pre[1];
                volatile unsigned long long *q = entry(ii);
                int spin = 0;
                while (mb_empty(v0) || mb_empty(v1)) {
                    v0 = q[0]; v1 = q[1];
                    if (++spin > LU_SPIN_MAX || (((spin & 1023) == 0) && *(volatile int *)a.err)) { *a.err = 1; break; }
                }
                double *o = mout + (size_t)(X & 1) * NIF * MBW + c * EX;
                o[0] = __longlong_as_double((long long)v0); o[1] = __longlong_as_double((long long)v1);
                q[0] = LU_EMPTY; q[1] = LU_EMPTY;                              // self-cleaning: ready for the next sweep
            }
        };
        unsigned long long wa0 = 0, wa1 = 0, wb0 = 0, wb1 = 0;
        auto load = [&](int X) {
            volatile unsigned long long *qa = entry(cell_of(X, 0)), *qb = entry(cell_of(X, 1));
            wa0 = qa[0]; wa1 = qa[1]; wb0 = qb[0]; wb1 = qb[1];
        };
        auto park = [&](int X) {
            unsigned long long *d = mpre + (size_t)(X & 3) * NIF * MBW;
            d[0] = wa0; d[1] = wa1; d[EX] = wb0; d[EX + 1] = wb1;
        };
        if (st == 0) {
            for (int q = 0; q < LC_SLOTS - 1; q++) fetch(q);
        }
        if (mth) { load(0); park(0); load(1); park(1); load(2); }
        deliver(0);
        long long c_wb = 0, c_sb = 0, c_mb = 0, c_dl = 0, c_fw = 0, c_b0 = 0;
        for (int S = 0; S < nst; S++) {
            const long long q0 = LU_CLOCK();
            __syncthreads();                                   // B_S: compute has finished super-step S-1
            const long long q1 = LU_CLOCK();
            writeback(S - 1);
            // every line of slot S-1 has been read (barrier.sync, not the .aligned bar.sync: the lanes of a service warp leave
            // writeback() on different paths and the compiler does not reconverge them in front of inline assembly)
            __syncwarp();
            const long long q2 = LU_CLOCK();
            asm volatile("barrier.sync 1, %0;" ::"n"(LC_NSV) : "memory");
            const long long q3 = LU_CLOCK();
            if (st == 0) fetch(S + LC_SLOTS - 1);          // into the slot that was just written back
            if (mth) { park(S + 2); load(S + 3); }
            const long long q4 = LU_CLOCK();
            deliver(S + 1);
            const long long q5 = LU_CLOCK();
            // every box that was asked for is seen to land before the CTA can leave (a ragged tile may have no compute thread
            // left waiting on the last slots)
            if (st == 0) lc_mbar_wait(full + S % LC_SLOTS, (unsigned)(S / LC_SLOTS) & 1u);
            const long long q6 = LU_CLOCK();
            c_b0 += q1 - q0; c_wb += q2 - q1; c_sb += q3 - q2; c_mb += q4 - q3; c_dl += q5 - q4; c_fw += q6 - q5;
        }
        if (a.dbg && (st == 0 || st == 40)) {
            unsigned long long *q = a.dbg + 16 * (K * a.NJ + J) + (st == 0 ? 4 : 10);
            q[0] = c_b0; q[1] = c_wb; q[2] = c_sb; q[3] = c_mb; q[4] = c_dl; q[5] = c_fw;
        }
        __syncthreads();
        writeback(nst - 1);
        return;
    }

    // ---------------------------------------------------------------------------------------- compute warps: one thread per line
    const int lam = tid;
    const int jj = lam % TJ, kk = lam / TJ;
    const int jjg = J * TJ + jj, kkg = K * TK + kk;
    const bool lvalid = jjg < bj && kkg < bk;
    const int sigma = jj + kk;
    const int lam_s = UPPER ? L - 1 - lam : lam;
    const bool has_j = jjg > 0, has_k = kkg > 0;
    volatile unsigned long long *mbj = nullptr, *mbk = nullptr;
    if (lvalid && jj == TJ - 1 && J + 1 < a.NJ) mbj = (unsigned long long *)(a.ifj + ((size_t)(J * bk + kkg) * bi) * LU_MB);
    if (lvalid && kk == TK - 1 && K + 1 < a.NK) mbk = (unsigned long long *)(a.ifk + ((size_t)(K * bj + jjg) * bi) * LU_MB);
    const int jg = UPPER ? a.box[3] - jjg : a.box[2] + jjg, kg = UPPER ? a.box[5] - kkg : a.box[4] + kkg;
    const bool line_in_nb = jg >= a.nbx[2] && jg <= a.nbx[3] && kg >= a.nbx[4] && kg <= a.nbx[5];
    const double signe = UPPER ? -0.5 : 0.5;
    const double gam = z.gam, cp = z.gam * z.cv;
    // incoming j / k products: neighbour line of the tile (exb, other parity) or neighbour tile (mbx); offsets of cell A
    const int offj = jj > 0 ? (lam - 1) * EX : -1, offk = kk > 0 ? (2 * L + lam - TJ) * EX : -1;
    const int mofj = kk * MBW, mofk = (TK + jj) * MBW;
    const int jj_s = UPPER ? TJ - 1 - jj : jj;
    double pi_[EX];
#pragma unroll
    for (int e = 0; e < EX; e++) pi_[e] = 0.;
    if (a.dbg && tid == 0) a.dbg[16 * (K * a.NJ + J)] = gtime();
    long long tcomp = 0, tmbar = 0;
    for (int S = 0; S < nst; S++) {
        __syncthreads();                                       // B_S
        // every compute thread observes every phase of the slot barriers, active or not: the normals ring of a 3dhomo zone is
        // read up to TK - 1 super-steps after it landed
        const long long tm0 = LU_CLOCK();
        lc_mbar_wait(full + S % LC_SLOTS, (unsigned)(S / LC_SLOTS) & 1u);
        tmbar += LU_CLOCK() - tm0;
        const int m = S - sigma;
        if (!(lvalid && m >= 0 && m < nm)) continue;
        const long long tc0 = LU_CLOCK();
        double *R0 = ring + (size_t)(S % LC_SLOTS) * C::SLOT + (size_t)lam_s * 2;
        const double *N0 = nring + (size_t)((S - kk + LC_NSL) % LC_NSL) * C::NRS + jj_s * 2;
        const double *xin = exb + (size_t)((S + 1) & 1) * 4 * L * EX;
        double *xout = exb + (size_t)(S & 1) * 4 * L * EX;
        const double *xmb = mbx + (size_t)(S & 1) * NIF * MBW;
        // both cells of the granule at once: one 16-byte load per staged array (conflict-free: 32 lines x 16 bytes), all of them
        // issued before the first dependent instruction
        double2 g[C::A];
#pragma unroll
        for (int q = 0; q < C::A; q++) g[q] = *(const double2 *)(R0 + (size_t)q * AS);
        double2 gn[5];
        if (MESH == 1) {
#pragma unroll
            for (int q = 0; q < 5; q++) gn[q] = *(const double2 *)(N0 + q * TJ * 2);
        }
        double2 pj[2][3], pk[2][3];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const double2 *pj_ = (const double2 *)(offj >= 0 ? xin + offj + c * L * EX : xmb + mofj + c * EX);
            const double2 *pk_ = (const double2 *)(offk >= 0 ? xin + offk + c * L * EX : xmb + mofk + c * EX);
            pj[c][0] = pj_[0]; pj[c][1] = pj_[1]; pj[c][2] = pj_[2]; pk[c][0] = pk_[0]; pk[c][1] = pk_[1]; pk[c][2] = pk_[2];
        }
#define GET(v, POS) ((POS) ? (v).y : (v).x)
#define PUT(v, POS, val) do { if (POS) (v).y = (val); else (v).x = (val); } while (0)
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int ii = 2 * m - p + c;
            if (ii < 0 || ii >= bi) continue;
            const int pos = UPPER ? 1 - c : c;            // compile-time after unrolling
#define RING(arr) GET(g[arr], pos)
            double acc[NEQ];
#pragma unroll
            for (int e = 0; e < NEQ; e++) acc[e] = RING(C::A_DW + e);
            const double coe1 = RING(C::A_COE), coe5 = RING(C::A_COE + 4);
            double coe6 = 1.;
            if (NEQ == 6) coe6 = RING(C::A_COE + 5);
            double ro = 0., u, v, w, t, nut = 0.;
            if (UPPER) {
                ro = RING(C::A_ST); u = RING(C::A_ST + 1); v = RING(C::A_ST + 2); w = RING(C::A_ST + 3); t = RING(C::A_ST + 4);
                if (NEQ == 6) nut = RING(C::A_ST + NEQ - 1);
            } else { u = RING(C::A_ST); v = RING(C::A_ST + 1); w = RING(C::A_ST + 2); t = RING(C::A_ST + 3); }
            double ti0 = 0., ti1 = 0., ti2 = 0., tj0 = 0., tj1 = 0., tj2 = 0., tk0 = 0., tk1 = 0., tk2 = 0.;
            if (MESH == 0) {
                ti0 = RING(C::A_T); ti1 = RING(C::A_T + 1); ti2 = RING(C::A_T + 2); tj0 = RING(C::A_T + 3); tj1 = RING(C::A_T + 4);
                tj2 = RING(C::A_T + 5); tk0 = RING(C::A_T + 6); tk1 = RING(C::A_T + 7); tk2 = RING(C::A_T + 8);
            } else if (MESH == 1) {
                ti0 = GET(gn[0], pos); ti1 = GET(gn[1], pos); tj0 = GET(gn[2], pos); tj1 = GET(gn[3], pos); tk2 = GET(gn[4], pos);
            } else { ti0 = z.tc0[0]; tj1 = z.tc0[1]; tk2 = z.tc0[2]; }
            const double cdi = RING(C::A_COE + 1) * signe, cdj = RING(C::A_COE + 2) * signe, cdk = RING(C::A_COE + 3) * signe;
            const double2 j0 = pj[c][0], j1 = pj[c][1], j2 = pj[c][2], k0 = pk[c][0], k1 = pk[c][1], k2 = pk[c][2];
            const double xal = coe1 * signe;
            const double xi = ii > 0 ? xal : 0., xj = has_j ? xal : 0., xk = has_k ? xal : 0.;
            if (UPPER) {
#pragma unroll
                for (int e = 0; e < 5; e++) acc[e] = acc[e] * coe5;
                if (NEQ == 6) acc[NEQ - 1] = acc[NEQ - 1] * coe6;
            }
            // the sum follows the reference's k, j, i order (invlu_l.for / invlu_u.for)
            acc[0] = ((acc[0] + k0.x * xk) + j0.x * xj) + pi_[0] * xi;
            acc[1] = ((acc[1] + k0.y * xk) + j0.y * xj) + pi_[1] * xi;
            acc[2] = ((acc[2] + k1.x * xk) + j1.x * xj) + pi_[2] * xi;
            acc[3] = ((acc[3] + k1.y * xk) + j1.y * xj) + pi_[3] * xi;
            acc[4] = ((acc[4] + k2.x * xk) + j2.x * xj) + pi_[4] * xi;
            if (NEQ == 6) acc[NEQ - 1] = ((acc[NEQ - 1] + k2.y * xk) + j2.y * xj) + pi_[EX - 1] * xi;
            const double diag = lu_rcp(coe5);
            double dwn[NEQ];
#pragma unroll
            for (int e = 0; e < 5; e++) dwn[e] = acc[e] * diag;
            if (NEQ == 6) dwn[NEQ - 1] = acc[NEQ - 1] * lu_rcp(coe6);
            double uk = u, vk = v, wk = w, tk_ = t;     // state of the k product
            if (UPPER && a.newton) {
                // mjr_newton[_SA].for, once per cell: the updated state goes back through the slot (write-back role) and, when
                // rop aliases rop_out, into the k product -- the reference has already updated plane k+1 when it sweeps plane k
                const int ig = a.box[1] - ii;
                const bool in_nb = line_in_nb && ig >= a.nbx[0] && ig <= a.nbx[1];
                const double cvinv = 1. / z.cv, cvinv2 = 0.5 * cvinv;
                const double rn = ro + dwn[0];
                const double r_1 = lu_rcp(rn);
                const double un = (ro * u + dwn[1]) * r_1, vn = (ro * v + dwn[2]) * r_1, wn = (ro * w + dwn[3]) * r_1;
                const double roe_old = ro * (z.cv * t + 0.5 * (u * u + v * v + w * w));
                const double tn = (roe_old + dwn[4]) * r_1 * cvinv - cvinv2 * (un * un + vn * vn + wn * wn);
                if (in_nb) {
                    PUT(g[C::A_ST], pos, rn); PUT(g[C::A_ST + 1], pos, un); PUT(g[C::A_ST + 2], pos, vn); PUT(g[C::A_ST + 3], pos, wn);
                    PUT(g[C::A_ST + 4], pos, tn);
                    if (NEQ == 6) {
                        double anulam = z.coesut * sqrt(tn * z.temp01) * tn * lu_rcp(tn + z.cmus1);   // = coesut sqrt(T/T0) / (1 + Cs/T)
                        anulam = anulam * r_1 * z.ratiom;
                        double tt = (ro * nut + dwn[NEQ - 1]) * r_1;
                        tt = fmin(tt, anulam);
                        PUT(g[C::A_ST + NEQ - 1], pos, fmax(tt, 0.));
                    }
                    if (a.alias) { uk = un; vk = vn; wk = wn; tk_ = tn; }
                }
            } else {
#pragma unroll
                for (int e = 0; e < NEQ; e++) PUT(g[C::A_DW + e], pos, dwn[e]);
            }
            // outgoing products
            double bj_[EX], bk_[EX];
            pi_[EX - 1] = 0.; bj_[EX - 1] = 0.; bk_[EX - 1] = 0.;
            lu_prod<NEQ>(ti0, ti1, ti2, u, v, w, t, cdi, gam, cp, dwn, pi_);
            lu_prod<NEQ>(tj0, tj1, tj2, u, v, w, t, cdj, gam, cp, dwn, bj_);
            lu_prod<NEQ>(tk0, tk1, tk2, uk, vk, wk, tk_, cdk, gam, cp, dwn, bk_);
            {
                double2 *q = (double2 *)(xout + (size_t)(c * L + lam) * EX);
                q[0] = make_double2(bj_[0], bj_[1]); q[1] = make_double2(bj_[2], bj_[3]); q[2] = make_double2(bj_[4], bj_[5]);
                double2 *r = (double2 *)(xout + (size_t)(2 * L + c * L + lam) * EX);
                r[0] = make_double2(bk_[0], bk_[1]); r[1] = make_double2(bk_[2], bk_[3]); r[2] = make_double2(bk_[4], bk_[5]);
            }
            if (mbj) {
                volatile unsigned long long *o = mbj + (size_t)ii * LU_MB;
#pragma unroll
                for (int e = 0; e < LU_MB; e++) o[e] = (unsigned long long)__double_as_longlong(bj_[e]);
            }
            if (mbk) {
                volatile unsigned long long *o = mbk + (size_t)ii * LU_MB;
#pragma unroll
                for (int e = 0; e < LU_MB; e++) o[e] = (unsigned long long)__double_as_longlong(bk_[e]);
            }
#undef RING
        }
        // results of both cells back into the slot (cells outside the line keep what was staged; the write-back role masks them)
        if (UPPER && a.newton) {
#pragma unroll
            for (int e = 0; e < NEQ; e++) *(double2 *)(R0 + (size_t)(C::A_ST + e) * AS) = g[C::A_ST + e];
        } else {
#pragma unroll
            for (int e = 0; e < NEQ; e++) *(double2 *)(R0 + (size_t)(C::A_DW + e) * AS) = g[C::A_DW + e];
        }
#undef GET
#undef PUT
        // the results were written into the slot through the generic proxy; the slot is re-armed by TMA (async proxy) two
        // barriers from here
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tcomp += LU_CLOCK() - tc0;
    }
    __syncthreads();
    if (a.dbg && tid == 0) { a.dbg[16 * (K * a.NJ + J) + 1] = gtime(); a.dbg[16 * (K * a.NJ + J) + 2] = tcomp; a.dbg[16 * (K * a.NJ + J) + 3] = tmbar; }
}

template <int MESH, int NEQ, bool UPPER>
__global__ void __launch_bounds__(LcCfg<MESH, NEQ, UPPER>::NTHR, 1) k_lu_cell(ZoneDev z, LuTileArgs a)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ int s_tile;
    if (threadIdx.x == 0) s_tile = a.order[atomicAdd(a.ticket, 1)];
    __syncthreads();
    lu_cell_body<MESH, NEQ, UPPER>(z, a, s_tile, smem);
}

template <int MESH, int NEQ, bool UPPER>
__global__ void __launch_bounds__(LcCfg<MESH, NEQ, UPPER>::NTHR, 1) k_lu_cell_b(const LuBatchItem *items, const int2 *order, int *ticket)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ int2 s_t;
    __shared__ __align__(16) LuBatchItem s_it;
    if (threadIdx.x == 0) s_t = order[atomicAdd(ticket, 1)];
    __syncthreads();
    const unsigned long long *src = (const unsigned long long *)(items + s_t.x);
    for (int i = threadIdx.x; i < (int)(sizeof(LuBatchItem) / 8); i += LcCfg<MESH, NEQ, UPPER>::NTHR) ((unsigned long long *)&s_it)[i] = src[i];
    __syncthreads();
    lu_cell_body<MESH, NEQ, UPPER>(s_it.z, s_it.a, s_t.y, smem);
}

// the skewed tensor maps need even strides (ni, NDIMDX); FAST_B200_LU=tile forces the 8-byte kernel
bool lu_use_cell(const ZoneDev &z)
{
    static const bool force_tile = [] { const char *e = getenv("FAST_B200_LU"); return e && e[0] == 't'; }();
    if (force_tile) return false;
    return z.ni % 2 == 0 && z.nd % 2 == 0 && z.nfk > 0;
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
bool fb_tma_encode_skew(CUtensorMap *out, const double *base, int ni, int nj, int nk, long long estride, int nvar, int box_k, int box_vars, int rank3);
int lu_cell_tk(int mesh, int neq);
bool lu_use_cell(const ZoneDev &z);

int lu_plan_build(LuPlan &p, const ZoneDev &z, const int box[6])
{
    lu_plan_free(p);
    for (int i = 0; i < 6; i++) p.box[i] = box[i];
    const int bi = box[1] - box[0] + 1, bj = box[3] - box[2] + 1, bk = box[5] - box[4] + 1;
    p.nplanes = bi + bj + bk - 2;
    p.cell = lu_use_cell(z);
    if (p.cell) {       // skewed tensor maps of the arrays whose address never changes; the rop levels follow on first use
        const int tk = lu_cell_tk(z.typ, z.neq);
        CUtensorMap h[2];
        p.cell = fb_tma_encode_skew(&h[0], z.drodm, z.ni, z.nj, z.nk, z.nd, z.neq, tk, z.neq, 0) &&
                 fb_tma_encode_skew(&h[1], z.coe, z.ni, z.nj, z.nk, z.nd, z.neq_coe, tk, z.neq_coe, 0) &&
                 cudaMalloc(&p.d_tm, 9 * sizeof(CUtensorMap)) == cudaSuccess;
        if (p.cell) cudaMemcpy(p.d_tm, h, sizeof(h), cudaMemcpyHostToDevice);
        for (int q = 0; q < 3; q++) p.rop_seen[q] = nullptr;
    }
    p.tj = LU_TJ; p.tk = p.cell ? lu_cell_tk(z.typ, z.neq) : lu_tile_tk(z.typ, z.neq);
    p.NJ = (bj + p.tj - 1) / p.tj; p.NK = (bk + p.tk - 1) / p.tk;
    std::vector<int> ord;
    for (int K = 0; K < p.NK; K++) for (int J = 0; J < p.NJ; J++) ord.push_back(J | (K << 16));
    std::stable_sort(ord.begin(), ord.end(), [](int a, int b) { return (a & 0xffff) + (a >> 16) < (b & 0xffff) + (b >> 16); });
    if (cudaMalloc(&p.d_order, (ord.size() + 16) * sizeof(int)) != cudaSuccess) return 1;   // + padding: idle mailbox threads read it as a dummy
    cudaMemset(p.d_order, 0, (ord.size() + 16) * sizeof(int));
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
    if (p.nrelax >= 2) {      // rhs of the Newton step and U-sum of the previous relaxation (the reference's ssor arrays)
        const size_t n = (size_t)z.nd * z.neq * sizeof(double);
        if (cudaMalloc(&p.d_rhs, n) != cudaSuccess || cudaMalloc(&p.d_us, n) != cudaSuccess) return 1;
        cudaMemset(p.d_us, 0, n);
    }
    if (z.typ != 2) {
        p.tstride = z.typ == 0 ? (long long)z.nd : (long long)z.ni * z.nj;
        const int nt = z.typ == 0 ? 9 : 5;
        if (cudaMalloc(&p.d_tcm, (size_t)nt * p.tstride * sizeof(double)) != cudaSuccess) return 1;
        cudaMemset(p.d_tcm, 0, (size_t)nt * p.tstride * sizeof(double));
        if (p.cell) {
            CUtensorMap h;
            const bool ok = z.typ == 0 ? fb_tma_encode_skew(&h, p.d_tcm, z.ni, z.nj, z.nk, p.tstride, 9, p.tk, 9, 0)
                                       : fb_tma_encode_skew(&h, p.d_tcm, z.ni, z.nj, z.nk, p.tstride, 5, p.tk, 5, 1);
            if (!ok) return 1;
            cudaMemcpy((CUtensorMap *)p.d_tm + 2, &h, sizeof(h), cudaMemcpyHostToDevice);
        }
    }
    // the fills above ran on the legacy default stream; the solver's streams are non-blocking
    return (cudaDeviceSynchronize() == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : 1;
}

void lu_plan_free(LuPlan &p)
{
    cudaFree(p.d_order); cudaFree(p.d_ctl); cudaFree(p.d_mb); cudaFree(p.d_tcm); cudaFree(p.d_rhs); cudaFree(p.d_us); cudaFree(p.d_tm);
    p.d_tm = nullptr; p.d_order = nullptr; p.d_ctl = nullptr; p.d_mb = nullptr; p.d_tcm = nullptr; p.d_rhs = nullptr; p.d_us = nullptr;
}

void lu_plan_metrics(LuPlan &p, const ZoneDev &z, cudaStream_t st)
{
    if (z.typ == 2 || !p.d_tcm) return;
    dim3 b(64, 4, 1), g((z.iv + 2 * z.nf + 63) / 64, (z.jv + 2 * z.nf + 3) / 4, z.typ == 0 ? z.kv + 2 * z.nfk : 1);
    if (z.typ == 0) k_lunorm<0><<<g, b, 0, st>>>(z, p.d_tcm, p.tstride);
    else k_lunorm<1><<<g, b, 0, st>>>(z, p.d_tcm, p.tstride);
}

static void lu_fill_args(LuTileArgs &a, const ZoneDev &z, const LuPlan &p, const double *rop, double *rop_out, int lu_match)
{
    memset(&a, 0, sizeof(a));
    a.rop = rop; a.rop_out = rop_out; a.dw = z.drodm; a.coe = z.coe; a.tcm = p.d_tcm; a.tstride = p.tstride;
    a.ifj = p.d_mb; a.ifk = p.d_mb ? p.d_mb + p.n_ifj : nullptr;
    a.order = p.d_order; a.err = p.d_ctl + 2;
    for (int i = 0; i < 6; i++) { a.box[i] = p.box[i]; a.nbx[i] = p.box[i]; }
    if (lu_match) { a.nbx[0] = 1; a.nbx[1] = z.iv; a.nbx[2] = 1; a.nbx[3] = z.jv; a.nbx[4] = 1; a.nbx[5] = z.kv; }
    a.bi = p.box[1] - p.box[0] + 1; a.bj = p.box[3] - p.box[2] + 1; a.bk = p.box[5] - p.box[4] + 1;
    a.NJ = p.NJ; a.NK = p.NK; a.alias = (rop == rop_out) ? 1 : 0; a.plane = z.ni * z.nj; a.newton = 1;
    if (p.cell) {
        a.tm = (const CUtensorMap *)p.d_tm;
        int q = 0;
        while (q < 3 && p.rop_seen[q] != rop && p.rop_seen[q] != nullptr) q++;
        if (q == 3) q = 0;                                   // (cannot happen: a zone has three time levels)
        if (p.rop_seen[q] != rop) {
            CUtensorMap h[2];
            fb_tma_encode_skew(&h[0], rop, z.ni, z.nj, z.nk, z.nd, z.neq, p.tk, 4, 0);
            fb_tma_encode_skew(&h[1], rop, z.ni, z.nj, z.nk, z.nd, z.neq, p.tk, z.neq, 0);
            cudaMemcpy((CUtensorMap *)p.d_tm + 3 + 2 * q, h, sizeof(h), cudaMemcpyHostToDevice);   // blocking, three times per zone
            p.rop_seen[q] = rop;
        }
        a.tm_rop = (const CUtensorMap *)p.d_tm + 3 + 2 * q;  // + 1: upper sweep (lu_cell_body)
    }
}

// ------------------------------------------------------------------------------------------------ multi-zone launch
struct LuBatch {
    int nz = 0, typ = 0, neq = 0, ntiles = 0;
    bool cell = false;
    int2 *d_order = nullptr;          // ticket -> (zone, J | K << 16), ordered by J + K with the zones interleaved
    int *d_ticket = nullptr;          // [lower, upper]
    // device tables of {ZoneDev, LuTileArgs} per zone, one per combination of state pointers the time-level rotation and
    // the sub-iteration number produce (six in a run), built on first use
    std::map<std::vector<const void *>, LuBatchItem *> tables;
};

LuBatch *lu_batch_create(int nz, const ZoneDev *const *zs, const LuPlan *const *plans)
{
    { const char *e = getenv("FAST_B200_LU"); if (nz < 2 || (e && e[0] == 'p')) return nullptr; }   // the hyperplane A/B kernels keep one launch per zone
    for (int n = 0; n < nz; n++)
        if (zs[n]->typ != zs[0]->typ || zs[n]->neq != zs[0]->neq || zs[n]->itypcp > 1 || !plans[n]->d_order || plans[n]->nrelax >= 2 ||
            plans[n]->cell != plans[0]->cell) return nullptr;
    LuBatch *b = new LuBatch();
    b->nz = nz; b->typ = zs[0]->typ; b->neq = zs[0]->neq; b->cell = plans[0]->cell;
    std::vector<std::pair<int, int2>> ord;     // (J + K, (zone, tile))
    for (int n = 0; n < nz; n++)
        for (int K = 0; K < plans[n]->NK; K++) for (int J = 0; J < plans[n]->NJ; J++) ord.push_back({J + K, make_int2(n, J | (K << 16))});
    std::stable_sort(ord.begin(), ord.end(), [](const std::pair<int, int2> &x, const std::pair<int, int2> &y) { return x.first < y.first; });
    std::vector<int2> o2(ord.size());
    for (size_t i = 0; i < ord.size(); i++) o2[i] = ord[i].second;
    b->ntiles = (int)o2.size();
    if (cudaMalloc(&b->d_order, o2.size() * sizeof(int2)) != cudaSuccess || cudaMalloc(&b->d_ticket, 2 * sizeof(int)) != cudaSuccess) {
        cudaGetLastError(); lu_batch_free(b); return nullptr;
    }
    cudaMemcpy(b->d_order, o2.data(), o2.size() * sizeof(int2), cudaMemcpyHostToDevice);
    return b;
}

void lu_batch_free(LuBatch *b)
{
    if (!b) return;
    for (auto &t : b->tables) cudaFree(t.second);
    cudaFree(b->d_order); cudaFree(b->d_ticket);
    delete b;
}

template <int MESH, int NEQ>
static void run_lu_batch(const LuBatch *b, const LuBatchItem *items, cudaStream_t st)
{
    static std::atomic<unsigned long long> attr_done{0};
    if (fb_first_on_device(attr_done)) {
        cudaFuncSetAttribute(k_lu_tile_b<MESH, NEQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LuCfg<MESH, NEQ, false>::SMEM);
        cudaFuncSetAttribute(k_lu_tile_b<MESH, NEQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LuCfg<MESH, NEQ, true>::SMEM);
        cudaFuncSetAttribute(k_lu_cell_b<MESH, NEQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LcCfg<MESH, NEQ, false>::SMEM);
        cudaFuncSetAttribute(k_lu_cell_b<MESH, NEQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LcCfg<MESH, NEQ, true>::SMEM);
    }
    cudaMemsetAsync(b->d_ticket, 0, 2 * sizeof(int), st);
    if (b->cell) {
        g_lu_cell_launches += 2;
        k_lu_cell_b<MESH, NEQ, false><<<b->ntiles, LcCfg<MESH, NEQ, false>::NTHR, LcCfg<MESH, NEQ, false>::SMEM, st>>>(items, b->d_order, b->d_ticket);
        k_lu_cell_b<MESH, NEQ, true><<<b->ntiles, LcCfg<MESH, NEQ, true>::NTHR, LcCfg<MESH, NEQ, true>::SMEM, st>>>(items, b->d_order, b->d_ticket + 1);
        return;
    }
    k_lu_tile_b<MESH, NEQ, false><<<b->ntiles, LU_NT, LuCfg<MESH, NEQ, false>::SMEM, st>>>(items, b->d_order, b->d_ticket);
    k_lu_tile_b<MESH, NEQ, true><<<b->ntiles, LU_NT, LuCfg<MESH, NEQ, true>::SMEM, st>>>(items, b->d_order, b->d_ticket + 1);
}

int lu_batch_launch(LuBatch *b, const ZoneDev *const *zs, const LuPlan *const *plans, const double *const *rop, double *const *rop_out,
                    int lu_match, cudaStream_t st)
{
    std::vector<const void *> key;
    for (int n = 0; n < b->nz; n++) { key.push_back(rop[n]); key.push_back(rop_out[n]); }
    key.push_back((const void *)(size_t)lu_match);
    auto it = b->tables.find(key);
    if (it == b->tables.end()) {
        std::vector<LuBatchItem> h(b->nz);
        for (int n = 0; n < b->nz; n++) { h[n].z = *zs[n]; lu_fill_args(h[n].a, *zs[n], *plans[n], rop[n], rop_out[n], lu_match); }
        LuBatchItem *d = nullptr;
        if (cudaMalloc(&d, h.size() * sizeof(LuBatchItem)) != cudaSuccess) { cudaGetLastError(); return -1; }
        cudaMemcpy(d, h.data(), h.size() * sizeof(LuBatchItem), cudaMemcpyHostToDevice);   // blocking, six times in a run
        it = b->tables.emplace(key, d).first;
    }
    const LuBatchItem *items = it->second;
    if (b->neq == 6) {
        if (b->typ == 0) run_lu_batch<0, 6>(b, items, st); else if (b->typ == 1) run_lu_batch<1, 6>(b, items, st); else run_lu_batch<2, 6>(b, items, st);
    } else {
        if (b->typ == 0) run_lu_batch<0, 5>(b, items, st); else if (b->typ == 1) run_lu_batch<1, 5>(b, items, st); else run_lu_batch<2, 5>(b, items, st);
    }
    return 3;
}

template <int MESH, int NEQ, bool UPPER>
static void launch_sweep(const ZoneDev &z, const LuPlan &p, const LuTileArgs &a, cudaStream_t st)
{
    static std::atomic<unsigned long long> attr_done{0};
    if (fb_first_on_device(attr_done)) {
        cudaFuncSetAttribute(k_lu_tile<MESH, NEQ, UPPER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LuCfg<MESH, NEQ, UPPER>::SMEM);
        cudaFuncSetAttribute(k_lu_cell<MESH, NEQ, UPPER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LcCfg<MESH, NEQ, UPPER>::SMEM);
    }
    const int ntiles = p.NJ * p.NK;
    if (p.cell) g_lu_cell_launches++;
    if (p.cell) k_lu_cell<MESH, NEQ, UPPER><<<ntiles, LcCfg<MESH, NEQ, UPPER>::NTHR, LcCfg<MESH, NEQ, UPPER>::SMEM, st>>>(z, a);
    else k_lu_tile<MESH, NEQ, UPPER><<<ntiles, LU_NT, LuCfg<MESH, NEQ, UPPER>::SMEM, st>>>(z, a);
}

template <int MESH, int NEQ>
static int run_lu_tiles(const ZoneDev &z, const LuPlan &p, const double *rop, double *rop_out, int lu_match, cudaStream_t st)
{
    LuTileArgs a;
    lu_fill_args(a, z, p, rop, rop_out, lu_match);
    cudaMemsetAsync(p.d_ctl, 0, 2 * sizeof(int), st);
    // FAST_B200_LU_DEBUG (k_lu_cell, build with -DLU_CLK for the cycle counters): per tile start / end (globaltimer) and cycles per role
    static const bool dbg = getenv("FAST_B200_LU_DEBUG") != nullptr;
    static unsigned long long *d_dbg = nullptr; static int dbg_calls = 0;
    const int ntiles = p.NJ * p.NK;
    a.dbg = nullptr;
    if (dbg && p.cell && dbg_calls < 3) { if (!d_dbg) cudaMalloc(&d_dbg, 2 * 16 * 65536 * sizeof(unsigned long long)); cudaMemsetAsync(d_dbg, 0, 2 * 16 * 65536 * 8, st); a.dbg = d_dbg; }
    a.ticket = p.d_ctl; a.order_mode = lu_match ? 1 : (MESH == 2 ? 2 : 0);
    launch_sweep<MESH, NEQ, false>(z, p, a, st);
    a.ticket = p.d_ctl + 1; a.order_mode = lu_match ? 1 : 0;
    if (a.dbg) a.dbg += 16 * 65536;
    launch_sweep<MESH, NEQ, true>(z, p, a, st);
    if (a.dbg) {
        cudaStreamSynchronize(st);
        std::vector<unsigned long long> h(2 * 16 * 65536);
        cudaMemcpy(h.data(), d_dbg, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        for (int sw = 0; sw < 2; sw++) {
            const unsigned long long *q = h.data() + sw * 16 * 65536; unsigned long long t0 = ~0ULL, t1 = 0;
            for (int t = 0; t < ntiles; t++) { t0 = std::min(t0, q[16 * t]); t1 = std::max(t1, q[16 * t + 1]); }
            fprintf(stderr, "LUDBG sweep %d call %d: NJ %d NK %d tiles %d bi %d total %.1f us\n", sw, dbg_calls, p.NJ, p.NK, ntiles, a.bi, (t1 - t0) * 1e-3);
            const double n = (a.bi + 1) / 2 + 22.;        // super-steps of a tile
            for (int t = 0; t < ntiles; t++) {
                const int J = t % p.NJ, K = t / p.NJ;
                if (dbg_calls == 1 && (J % 8 == 0 || J == p.NJ - 1))
                    fprintf(stderr, "LUDBG   tile J %2d K %2d start %8.1f end %8.1f dur %8.1f us | cycles / super-step: compute %6.0f (+ wait for the slot %5.0f) | service thread 0: B_S %5.0f wb %5.0f svc-barrier %5.0f mbox %5.0f deliver %5.0f slot-wait %5.0f | thread 40: %5.0f %5.0f %5.0f %5.0f %5.0f %5.0f\n",
                            J, K, (q[16 * t] - t0) * 1e-3, (q[16 * t + 1] - t0) * 1e-3, (q[16 * t + 1] - q[16 * t]) * 1e-3, q[16 * t + 2] / n, q[16 * t + 3] / n,
                            q[16 * t + 4] / n, q[16 * t + 5] / n, q[16 * t + 6] / n, q[16 * t + 7] / n, q[16 * t + 8] / n, q[16 * t + 9] / n,
                            q[16 * t + 10] / n, q[16 * t + 11] / n, q[16 * t + 12] / n, q[16 * t + 13] / n, q[16 * t + 14] / n, q[16 * t + 15] / n);
            }
        }
        dbg_calls++;
    }
    return 3;
}

// ------------------------------------------------------------------------------------------------ nb_relax >= 2
// LU/invlu.for:163-240 -> invlussor1_l / invlussor_l / invlussor_u[_SA].for: symmetric Gauss-Seidel relaxations of
// (D + L + U) dW = rhs.  Relaxation m >= 2 is run with the SAME sweep kernels by moving the coupling to the previous
// relaxation into their inputs:
//   lower:  (D + L) dW* = rhs + Usum(dW_prev)               -> input  rhs' = rhs + Usum                     (k_ssor_lower_in)
//   upper:  (D + U) dW  = rhs + Lsum(dW*) = D dW* - Usum    -> input  dW*' = dW* - Usum / coe5 (coe6)        (k_ssor_upper_in)
// with Usum(l) = xal(l) * sum_d B_d(l+d) dW_prev(l+d) over the upper neighbours inside the box (k_ssor_usum; the reference
// accumulates the same sum in its ssor array while it sweeps, mjr_drodm_ssor.for).  Only the last upper sweep applies the
// Newton update (lussor_end, invlussor_u.for:155-170), over the interior ind_loop_sdm; the others store dW.
template <int MESH, int NEQ>
__global__ void k_ssor_usum(ZoneDev z, const double *__restrict__ rop, int i0, int i1, int j0, int j1, int k0, int k1, double *__restrict__ us)
{
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x, j = j0 + blockIdx.y * blockDim.y + threadIdx.y, k = k0 + blockIdx.z;
    if (i > i1 || j > j1) return;
    const int nd = z.nd, l = inddm(z, i, j, k), lt = MESH == 2 ? 0 : indmtr(z, i, j, k);
    const double xal = z.coe[l] * -0.5;
    const int inc[3] = {1, z.incj, z.inck}, incm[3] = {z.mi, z.mj, z.mk};
    const bool has[3] = {i < i1, j < j1, k < k1};
    double s[NEQ], b[NEQ];
#pragma unroll
    for (int e = 0; e < NEQ; e++) s[e] = 0.;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        if (!has[d]) continue;
        lu_term<MESH, NEQ>(z, rop, d, l + inc[d], lt + incm[d], (MESH == 1 && d == 2) ? lt : lt + 2 * incm[d], -0.5, false, b);
#pragma unroll
        for (int e = 0; e < NEQ; e++) s[e] = s[e] + b[e] * xal;
    }
#pragma unroll
    for (int e = 0; e < NEQ; e++) us[l + e * nd] = s[e];
}
template <int NEQ, bool UPPER>
__global__ void k_ssor_in(ZoneDev z, int i0, int i1, int j0, int j1, int k0, const double *__restrict__ rhs, const double *__restrict__ us)
{
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x, j = j0 + blockIdx.y * blockDim.y + threadIdx.y, k = k0 + blockIdx.z;
    if (i > i1 || j > j1) return;
    const int nd = z.nd, l = inddm(z, i, j, k);
    if (!UPPER) {
#pragma unroll
        for (int e = 0; e < NEQ; e++) z.drodm[l + e * nd] = rhs[l + e * nd] + us[l + e * nd];
    } else {
        const double d5 = 1. / z.coe[l + 4 * nd];
#pragma unroll
        for (int e = 0; e < 5; e++) z.drodm[l + e * nd] = z.drodm[l + e * nd] - us[l + e * nd] * d5;
        if (NEQ == 6) z.drodm[l + 5 * nd] = z.drodm[l + 5 * nd] - us[l + 5 * nd] / z.coe[l + 5 * nd];
    }
}

template <int MESH, int NEQ>
static int run_lu_ssor(const ZoneDev &z, const LuPlan &p, const double *rop, double *rop_out, cudaStream_t st)
{
    LuTileArgs a;
    lu_fill_args(a, z, p, rop, rop_out, 1);       // Newton box = interior (ind_loop_sdm)
    const int alias = a.alias;
    const int *bx = p.box;
    dim3 b(64, 4, 1), g((bx[1] - bx[0] + 64) / 64, (bx[3] - bx[2] + 4) / 4, bx[5] - bx[4] + 1);
    cudaMemcpyAsync(p.d_rhs, z.drodm, (size_t)z.nd * NEQ * sizeof(double), cudaMemcpyDeviceToDevice, st);
    int launches = 1;
    for (int m = 1; m <= p.nrelax; m++) {
        if (m > 1) {
            k_ssor_usum<MESH, NEQ><<<g, b, 0, st>>>(z, rop, bx[0], bx[1], bx[2], bx[3], bx[4], bx[5], p.d_us);
            k_ssor_in<NEQ, false><<<g, b, 0, st>>>(z, bx[0], bx[1], bx[2], bx[3], bx[4], p.d_rhs, p.d_us);
            launches += 2;
        }
        cudaMemsetAsync(p.d_ctl, 0, 2 * sizeof(int), st);
        a.ticket = p.d_ctl; a.newton = 0; a.alias = 0;
        launch_sweep<MESH, NEQ, false>(z, p, a, st);
        if (m > 1) { k_ssor_in<NEQ, true><<<g, b, 0, st>>>(z, bx[0], bx[1], bx[2], bx[3], bx[4], p.d_rhs, p.d_us); launches++; }
        a.ticket = p.d_ctl + 1; a.newton = m == p.nrelax ? 1 : 0; a.alias = a.newton ? alias : 0;
        launch_sweep<MESH, NEQ, true>(z, p, a, st);
        launches += 3;
    }
    return launches;
}

int launch_lu_tiles(const ZoneDev &z, const LuPlan &p, const double *rop, double *rop_out, int lu_match, cudaStream_t st)
{
    if (p.nrelax >= 2) {      // LU-SSOR relaxations: always the tile kernel
        if (z.neq == 6) return z.typ == 0 ? run_lu_ssor<0, 6>(z, p, rop, rop_out, st) : (z.typ == 1 ? run_lu_ssor<1, 6>(z, p, rop, rop_out, st) : run_lu_ssor<2, 6>(z, p, rop, rop_out, st));
        return z.typ == 0 ? run_lu_ssor<0, 5>(z, p, rop, rop_out, st) : (z.typ == 1 ? run_lu_ssor<1, 5>(z, p, rop, rop_out, st) : run_lu_ssor<2, 5>(z, p, rop, rop_out, st));
    }
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
{   // reports the spin-timeout flag once and clears it (the caller raises; a later sweep starts clean)
    int e = 0;
    if (p.d_ctl) {
        cudaMemcpy(&e, p.d_ctl + 2, sizeof(int), cudaMemcpyDeviceToHost);
        if (e) cudaMemset(p.d_ctl + 2, 0, sizeof(int));
    }
    return e;
}
