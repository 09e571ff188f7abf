// fb_lu2.cu -- LU-SGS lower / upper sweeps + Newton update, tile-pipelined wavefront, ONE THREAD PER GRID LINE,
// TWO TILES PER SM.
//
// Reference: S/Compute/LU/invlu.for:6-344 -> invlu_l[_SA].for:79-215, invlu_u[_SA].for:60-215, lu_{i,j,k}_*.for,
// mjr_newton[_SA].for.  Same dependency-preserving reordering, mailboxes and staging ring as k_lu_tile (fb_lu.cu, which
// stays selectable with FAST_B200_LU=tile), re-balanced after measuring that kernel (profiles/r01_summary.md): its
// compute warps -- 4 lanes per line, every lane redoing dW* -- are busy 1130 / 2070 cycles per step on a chain of
// dependent FP64 instructions, memory is idle, one 172-KB tile fits per SM and 264 tiles need two waves.  Here
//   * thread lam = jj + TJ*kk owns line (jj,kk) of a TJ x TK = 8 x 8 tile (8 x 6 for 3dfull SA): it keeps dW*, the state
//     and its i product in registers, evaluates the three products B_d(l).dW(l) itself (three independent chains,
//     common sub-expressions shared) and applies the Newton update of the upper sweep in place;
//   * 2 compute warps + 4 refill / write-back warps + 2 mailbox warps = 256 threads, one __syncthreads per step;
//   * the staging ring holds NB = 2 chunks of CI = 4 cells per line and array (half of k_lu_tile's), so that TWO CTAs
//     are resident per SM: 296 tile slots for the 264 tiles of C3 (one wave), and the second tile's warps fill the
//     issue slots the first one leaves idle while it waits on its dependent chain.
// HBM traffic per cell is unchanged: lower reads rhs+coe+(u,v,w,T)+normals, writes dW*; upper reads dW*+coe+state+
// normals and writes ro_P1 (dW itself is never stored).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "fb_implicit.cuh"
#include "fb_lu_prod.cuh"

#define L2_CI 4
#define L2_NB 2
#define L2_TJ 8
#define L2_MB 6
#define L2_EMPTY 0xFFF7DEADBEEF5A5AULL   // signalling NaN: never produced by arithmetic
#define L2_SPIN_MAX (1 << 21)
#define L2_NCT 64          // compute threads (one per line)
#define L2_NIO 128         // refill + write-back threads (two halves sharing the staged arrays)
#define L2_NMT 64          // mailbox threads (3 per interface line)

template <int MESH, int NEQ, bool UPPER>
struct Lu2Cfg {
    static constexpr int TJ = L2_TJ;
    static constexpr int TK = (MESH == 0 && NEQ == 6) ? 6 : 8;
    static constexpr int L = TJ * TK;
    static constexpr int NCT = L2_NCT;
    static constexpr int NTH = L2_NCT + L2_NIO + L2_NMT;
    static constexpr int NC = NEQ == 6 ? 6 : 5;            // coe columns staged
    static constexpr int NS = UPPER ? NEQ : 4;             // state variables staged (upper: all, lower: u,v,w,T)
    static constexpr int S0 = UPPER ? 0 : 1;
    static constexpr int NT = MESH == 0 ? 9 : (MESH == 1 ? 5 : 0);
    static constexpr int A_DW = 0, A_COE = NEQ, A_ST = NEQ + NC, A_T = NEQ + NC + NS;
    static constexpr int A = NEQ + NC + NS + NT;
    static constexpr int NIL = TJ + TK;                    // interface lines: TK on the j side, TJ on the k side
    static constexpr int RING = L2_NB * A * L * L2_CI, EXB = 2 * 2 * L * L2_MB, MBS = 4 * NIL * L2_MB, MBX = 2 * NIL * L2_MB;
    static constexpr size_t SMEM = (size_t)(RING + EXB + MBS + MBX) * sizeof(double);
    static_assert(L <= L2_NCT && (L / L2_CI) * L2_CI <= L2_NIO / 2 && 3 * NIL <= L2_NMT, "thread roles too small for the tile");
};

void lu2_tile_dims(int mesh, int neq, int &tj, int &tk) { tj = L2_TJ; tk = (mesh == 0 && neq == 6) ? 6 : 8; }

struct Lu2Args {
    const double *rop;      // rop_ssiter
    double *rop_out;        // ro_P1
    double *dw;             // drodm: rhs -> dW*
    const double *coe;
    const double *tcm;      // cell-centred mean normals (k_lunorm)
    long long tstride;
    double *ifj, *ifk;      // mailboxes
    const int *order;       // ticket -> J | K << 16
    int *ticket;
    int *err;
    int box[6], nbx[6];
    int bi, bj, bk, NJ, NK;
    int alias;
    int plane;              // ni*nj (3dhomo: tcm is a plane array)
    unsigned long long *dbg;
};

__device__ __forceinline__ void l2_cp_async8(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void l2_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void l2_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ unsigned long long l2_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// ring address of (chunk buffer, array, line, cell slot).  The line's block of CI cells is XOR-permuted inside its kk row
// (jj ^ 2kk): the 32 lines of a compute warp, whose slots follow the line skew, then hit every 8-byte bank exactly twice
// (2 wavefronts per LDS.64, the minimum), and the 8 lines a refill warp fills at once spread over two bank groups.
template <int A, int L>
__device__ __forceinline__ int ring_at(int buf, int arr, int lam, int slot)
{
    return ((buf * A + arr) * L + (lam ^ ((lam >> 2) & 6))) * L2_CI + slot;
}

template <int MESH, int NEQ, bool UPPER>
__global__ void __launch_bounds__((Lu2Cfg<MESH, NEQ, UPPER>::NTH), 2) k_lu_line(ZoneDev z, Lu2Args a)
{
    using C = Lu2Cfg<MESH, NEQ, UPPER>;
    constexpr int TJ = C::TJ, TK = C::TK, L = C::L, A = C::A, CI = L2_CI, NB = L2_NB, EX = L2_MB, NIL = C::NIL;
    constexpr int LPK = TJ / CI;                      // lines per kk row whose chunk turns over at a given step
    extern __shared__ __align__(16) double smem[];
    double *ring = smem;                              // [NB][A][L][CI]
    double *exb = smem + C::RING;                     // [2][2][L][EX] products handed to (jj+1) and (kk+1)
    unsigned long long *mbs = (unsigned long long *)(exb + C::EXB);   // [4][NIL][EX] prefetched mailbox words
    double *mbx = (double *)(mbs + C::MBS);           // [2][NIL][EX] validated products of the neighbour tiles
    __shared__ int s_tile;
    const int tid = threadIdx.x;
    if (tid == 0) s_tile = a.order[atomicAdd(a.ticket, 1)];
    // exchange buffers start at zero: products that are multiplied by a zero weight (box boundary) must be finite
    for (int i = tid; i < C::EXB + C::MBS + C::MBX; i += C::NTH) exb[i] = 0.;
    __syncthreads();
    const int J = s_tile & 0xffff, K = s_tile >> 16;
    if (a.dbg && tid == 0) a.dbg[8 * (K * a.NJ + J)] = l2_gtime();
    const int bi = a.bi, bj = a.bj, bk = a.bk;
    const int nst = bi + TJ + TK - 2;
    const int nd = z.nd;
    const int sg = UPPER ? -1 : 1;
    // cell address of sweep coordinates (ii, jjg, kkg): l0 + sg * (ii + jjg * incj + kkg * inck)
    const int l0 = UPPER ? inddm(z, a.box[1], a.box[3], a.box[5]) : inddm(z, a.box[0], a.box[2], a.box[4]);

    if (tid >= C::NCT) {
        // ------------------------------------------------------------------------------------ service warps
        if (tid < C::NCT + L2_NIO) {
            // ---- refill + write-back: 2 x 64 threads, lanes along i inside a chunk, then across the 16 lines whose chunk
            // turns over at this step; the two halves share the staged arrays
            const int st = tid - C::NCT, half = st / 64, s2 = st % 64;
            const int q = s2 / CI, cell = s2 % CI;
            const int kk = q / LPK, kkg = K * TK + kk;
            struct Ln { int lam, lrow, m, jjg; bool ok; };
            auto line_of = [&](int t) {
                Ln o;
                const int jj = ((t - kk) % CI + CI) % CI + CI * (q % LPK);
                o.jjg = J * TJ + jj;
                o.ok = kk < TK && o.jjg < bj && kkg < bk;
                o.m = (t - (jj + kk)) / CI;                         // exact: a multiple of CI (may be < 0)
                o.lam = kk * TJ + jj;
                o.lrow = l0 + sg * (o.jjg * z.incj + kkg * z.inck);
                return o;
            };
            auto refill = [&](int t) {
                const Ln o = line_of(t);
                const int c = o.m - 1 + NB, ii = c * CI + cell;
                if (!(o.ok && c >= 0 && ii < bi)) return;
                const int l = o.lrow + sg * ii;
                double *dst = ring + ring_at<A, L>(c % NB, 0, o.lam, cell);
                if (half == 0) {
#pragma unroll
                    for (int e = 0; e < NEQ; e++) l2_cp_async8(dst + (size_t)(C::A_DW + e) * L * CI, a.dw + l + e * nd);
#pragma unroll
                    for (int e = 0; e < C::NC; e++) l2_cp_async8(dst + (size_t)(C::A_COE + e) * L * CI, a.coe + l + e * nd);
                } else {
#pragma unroll
                    for (int e = 0; e < C::NS; e++) l2_cp_async8(dst + (size_t)(C::A_ST + e) * L * CI, a.rop + l + (C::S0 + e) * nd);
                    if (MESH != 2) {
                        const int lt = MESH == 1 ? l % a.plane : l;
#pragma unroll
                        for (int e = 0; e < C::NT; e++) l2_cp_async8(dst + (size_t)(C::A_T + e) * L * CI, a.tcm + lt + e * a.tstride);
                    }
                }
            };
            // write-back: dW* (lower) or the Newton-updated state the compute thread left in the ring (upper); each half
            // takes every other variable
            auto writeback = [&](int t) {
                const Ln o = line_of(t);
                const int c = o.m - 1, ii = c * CI + cell;
                if (!(o.ok && c >= 0 && ii < bi)) return;
                const int l = o.lrow + sg * ii;
                const double *src = ring + ring_at<A, L>(c % NB, 0, o.lam, cell);
                if (!UPPER) {
#pragma unroll
                    for (int e = 0; e < NEQ; e++) if ((e & 1) == half) a.dw[l + e * nd] = src[(size_t)(C::A_DW + e) * L * CI];
                } else {
                    const int i = a.box[1] - ii, j = a.box[3] - o.jjg, k = a.box[5] - kkg;
                    if (i >= a.nbx[0] && i <= a.nbx[1] && j >= a.nbx[2] && j <= a.nbx[3] && k >= a.nbx[4] && k <= a.nbx[5]) {
#pragma unroll
                        for (int e = 0; e < NEQ; e++) if ((e & 1) == half) a.rop_out[l + e * nd] = src[(size_t)(C::A_ST + e) * L * CI];
                    }
                }
            };
            // rule(t): the lines with (t - sigma) % CI == 0 finished chunk m-1 at step t-1 (m = (t - sigma) / CI).  That
            // chunk is written back one iteration later (t = s-1: the barrier of iteration s-1 orders it after the compute
            // step that produced it) and its buffer refilled with chunk m-1+NB two iterations later (t = s-2).
            for (int t = (1 - NB) * CI; t <= -3; t++) { refill(t); l2_commit(); }
            for (int s = 0; s < nst; s++) {
                refill(s - 2);
                l2_commit();
                writeback(s - 1);
                l2_wait<(NB - 1) * CI - 2>();
                __syncthreads();
            }
            __syncthreads();
            l2_wait<0>();
            for (int s = nst; s <= nst + CI; s++) writeback(s - 1);
            return;
        }
        // ---- mailboxes: thread sw owns 16 bytes of the entry of interface line il (il < TK: j interface, row kk = il;
        // otherwise k interface, column jj = il - TK).  Words needed at step X are loaded at iteration X-3 (registers
        // carried over the loop edge), parked in shared memory at X-2 and, before the barrier of step X, delivered
        // validated into mbx (polling global memory while the producer tile has not delivered); the entry is re-armed.
        const int sw = tid - C::NCT - L2_NIO;
        const int il = sw / 3, part = sw % 3;
        const bool isj = il < TK;
        const int mjj = isj ? 0 : il - TK, mkk = isj ? il : 0;
        const int mjg = J * TJ + mjj, mkg = K * TK + mkk;
        const bool mth = il < NIL;
        const bool mvalid = mth && mjg < bj && mkg < bk && (isj ? J > 0 : K > 0);
        volatile unsigned long long *mrow = (volatile unsigned long long *)a.order;   // harmless readable dummy
        if (mvalid) mrow = (unsigned long long *)(isj ? a.ifj + ((size_t)((J - 1) * bk + mkg) * bi) * L2_MB
                                                        : a.ifk + ((size_t)((K - 1) * bj + mjg) * bi) * L2_MB) + 2 * part;
        const int msig = mjj + mkk;
        unsigned long long *mpre = mbs + (size_t)(mth ? il : 0) * EX + 2 * part;   // + (step & 3) * NIL * EX
        double *mout = mbx + (size_t)(mth ? il : 0) * EX + 2 * part;                // + (step & 1) * NIL * EX
        auto entry = [&](int X) {    // clamped: always a readable address
            int ii = X - msig; ii = ii < 0 ? 0 : (ii >= bi ? bi - 1 : ii);
            return mrow + (mvalid ? (size_t)ii * L2_MB : 0);
        };
        auto needed = [&](int X) { const int ii = X - msig; return mvalid && ii >= 0 && ii < bi; };
        auto deliver = [&](int X) {
            if (!needed(X)) return;
            const unsigned long long *pre = mpre + (size_t)(X & 3) * NIL * EX;
            unsigned long long v0 = pre[0], v1 = pre[1];
            volatile unsigned long long *p = entry(X);
            int spin = 0;
            while (v0 == L2_EMPTY || v1 == L2_EMPTY) {
                v0 = p[0]; v1 = p[1];
                if (++spin > L2_SPIN_MAX || (((spin & 1023) == 0) && *(volatile int *)a.err)) { *a.err = 1; break; }
            }
            double *o = mout + (size_t)(X & 1) * NIL * EX;
            o[0] = __longlong_as_double((long long)v0); o[1] = __longlong_as_double((long long)v1);
            p[0] = L2_EMPTY; p[1] = L2_EMPTY;                                  // self-cleaning: ready for the next sweep
        };
        for (int X = 0; X < 2; X++) {
            volatile unsigned long long *p = entry(X);
            unsigned long long *d = mpre + (size_t)(X & 3) * NIL * EX;
            if (mth) { d[0] = p[0]; d[1] = p[1]; }
        }
        unsigned long long w0, w1;
        { volatile unsigned long long *p = entry(2); w0 = p[0]; w1 = p[1]; }
        for (int s = 0; s < nst; s++) {
            if (mth) { unsigned long long *d = mpre + (size_t)((s + 2) & 3) * NIL * EX; d[0] = w0; d[1] = w1; }
            { volatile unsigned long long *p = entry(s + 3); w0 = p[0]; w1 = p[1]; }
            deliver(s);
            __syncthreads();
        }
        __syncthreads();
        return;
    }

    // ---------------------------------------------------------------------------------------- compute warps
    const int lam = tid;
    const int jj = lam % TJ, kk = lam / TJ;
    const int jjg = J * TJ + jj, kkg = K * TK + kk;
    const bool lvalid = lam < L && jjg < bj && kkg < bk;
    const int sigma = jj + kk;
    const bool has_j = jjg > 0, has_k = kkg > 0;
    volatile unsigned long long *mb_wj = nullptr, *mb_wk = nullptr;
    if (lvalid && jj == TJ - 1 && J + 1 < a.NJ) mb_wj = (unsigned long long *)(a.ifj + ((size_t)(J * bk + kkg) * bi) * L2_MB);
    if (lvalid && kk == TK - 1 && K + 1 < a.NK) mb_wk = (unsigned long long *)(a.ifk + ((size_t)(K * bj + jjg) * bi) * L2_MB);
    // Newton box test of the own line (upper sweep)
    const int jg = UPPER ? a.box[3] - jjg : a.box[2] + jjg, kg = UPPER ? a.box[5] - kkg : a.box[4] + kkg;
    const bool line_in_nb = jg >= a.nbx[2] && jg <= a.nbx[3] && kg >= a.nbx[4] && kg <= a.nbx[5];
    const double signe = UPPER ? -0.5 : 0.5;
    const double gam = z.gam, cp = z.gam * z.cv;
    const double cvinv = 1. / z.cv, cvinv2 = 0.5 * cvinv;
    // where the j / k products come from: neighbour line of the tile (exb) or neighbour tile (mbx); offsets in doubles
    const int offj = jj > 0 ? (lam - 1) * EX : -1, offk = kk > 0 ? (L + lam - TJ) * EX : -1;
    const int mofj = kk * EX, mofk = (TK + jj) * EX;
    double pi_[EX];                                    // own i product of the previous cell
#pragma unroll
    for (int e = 0; e < EX; e++) pi_[e] = 0.;
    if (a.dbg && tid == 0) a.dbg[8 * (K * a.NJ + J) + 1] = l2_gtime();
    for (int s = 0; s < nst; s++) {
        __syncthreads();
        if (a.dbg && tid == 0 && (s == 16 || s == 128 || s == 256 || s == 384)) a.dbg[8 * (K * a.NJ + J) + 4 + (s >> 7)] = l2_gtime();
        const int ii = s - sigma;
        if (!(lvalid && ii >= 0 && ii < bi)) continue;
        const int c = ii / CI, slot = ii % CI;
        double *R = ring + ring_at<A, L>(c % NB, 0, lam, slot);
#define RING(arr) R[(size_t)(arr) * L * CI]
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
        // cell-centred mean normals of the three directions
        double nix = 0., niy = 0., niz = 0., njx = 0., njy = 0., njz = 0., nkx = 0., nky = 0., nkz = 0.;
        if (MESH == 0) {
            nix = RING(C::A_T); niy = RING(C::A_T + 1); niz = RING(C::A_T + 2); njx = RING(C::A_T + 3); njy = RING(C::A_T + 4);
            njz = RING(C::A_T + 5); nkx = RING(C::A_T + 6); nky = RING(C::A_T + 7); nkz = RING(C::A_T + 8);
        } else if (MESH == 1) {
            nix = RING(C::A_T); niy = RING(C::A_T + 1); njx = RING(C::A_T + 2); njy = RING(C::A_T + 3); nkz = RING(C::A_T + 4);
        } else { nix = z.tc0[0]; njy = z.tc0[1]; nkz = z.tc0[2]; }
        const double cdi = RING(C::A_COE + 1) * signe, cdj = RING(C::A_COE + 2) * signe, cdk = RING(C::A_COE + 3) * signe;
        // incoming products; the sum follows the reference's k, j, i order (invlu_l.for / invlu_u.for)
        const double *xin = exb + (size_t)((s + 1) & 1) * 2 * L * EX;
        const double *xmb = mbx + (size_t)(s & 1) * NIL * EX;
        const double2 *pj_ = (const double2 *)(offj >= 0 ? xin + offj : xmb + mofj);
        const double2 *pk_ = (const double2 *)(offk >= 0 ? xin + offk : xmb + mofk);
        const double2 j0 = pj_[0], j1 = pj_[1], j2 = pj_[2], k0 = pk_[0], k1 = pk_[1], k2 = pk_[2];
        const double xal = coe1 * signe;
        const double xi = ii > 0 ? xal : 0., xj = has_j ? xal : 0., xk = has_k ? xal : 0.;
        if (UPPER) {
#pragma unroll
            for (int e = 0; e < 5; e++) acc[e] = acc[e] * coe5;
            if (NEQ == 6) acc[NEQ - 1] = acc[NEQ - 1] * coe6;
        }
        acc[0] = ((acc[0] + k0.x * xk) + j0.x * xj) + pi_[0] * xi;
        acc[1] = ((acc[1] + k0.y * xk) + j0.y * xj) + pi_[1] * xi;
        acc[2] = ((acc[2] + k1.x * xk) + j1.x * xj) + pi_[2] * xi;
        acc[3] = ((acc[3] + k1.y * xk) + j1.y * xj) + pi_[3] * xi;
        acc[4] = ((acc[4] + k2.x * xk) + j2.x * xj) + pi_[4] * xi;
        if (NEQ == 6) acc[NEQ - 1] = ((acc[NEQ - 1] + k2.y * xk) + j2.y * xj) + pi_[5] * xi;
        const double diag = lu_rcp(coe5);
        double dwn[NEQ];
#pragma unroll
        for (int e = 0; e < 5; e++) dwn[e] = acc[e] * diag;
        if (NEQ == 6) dwn[NEQ - 1] = acc[NEQ - 1] * lu_rcp(coe6);
        double uk = u, vk = v, wk = w, tk_ = t;          // state the k product is evaluated with
        if (!UPPER) {
#pragma unroll
            for (int e = 0; e < NEQ; e++) RING(C::A_DW + e) = dwn[e];
        } else {
            // mjr_newton[_SA].for: ro_P1 = prim(W(rop_ssiter) + dW); left in the ring for the write-back warps
            const double rn = ro + dwn[0];
            const double r_1 = lu_rcp(rn);
            const double un = (ro * u + dwn[1]) * r_1, vn = (ro * v + dwn[2]) * r_1, wn = (ro * w + dwn[3]) * r_1;
            const double roe_old = ro * (z.cv * t + 0.5 * (u * u + v * v + w * w));
            const double tn = (roe_old + dwn[4]) * r_1 * cvinv - cvinv2 * (un * un + vn * vn + wn * wn);
            RING(C::A_ST) = rn; RING(C::A_ST + 1) = un; RING(C::A_ST + 2) = vn; RING(C::A_ST + 3) = wn; RING(C::A_ST + 4) = tn;
            if (NEQ == 6) {
                double anulam = z.coesut * sqrt(tn * z.temp01) * tn * lu_rcp(tn + z.cmus1);   // = coesut sqrt(T/T0) / (1 + Cs/T)
                anulam = anulam * r_1 * z.ratiom;
                double tt = (ro * nut + dwn[NEQ - 1]) * r_1;
                tt = fmin(tt, anulam);
                RING(C::A_ST + NEQ - 1) = fmax(tt, 0.);
            }
            // with rop aliased to rop_out the reference has already Newton-updated plane k+1 when it sweeps plane k
            // (invlu_u.for), i.e. the k product of a cell uses its UPDATED state
            if (a.alias) {
                const int ig = a.box[1] - ii;
                const bool in_nb = line_in_nb && ig >= a.nbx[0] && ig <= a.nbx[1];
                uk = in_nb ? un : u; vk = in_nb ? vn : v; wk = in_nb ? wn : w; tk_ = in_nb ? tn : t;
            }
        }
        // outgoing products
        double bj_[EX], bk_[EX];
        pi_[EX - 1] = 0.; bj_[EX - 1] = 0.; bk_[EX - 1] = 0.;
        lu_prod<NEQ>(nix, niy, niz, u, v, w, t, cdi, gam, cp, dwn, pi_);
        lu_prod<NEQ>(njx, njy, njz, u, v, w, t, cdj, gam, cp, dwn, bj_);
        lu_prod<NEQ>(nkx, nky, nkz, uk, vk, wk, tk_, cdk, gam, cp, dwn, bk_);
        double2 *xo = (double2 *)(exb + (size_t)((s & 1) * 2 * L + lam) * EX);
        xo[0] = make_double2(bj_[0], bj_[1]); xo[1] = make_double2(bj_[2], bj_[3]); xo[2] = make_double2(bj_[4], bj_[5]);
        xo += L * EX / 2;
        xo[0] = make_double2(bk_[0], bk_[1]); xo[1] = make_double2(bk_[2], bk_[3]); xo[2] = make_double2(bk_[4], bk_[5]);
        if (mb_wj) {
            volatile unsigned long long *o = mb_wj + (size_t)ii * L2_MB;
#pragma unroll
            for (int e = 0; e < L2_MB; e++) o[e] = (unsigned long long)__double_as_longlong(bj_[e]);
        }
        if (mb_wk) {
            volatile unsigned long long *o = mb_wk + (size_t)ii * L2_MB;
#pragma unroll
            for (int e = 0; e < L2_MB; e++) o[e] = (unsigned long long)__double_as_longlong(bk_[e]);
        }
#undef RING
    }
    __syncthreads();
    if (a.dbg && tid == 0) a.dbg[8 * (K * a.NJ + J) + 2] = l2_gtime();
}

// ------------------------------------------------------------------------------------------------ host side
template <int MESH, int NEQ>
static int run_lu_lines(const ZoneDev &z, const LuPlan &p, const double *rop, double *rop_out, int lu_match, cudaStream_t st)
{
    Lu2Args a;
    a.rop = rop; a.rop_out = rop_out; a.dw = z.drodm; a.coe = z.coe; a.tcm = p.d_tcm; a.tstride = p.tstride;
    a.ifj = p.d_mb; a.ifk = p.d_mb ? p.d_mb + p.n_ifj : nullptr;
    a.order = p.d_order; a.err = p.d_ctl + 2;
    for (int i = 0; i < 6; i++) { a.box[i] = p.box[i]; a.nbx[i] = p.box[i]; }
    if (lu_match) { a.nbx[0] = 1; a.nbx[1] = z.iv; a.nbx[2] = 1; a.nbx[3] = z.jv; a.nbx[4] = 1; a.nbx[5] = z.kv; }
    a.bi = p.box[1] - p.box[0] + 1; a.bj = p.box[3] - p.box[2] + 1; a.bk = p.box[5] - p.box[4] + 1;
    a.NJ = p.NJ; a.NK = p.NK; a.alias = (rop == rop_out) ? 1 : 0; a.plane = z.ni * z.nj;
    using CL = Lu2Cfg<MESH, NEQ, false>; using CU = Lu2Cfg<MESH, NEQ, true>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_lu_line<MESH, NEQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CL::SMEM);
        cudaFuncSetAttribute(k_lu_line<MESH, NEQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CU::SMEM);
        cudaFuncSetAttribute(k_lu_line<MESH, NEQ, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(k_lu_line<MESH, NEQ, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        attr_done = true;
    }
    cudaMemsetAsync(p.d_ctl, 0, 2 * sizeof(int), st);
    const int ntiles = p.NJ * p.NK;
    static const bool dbg = getenv("FAST_B200_LU_DEBUG") != nullptr;
    static unsigned long long *d_dbg = nullptr; static int dbg_calls = 0;
    a.dbg = nullptr;
    if (dbg && dbg_calls < 3) { if (!d_dbg) cudaMalloc(&d_dbg, 2 * 8 * 65536 * sizeof(unsigned long long)); a.dbg = d_dbg; }
    a.ticket = p.d_ctl;
    k_lu_line<MESH, NEQ, false><<<ntiles, CL::NTH, CL::SMEM, st>>>(z, a);
    a.ticket = p.d_ctl + 1;
    if (a.dbg) a.dbg += 8 * 65536;
    k_lu_line<MESH, NEQ, true><<<ntiles, CU::NTH, CU::SMEM, st>>>(z, a);
    if (a.dbg) {
        cudaStreamSynchronize(st);
        std::vector<unsigned long long> h(2 * 8 * 65536);
        cudaMemcpy(h.data(), d_dbg, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        for (int sw = 0; sw < 2; sw++) {
            const unsigned long long *q = h.data() + sw * 8 * 65536; unsigned long long t0 = ~0ULL, t1 = 0;
            for (int t = 0; t < ntiles; t++) { t0 = std::min(t0, q[8 * t]); t1 = std::max(t1, q[8 * t + 2]); }
            fprintf(stderr, "LUDBG sweep %d call %d: NJ %d NK %d tiles %d bi %d total %.1f us\n", sw, dbg_calls, p.NJ, p.NK, ntiles, a.bi, (t1 - t0) * 1e-3);
            for (int t = 0; t < ntiles; t++) {
                const int J = t % p.NJ, K = t / p.NJ;
                if (dbg_calls == 1 && (J % 4 == 0) && (K % 2 == 0 || K == p.NK - 1))
                    fprintf(stderr, "LUDBG   tile J %2d K %2d start %8.1f s16 %8.1f s128 %8.1f s256 %8.1f s384 %8.1f end %8.1f dur %8.1f\n", J, K,
                            (q[8 * t] - t0) * 1e-3, (q[8 * t + 4] - t0) * 1e-3, (q[8 * t + 5] - t0) * 1e-3, (q[8 * t + 6] - t0) * 1e-3, (q[8 * t + 7] - t0) * 1e-3,
                            (q[8 * t + 2] - t0) * 1e-3, (q[8 * t + 2] - q[8 * t]) * 1e-3);
            }
        }
        dbg_calls++;
    }
    return 3;
}

int launch_lu_lines(const ZoneDev &z, const LuPlan &p, const double *rop, double *rop_out, int lu_match, cudaStream_t st)
{
    if (z.neq == 6) {
        if (z.typ == 0) return run_lu_lines<0, 6>(z, p, rop, rop_out, lu_match, st);
        if (z.typ == 1) return run_lu_lines<1, 6>(z, p, rop, rop_out, lu_match, st);
        return run_lu_lines<2, 6>(z, p, rop, rop_out, lu_match, st);
    }
    if (z.typ == 0) return run_lu_lines<0, 5>(z, p, rop, rop_out, lu_match, st);
    if (z.typ == 1) return run_lu_lines<1, 5>(z, p, rop, rop_out, lu_match, st);
    return run_lu_lines<2, 5>(z, p, rop, rop_out, lu_match, st);
}
