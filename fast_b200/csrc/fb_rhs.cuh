// fb_rhs.cuh -- fused rhs kernel templates; instantiated per (MESH, FLOW) in fb_rhs_m*f*.cu so that the 81 x 2
// variants compile in parallel.
//   k_rhs3: k-marching, state planes staged by TMA into a shared-memory ring (default whenever the zone's strides
//           can be encoded in a tensor map: ni and NDIMDX even);
//   k_rhs2: k-marching, state read through L1 (any zone size).
// Both evaluate every face flux once and share it between its two cells, and sum in the reference's order
// +F_k +F_j +F_i -F_i(i+1) -F_j(j+1) -F_k(k+1) (template_FluxAndBalance.for:192-271), so they agree bit for bit.
#pragma once
#include <cstdlib>
#include <type_traits>
#include <cuda.h>
#include "fb_kernels.cuh"
#include "fb_flux.cuh"
#include "fb_source.cuh"
#include "fb_rhs_decl.cuh"


static inline dim3 grid3r(int ni, int nj, int nk, dim3 b) { return dim3((ni + b.x - 1) / b.x, (nj + b.y - 1) / b.y, nk); }

// SENSOR byte store.  Written as PTX without a memory clobber: a C++ store through unsigned char* may alias everything,
// which makes the compiler re-load the staged state around it.
__device__ __forceinline__ void st_wig(unsigned char *p, unsigned v) { asm("st.global.u8 [%0], %1;" ::"l"(p), "r"(v)); }

// Halo-exchange overlap (RhsArgs::gx != 0): the stage is launched in two parts with 1-D grids, CTA c of a part working on
// tile rhs_face_first(tile0 + c) -- a bijection of the gx x gy x gz tile grid that puts the tiles touching a zone face with
// cross-rank joins first (their cells are the donors of the exchange).  Part 1 = the face tiles, part 2 = the interior tiles on another
// stream; the exchange starts behind an event recorded after part 1.  Pure arithmetic on kernel parameters, and nothing
// in the kernel tells anybody that a tile is finished: two earlier variants -- tile coordinates read from a table, and a
// per-tile completion count (fence + atomic, even predicated off) at the end of the kernel -- cost the TGV stage 40 %
// (1.0 -> 1.43 ms, profiles/r01_rhs_overlap_variants.md) with the overlap switched OFF.
#include "fb_tiles.cuh"   // rhs_face_first, rhs_face_flags, rhs_face_tiles
// which directions end in a tile of a single cell (the donors of a join are the two cell layers behind the face)
static inline int rhs_thin(const ZoneDev &z, dim3 g, int ti, int tj, int tk)
{
    return ((z.iv - (int)(g.x - 1) * ti < 2) ? 1 : 0) | ((z.jv - (int)(g.y - 1) * tj < 2) ? 2 : 0) | ((z.kv - (int)(g.z - 1) * tk < 2) ? 4 : 0);
}
// launch grid of a part (RhsArgs::part: 0 whole zone as a 3-D grid, 1 face tiles, 2 interior tiles); false = nothing to launch
static inline bool rhs_part_grid(RhsArgs &a, dim3 &g)
{
    a.gx = a.gy = a.gz = a.tile0 = 0;
    if (a.part == 0) return true;
    a.faces = (int)rhs_face_flags((unsigned)a.faces, (unsigned)a.thin, g.x, g.y, g.z);
    const unsigned nb = rhs_face_tiles(g.x, g.y, g.z, (unsigned)a.faces), total = g.x * g.y * g.z;
    a.gx = (int)g.x; a.gy = (int)g.y; a.gz = (int)g.z;
    a.tile0 = a.part == 1 ? 0 : (int)nb;
    const unsigned n = a.part == 1 ? nb : total - nb;
    g = dim3(n, 1, 1);
    return n > 0;
}

// ------------------------------------------------------------------------------------------------ wall flux correction
// S/BC/correct_flux.for:6-190 (lapply = 4: Roe scheme + wall windows), between the flux balance and the update
// (navier_stokes_struct.for:376-383): the interior Euler flux of a wall face is taken back out of the balance
// (template_correction.for:252-300, assemble_drodm_corrflu.for; only the five Euler components, iflow = 1) and the wall
// pressure flux put in (CL_correction_flu.for -> bflwall.for:96-131, bflwallslip.for:104-190; non-ALE: qen = 0).
// Only cells that touch a wall window do any work; they read the state through global memory.
template <int MESH, int SLOPE, int NEQ, int D>
__device__ __forceinline__ void wall_corr_dir(const ZoneDev &z, const GView &g, const CorrWin &c, int fi, int fj, int fk, double *acc)
{
    const int nd = z.nd;
    const int l = inddm(z, fi, fj, fk), lt = MESH == 2 ? 0 : indmtr(z, fi, fj, fk);
    const int ia = g.template dn<D>(l);
    const bool up = (c.idir % 2) == 0;
    double flu[5]; unsigned wg = 0;
    const GMet1 mt(z);
    face_flux<MESH, 1, 5, SLOPE, D>(z, g, mt, l, wg, false, lt, flu);
    const double sens_c = up ? 1. : -1., sens_b = up ? -1. : 1.;
#pragma unroll
    for (int e = 0; e < 5; e++) acc[e] += flu[e] * sens_c;
    double tc[3], s;
    face_normal<MESH, D>(z, mt, lt, tc, s);
    const double *rop = g.rop;
    double r = 0.5 * (rop[l] + rop[ia]), u = 0.5 * (rop[l + nd] + rop[ia + nd]);
    double v = 0.5 * (rop[l + 2 * nd] + rop[ia + 2 * nd]), w = 0.5 * (rop[l + 3 * nd] + rop[ia + 3 * nd]);
    const double u_int = tc[0] * u + tc[1] * v + tc[2] * w;
    double f[NEQ];
#pragma unroll
    for (int e = 0; e < NEQ; e++) f[e] = 0.;
    if (!c.slip) {
        const double p = 0.5 * (rop[l + 4 * nd] * rop[l] + rop[ia + 4 * nd] * rop[ia]) * z.rgp;
        f[1] = tc[0] * p + u_int * u * r; f[2] = tc[1] * p + u_int * v * r; f[3] = tc[2] * p + u_int * w * r;
    } else {
        if (NEQ == 6) { r = rop[ia]; u = rop[ia + nd]; v = rop[ia + 2 * nd]; w = rop[ia + 3 * nd]; }
        double p = rop[ia + 4 * nd] * rop[ia] * z.rgp;
        const double h = z.gam1 * p + r * (u * u + v * v + w * w) * 0.5;
        p = 0.5 * (rop[l + 4 * nd] * rop[l] + rop[ia + 4 * nd] * rop[ia]) * z.rgp;
        f[0] = u_int * r;
        f[1] = tc[0] * p + u_int * r * u; f[2] = tc[1] * p + u_int * r * v; f[3] = tc[2] * p + u_int * r * w;
        f[4] = u_int * h;
        if (NEQ == 6) f[NEQ - 1] = u_int * r * rop[ia + 5 * nd];
    }
#pragma unroll
    for (int e = 0; e < NEQ; e++) acc[e] += f[e] * sens_b;
}

template <int MESH, int SLOPE, int NEQ>
__device__ __noinline__ void wall_corr(const ZoneDev &z, const double *rop, int i, int j, int k, double *acc)
{
    const GView g(z, rop);
    for (int n = 0; n < z.ncorr; n++) {
        const CorrWin &c = z.corr[n];
        const int d = (c.idir - 1) / 2, up = (c.idir % 2) == 0 ? 1 : 0;
        const int fi = i + (d == 0 ? up : 0), fj = j + (d == 1 ? up : 0), fk = k + (d == 2 ? up : 0);   // cell whose lower face is the wall face
        if (fi < c.w[0] || fi > c.w[1] || fj < c.w[2] || fj > c.w[3] || fk < c.w[4] || fk > c.w[5]) continue;
        if (d == 0) wall_corr_dir<MESH, SLOPE, NEQ, 0>(z, g, c, fi, fj, fk, acc);
        else if (d == 1) wall_corr_dir<MESH, SLOPE, NEQ, 1>(z, g, c, fi, fj, fk, acc);
        else wall_corr_dir<MESH, SLOPE, NEQ, 2>(z, g, c, fi, fj, fk, acc);
    }
}

// ------------------------------------------------------------------------------------------------ cell epilogue
// explicit: core3ark3.for:108-154 (+203-236 for nutilde) -- the stage is updated from its own input
// (navier_stokes_struct.for:477-483); implicit: core3as2.for:70-100.  l: cell in the view, lg: global cell index.
template <int MESH, int FLOW, bool IMPL, class V>
__device__ __forceinline__ void rhs_finish(const ZoneDev &z, const RhsArgs &a, const V &v, int l, int lg, double coe1, double *acc,
                                           const double *npre = nullptr)
{
    constexpr int NEQ = FLOW == 3 ? 6 : 5;
    const int nd = z.nd;
    const double cv = z.cv;
    if (!IMPL) {
#pragma unroll
        for (int e = 0; e < NEQ; e++) z.drodm[lg + e * nd] = acc[e];
        double r_1 = a.coefW * coe1;
        const double ro_old = v.q(0, l), u_old = v.q(1, l), v_old = v.q(2, l), w_old = v.q(3, l), t_old = v.q(4, l);
        const double tab1 = acc[0] * r_1, tab2 = acc[1] * r_1, tab3 = acc[2] * r_1, tab4 = acc[3] * r_1, tab5 = acc[4] * r_1;
        const double cvinv = 1. / cv, cvinv2 = 0.5 * cvinv;
        const double rn = ro_old + tab1;
        const double r1n = 1. / rn;
        const double un = (ro_old * u_old + tab2) * r1n, vn = (ro_old * v_old + tab3) * r1n, wn = (ro_old * w_old + tab4) * r1n;
        const double roe_old = ro_old * (cv * t_old + 0.5 * (u_old * u_old + v_old * v_old + w_old * w_old));
        const double tn = (roe_old + tab5) * r1n * cvinv - cvinv2 * (un * un + vn * vn + wn * wn);
        double *__restrict__ out = a.rop_out;
        out[lg] = rn; out[lg + nd] = un; out[lg + 2 * nd] = vn; out[lg + 3 * nd] = wn; out[lg + 4 * nd] = tn;
        if (FLOW == 3) {
            const double tab6 = acc[NEQ - 1] * r_1;
            double anulam = z.coesut * sqrt(tn * z.temp01) / (1. + z.cmus1 / tn);
            anulam = anulam * r1n * z.ratiom;
            double t = (ro_old * v.q(5, l) + tab6) * r1n;
            t = fmin(t, anulam);
            out[lg + 5 * nd] = fmax(t, 0.);
        }
    } else {
        implicit_assemble<NEQ>(z, v, l, a.rop, a.ro_n, a.ro_n1, lg, coe1, a.c1, a.c2, a.c3, acc, npre);
#pragma unroll
        for (int e = 0; e < NEQ; e++) z.drodm[lg + e * nd] = acc[e];
    }
}

// ------------------------------------------------------------------------------------------------ fused rhs, L1 path
// k-marching kernel: every face flux is evaluated ONCE and shared by its two cells.
//   CTA = 32 x (RJ+1) threads over a tile of 31 x RJ cells, marching over a chunk of k planes.
//   thread (li, m): lower i face and lower j face of cell (i0+li, j0+m, k) and its lower k face;
//     F_i(i+1) comes from lane li+1 (warp shuffle; lane 31 only owns a face), F_j(j+1) from row m+1 (shared memory,
//     double-buffered by plane parity: one __syncthreads per plane; row RJ only owns faces), F_k(k+1) is the lower k
//     face of the next plane: the cell of plane k-1 is finished (and written) when plane k has produced it.
#define RHS2_RJ 8
#define RHS2_KC 32

template <int MESH, int FLOW, int SCHEME, int SLOPE, bool IMPL, bool WC>
__global__ void __launch_bounds__(32 * (RHS2_RJ + 1)) k_rhs2(ZoneDev z, RhsArgs a)
{
    constexpr int NEQ = FLOW == 3 ? 6 : 5, RJ = RHS2_RJ;
    // (unsigned on purpose, like blockIdx: with signed tile coordinates the address arithmetic of the whole kernel grows by a quarter)
    unsigned bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z;
    if (!IMPL && a.gx) rhs_face_first(blockIdx.x + (unsigned)a.tile0, a.gx, a.gy, a.gz, a.faces, bx, by, bz);   // explicit stages only
    __shared__ double sFj[2][RJ + 1][NEQ][32];
    const int li = threadIdx.x, m = threadIdx.y;
    const int i = 1 + bx * 31 + li, j = 1 + by * RJ + m;
    const int k0 = 1 + bz * RHS2_KC, k1 = min(k0 + RHS2_KC - 1, z.kv);
    const bool cell = li < 31 && m < RJ && i <= z.iv && j <= z.jv;      // owns a cell
    const bool fi = m < RJ && j <= z.jv && i <= z.iv + 1;               // owns a lower i face that somebody needs
    const bool fj = li < 31 && i <= z.iv && j <= z.jv + 1;              // owns a lower j face that somebody needs
    const int nd = z.nd;
    constexpr bool wcomp = WC;   // first sub-iteration: the SENSOR switch is evaluated in the flux and stored
    const GView v(z, a.rop);
    const GMet1 mt(z);
    double acc[NEQ], f[NEQ];
#pragma unroll
    for (int e = 0; e < NEQ; e++) acc[e] = 0.;
    double coe1p = 0.;        // dt/vol of the cell being finished (plane k-1)
    for (int k = k0; k <= k1 + 1; k++) {
        const int l = inddm(z, i, j, k);
        const int lt = MESH == 2 ? 0 : indmtr(z, i, j, k);
        const bool last = k == k1 + 1;
        // early global loads of the cell of plane k: consumed after the k-face flux, which hides their latency
        double accn[NEQ], coe1 = 0.;
#pragma unroll
        for (int e = 0; e < NEQ; e++) accn[e] = 0.;
        if (cell && !last) {
            if (!(IMPL || a.nitcfg == 1)) {
#pragma unroll
                for (int e = 0; e < NEQ; e++) accn[e] = FB_LD(z.drodm + l + e * nd);
            }
            if (IMPL) coe1 = FB_LD(z.coe + l);
            else coe1 = MESH == 2 ? z.dtc / z.vol0 : z.dtc / __ldg(z.vol + lt);
        }
        unsigned wg = (SCHEME == 2 && !wcomp && (cell || fi || fj)) ? z.wigb[l] : 0u;
        // lower k face: finishes the cell of plane k-1, starts the cell of plane k
        if (cell) {
            face_flux<MESH, FLOW, SCHEME, SLOPE, 2>(z, v, mt, l, wg, wcomp, lt, f);
            if (k > k0) {
#pragma unroll
                for (int e = 0; e < NEQ; e++) acc[e] -= f[e];
                if (SCHEME == 5 && z.ncorr) wall_corr<MESH, SLOPE, NEQ>(z, a.rop, i, j, k - 1, acc);
                rhs_finish<MESH, FLOW, IMPL>(z, a, v, l - z.inck, l - z.inck, coe1p, acc);
            }
        }
        if (last) {           // uniform
            if (wcomp && cell && k == z.kv + 1) st_wig(z.wigb + l, wg);   // k faces of the plane behind the last cells
            break;
        }
        if (cell) {
            // init_rhs.for:85-129
#pragma unroll
            for (int e = 0; e < NEQ; e++) acc[e] = (IMPL || a.nitcfg == 1) ? 0. : accn[e] * a.coefH;
            if (FLOW == 3) sa_source<MESH, IMPL>(z, v, mt, l, l, lt, a.nitcfg, coe1, acc);
#pragma unroll
            for (int e = 0; e < NEQ; e++) acc[e] += f[e];
        }
        coe1p = coe1;
        // lower j face -> shared memory
        if (fj) {
            face_flux<MESH, FLOW, SCHEME, SLOPE, 1>(z, v, mt, l, wg, wcomp, lt, f);
#pragma unroll
            for (int e = 0; e < NEQ; e++) sFj[k & 1][m][e][li] = f[e];
            if (cell) {
#pragma unroll
                for (int e = 0; e < NEQ; e++) acc[e] += f[e];
            }
        }
        // lower i face; the upper one comes from the next lane
        if (fi) face_flux<MESH, FLOW, SCHEME, SLOPE, 0>(z, v, mt, l, wg, wcomp, lt, f);
        else {
#pragma unroll
            for (int e = 0; e < NEQ; e++) f[e] = 0.;
        }
#pragma unroll
        for (int e = 0; e < NEQ; e++) {
            const double up = __shfl_down_sync(0xffffffffu, f[e], 1);
            if (cell) acc[e] = (acc[e] + f[e]) - up;
        }
        // SENSOR byte of the cell for the later stages; face-only threads store where no tile owns the cell
        if (wcomp && (cell || (fi && i == z.iv + 1) || (fj && j == z.jv + 1))) st_wig(z.wigb + l, wg);
        __syncthreads();
        if (cell) {
#pragma unroll
            for (int e = 0; e < NEQ; e++) acc[e] -= sFj[k & 1][m + 1][e][li];
        }
    }
}

// ------------------------------------------------------------------------------------------------ fused rhs, TMA path
// Same thread <-> face mapping as k_rhs2, but the primitive state (and mu+mut) of the tile, halo included, is staged
// plane by plane into a shared-memory ring by TMA (one cp.async.bulk.tensor box of 36 x (RJ+4) x 1 x NEQ doubles per
// plane and array, completion on an mbarrier).  Iteration k works on planes k-1, k, k+1 (the own column at k-2, which
// only the k-face MUSCL stencil needs, is carried in registers) while plane k+2 is in flight; the slot of plane k-1
// is re-armed for plane k+3 behind the end-of-plane barrier (4-slot ring; the 3-slot ring of the 3dhomo SA kernel asks for
// plane k+2 there and has sent plane k+3 to L2).  All stencil reads are LDS with immediate offsets: no
// address arithmetic, no L1 tag traffic, HBM sees every plane once per tile (+ halo, served by L2).
#ifndef RHS3_NS
#define RHS3_NS 4           // ring slots
#endif
#ifndef RHS3_RJ_NS
#define RHS3_RJ_NS 8        // tile rows, Euler / laminar
#endif
#ifndef RHS3_RJ_SA
#define RHS3_RJ_SA 7        // tile rows, Spalart-Allmaras (7 staged variables): 8 warps of 128 registers, so that two CTAs fill the
                            // four 16 K-register sub-partitions of an SM exactly (4 warps each); RJ = 6 left a warp slot empty
#endif
#ifndef RHS3_MINB
#define RHS3_MINB 2         // resident CTAs per SM the register allocation aims at
#endif
#ifndef RHS3_SMET_SA
#define RHS3_SMET_SA 1      // 3dhomo SA zones: 1 = k-invariant metric tile in shared memory (19 KB per CTA), paid for with one ring
                            // slot (RHS3_NS_SA = 3) so that two CTAs of RJ = 7 stay resident; 0 = metrics through L1 and a 4-slot ring
                            // (measured on one B200: C5 rhs 0.752 -> 0.673 ms per block, plate 1.52 -> 1.35 ms with the tile)
#endif
#ifndef RHS3_NS_SA
#define RHS3_NS_SA (RHS3_SMET_SA ? 3 : RHS3_NS)
#endif
#ifndef RHS3_NS_HOMO
#define RHS3_NS_HOMO 3      // ring slots of the 3dhomo Euler / laminar kernels: with the metric tile, four slots leave one CTA per SM
                            // (measured on a 256x128x64 laminar plate, tools/homo_lam_time.py: rhs 0.236 -> 0.198 ms per launch with three)
#endif
#ifndef RHS3_RJ_1B
#define RHS3_RJ_1B 11       // tile rows of the kernels that run one CTA per SM (Roe, 3dfull: 168 registers): 12 warps fill the
                            // register file (measured: C4 rhs 0.625 -> 0.550 ms per 128^3 block against 8 / 7 rows); 0 = as the others
#endif
#ifndef RHS3_NSF_SA
#define RHS3_NSF_SA 2       // F_j exchange buffers of the 3dhomo SA kernel (1 = second barrier per plane, 12 KB less shared memory)
#endif
#ifndef RHS3_KC
#define RHS3_KC 32          // k planes per CTA (measured: 32 < 64 < 128 on TGV 256^3 and the plate)
#endif
#ifndef RHS3_KC_SA
#define RHS3_KC_SA 16       // ... of the 3dhomo SA kernel (measured: C5 rhs 0.674 -> 0.664 ms per block, plate 1.349 -> 1.290 ms)
#endif
#ifndef RHS3_KC_1B
#define RHS3_KC_1B RHS3_KC  // ... of the kernels that run one CTA per SM
#endif

template <int MESH, int FLOW> struct Rhs3SMet { static constexpr bool V = MESH == 1 && (FLOW != 3 || RHS3_SMET_SA != 0); };
// register budget: two resident CTAs per SM where the flux fits 96-128 registers; Roe and the 3dfull metrics need more
template <int MESH, int SCHEME> struct Rhs3MinB { static constexpr int V = (SCHEME == 5 || MESH == 0) ? 1 : RHS3_MINB; };
template <int MESH, int FLOW, int SCHEME> struct Rhs3NS {   // ring slots (a kernel alone on its SM has room for the deeper ring)
    static constexpr int V = (MESH == 1 && Rhs3MinB<MESH, SCHEME>::V > 1) ? (FLOW == 3 ? RHS3_NS_SA : RHS3_NS_HOMO) : RHS3_NS;
};
template <bool SM, int RS, int VS>
__device__ __forceinline__ typename std::conditional<SM, SMet1<RS, VS>, GMet1>::type make_met(const ZoneDev &z, const double *smet)
{
    if constexpr (SM) { SMet1<RS, VS> m; m.s = smet; return m; }
    else return GMet1(z);
}

template <int FLOW, int RJ>
struct Rhs3Cfg {
    static constexpr int NEQ = FLOW == 3 ? 6 : 5;
    static constexpr int NV = NEQ + (FLOW >= 2 ? 1 : 0);
    static constexpr int BI = 36, ROWS = RJ + 4;          // box: 2 halo columns / rows on each side (+1 face lane, +1 pad)
    static constexpr int RS = BI, VS = RS * ROWS;
    static constexpr int MUOFF = (NEQ * VS + 15) / 16 * 16;                    // TMA destinations must be 128-byte aligned
    static constexpr int PS = (MUOFF + (NV - NEQ) * VS + 15) / 16 * 16;       // slot stride
    static constexpr int NT = 32 * (RJ + 1);
    static constexpr unsigned TX = (unsigned)(NV * VS * sizeof(double));   // bytes one plane brings in
    // F_j exchange: double-buffered by plane parity (one barrier per plane); 3dhomo SA keeps a single buffer and pays a
    // second barrier per plane so that ring + metric tile leave room for two CTAs per SM
    template <int MESH> __host__ __device__ static constexpr int nsf() { return (MESH == 1 && FLOW == 3) ? RHS3_NSF_SA : 2; }
    template <int MESH> __host__ __device__ static constexpr int sfj() { return nsf<MESH>() * (RJ + 1) * NEQ * 32; }
    template <int MESH> __host__ __device__ static constexpr int met() { return Rhs3SMet<MESH, FLOW>::V ? 6 * VS : 0; }   // 3dhomo: metric tile (k-invariant)
    template <int MESH, int SCHEME> __host__ __device__ static constexpr size_t smem() { return (size_t)(Rhs3NS<MESH, FLOW, SCHEME>::V * PS + sfj<MESH>() + met<MESH>()) * sizeof(double) + Rhs3NS<MESH, FLOW, SCHEME>::V * sizeof(unsigned long long); }
};

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(b), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <int MESH, int FLOW, int SCHEME> struct Rhs3KC {   // k planes per CTA
    static constexpr int V = Rhs3MinB<MESH, SCHEME>::V == 1 ? RHS3_KC_1B : ((MESH == 1 && FLOW == 3) ? RHS3_KC_SA : RHS3_KC);
};

// the box of a plane on its way to L2 (no shared-memory destination, no completion): used one plane ahead of the 3-slot ring
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap *tm, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <int MESH, int FLOW, int SCHEME, int SLOPE, bool IMPL, int RJ, bool WC>
__global__ void __launch_bounds__(32 * (RJ + 1), Rhs3MinB<MESH, SCHEME>::V)
k_rhs3(const CUtensorMap *__restrict__ tm_rop, const CUtensorMap *__restrict__ tm_mu, ZoneDev z, RhsArgs a)
{
    using C = Rhs3Cfg<FLOW, RJ>;
    constexpr int NEQ = C::NEQ, NS = Rhs3NS<MESH, FLOW, SCHEME>::V, PS = C::PS, RS = C::RS, VS = C::VS;
    // (unsigned on purpose, like blockIdx: with signed tile coordinates the address arithmetic of the whole kernel grows by a quarter)
    unsigned bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z;
    if (!IMPL && a.gx) rhs_face_first(blockIdx.x + (unsigned)a.tile0, a.gx, a.gy, a.gz, a.faces, bx, by, bz);   // explicit stages only
    // the rhs kernel of the next zone (launched with the programmatic-serialization attribute) may be scheduled as soon as
    // every CTA of this grid has got here or finished, i.e. while the last wave drains
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ __align__(128) double ring[];
    double (*sFj)[RJ + 1][NEQ][32] = (double (*)[RJ + 1][NEQ][32])(ring + NS * PS);
    constexpr int NSF = C::template nsf<MESH>();
    double *smet = ring + NS * PS + C::template sfj<MESH>();   // 3dhomo: [ti.x, ti.y, tj.x, tj.y, tk, vol][rows][36]
    unsigned long long *bar = (unsigned long long *)(smet + C::template met<MESH>());
    const int li = threadIdx.x, m = threadIdx.y;
    const bool leader = li == 0 && m == 0;
    const int i = 1 + bx * 31 + li, j = 1 + by * RJ + m;
    constexpr int KC = Rhs3KC<MESH, FLOW, SCHEME>::V;
    const int k0 = 1 + bz * KC, k1 = min(k0 + KC - 1, z.kv);
    const bool cell = li < 31 && m < RJ && i <= z.iv && j <= z.jv;
    const bool fi = m < RJ && j <= z.jv && i <= z.iv + 1;
    const bool fj = li < 31 && i <= z.iv && j <= z.jv + 1;
    const int nd = z.nd;
    constexpr bool wcomp = WC;   // first sub-iteration: the SENSOR switch is evaluated in the flux and stored
    // array coordinates of the box origin (cell i0-2, j0-2) and planes k0-1 .. k1+2
    // (TMA needs a 16-byte aligned global start: an odd first column is shifted one to the left -- the 36-wide box
    // still covers the 35 columns i0-2 .. i0+32 the tile reads)
    const int ca0 = bx * 31 + z.nf - 2, cb = by * RJ + z.nf - 2;
    const int sh = ca0 & 1, ca = ca0 - sh;
    const int pfirst = k0 - 1, plast = k1 + 2;
    auto issue = [&](int p) {
        const int slot = (p - pfirst) % NS;
        mbar_expect_tx(bar + slot, C::TX);
        tma_load_4d(ring + slot * PS, tm_rop, ca, cb, p + z.nfk - 1, 0, bar + slot);
        if (FLOW >= 2) tma_load_4d(ring + slot * PS + C::MUOFF, tm_mu, ca, cb, p + z.nfk - 1, 0, bar + slot);
    };
    auto prefetch = [&](int p) {
        tma_prefetch_4d(tm_rop, ca, cb, p + z.nfk - 1, 0);
        if (FLOW >= 2) tma_prefetch_4d(tm_mu, ca, cb, p + z.nfk - 1, 0);
    };
    if (leader) {
        for (int s = 0; s < NS; s++) mbar_init(bar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (Rhs3SMet<MESH, FLOW>::V) {
        // the metrics of a 3dhomo zone do not depend on k: the tile's neighbourhood is copied once and serves every plane
        const int tid = m * 32 + li, mlast = z.nm - 1;
#pragma unroll
        for (int it = 0; it < (6 * VS + C::NT - 1) / C::NT; it++) {      // independent loads, all in flight together
            const int n = tid + it * C::NT;
            if (n >= 6 * VS) break;
            const int arr = n / VS, r = (n % VS) / RS, c = n % RS;
            int lt = indmtr(z, 1 + bx * 31 - 2 - sh + c, 1 + by * RJ - 2 + r, 1);
            lt = lt < 0 ? 0 : (lt > mlast ? mlast : lt);        // outside the arrays: never used, any finite value will do
            const double *src = arr < 2 ? z.ti + arr * z.nm : (arr < 4 ? z.tj + (arr - 2) * z.nm : (arr == 4 ? z.tk : z.vol));
            smet[n] = __ldg(src + lt);
        }
    }
    __syncthreads();
    if (leader)
        for (int p = pfirst; p < pfirst + NS && p <= plast; p++) issue(p);
    // own column at plane k0-2 (registers)
    double km2[NEQ];
    {
        const int l = inddm(z, cell ? i : 1, cell ? j : 1, k0 - 2);
#pragma unroll
        for (int e = 0; e < NEQ; e++) km2[e] = FB_LD(a.rop + l + e * nd);
    }
    double acc[NEQ], f[NEQ];
#pragma unroll
    for (int e = 0; e < NEQ; e++) acc[e] = 0.;
    double coe1p = 0.;
    // implicit_assemble reads W(n) unless rop is W(n) itself (first sub-iteration); loaded ahead of the flux
    const bool ldn = IMPL && a.ro_n != a.rop;
    mbar_wait(bar + 0, 0);
    mbar_wait(bar + 1, 0);
    const int lcol = (m + 2) * RS + li + 2 + sh;
    typename std::conditional<Rhs3SMet<MESH, FLOW>::V, SMet1<RS, VS>, GMet1>::type mt = make_met<Rhs3SMet<MESH, FLOW>::V, RS, VS>(z, smet);
    for (int k = k0; k <= k1 + 1; k++) {
        const int pk = k - pfirst;                       // >= 1
        const int sk = pk % NS, su = (pk + 1) % NS, sd = (pk - 1) % NS;
        // early global loads of the cell of plane k (old rhs for the RK accumulation, dt/vol, SENSOR byte): issued
        // before the wait and the k-face flux, which hide their latency
        double accn[NEQ], coe1 = 0., distk = -1., wn[NEQ];
#pragma unroll
        for (int e = 0; e < NEQ; e++) { accn[e] = 0.; wn[e] = 0.; }
        {
            const int lg = inddm(z, i, j, k);
            if (IMPL && ldn && cell && k > k0) {       // W(n) of cell k-1, finished in this iteration
#pragma unroll
                for (int e = 0; e < NEQ; e++) wn[e] = FB_LD(a.ro_n + lg - z.inck + e * nd);
            }
            if (cell && k <= k1) {
                if (FLOW == 3) distk = z.xmut[lg + nd];
                if (!(IMPL || a.nitcfg == 1)) {
#pragma unroll
                    for (int e = 0; e < NEQ; e++) accn[e] = FB_LD(z.drodm + lg + e * nd);
                }
                if (IMPL) coe1 = FB_LD(z.coe + lg);
                else coe1 = MESH == 2 ? z.dtc / z.vol0 : z.dtc / __ldg(z.vol + indmtr(z, i, j, k));
            }
        }
        unsigned wg = (SCHEME == 2 && !wcomp && (cell || fi || fj)) ? z.wigb[inddm(z, i, j, k)] : 0u;
        mbar_wait(bar + su, ((pk + 1) / NS) & 1);        // plane k+1 has landed
        SView<RS, VS, NEQ, C::MUOFF> v;
        v.s = ring; v.kup = (su - sk) * PS; v.kdn = (sd - sk) * PS; v.km2 = km2;
        const int ls = sk * PS + lcol;
        const int l = inddm(z, i, j, k);
        const int lt = MESH == 2 ? 0 : (Rhs3SMet<MESH, FLOW>::V ? lcol : indmtr(z, i, j, k));   // 3dhomo: index in the shared metric tile
        const bool last = k == k1 + 1;
        if (cell) {
            face_flux<MESH, FLOW, SCHEME, SLOPE, 2>(z, v, mt, ls, wg, wcomp, lt, f);
            if (k > k0) {
#pragma unroll
                for (int e = 0; e < NEQ; e++) acc[e] -= f[e];
                if (SCHEME == 5 && z.ncorr) wall_corr<MESH, SLOPE, NEQ>(z, a.rop, i, j, k - 1, acc);
                rhs_finish<MESH, FLOW, IMPL>(z, a, v, ls + v.kdn, l - z.inck, coe1p, acc, (IMPL && ldn) ? wn : nullptr);
            }
        }
        if (last) {           // uniform; no TMA is in flight any more (plane k1+2 was waited for above)
            if (wcomp && cell && k == z.kv + 1) st_wig(z.wigb + l, wg);
            break;
        }
        if (cell) {
#pragma unroll
            for (int e = 0; e < NEQ; e++) acc[e] = (IMPL || a.nitcfg == 1) ? 0. : accn[e] * a.coefH;
            if (FLOW == 3) sa_source<MESH, IMPL>(z, v, mt, ls, l, lt, a.nitcfg, coe1, acc, distk);
#pragma unroll
            for (int e = 0; e < NEQ; e++) acc[e] += f[e];
        }
        coe1p = coe1;
        if (fj) {
            face_flux<MESH, FLOW, SCHEME, SLOPE, 1>(z, v, mt, ls, wg, wcomp, lt, f);
#pragma unroll
            for (int e = 0; e < NEQ; e++) sFj[NSF == 2 ? (k & 1) : 0][m][e][li] = f[e];
            if (cell) {
#pragma unroll
                for (int e = 0; e < NEQ; e++) acc[e] += f[e];
            }
        }
        if (fi) face_flux<MESH, FLOW, SCHEME, SLOPE, 0>(z, v, mt, ls, wg, wcomp, lt, f);
        else {
#pragma unroll
            for (int e = 0; e < NEQ; e++) f[e] = 0.;
        }
#pragma unroll
        for (int e = 0; e < NEQ; e++) {
            const double up = __shfl_down_sync(0xffffffffu, f[e], 1);
            if (cell) acc[e] = (acc[e] + f[e]) - up;
        }
        if (wcomp && (cell || (fi && i == z.iv + 1) || (fj && j == z.jv + 1))) st_wig(z.wigb + l, wg);
        // plane k-1 becomes plane k-2 of the next iteration
#pragma unroll
        for (int e = 0; e < NEQ; e++) km2[e] = v.q(e, ls + v.kdn);
        __syncthreads();
        // every thread is done with plane k-1: its slot takes plane k+3
        if (leader && k + NS - 1 <= plast) issue(k + NS - 1);
        if (NS < 4 && leader && k + NS <= plast) prefetch(k + NS);   // short ring: the box after that one is on its way to L2
        if (cell) {
#pragma unroll
            for (int e = 0; e < NEQ; e++) acc[e] -= sFj[NSF == 2 ? (k & 1) : 0][m + 1][e][li];
        }
        if (NSF == 1) __syncthreads();      // single F_j buffer: everybody has read it before the next plane overwrites it
    }
}

// ------------------------------------------------------------------------------------------------ launchers
// tensor maps (in device memory) of rop(ni,nj,nk,neq) / xmut(ni,nj,nk,naux) with a (36, rows, 1, nvar) box; nullptr when the
// strides cannot be encoded (odd ni or NDIMDX) or the driver entry point is missing (fb_tma.cu)
const CUtensorMap *fb_tma_map(const double *base, int ni, int nj, int nk, long long var_stride, int nvar_total, int box_rows, int box_vars);

template <int MESH, int FLOW, int SCHEME, int SLOPE>
static void launch_rhs2_t(const ZoneDev &z, RhsArgs &a, bool impl, cudaStream_t st)
{
    dim3 b(32, RHS2_RJ + 1, 1);
    dim3 g((z.iv + 30) / 31, (z.jv + RHS2_RJ - 1) / RHS2_RJ, (z.kv + RHS2_KC - 1) / RHS2_KC);
    a.thin = rhs_thin(z, g, 31, RHS2_RJ, RHS2_KC);
    if (!rhs_part_grid(a, g)) return;
    // WC: first sub-iteration of a SENSOR zone -- the switch is evaluated inside the flux and stored
    if (SCHEME == 2 && a.nitcfg == 1 && a.wfuse) {
        if (impl) k_rhs2<MESH, FLOW, SCHEME, SLOPE, true, SCHEME == 2><<<g, b, 0, st>>>(z, a);
        else k_rhs2<MESH, FLOW, SCHEME, SLOPE, false, SCHEME == 2><<<g, b, 0, st>>>(z, a);
    } else {
        if (impl) k_rhs2<MESH, FLOW, SCHEME, SLOPE, true, false><<<g, b, 0, st>>>(z, a);
        else k_rhs2<MESH, FLOW, SCHEME, SLOPE, false, false><<<g, b, 0, st>>>(z, a);
    }
}

template <int MESH, int FLOW, int SCHEME, int SLOPE>
static bool launch_rhs3_t(const ZoneDev &z, RhsArgs &a, bool impl, cudaStream_t st)
{
    constexpr int RJ = (RHS3_RJ_1B && Rhs3MinB<MESH, SCHEME>::V == 1) ? RHS3_RJ_1B : (FLOW == 3 ? RHS3_RJ_SA : RHS3_RJ_NS);
    using C = Rhs3Cfg<FLOW, RJ>;
    // the configuration macros above must leave the residency they were measured with: 227 KB per CTA, 228 KB (+ 1 KB per CTA) per SM
    static_assert(C::template smem<MESH, SCHEME>() <= 227 * 1024, "k_rhs3: shared memory of one CTA");
    static_assert(Rhs3MinB<MESH, SCHEME>::V * (C::template smem<MESH, SCHEME>() + 1024) <= 228 * 1024, "k_rhs3: shared memory of the resident CTAs");
    static_assert(Rhs3MinB<MESH, SCHEME>::V * 32 * (RJ + 1) * (Rhs3MinB<MESH, SCHEME>::V == 1 ? 168 : 128) <= 65536 || FLOW != 3, "k_rhs3: registers of the resident CTAs");
    if (z.nf != 2 || z.nfk != 2) return false;
    const CUtensorMap *tr = fb_tma_map(a.rop, z.ni, z.nj, z.nk, z.nd, C::NEQ, C::ROWS, C::NEQ);
    const CUtensorMap *tm = FLOW >= 2 ? fb_tma_map(z.xmut, z.ni, z.nj, z.nk, z.nd, FLOW == 3 ? 2 : 1, C::ROWS, 1) : tr;
    if (!tr || !tm) return false;
    static std::atomic<unsigned long long> attr{0};
    if (fb_first_on_device(attr)) {
        cudaFuncSetAttribute(k_rhs3<MESH, FLOW, SCHEME, SLOPE, true, RJ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::template smem<MESH, SCHEME>());
        cudaFuncSetAttribute(k_rhs3<MESH, FLOW, SCHEME, SLOPE, false, RJ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::template smem<MESH, SCHEME>());
        cudaFuncSetAttribute(k_rhs3<MESH, FLOW, SCHEME, SLOPE, true, RJ, SCHEME == 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::template smem<MESH, SCHEME>());
        cudaFuncSetAttribute(k_rhs3<MESH, FLOW, SCHEME, SLOPE, false, RJ, SCHEME == 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::template smem<MESH, SCHEME>());
    }
    dim3 b(32, RJ + 1, 1);
    constexpr int KC = Rhs3KC<MESH, FLOW, SCHEME>::V;
    dim3 g((z.iv + 30) / 31, (z.jv + RJ - 1) / RJ, (z.kv + KC - 1) / KC);
    a.thin = rhs_thin(z, g, 31, RJ, KC);
    if (!rhs_part_grid(a, g)) return true;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = g; cfg.blockDim = b; cfg.dynamicSmemBytes = C::template smem<MESH, SCHEME>(); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = a.pdl ? 1 : 0;
    if (SCHEME == 2 && a.nitcfg == 1 && a.wfuse) {   // first sub-iteration of a SENSOR zone: the switch is evaluated in the flux
        if (impl) cudaLaunchKernelEx(&cfg, k_rhs3<MESH, FLOW, SCHEME, SLOPE, true, RJ, SCHEME == 2>, tr, tm, z, a);
        else cudaLaunchKernelEx(&cfg, k_rhs3<MESH, FLOW, SCHEME, SLOPE, false, RJ, SCHEME == 2>, tr, tm, z, a);
    } else {
        if (impl) cudaLaunchKernelEx(&cfg, k_rhs3<MESH, FLOW, SCHEME, SLOPE, true, RJ, false>, tr, tm, z, a);
        else cudaLaunchKernelEx(&cfg, k_rhs3<MESH, FLOW, SCHEME, SLOPE, false, RJ, false>, tr, tm, z, a);
    }
    return true;
}

extern long long g_fb_rhs_launches[2];

template <int MESH, int FLOW, int SCHEME, int SLOPE>
static void launch_rhs_t(const ZoneDev &z, RhsArgs &a, bool impl, cudaStream_t st)
{
    static const bool l1path = [] { const char *e = getenv("FAST_B200_RHS"); return e && e[0] == '2'; }();   // A/B switch
    if (!l1path && launch_rhs3_t<MESH, FLOW, SCHEME, SLOPE>(z, a, impl, st)) { g_fb_rhs_launches[0]++; return; }
    launch_rhs2_t<MESH, FLOW, SCHEME, SLOPE>(z, a, impl, st);
    g_fb_rhs_launches[1]++;
}
template <int MESH, int FLOW, int SCHEME>
static void launch_rhs_s(const ZoneDev &z, RhsArgs &a, bool impl, cudaStream_t st)
{
    switch (z.slope) {
    case 2: launch_rhs_t<MESH, FLOW, SCHEME, 2>(z, a, impl, st); break;
    case 3: launch_rhs_t<MESH, FLOW, SCHEME, 3>(z, a, impl, st); break;
    default: launch_rhs_t<MESH, FLOW, SCHEME, 1>(z, a, impl, st); break;
    }
}
template <int MESH, int FLOW>
void launch_rhs_k(const ZoneDev &z, RhsArgs &a, bool impl, cudaStream_t st)
{
    switch (z.kflu) {
    case 1: launch_rhs_s<MESH, FLOW, 1>(z, a, impl, st); break;
    case 2: launch_rhs_s<MESH, FLOW, 2>(z, a, impl, st); break;
    default: launch_rhs_s<MESH, FLOW, 5>(z, a, impl, st); break;
    }
}
