// fb_api.cu -- C ABI (include/fastb200.h) and the sub-iteration driver.
//
// The driver restates the ORDER of S/Compute/gsdr3.cpp:152-444 and computePT.cpp:534-679 for one task per zone:
//   rhs of every zone -> [lu_match: drodm transfers] -> BC on rhs + corners -> residual norms -> LU + Newton update
//   -> match-join transfers on rop_CL -> BC ghost fill -> corners, then time-level rotation.
// Everything is enqueued on one CUDA stream per solver (zones are independent inside a phase; phases are ordered by
// the stream), no host synchronisation inside fb_compute.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cmath>
#include <dlfcn.h>
#include "../../include/fastb200.h"
#include "fb_kernels.cuh"
#include "fb_implicit.cuh"
#include "fb_tiles.cuh"

static thread_local std::string g_err;
static int fail(int code, const std::string &m) { g_err = m; return code; }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(FB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

struct Zone {
    std::vector<int> pi; std::vector<double> pr;
    ZoneDev d;
    double *rop[3] = {nullptr, nullptr, nullptr};   // ro, ro_M1, ro_P1 (rotated on the host side of the handle)
    double *xyz = nullptr, *mtr = nullptr, *aux = nullptr, *drodm = nullptr, *coe = nullptr;
    unsigned char *wigb = nullptr;
    double *rdm = nullptr;          // device: nssiter*neq*2
    double *cflv = nullptr;         // device: CFL max, min, sum (cpcfl3) + number of non-finite cells / residuals seen by the checks
    BcWin *wins = nullptr, *wins_rhs = nullptr;
    int nwin = 0, nwin_rhs = 0, maxc = 0, maxc_rhs = 0;
    int lu_box[6];                  // ishift_lu (BCzone.cpp:136-143)
    int lcorner = 1;
    LuPlan lu;                      // hyperplane bookkeeping (fb_implicit.cu)
    int xfaces = 0;                 // faces whose interior cells are donors of cross-rank joins (bit 0..5 = imin .. kmax)
    bool has_params = false, has_xyz = false;
    cudaEvent_t ev_up = nullptr; bool up_pending = false, up_pending_first = false;   // fb_compute_host: this zone's input levels are still on their way (copy stream)
    size_t n_state = 0, n_aux = 0;
};

struct Timer { cudaEvent_t a, b; double ms = 0; long long n = 0; std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pend; };

struct fb_solver {
    int dev = 0, nz = 0;
    cudaStream_t st = nullptr;
    cudaStream_t st_h2d = nullptr;    // fb_compute_host: uploads of the zones' input levels, overlapped with the first kernels of the zones already there
    cudaStream_t st2 = nullptr;       // cross-rank halo exchange (pack, host transport, unpack): overlaps the interior tiles
    cudaEvent_t ev1 = nullptr, ev2 = nullptr;
    // multi-zone solvers: the per-zone kernels of a phase fan out over st + 3 lanes so that the tail of one zone's grid
    // (a 128^3 block is 1.08 waves of rhs CTAs) is filled by the next zone's CTAs; the phase joins back on st
    cudaStream_t lane[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_lane[3] = {nullptr, nullptr, nullptr};
    std::vector<Zone> z;
    bool metric_done = false;
    JoinSet joins;
    fb_exchange_fn xcb = nullptr; void *xuser = nullptr;   // cross-rank halo exchange through a host transport (fb_set_exchange_cb)
    void *nccl = nullptr;                                  // ... or natively: ncclComm_t created by fb_nccl_init
    struct Peer { int rank; long long so, sn, ro, rn; };
    std::vector<Peer> peers;                               // fb_set_peers: slices of the send / receive buffers per neighbour rank
    double *halo_send = nullptr, *halo_recv = nullptr;
    bool timing = false;
    Timer tm[6];     // rhs, lu, bc, transfer, coe, all
    long long launches = 0;
    LuBatch *lub = nullptr; bool lub_tried = false;   // multi-zone LU launch (fb_lu.cu: k_lu_tile_b), built on first use
};

static const char *TNAMES[6] = {"rhs", "lu", "bc", "transfer", "coe", "all"};

struct Scope {   // CUDA-event bracket around one kernel family, resolved lazily (no sync in the hot loop)
    fb_solver *s; int id; cudaEvent_t a, b; bool on;
    Scope(fb_solver *s_, int id_) : s(s_), id(id_), on(s_->timing) {
        if (on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s->st); }
    }
    ~Scope() { if (on) { cudaEventRecord(b, s->st); s->tm[id].pend.push_back({a, b}); } }
};
static void resolve(fb_solver *s)
{
    for (int i = 0; i < 6; i++) {
        for (auto &p : s->tm[i].pend) {
            cudaEventSynchronize(p.second); float ms = 0; cudaEventElapsedTime(&ms, p.first, p.second);
            s->tm[i].ms += ms; cudaEventDestroy(p.first); cudaEventDestroy(p.second);
        }
        s->tm[i].pend.clear();
    }
}

namespace {
typedef struct { char internal[128]; } fb_ncclUniqueId;
typedef int (*nccl_getid_t)(fb_ncclUniqueId *);
typedef int (*nccl_init_t)(void **, int, fb_ncclUniqueId, int);
typedef int (*nccl_sr_t)(const void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_rv_t)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_v_t)(void);
typedef int (*nccl_destroy_t)(void *);
typedef const char *(*nccl_err_t)(int);
struct NcclApi {
    void *h = nullptr; nccl_getid_t getid = nullptr; nccl_init_t init = nullptr; nccl_sr_t send = nullptr; nccl_rv_t recv = nullptr;
    nccl_v_t gstart = nullptr, gend = nullptr; nccl_destroy_t destroy = nullptr; nccl_err_t errstr = nullptr;
};
NcclApi *nccl_api()
{
    static NcclApi api; static bool tried = false;
    if (!tried) {
        tried = true;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) { api.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (api.h) break; }
        if (api.h) {
            api.getid = (nccl_getid_t)dlsym(api.h, "ncclGetUniqueId"); api.init = (nccl_init_t)dlsym(api.h, "ncclCommInitRank");
            api.send = (nccl_sr_t)dlsym(api.h, "ncclSend"); api.recv = (nccl_rv_t)dlsym(api.h, "ncclRecv");
            api.gstart = (nccl_v_t)dlsym(api.h, "ncclGroupStart"); api.gend = (nccl_v_t)dlsym(api.h, "ncclGroupEnd");
            api.destroy = (nccl_destroy_t)dlsym(api.h, "ncclCommDestroy"); api.errstr = (nccl_err_t)dlsym(api.h, "ncclGetErrorString");
            if (!api.getid || !api.init || !api.send || !api.recv || !api.gstart || !api.gend) api.h = nullptr;
        }
    }
    return api.h ? &api : nullptr;
}
const int FB_NCCL_DOUBLE = 8;   // ncclFloat64 (nccl.h: ncclDataType_t)
}

extern "C" {

int fb_version(void) { return 100; }
const char *fb_last_error(void) { return g_err.c_str(); }
int fb_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

int fb_solver_create(int device, int nzones, fb_solver **out)
{
    if (!out || nzones <= 0) return fail(FB_EINVAL, "fb_solver_create: bad arguments");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(FB_ECUDA, "fb_solver_create: no CUDA device visible -- fast_b200 has no CPU fallback");
    }
    CK(cudaSetDevice(device));
    fb_solver *s = new fb_solver();
    s->dev = device; s->nz = nzones; s->z.resize(nzones);
    CK(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s->st_h2d, cudaStreamNonBlocking));
    {   // the exchange stream outranks the others: its pack / unpack CTAs go ahead of the interior tiles still queued
        int lo = 0, hi = 0; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&s->st2, cudaStreamNonBlocking, hi));
    }
    CK(cudaEventCreateWithFlags(&s->ev1, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&s->ev2, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
    for (int q = 0; q < 3; q++) { CK(cudaStreamCreateWithFlags(&s->lane[q], cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&s->ev_lane[q], cudaEventDisableTiming)); }
    *out = s;
    return FB_OK;
}

void fb_solver_destroy(fb_solver *s)
{
    if (!s) return;
    cudaSetDevice(s->dev);
    for (auto &z : s->z) {
        for (int l = 0; l < 3; l++) cudaFree(z.rop[l]);
        cudaFree(z.xyz); cudaFree(z.mtr); cudaFree(z.aux); cudaFree(z.drodm); cudaFree(z.coe); cudaFree(z.wigb);
        cudaFree(z.rdm); cudaFree(z.cflv); cudaFree(z.wins); cudaFree(z.wins_rhs);
        lu_plan_free(z.lu);
    }
    joins_free(s->joins);
    cudaFree(s->halo_send); cudaFree(s->halo_recv);
    if (s->st) cudaStreamDestroy(s->st);
    if (s->st_h2d) cudaStreamDestroy(s->st_h2d);
    for (auto &z : s->z) if (z.ev_up) cudaEventDestroy(z.ev_up);
    if (s->st2) cudaStreamDestroy(s->st2);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->ev2) cudaEventDestroy(s->ev2);
    for (int q = 0; q < 3; q++) { if (s->lane[q]) cudaStreamDestroy(s->lane[q]); if (s->ev_lane[q]) cudaEventDestroy(s->ev_lane[q]); }
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    lu_batch_free(s->lub);
    if (s->nccl) { NcclApi *n = nccl_api(); if (n && n->destroy) n->destroy(s->nccl); }
    delete s;
}

// ghost range of one BC window for a whole-zone sub-domain, npass = 0 (C/BC/indice_cl_sdm.for:40-90)
static bool bc_window(const int *pi, int idir, const int *fen, int out[6])
{
    const int nf = pi[PI_NIJK + 3], nfk = pi[PI_NIJK + 4];
    const int sdm[6] = {1, pi[PI_IJKV], 1, pi[PI_IJKV + 1], 1, pi[PI_IJKV + 2]};
    for (int d = 0; d < 3; d++) if (!(sdm[2 * d] <= fen[2 * d + 1] && sdm[2 * d + 1] >= fen[2 * d])) return false;
    for (int d = 0; d < 3; d++) {
        out[2 * d] = fen[2 * d] > sdm[2 * d] ? fen[2 * d] : sdm[2 * d];
        out[2 * d + 1] = fen[2 * d + 1] < sdm[2 * d + 1] ? fen[2 * d + 1] : sdm[2 * d + 1];
    }
    const int d = (idir - 1) / 2, fic = d == 2 ? nfk : nf;
    if (idir % 2 == 1) { out[2 * d + 1] = out[2 * d] - 1; out[2 * d] = out[2 * d] - fic; }
    else { out[2 * d + 1] = out[2 * d] + fic; out[2 * d] = out[2 * d] + 1; }
    return true;
}
static int win_cells(const BcWin &w)
{
    const int d = (w.idir - 1) / 2, t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
    return (w.w[2 * t1 + 1] - w.w[2 * t1] + 1) * (w.w[2 * t2 + 1] - w.w[2 * t2] + 1);
}

// BC tables for lrhs=0 (state) and lrhs=1 (rhs before the LU sweep), following S/BC/BCzone.cpp:86-360
static int build_bc_tables(Zone &z)
{
    const int *pi = z.pi.data();
    const int pt = pi[PI_PT_BC];
    if (pt <= 0 || pt >= (int)z.pi.size()) return fail(FB_EINVAL, "Parameter_int has no BC table (PT_BC); run _BCcompact first");
    const int nb = pi[pt];
    const int ijkv[3] = {pi[PI_IJKV], pi[PI_IJKV + 1], pi[PI_IJKV + 2]};
    std::vector<BcWin> w0, w1;
    int lskip[6] = {1, 1, 1, 1, 1, 1}, ificmax[6];
    const int lum = pi[PI_LU_MATCH] != 0;
    for (int p = 0; p < 4; p++) ificmax[p] = pi[PI_NIJK + 3] * lum;
    for (int p = 4; p < 6; p++) ificmax[p] = pi[PI_NIJK + 4] * lum;
    z.lcorner = 1;
    for (int b = 0; b < nb; b++) {
        const int pb = pi[pt + 1 + b], ty = pi[pb + BC_TYPE], ip = pi[pb + BC_IDIR] - 1;
        if (ip < 0 || ip > 5) return fail(FB_EINVAL, "BC window with idir outside 1..6");
        const bool ok = ty == 0 || ty == 1 || ty == 3 || ty == 4 || ty == 6 || ty == 11 || ty == 22;
        if (!ok) return fail(FB_EINVAL, "BC type " + std::to_string(ty) + " is outside the hot-path scope (0,1,3,4,6,11,22)");
        if ((ty >= 3 && ty <= 7) || ty == 12 || ty == 18 || ty == 30) {
            lskip[ip] = 0; z.lcorner = 0;
            if ((ty >= 3 && ty <= 7) || ty == 12) ificmax[ip] = 1;
        }
    }
    for (int p = 0; p < 6; p++) {
        const int base = (p % 2) ? ijkv[p / 2] : 1;
        z.lu_box[p] = base + ((p % 2) ? 1 : -1) * (1 - lskip[p]) * ificmax[p];
    }
    // Roe + wall windows: faces whose interior Euler flux is replaced by the wall pressure flux (correct_flux.for:60-131)
    z.d.ncorr = 0;
    if (pi[PI_KFLUDOM] == 5)
        for (int b = 0; b < nb; b++) {
            const int pb = pi[pt + 1 + b], ty = pi[pb + BC_TYPE], idir = pi[pb + BC_IDIR];
            if (!(ty == 3 || ty == 4 || ty == 6)) continue;
            CorrWin c; c.idir = idir; c.slip = ty == 3 ? 1 : 0;
            if (!bc_window(pi, idir, pi + pb + BC_FEN, c.w)) continue;
            const int d = (idir - 1) / 2;
            if (idir % 2 == 1) { c.w[2 * d] = 1; c.w[2 * d + 1] = 1; } else c.w[2 * d + 1] = c.w[2 * d];
            if (z.d.ncorr >= FB_MAXCORR) return fail(FB_EINVAL, "more than 8 wall windows on a Roe zone (flux correction table)");
            z.d.corr[z.d.ncorr++] = c;
        }
    for (int idir = 0; idir < 6; idir++) {      // default rhs extrapolation on the implicit-ed faces (BCzone.cpp:190-256)
        if (lskip[idir]) continue;
        int fen[6] = {1, ijkv[0], 1, ijkv[1], 1, ijkv[2]};
        const int d = idir / 2;
        if (idir % 2 == 0) fen[2 * d + 1] = 1; else fen[2 * d] = fen[2 * d + 1];
        BcWin w; w.type = 0; w.idir = idir + 1;
        if (bc_window(pi, idir + 1, fen, w.w)) w1.push_back(w);
    }
    for (int b = 0; b < nb; b++) {
        const int pb = pi[pt + 1 + b];
        BcWin w; w.type = pi[pb + BC_TYPE]; w.idir = pi[pb + BC_IDIR];
        if (!bc_window(pi, w.idir, pi + pb + BC_FEN, w.w)) continue;
        w0.push_back(w);
        if (lskip[w.idir - 1] == 0) {
            BcWin r = w;
            const int ty = r.type;
            if (ty == 0 || ty == 1 || ty == 2 || ty == 10 || ty == 11 || ty == 14 || ty == 16 || ty == 17 || ty == 20 || ty == 21) r.type = 0;
            w1.push_back(r);
        }
    }
    // The reference applies the windows one after the other; windows of one zone never overlap in scope
    // (each face ghost cell belongs to one window), except the default rhs extrapolation which must run FIRST.
    // => rhs table is applied in two launches when both kinds are present (see apply_bc).
    auto upload = [&](const std::vector<BcWin> &v, BcWin *&dptr, int &n, int &maxc) -> cudaError_t {
        n = (int)v.size(); maxc = 0;
        for (auto &w : v) { const int c = win_cells(w); if (c > maxc) maxc = c; }
        if (n == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(&dptr, n * sizeof(BcWin)); if (e != cudaSuccess) return e;
        return cudaMemcpy(dptr, v.data(), n * sizeof(BcWin), cudaMemcpyHostToDevice);
    };
    CK(upload(w0, z.wins, z.nwin, z.maxc));
    CK(upload(w1, z.wins_rhs, z.nwin_rhs, z.maxc_rhs));
    // number of leading default-extrapolation windows in the rhs table
    int ndef = 0; for (int idir = 0; idir < 6; idir++) if (!lskip[idir]) ndef++;
    z.lu.ndefault_rhs = ndef;
    return FB_OK;
}

int fb_zone_set_params(fb_solver *s, int nd, const int *pi, int n_int, const double *pr, int n_real)
{
    if (!s || nd < 0 || nd >= s->nz || !pi || !pr || n_int < PI_SIZE || n_real < PR_SIZE)
        return fail(FB_EINVAL, "fb_zone_set_params: bad arguments");
    CK(cudaSetDevice(s->dev));
    Zone &z = s->z[nd];
    z.pi.assign(pi, pi + n_int); z.pr.assign(pr, pr + n_real);
    ZoneDev &d = z.d; memset(&d, 0, sizeof(d));
    d.ni = pi[PI_NIJK]; d.nj = pi[PI_NIJK + 1]; d.nk = pi[PI_NIJK + 2]; d.nf = pi[PI_NIJK + 3]; d.nfk = pi[PI_NIJK + 4];
    d.iv = pi[PI_IJKV]; d.jv = pi[PI_IJKV + 1]; d.kv = pi[PI_IJKV + 2]; d.nd = pi[PI_NDIMDX];
    d.incj = d.ni; d.inck = d.ni * d.nj;
    d.mi = pi[PI_NIJK_MTR]; d.mj = pi[PI_NIJK_MTR + 1]; d.mk = pi[PI_NIJK_MTR + 2]; d.mf = pi[PI_NIJK_MTR + 3];
    d.mfk = pi[PI_NIJK_MTR + 4]; d.nm = pi[PI_NDIMDX_MTR];
    d.xi = pi[PI_NIJK_XYZ]; d.xj = pi[PI_NIJK_XYZ + 1]; d.xf = pi[PI_NIJK_XYZ + 3]; d.xfk = pi[PI_NIJK_XYZ + 4];
    d.ndx = pi[PI_NDIMDX_XYZ];
    d.typ = pi[PI_ITYPZONE]; d.iflow = pi[PI_IFLOW]; d.kflu = pi[PI_KFLUDOM]; d.slope = pi[PI_SLOPE]; d.neq = pi[PI_NEQ];
    d.itypcp = pi[PI_ITYPCP]; d.neq_coe = pi[PI_NEQ_COE]; d.rk = pi[PI_RK]; d.dtloc = pi[PI_DTLOC];
    d.sensortype = pi[PI_SENSORTYPE];
    if (d.typ < 0 || d.typ > 2) return fail(FB_EINVAL, "zone type " + std::to_string(d.typ) + " not built (3dfull/3dhomo/3dcart only)");
    if (d.iflow < 1 || d.iflow > 3) return fail(FB_EINVAL, "IFLOW outside 1..3");
    if (!(d.kflu == 1 || d.kflu == 2 || d.kflu == 5)) return fail(FB_EINVAL, "scheme (KFLUDOM) outside ausmpred/senseur/roe");
    if (d.slope < 1 || d.slope > 3) return fail(FB_EINVAL, "slope outside o1/o3/minmod");
    if (pi[PI_LALE] != 0 || pi[PI_EXPLOC] != 0 || pi[PI_ILES] != 0 || pi[PI_IMPLICITSOLVER] != 0 || pi[PI_NB_RELAX] < 1)
        return fail(FB_EINVAL, "ALE / local explicit / LES / GMRES are outside the hot-path scope");
    if (d.nf != 2) return fail(FB_EINVAL, "only 2 ghost layers are supported");
    // lu_match with SA (invlu_overlap_{l,u}_SA.for): same sweeps with six equations; extrap_coe.for only acts where the sweep
    // box reaches the second ghost row, which needs a BC of type 18 / 30 -- rejected above -- so it is a no-op here
    if (d.dtloc == 1 && d.iflow == 1) return fail(FB_EINVAL, "local time step needs a viscous model (dtloc.for divides by mu)");
    if (d.dtloc == 1 && d.itypcp == 2) return fail(FB_EINVAL, "explicit local time step (tstb3_expli) is not built");
    // constants of template_FluxAndBalance.for:121-153
    d.gam = pr[PR_GAMMA]; d.cv = pr[PR_CVINF]; d.rgp = d.cv * (d.gam - 1.); d.gam1 = d.gam / (d.gam - 1.);
    d.gam3 = d.gam1 / pr[PR_PRANDT] * d.rgp; d.gam4 = d.gam1 / pr[PR_PRANDT_TUR] * d.rgp;
    d.cmus1 = pr[PR_CS]; d.temp01 = 1. / pr[PR_TEMP0];
    d.coesut = pr[PR_XMUL0] * (1. + d.cmus1 * d.temp01);
    d.coesut_vis = pr[PR_XMUL0] * (1. + d.cmus1 * d.temp01) * sqrt(d.temp01);
    d.sigma_1 = 1. / SA_SIGMA;
    d.c1 = 0.02 * pr[PR_VINF]; d.c2 = 0.02 / (pr[PR_VINF] * pr[PR_ROINF]);
    d.opt0 = (double)pi[PI_SENSORTYPE]; d.wig_cte = pr[PR_WIG_AMP]; d.psiroe = pr[PR_PSIROE];
    d.tmin_1 = 100. / pr[PR_TINF]; d.dtc = pr[PR_DTC]; d.ratiom = pr[PR_RATIOM]; d.cfl = pr[PR_CFL];
    d.roinf = pr[PR_ROINF]; d.xmul0 = pr[PR_XMUL0]; d.prandt = pr[PR_PRANDT];
    d.vinf[0] = pr[PR_VXINF]; d.vinf[1] = pr[PR_VYINF]; d.vinf[2] = pr[PR_VZINF]; d.pinf = pr[PR_PINF];
    d.ronutildeinf = pr[PR_RONUTILDEINF];
    // HBM arrays
    z.n_state = (size_t)d.nd * d.neq;
    const int naux = d.iflow == 3 ? 2 : (d.iflow == 2 ? 1 : 0);
    z.n_aux = (size_t)d.nd * naux;
    for (int l = 0; l < 3; l++) { cudaFree(z.rop[l]); CK(cudaMalloc(&z.rop[l], z.n_state * sizeof(double))); CK(cudaMemset(z.rop[l], 0, z.n_state * sizeof(double))); }
    cudaFree(z.drodm); CK(cudaMalloc(&z.drodm, z.n_state * sizeof(double))); CK(cudaMemset(z.drodm, 0, z.n_state * sizeof(double)));
    cudaFree(z.aux); z.aux = nullptr;
    if (naux) { CK(cudaMalloc(&z.aux, z.n_aux * sizeof(double))); CK(cudaMemset(z.aux, 0, z.n_aux * sizeof(double))); }
    cudaFree(z.coe); z.coe = nullptr;
    if (d.itypcp <= 1) { CK(cudaMalloc(&z.coe, (size_t)d.nd * d.neq_coe * sizeof(double))); CK(cudaMemset(z.coe, 0, (size_t)d.nd * d.neq_coe * sizeof(double))); }
    cudaFree(z.wigb); z.wigb = nullptr;
    if (d.kflu == 2) { CK(cudaMalloc(&z.wigb, d.nd)); CK(cudaMemset(z.wigb, 0, d.nd)); }
    const int neq_ij = pi[PI_NEQ_IJ], neq_k = pi[PI_NEQ_K];
    cudaFree(z.mtr); CK(cudaMalloc(&z.mtr, (size_t)d.nm * (2 * neq_ij + neq_k + 1) * sizeof(double)));
    d.ti = z.mtr; d.tj = d.ti + (size_t)d.nm * neq_ij; d.tk = d.tj + (size_t)d.nm * neq_ij; d.vol = d.tk + (size_t)d.nm * neq_k;
    cudaFree(z.xyz); CK(cudaMalloc(&z.xyz, (size_t)3 * d.ndx * sizeof(double)));
    d.x = z.xyz; d.y = z.xyz + d.ndx; d.z = z.xyz + 2 * (size_t)d.ndx;
    const int nss = pi[PI_NSSITER] > 0 ? pi[PI_NSSITER] : 1;
    cudaFree(z.rdm); CK(cudaMalloc(&z.rdm, (size_t)nss * d.neq * 2 * sizeof(double))); CK(cudaMemset(z.rdm, 0, (size_t)nss * d.neq * 2 * sizeof(double)));
    cudaFree(z.cflv); CK(cudaMalloc(&z.cflv, 4 * sizeof(double))); CK(cudaMemset(z.cflv, 0, 4 * sizeof(double)));
    d.xmut = z.aux; d.drodm = z.drodm; d.coe = z.coe; d.wigb = z.wigb;
    cudaFree(z.wins); cudaFree(z.wins_rhs); z.wins = z.wins_rhs = nullptr;
    int rc = build_bc_tables(z); if (rc) return rc;
    lu_batch_free(s->lub); s->lub = nullptr; s->lub_tried = false;
    z.lu.nrelax = pi[PI_NB_RELAX];     // >= 2: LU-SSOR relaxations (invlussor*.for)
    if (d.itypcp <= 1) { rc = lu_plan_build(z.lu, d, z.lu_box); if (rc) return fail(FB_ECUDA, "lu_plan_build failed"); }
    z.has_params = true; s->metric_done = false;
    return FB_OK;
}

int fb_zone_set_coords(fb_solver *s, int nd, const double *x, const double *y, const double *zc)
{
    if (!s || nd < 0 || nd >= s->nz || !x || !y || !zc) return fail(FB_EINVAL, "fb_zone_set_coords: bad arguments");
    Zone &z = s->z[nd]; if (!z.has_params) return fail(FB_ESTATE, "fb_zone_set_coords before fb_zone_set_params");
    CK(cudaSetDevice(s->dev));
    const size_t n = (size_t)z.d.ndx * sizeof(double);
    CK(cudaMemcpy(z.xyz, x, n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(z.xyz + z.d.ndx, y, n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(z.xyz + 2 * (size_t)z.d.ndx, zc, n, cudaMemcpyHostToDevice));
    z.has_xyz = true;
    return FB_OK;
}

int fb_init_metric(fb_solver *s)
{
    if (!s) return fail(FB_EINVAL, "null solver");
    CK(cudaSetDevice(s->dev));
    for (auto &z : s->z) {
        if (!z.has_params || !z.has_xyz) return fail(FB_ESTATE, "fb_init_metric: every zone needs params and coordinates");
        launch_metric(z.d, s->st);
        if (z.d.itypcp <= 1) lu_plan_metrics(z.lu, z.d, s->st);
    }
    CK(cudaStreamSynchronize(s->st));
    for (auto &z : s->z) {   // 3dcart constants live in the descriptor (template_FluxAndBalance.for:112-119)
        if (z.d.typ == 2) {
            double h[4];
            CK(cudaMemcpy(h, z.mtr, 4 * sizeof(double), cudaMemcpyDeviceToHost));   // ti|tj|tk|vol, one value each
            z.d.tc0[0] = h[0]; z.d.tc0[1] = h[1]; z.d.tc0[2] = h[2]; z.d.vol0 = h[3];
        }
    }
    CK(cudaGetLastError());
    s->metric_done = true;
    return FB_OK;
}

int fb_zone_get_metric(fb_solver *s, int nd, double *ti, double *tj, double *tk, double *vol)
{
    if (!s || nd < 0 || nd >= s->nz) return fail(FB_EINVAL, "bad zone");
    Zone &z = s->z[nd]; if (!s->metric_done) return fail(FB_ESTATE, "metrics not computed");
    CK(cudaSetDevice(s->dev));
    const int nij = z.pi[PI_NEQ_IJ], nk = z.pi[PI_NEQ_K]; const size_t nm = z.d.nm;
    if (ti) CK(cudaMemcpy(ti, z.d.ti, nm * nij * sizeof(double), cudaMemcpyDeviceToHost));
    if (tj) CK(cudaMemcpy(tj, z.d.tj, nm * nij * sizeof(double), cudaMemcpyDeviceToHost));
    if (tk) CK(cudaMemcpy(tk, z.d.tk, nm * nk * sizeof(double), cudaMemcpyDeviceToHost));
    if (vol) CK(cudaMemcpy(vol, z.d.vol, nm * sizeof(double), cudaMemcpyDeviceToHost));
    return FB_OK;
}

static int check_lu(fb_solver *s);
static int chk_zone(fb_solver *s, int nd, int level)
{
    if (!s || nd < 0 || nd >= s->nz || level < 0 || level > 2) return fail(FB_EINVAL, "bad zone / level");
    if (!s->z[nd].has_params) return fail(FB_ESTATE, "zone has no parameters yet");
    return FB_OK;
}
int fb_zone_upload_state(fb_solver *s, int nd, int level, const double *rop)
{
    int rc = chk_zone(s, nd, level); if (rc) return rc;
    CK(cudaSetDevice(s->dev));
    CK(cudaMemcpyAsync(s->z[nd].rop[level], rop, s->z[nd].n_state * sizeof(double), cudaMemcpyHostToDevice, s->st));
    CK(cudaStreamSynchronize(s->st));
    return FB_OK;
}
int fb_zone_download_state(fb_solver *s, int nd, int level, double *rop)
{
    int rc = chk_zone(s, nd, level); if (rc) return rc;
    CK(cudaSetDevice(s->dev));
    CK(cudaMemcpyAsync(rop, s->z[nd].rop[level], s->z[nd].n_state * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    CK(cudaStreamSynchronize(s->st));
    return check_lu(s);
}
int fb_zone_upload_aux(fb_solver *s, int nd, const double *aux, int n)
{
    int rc = chk_zone(s, nd, 0); if (rc) return rc;
    Zone &z = s->z[nd]; if ((size_t)n != z.n_aux) return fail(FB_EINVAL, "aux size mismatch");
    if (n == 0) return FB_OK;
    CK(cudaSetDevice(s->dev));
    CK(cudaMemcpyAsync(z.aux, aux, z.n_aux * sizeof(double), cudaMemcpyHostToDevice, s->st));
    CK(cudaStreamSynchronize(s->st));
    return FB_OK;
}
int fb_zone_download_aux(fb_solver *s, int nd, double *aux, int n)
{
    int rc = chk_zone(s, nd, 0); if (rc) return rc;
    Zone &z = s->z[nd]; if ((size_t)n != z.n_aux) return fail(FB_EINVAL, "aux size mismatch");
    if (n == 0) return FB_OK;
    CK(cudaSetDevice(s->dev));
    CK(cudaMemcpyAsync(aux, z.aux, z.n_aux * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    CK(cudaStreamSynchronize(s->st));
    return FB_OK;
}
int fb_zone_download_work(fb_solver *s, int nd, double *drodm, double *coe)
{
    int rc = chk_zone(s, nd, 0); if (rc) return rc;
    Zone &z = s->z[nd];
    CK(cudaSetDevice(s->dev));
    CK(cudaStreamSynchronize(s->st));
    if (drodm) CK(cudaMemcpy(drodm, z.drodm, z.n_state * sizeof(double), cudaMemcpyDeviceToHost));
    if (coe && z.coe) CK(cudaMemcpy(coe, z.coe, (size_t)z.d.nd * z.d.neq_coe * sizeof(double), cudaMemcpyDeviceToHost));
    return FB_OK;
}

int fb_zone_download_wig(fb_solver *s, int nd, unsigned char *wig, int n)
{
    int rc = chk_zone(s, nd, 0); if (rc) return rc;
    Zone &z = s->z[nd];
    if (!wig || n != z.d.nd) return fail(FB_EINVAL, "fb_zone_download_wig: size must be NDIMDX");
    if (!z.wigb) return fail(FB_ESTATE, "zone has no SENSOR switch (scheme != senseur)");
    CK(cudaSetDevice(s->dev)); CK(cudaStreamSynchronize(s->st));
    CK(cudaMemcpy(wig, z.wigb, (size_t)z.d.nd, cudaMemcpyDeviceToHost));
    return FB_OK;
}

int fb_set_joins(fb_solver *s, const int *ji, int n)
{
    if (!s || !ji || n < 1) return fail(FB_EINVAL, "fb_set_joins: bad arguments");
    CK(cudaSetDevice(s->dev));
    std::vector<int> nds(s->nz), neqs(s->nz);
    for (int i = 0; i < s->nz; i++) { nds[i] = s->z[i].d.nd; neqs[i] = s->z[i].d.neq; }
    std::string err;
    if (joins_build(s->joins, ji, n, s->nz, nds.data(), neqs.data(), err)) return fail(FB_EINVAL, "fb_set_joins: " + err);
    // which faces of each zone feed cross-rank joins (kind 1 = pack): the explicit stages compute the tiles on those faces first
    for (auto &z : s->z) z.xfaces = 0;
    {
        const int nj = ji[0]; const int *idx = ji + 2 + 7 * nj;
        for (int q = 0; q < nj; q++) {
            const int *d = ji + 2 + 7 * q;
            if (d[0] != 1 || d[3] <= 0) continue;
            Zone &z = s->z[d[1]]; const ZoneDev &zd = z.d;
            int lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {-(1 << 30), -(1 << 30), -(1 << 30)};
            for (int p = 0; p < d[3]; p++) {
                const int l = idx[d[4] + p];
                const int c[3] = {l % zd.ni - zd.nf + 1, (l / zd.ni) % zd.nj - zd.nf + 1, l / (zd.ni * zd.nj) - zd.nfk + 1};
                for (int e = 0; e < 3; e++) { lo[e] = std::min(lo[e], c[e]); hi[e] = std::max(hi[e], c[e]); }
            }
            const int nv[3] = {zd.iv, zd.jv, zd.kv};
            int flags = 0;
            for (int e = 0; e < 3; e++) {
                if (hi[e] - lo[e] + 1 >= nv[e]) continue;                   // spans the direction: not a face in it
                if (hi[e] <= 2) flags |= 1 << (2 * e);                         // two layers behind the lower face
                else if (lo[e] >= nv[e] - 1) flags |= 2 << (2 * e);            // ... the upper face
                else flags = 63;                                               // not a face window: all faces first
            }
            z.xfaces |= flags ? flags : 63;
        }
    }
    cudaFree(s->halo_send); cudaFree(s->halo_recv); s->halo_send = s->halo_recv = nullptr;
    if (s->joins.send_doubles) CK(cudaMalloc(&s->halo_send, s->joins.send_doubles * sizeof(double)));
    if (s->joins.recv_doubles) CK(cudaMalloc(&s->halo_recv, s->joins.recv_doubles * sizeof(double)));
    return FB_OK;
}

int fb_set_joins_tc(fb_solver *s, const int *param_int_tc, int n_int, const double *param_real_tc, int n_real)
{
    if (!s || !param_int_tc || !param_real_tc || n_int < 1 || n_real < 0) return fail(FB_EINVAL, "fb_set_joins_tc: bad arguments");
    CK(cudaSetDevice(s->dev));
    std::vector<int> nds(s->nz), neqs(s->nz);
    for (int i = 0; i < s->nz; i++) { nds[i] = s->z[i].d.nd; neqs[i] = s->z[i].d.neq; }
    std::string err;
    if (joins_build_tc(s->joins, param_int_tc, n_int, param_real_tc, n_real, s->nz, nds.data(), neqs.data(), err))
        return fail(FB_EINVAL, "fb_set_joins_tc: " + err);
    for (auto &z : s->z) z.xfaces = 0;                 // no cross-rank entries in this format (NoTransfert = 1: intra-process joins)
    cudaFree(s->halo_send); cudaFree(s->halo_recv); s->halo_send = s->halo_recv = nullptr;
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------------- NCCL (native exchange)
// The cross-rank halo exchange of C/Com/setInterpTransfersFast.cpp:962-1790 (one message per neighbour rank and sub-iteration)
// runs inside the library: ncclSend / ncclRecv of the packed donor values on the exchange stream, grouped per exchange.  NCCL
// is bound at run time (dlopen) so that single-GPU users need no NCCL at all; whichever libnccl.so.2 the process has already
// loaded (e.g. torch's) is the one that answers.

int fb_nccl_unique_id(char id[128])
{
    NcclApi *n = nccl_api();
    if (!n || !id) return fail(FB_ECUDA, "fb_nccl_unique_id: libnccl.so.2 cannot be loaded");
    fb_ncclUniqueId u; const int rc = n->getid(&u);
    if (rc) return fail(FB_ECUDA, std::string("ncclGetUniqueId: ") + (n->errstr ? n->errstr(rc) : "error"));
    memcpy(id, u.internal, 128);
    return FB_OK;
}
int fb_nccl_init(fb_solver *s, int rank, int world, const char id[128])
{
    NcclApi *n = nccl_api();
    if (!s || !id || world < 1 || rank < 0 || rank >= world) return fail(FB_EINVAL, "fb_nccl_init: bad arguments");
    if (!n) return fail(FB_ECUDA, "fb_nccl_init: libnccl.so.2 cannot be loaded");
    CK(cudaSetDevice(s->dev));
    fb_ncclUniqueId u; memcpy(u.internal, id, 128);
    void *comm = nullptr; const int rc = n->init(&comm, world, u, rank);
    if (rc) return fail(FB_ECUDA, std::string("ncclCommInitRank: ") + (n->errstr ? n->errstr(rc) : "error"));
    s->nccl = comm;
    return FB_OK;
}
int fb_set_peers(fb_solver *s, int npeers, const int *rank, const long long *send_off, const long long *send_n,
                 const long long *recv_off, const long long *recv_n)
{
    if (!s || npeers < 0 || (npeers && (!rank || !send_off || !send_n || !recv_off || !recv_n))) return fail(FB_EINVAL, "fb_set_peers: bad arguments");
    s->peers.clear();
    for (int p = 0; p < npeers; p++) {
        if (send_off[p] < 0 || recv_off[p] < 0 || send_off[p] + send_n[p] > s->joins.send_doubles || recv_off[p] + recv_n[p] > s->joins.recv_doubles)
            return fail(FB_EINVAL, "fb_set_peers: slice outside the halo buffers (call fb_set_joins first)");
        s->peers.push_back({rank[p], send_off[p], send_n[p], recv_off[p], recv_n[p]});
    }
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------------- driver
static void apply_bc_zone(fb_solver *s, Zone &z, double *rop, int lrhs)
{
    if (lrhs == 0) {
        launch_bc(z.d, z.wins, z.nwin, z.maxc, rop, 0, s->st); s->launches += z.nwin > 0;
        int box[6] = {1, z.d.iv, 1, z.d.jv, 1, z.d.kv};
        launch_corners(z.d, rop, box, s->st); s->launches++;
    } else {
        const int nd0 = z.lu.ndefault_rhs;
        if (nd0 > 0) { launch_bc(z.d, z.wins_rhs, nd0, z.maxc_rhs, rop, 1, s->st); s->launches++; }
        if (z.nwin_rhs > nd0) { launch_bc(z.d, z.wins_rhs + nd0, z.nwin_rhs - nd0, z.maxc_rhs, rop, 1, s->st); s->launches++; }
        if (z.lcorner == 0) { launch_corners(z.d, rop, z.lu_box, s->st); s->launches++; }
    }
}

// Match-join transfers.  Cross-rank joins (C/Com/setInterpTransfersFast.cpp:962 setInterpTransfersInter) run on the exchange
// stream st2: pack the donor cells (ordered after everything queued on st so far), let the host transport move the send
// buffer to the peers' receive buffers (fb_exchange_stream), unpack into the ghost cells.  transfers_begin returns at once;
// whatever the caller queues on st before transfers_end overlaps the exchange.  Local joins run on st in transfers_end.
static bool cross_rank(fb_solver *s) { return (s->xcb || s->nccl) && (s->joins.send_doubles || s->joins.recv_doubles); }

// ev1_recorded: the caller has already recorded ev1 on st behind the kernels that produce the donor cells (the face tiles of
// an explicit stage, gsdr3) and may have queued more work since; otherwise the exchange waits for everything queued on st so far
static void transfers_begin(fb_solver *s, int which, double **target, bool ev1_recorded = false)
{
    if (s->joins.njoin == 0 || !cross_rank(s)) return;
    std::vector<double *> ptr(s->nz);
    for (int i = 0; i < s->nz; i++) ptr[i] = which == 1 ? s->z[i].drodm : target[i];
    if (!ev1_recorded) cudaEventRecord(s->ev1, s->st);
    cudaStreamWaitEvent(s->st2, s->ev1, 0);
    if (s->joins.send_doubles) { joins_pack(s->joins, ptr.data(), s->halo_send, s->joins.send_doubles * 8, s->st2); s->launches++; }
    if (s->nccl) {       // one grouped send / receive pair per neighbour rank, stream-ordered on the exchange stream: no host in the loop
        NcclApi *n = nccl_api();
        n->gstart();
        for (auto &p : s->peers) {
            if (p.sn) n->send(s->halo_send + p.so, (size_t)p.sn, FB_NCCL_DOUBLE, p.rank, s->nccl, s->st2);
            if (p.rn) n->recv(s->halo_recv + p.ro, (size_t)p.rn, FB_NCCL_DOUBLE, p.rank, s->nccl, s->st2);
        }
        n->gend();
    } else s->xcb(s->xuser, which);
    if (s->joins.recv_doubles) { joins_unpack(s->joins, ptr.data(), s->halo_recv, s->joins.recv_doubles * 8, s->st2); s->launches++; }
    cudaEventRecord(s->ev2, s->st2);
}
static void transfers_end(fb_solver *s, int which, double **target)
{
    if (s->joins.njoin == 0) return;
    Scope t(s, 3);
    std::vector<double *> ptr(s->nz);
    for (int i = 0; i < s->nz; i++) ptr[i] = which == 1 ? s->z[i].drodm : target[i];
    s->launches += joins_apply(s->joins, ptr.data(), s->nz, s->st);
    if (cross_rank(s)) cudaStreamWaitEvent(s->st, s->ev2, 0);
}
static void transfers(fb_solver *s, int which, int level_ssiter_dummy, double **target)
{
    (void)level_ssiter_dummy;
    transfers_begin(s, which, target);
    transfers_end(s, which, target);
}

// fork / join of a phase whose per-zone kernels are independent (fb_solver::lane)
struct Fan {
    fb_solver *s; bool on;
    Fan(fb_solver *s_, bool on_) : s(s_), on(on_) {
        if (!on) return;
        cudaEventRecord(s->ev_fork, s->st);
        for (int q = 0; q < 3; q++) cudaStreamWaitEvent(s->lane[q], s->ev_fork, 0);
    }
    cudaStream_t stream(int n) const { return on && (n & 3) ? s->lane[(n & 3) - 1] : s->st; }
    ~Fan() {
        if (!on) return;
        for (int q = 0; q < 3; q++) { cudaEventRecord(s->ev_lane[q], s->lane[q]); cudaStreamWaitEvent(s->st, s->ev_lane[q], 0); }
    }
};

static long long g_lu_batch_launches = 0;   // multi-zone LU sweeps (lower + upper pairs), fb_counter("lu_batch")
static int gsdr3(fb_solver *s, int nitcfg, int nssiter, int nitrun, int first_it, int lssiter_verif, int lexit_lu)
{
    (void)nitrun; (void)nssiter;
    const bool implicit = s->z[0].d.itypcp <= 1;
    static const bool no_overlap = [] { const char *e = getenv("FAST_B200_OVERLAP"); return e && e[0] == '0'; }();   // A/B switch
    const bool overlap = !implicit && lexit_lu == 0 && cross_rank(s) && !no_overlap;
    std::vector<double *> cl(s->nz);
    // bit 0: the nstep-1 kernels (viscosity, coe, CFL, SENSOR) of the zones fan out over the lanes; bit 1: the rhs kernels too.
    // Default 1: with the rhs kernels of several zones running concurrently from different streams the TMA-staged kernel
    // k_rhs3 is NOT deterministic -- about one 1e-5 perturbation per ten iterations of the 8-block C5 brick, none with one
    // stream, none with the L1-path kernel k_rhs2, whatever the descriptor placement or cache policy of its global loads
    // (profiles/r02_race.md) -- so they are launched back to back on the solver stream.  FAST_B200_FAN=3 restores the r01 behaviour.
    static const int fan_mode = [] { const char *e = getenv("FAST_B200_FAN"); return e ? atoi(e) : 1; }();
    const bool fan = s->nz > 1 && (fan_mode & 2), fan_coe = s->nz > 1 && (fan_mode & 1);
    {
        Scope all(s, 5);
        std::vector<double *> ss(s->nz);
        for (int n = 0; n < s->nz; n++) {
            Zone &z = s->z[n];
            // gsdr3.cpp:162-173: flux input and array receiving transfers + BC
            double *ssiter, *clp;
            if (nitcfg == 1) { ssiter = z.rop[0]; clp = z.rop[2]; }
            else if (implicit) { ssiter = z.rop[2]; clp = z.rop[2]; }
            else if (nitcfg % 2 == 0) { ssiter = z.rop[2]; clp = z.rop[0]; }
            else { ssiter = z.rop[0]; clp = z.rop[2]; }
            cl[n] = clp; ss[n] = ssiter;
        }
        if (nitcfg == 1) {   // navier_stokes_struct.for:216-301
            Scope c(s, 4);
            Fan f(s, fan_coe);
            for (int n = 0; n < s->nz; n++) {
                Zone &z = s->z[n]; double *ssiter = ss[n]; cudaStream_t zst = f.stream(n);
                if (z.up_pending) cudaStreamWaitEvent(zst, z.ev_up, 0);      // fb_compute_host: the zone's inputs have landed
                if (z.d.iflow >= 2) { launch_invist(z.d, ssiter, implicit || z.d.iflow == 3 ? z.d.nf : 1, zst); s->launches++; }
                if (implicit) { launch_coe(z.d, ssiter, first_it, zst); s->launches++; }
                if (lssiter_verif) { launch_cfl(z.d, ssiter, z.cflv, zst); s->launches += 2; }   // cptst3.for:113-120 -> cpcfl3
                // SENSOR switch (wiggle2001.for) at nstep 1 (src_term.for:258-266)
                if (z.d.kflu == 2 && !fb_sensor_fused()) { launch_sensor(z.d, ssiter, zst); s->launches++; }
            }
        }
        if (!overlap) {
            Scope r(s, 0);
            Fan f(s, fan);
            for (int n = 0; n < s->nz; n++) {
                Zone &z = s->z[n]; cudaStream_t zst = f.stream(n);
                z.up_pending_first = z.up_pending;
                if (z.up_pending) { cudaStreamWaitEvent(zst, z.ev_up, 0); z.up_pending = false; }
                const bool ring = implicit && z.d.iflow == 3 && nitcfg == 1;
                if (ring) { launch_sa_ring(z.d, ss[n], zst); s->launches++; }
                // explicit: the stage writes the array that then gets the BCs.
                // Zones are independent in this phase: on one stream the rhs kernel of zone n may start while that of zone n-1
                // drains its last wave (programmatic dependent launch) -- unless a kernel of zone n itself sits in between
                static const bool no_pdl = [] { const char *e = getenv("FAST_B200_PDL"); return e && e[0] == '0'; }();   // A/B switch
                const int pdl = (!fan && !no_pdl && n > 0 && !ring && !s->z[n].up_pending_first) ? 1 : 0;
                launch_rhs(z.d, ss[n], implicit ? nullptr : cl[n], z.rop[0], z.rop[1], nitcfg, first_it, zst, 0, 63, pdl); s->launches++;
            }
        } else {
            // explicit stage + cross-rank joins: the tiles that touch a zone face (their cells are the donors of the exchange)
            // run first on st; the exchange stream starts behind them (ev1) while the interior tiles, queued on a lane at the
            // same time, fill the SMs as the face tiles drain.  Everything joins back on st before the local joins and BCs.
            Scope r(s, 0);
            for (int n = 0; n < s->nz; n++) if (s->z[n].up_pending) { cudaStreamWaitEvent(s->st, s->z[n].ev_up, 0); s->z[n].up_pending = false; }
            cudaEventRecord(s->ev_fork, s->st); cudaStreamWaitEvent(s->lane[0], s->ev_fork, 0);
            for (int n = 0; n < s->nz; n++) {
                Zone &z = s->z[n];
                launch_rhs(z.d, ss[n], cl[n], z.rop[0], z.rop[1], nitcfg, first_it, s->st, 1, z.xfaces); s->launches++;
            }
            cudaEventRecord(s->ev1, s->st);
            for (int n = 0; n < s->nz; n++) {
                Zone &z = s->z[n];
                launch_rhs(z.d, ss[n], cl[n], z.rop[0], z.rop[1], nitcfg, first_it, s->lane[0], 2, z.xfaces); s->launches++;
            }
            transfers_begin(s, 0, cl.data(), true);
            cudaEventRecord(s->ev_lane[0], s->lane[0]); cudaStreamWaitEvent(s->st, s->ev_lane[0], 0);
        }
        if (implicit) {
            if (s->z[0].pi[PI_LU_MATCH] == 1) transfers(s, 1, 0, nullptr);   // lhs.cpp:1-15
            // several zones: one pair of launches sweeps them all, tiles of the zones interleaved (k_lu_tile_b)
            static const bool no_batch = [] { const char *e = getenv("FAST_B200_LU_BATCH"); return e && e[0] == '0'; }();   // A/B switch
            if (s->nz > 1 && !s->lub_tried && !no_batch && lexit_lu == 0) {
                std::vector<const ZoneDev *> zs(s->nz); std::vector<const LuPlan *> ps(s->nz);
                for (int n = 0; n < s->nz; n++) { zs[n] = &s->z[n].d; ps[n] = &s->z[n].lu; }
                bool same = true;
                for (int n = 1; n < s->nz; n++) same = same && s->z[n].pi[PI_LU_MATCH] == s->z[0].pi[PI_LU_MATCH];
                s->lub = same ? lu_batch_create(s->nz, zs.data(), ps.data()) : nullptr; s->lub_tried = true;
            }
            const bool batch = s->lub && lexit_lu == 0;
            for (int n = 0; n < s->nz; n++) {
                Zone &z = s->z[n];
                { Scope b(s, 2); apply_bc_zone(s, z, z.drodm, 1); }           // lhs.cpp:57-67
                if (lssiter_verif == 1) { launch_resid(z.d, z.rdm, z.cflv, nitcfg, z.pi[PI_NSSITER], s->st); s->launches += 2; }   // cprdu3s1
                if (lexit_lu == 0 && !batch) {
                    Scope l(s, 1);
                    double *ssiter = nitcfg == 1 ? z.rop[0] : z.rop[2];
                    s->launches += launch_lu(z.d, z.lu, ssiter, z.rop[2], z.pi[PI_LU_MATCH] == 1, s->st);   // invlu_ + mjr_newton
                }
            }
            if (batch) {
                Scope l(s, 1);
                std::vector<const ZoneDev *> zs(s->nz); std::vector<const LuPlan *> ps(s->nz);
                std::vector<const double *> rin(s->nz); std::vector<double *> rout(s->nz);
                for (int n = 0; n < s->nz; n++) {
                    zs[n] = &s->z[n].d; ps[n] = &s->z[n].lu; rin[n] = nitcfg == 1 ? s->z[n].rop[0] : s->z[n].rop[2]; rout[n] = s->z[n].rop[2];
                }
                const int nl = lu_batch_launch(s->lub, zs.data(), ps.data(), rin.data(), rout.data(), s->z[0].pi[PI_LU_MATCH] == 1, s->st);
                if (nl < 0) return fail(FB_ECUDA, "lu_batch_launch: out of device memory");
                s->launches += nl; g_lu_batch_launches++;
            }
        }
        if (lexit_lu == 0) {
            if (overlap) transfers_end(s, 0, cl.data());
            else transfers(s, 0, 0, cl.data());                                // transfert_multiblock.cpp
            Scope b(s, 2);
            for (int n = 0; n < s->nz; n++) apply_bc_zone(s, s->z[n], cl[n], 0);   // bcs.cpp:32-40
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(FB_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
    return FB_OK;
}

int fb_compute(fb_solver *s, int nitrun, int nstep_deb, int nstep_fin, int nit_c, int *first_it)
{
    if (!s || !first_it || nit_c < 1) return fail(FB_EINVAL, "fb_compute: bad arguments");
    if (!s->metric_done) return fail(FB_ESTATE, "fb_compute before fb_init_metric");
    CK(cudaSetDevice(s->dev));
    const int itypcp = s->z[0].d.itypcp, nssiter = s->z[0].pi[PI_NSSITER];
    for (auto &z : s->z) if ((z.d.itypcp <= 1) != (itypcp <= 1)) return fail(FB_EINVAL, "mixed explicit/implicit zones");
    int modulo_verif = s->z[0].lu.modulo_verif > 0 ? s->z[0].lu.modulo_verif : 200;
    for (int it = 0; it < nit_c; it++) {
        const int nitrun_loc = nitrun * nit_c + it;
        for (int nstep = nstep_deb; nstep <= nstep_fin; nstep++) {
            int verif = 0, lexit = 0;                       // C/Compute/souszones_list.cpp:36-83
            if (nitrun_loc % modulo_verif == 0) { verif = 1; if (nstep == nssiter && itypcp != 2) lexit = 1; }
            if (verif == 0 && nstep == nssiter && itypcp == 1) continue;   // computePT.cpp:582
            int rc = gsdr3(s, nstep, nssiter, nitrun_loc, *first_it, verif, lexit); if (rc) return rc;
        }
        *first_it = 1;
        // computePT.cpp:675-679 swaps its local pointers; what persists is the rotation FastS._compute applies to the tree
        // (S/PyTree.py:226-228): switchPointers__(zones, case, nitmax%2+2) is a no-op (order 2, C/PyTree.py:2299) for an explicit
        // scheme with an even number of stages -- the last stage of RK2 has written `ro` itself (gsdr3.cpp:162-173)
        if (!(itypcp == 2 && nssiter % 2 == 0))
            for (auto &z : s->z) { double *sv = z.rop[1]; z.rop[1] = z.rop[0]; z.rop[0] = z.rop[2]; z.rop[2] = sv; }
    }
    return FB_OK;
}

int fb_set_modulo_verif(fb_solver *s, int m) { if (!s || m < 1) return fail(FB_EINVAL, "bad modulo_verif"); for (auto &z : s->z) z.lu.modulo_verif = m; return FB_OK; }

int fb_apply_bc(fb_solver *s, int level)
{
    if (!s || level < 0 || level > 2) return fail(FB_EINVAL, "bad level");
    if (!s->metric_done) return fail(FB_ESTATE, "fb_apply_bc before fb_init_metric");
    CK(cudaSetDevice(s->dev));
    for (auto &z : s->z) apply_bc_zone(s, z, z.rop[level], 0);
    CK(cudaGetLastError());
    return FB_OK;
}
int fb_transfers(fb_solver *s, int level)
{
    if (!s || level < 0 || level > 2) return fail(FB_EINVAL, "bad level");
    CK(cudaSetDevice(s->dev));
    std::vector<double *> t(s->nz); for (int i = 0; i < s->nz; i++) t[i] = s->z[i].rop[level];
    transfers(s, 0, 0, t.data());
    CK(cudaGetLastError());
    return FB_OK;
}
int fb_compute_mut(fb_solver *s)
{
    if (!s) return fail(FB_EINVAL, "null solver");
    CK(cudaSetDevice(s->dev));
    for (auto &z : s->z) if (z.d.iflow >= 2) { launch_invist(z.d, z.rop[0], z.d.nf, s->st); s->launches++; }
    CK(cudaGetLastError());
    return FB_OK;
}

int fb_zone_get_residuals(fb_solver *s, int nd, double *rdm, int n)
{
    int rc = chk_zone(s, nd, 0); if (rc) return rc;
    Zone &z = s->z[nd];
    const int nss = z.pi[PI_NSSITER] > 0 ? z.pi[PI_NSSITER] : 1;
    if (n != nss * z.d.neq * 2) return fail(FB_EINVAL, "rdm size must be nssiter*neq*2");
    CK(cudaSetDevice(s->dev)); CK(cudaStreamSynchronize(s->st));
    CK(cudaMemcpy(rdm, z.rdm, n * sizeof(double), cudaMemcpyDeviceToHost));
    return FB_OK;
}
int fb_zone_get_cfl(fb_solver *s, int nd, double cfl[3])
{
    int rc = chk_zone(s, nd, 0); if (rc) return rc;
    CK(cudaSetDevice(s->dev)); CK(cudaStreamSynchronize(s->st));
    CK(cudaMemcpy(cfl, s->z[nd].cflv, 3 * sizeof(double), cudaMemcpyDeviceToHost));
    const double cells = (double)s->z[nd].d.iv * s->z[nd].d.jv * s->z[nd].d.kv;
    cfl[2] /= cells;   // computePT.cpp:645
    return FB_OK;
}
int fb_zone_isfinite(fb_solver *s, int nd, long long *bad)
{
    int rc = chk_zone(s, nd, 0); if (rc) return rc;
    if (!bad) return fail(FB_EINVAL, "fb_zone_isfinite: null output");
    CK(cudaSetDevice(s->dev)); CK(cudaStreamSynchronize(s->st));
    double v = 0.;
    CK(cudaMemcpy(&v, s->z[nd].cflv + 3, sizeof(double), cudaMemcpyDeviceToHost));
    *bad = (long long)v;
    return FB_OK;
}

int fb_compute_host(fb_solver *s, int nitrun, int nit_c, int *first_it, double **ro, double **ro_m1, double **ro_p1)
{
    if (!s || !ro || !ro_m1 || !ro_p1 || !first_it) return fail(FB_EINVAL, "fb_compute_host: bad arguments");
    if (nit_c != 1) return fail(FB_EINVAL, "fb_compute_host: nit_c must be 1 (the host rotates its pointers after every iteration)");
    CK(cudaSetDevice(s->dev));
    // inputs of one iteration: ro always; ro_M1 only where the BDF2 time term reads it (core3as2.for:37-46: c3 = 0 at the first
    // iteration and with local time stepping); ro_P1 is pure output.  The uploads run zone by zone on a copy stream, each
    // followed by an event: a zone's first kernels (viscosity, coe, rhs) start as soon as ITS levels are there, while the
    // next zones are still crossing the bus.  Everything that couples the zones (LU launch, joins) comes after all of them.
    const bool implicit = s->z[0].d.itypcp <= 1;
    cudaEventRecord(s->ev_fork, s->st); cudaStreamWaitEvent(s->st_h2d, s->ev_fork, 0);     // the previous iteration is done with the buffers
    for (int n = 0; n < s->nz; n++) {
        Zone &z = s->z[n];
        if (!z.ev_up) CK(cudaEventCreateWithFlags(&z.ev_up, cudaEventDisableTiming));
        CK(cudaMemcpyAsync(z.rop[0], ro[n], z.n_state * sizeof(double), cudaMemcpyHostToDevice, s->st_h2d));
        const bool bdf2 = implicit && *first_it != 0 && z.d.dtloc != 1;
        if (bdf2) CK(cudaMemcpyAsync(z.rop[1], ro_m1[n], z.n_state * sizeof(double), cudaMemcpyHostToDevice, s->st_h2d));
        CK(cudaEventRecord(z.ev_up, s->st_h2d));
        z.up_pending = true;
    }
    const int nssiter = s->z[0].pi[PI_NSSITER];
    int rc = fb_compute(s, nitrun, 1, nssiter, 1, first_it); if (rc) return rc;
    // after the device-side rotation level 0 is the new solution: it lands in the caller's ro_P1 buffer, and the caller
    // rotates its three pointers exactly like FastC.switchPointers__ (ro_M1 <- ro <- ro_P1 <- old ro_M1)
    // (explicit schemes with an even number of stages, i.e. RK2, do not rotate: the new solution is `ro` itself and goes back
    // into the caller's ro buffers)
    const bool norot = !implicit && nssiter % 2 == 0;
    for (int n = 0; n < s->nz; n++)
        CK(cudaMemcpyAsync(norot ? ro[n] : ro_p1[n], s->z[n].rop[0], s->z[n].n_state * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    CK(cudaStreamSynchronize(s->st));
    return check_lu(s);
}

int fb_timing_enable(fb_solver *s, int on) { if (!s) return fail(FB_EINVAL, "null"); s->timing = on != 0; return FB_OK; }
int fb_timing_reset(fb_solver *s)
{
    if (!s) return fail(FB_EINVAL, "null");
    cudaSetDevice(s->dev); cudaStreamSynchronize(s->st); resolve(s);
    for (int i = 0; i < 6; i++) { s->tm[i].ms = 0; s->tm[i].n = 0; }
    s->launches = 0;
    return FB_OK;
}
int fb_timing_get(fb_solver *s, const char *name, double *ms, long long *launches)
{
    if (!s || !name) return fail(FB_EINVAL, "null");
    cudaSetDevice(s->dev); cudaStreamSynchronize(s->st); resolve(s);
    for (int i = 0; i < 6; i++) if (!strcmp(name, TNAMES[i])) { if (ms) *ms = s->tm[i].ms; if (launches) *launches = s->launches; return FB_OK; }
    return fail(FB_EINVAL, "unknown timer name");
}
extern long long g_fb_rhs_launches[2];
extern long long g_lu_cell_launches;
// test hook (tests/test_host_logic.py): the tile order of the two-part explicit stage on a gx x gy x gz tile grid, computed
// by the very function the kernels use; out = 3 ints per tile in launch order, *nface = tiles of part 1
int fb_debug_tile_order(int gx, int gy, int gz, int faces, int thin, int *out, int *nface)
{
    if (gx < 1 || gy < 1 || gz < 1 || !out || !nface) return fail(FB_EINVAL, "fb_debug_tile_order: bad arguments");
    const unsigned fl = rhs_face_flags((unsigned)faces & 63u, (unsigned)thin & 7u, gx, gy, gz);
    *nface = (int)rhs_face_tiles(gx, gy, gz, fl);
    for (unsigned c = 0; c < (unsigned)(gx * gy * gz); c++) {
        unsigned x, y, z; rhs_face_first(c, gx, gy, gz, fl, x, y, z);
        out[3 * c] = (int)x; out[3 * c + 1] = (int)y; out[3 * c + 2] = (int)z;
    }
    return FB_OK;
}

int fb_counter(const char *name, long long *out)
{
    if (!name || !out) return fail(FB_EINVAL, "null");
    if (!strcmp(name, "rhs_tma")) { *out = g_fb_rhs_launches[0]; return FB_OK; }
    if (!strcmp(name, "rhs_l1")) { *out = g_fb_rhs_launches[1]; return FB_OK; }
    if (!strcmp(name, "lu_batch")) { *out = g_lu_batch_launches; return FB_OK; }
    if (!strcmp(name, "lu_cell")) { *out = g_lu_cell_launches; return FB_OK; }
    return fail(FB_EINVAL, "unknown counter");
}
static int check_lu(fb_solver *s)
{
    for (auto &z : s->z) if (z.d.itypcp <= 1 && lu_plan_error(z.lu))
        return fail(FB_ECUDA, "LU wavefront: a tile timed out waiting for its neighbour (results invalid)");
    return FB_OK;
}
int fb_sync(fb_solver *s) { if (!s) return fail(FB_EINVAL, "null"); CK(cudaSetDevice(s->dev)); CK(cudaStreamSynchronize(s->st)); CK(cudaGetLastError()); return check_lu(s); }
unsigned long long fb_zone_device_ptr(fb_solver *s, int nd, int level)
{
    if (!s || nd < 0 || nd >= s->nz || level < 0 || level > 3) return 0;
    return (unsigned long long)(level == 3 ? s->z[nd].drodm : s->z[nd].rop[level]);
}
int fb_halo_pack(fb_solver *s, int which, unsigned long long buf, long long nbytes)
{
    if (!s) return fail(FB_EINVAL, "null");
    CK(cudaSetDevice(s->dev));
    std::vector<double *> t(s->nz); for (int i = 0; i < s->nz; i++) t[i] = which == 3 ? s->z[i].drodm : s->z[i].rop[which];
    if (joins_pack(s->joins, t.data(), (double *)buf, nbytes, s->st)) return fail(FB_EINVAL, "halo buffer too small");
    s->launches++;
    return FB_OK;
}
int fb_halo_unpack(fb_solver *s, int which, unsigned long long buf, long long nbytes)
{
    if (!s) return fail(FB_EINVAL, "null");
    CK(cudaSetDevice(s->dev));
    std::vector<double *> t(s->nz); for (int i = 0; i < s->nz; i++) t[i] = which == 3 ? s->z[i].drodm : s->z[i].rop[which];
    if (joins_unpack(s->joins, t.data(), (const double *)buf, nbytes, s->st)) return fail(FB_EINVAL, "halo buffer too small");
    s->launches++;
    return FB_OK;
}
unsigned long long fb_stream(fb_solver *s) { return s ? (unsigned long long)s->st : 0; }
unsigned long long fb_exchange_stream(fb_solver *s) { return s ? (unsigned long long)s->st2 : 0; }
int fb_set_exchange_cb(fb_solver *s, fb_exchange_fn cb, void *user)
{
    if (!s) return fail(FB_EINVAL, "null");
    s->xcb = cb; s->xuser = user;
    return FB_OK;
}
int fb_halo_buffers(fb_solver *s, unsigned long long *send_ptr, unsigned long long *recv_ptr)
{
    if (!s) return fail(FB_EINVAL, "null");
    if (send_ptr) *send_ptr = (unsigned long long)s->halo_send;
    if (recv_ptr) *recv_ptr = (unsigned long long)s->halo_recv;
    return FB_OK;
}

}   // extern "C"
extern "C" int fb_halo_sizes(fb_solver *s, long long *send_doubles, long long *recv_doubles)
{
    if (!s) return -1;
    if (send_doubles) *send_doubles = s->joins.send_doubles;
    if (recv_doubles) *recv_doubles = s->joins.recv_doubles;
    return 0;
}
