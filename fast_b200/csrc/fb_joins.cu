// fb_joins.cu -- match-join ("raccord") ghost transfers.
//
// Reference: K_FASTC::setInterpTransfersIntra (C/Com/setInterpTransfersFast.cpp:440-955) computes, for every join,
//   rcv[var][rcvPts[n]] = sum_c coef[c] * dnr[var][donorPts[n] + off_c]
// For exact match joins the stencil is one donor cell with coefficient 1 (ST/tgvMpi.py:136-139), i.e. an indexed
// copy of the neq variables.  One launch handles all joins of a kind: blockIdx.y = join, threads over points.
// Join table handed to fb_set_joins: [njoin, total_idx, {kind, dz, rz, npts, off_d, off_r, boff} x njoin, idx...]
//   kind 0: donor and receiver zones on this GPU;  kind 1: donor here, packed into the send buffer at boff;
//   kind 2: receiver here, unpacked from the receive buffer at boff (buffers: var-major per join, FP64).
// off_d / off_r index the trailing idx array (0-based cell addresses l); unused side = -1.
#include <vector>
#include <string>
#include "fb_implicit.cuh"

#define JD 10   // device descriptor: kind, dz, rz, npts, off_d, off_r, boff, neq, nd_d, nd_r

__global__ void k_join(const int *__restrict__ desc, const int *__restrict__ idx, double *const *__restrict__ zp,
                       double *__restrict__ buf, int kind)
{
    const int *d = desc + JD * blockIdx.y;
    if (d[0] != kind) return;
    const int npts = d[3], neq = d[7];
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= npts) return;
    if (kind == 0) {
        const double *src = zp[d[1]]; double *dst = zp[d[2]];
        const int ls = idx[d[4] + n], lr = idx[d[5] + n];
        for (int e = 0; e < neq; e++) dst[lr + e * d[9]] = src[ls + e * d[8]];
    } else if (kind == 1) {
        const double *src = zp[d[1]];
        const int ls = idx[d[4] + n];
        for (int e = 0; e < neq; e++) buf[d[6] + e * npts + n] = src[ls + e * d[8]];
    } else {
        double *dst = zp[d[2]];
        const int lr = idx[d[5] + n];
        for (int e = 0; e < neq; e++) dst[lr + e * d[9]] = buf[d[6] + e * npts + n];
    }
}

int joins_build(JoinSet &j, const int *ji, int n, int nz, const int *ndimdx, const int *neqs, std::string &err)
{
    joins_free(j);
    if (n < 2) { err = "table too short"; return 1; }
    const int nj = ji[0], nidx = ji[1];
    if (nj < 0 || n != 2 + 7 * nj + nidx) { err = "size mismatch: expected 2+7*njoin+nidx ints"; return 1; }
    std::vector<int> desc((size_t)JD * (nj > 0 ? nj : 1), 0);
    j.maxpts = 0; j.send_doubles = j.recv_doubles = 0;
    for (int q = 0; q < nj; q++) {
        const int *s = ji + 2 + 7 * q; int *d = desc.data() + JD * q;
        for (int t = 0; t < 7; t++) d[t] = s[t];
        const int kind = s[0], dz = s[1], rz = s[2], npts = s[3];
        if (kind < 0 || kind > 2 || npts < 0) { err = "bad join kind / size"; return 1; }
        if (kind != 2 && (dz < 0 || dz >= nz || s[4] < 0 || s[4] + npts > nidx)) { err = "bad donor zone / offset"; return 1; }
        if (kind != 1 && (rz < 0 || rz >= nz || s[5] < 0 || s[5] + npts > nidx)) { err = "bad receiver zone / offset"; return 1; }
        const int zref = kind == 2 ? rz : dz;
        d[7] = neqs[zref];
        d[8] = kind != 2 ? ndimdx[dz] : 0; d[9] = kind != 1 ? ndimdx[rz] : 0;
        if (kind == 0 && neqs[dz] != neqs[rz]) { err = "donor and receiver zones have different NEQ"; return 1; }
        if (npts > j.maxpts) j.maxpts = npts;
        if (kind == 1 && s[6] + (long long)npts * d[7] > j.send_doubles) j.send_doubles = s[6] + (long long)npts * d[7];
        if (kind == 2 && s[6] + (long long)npts * d[7] > j.recv_doubles) j.recv_doubles = s[6] + (long long)npts * d[7];
    }
    j.njoin = nj; j.nz = nz; j.h_desc = desc;
    if (nj == 0) return 0;
    if (cudaMalloc(&j.d_desc, desc.size() * sizeof(int)) != cudaSuccess) { err = "cudaMalloc"; return 1; }
    cudaMemcpy(j.d_desc, desc.data(), desc.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (cudaMalloc(&j.d_idx, (size_t)(nidx > 0 ? nidx : 1) * sizeof(int)) != cudaSuccess) { err = "cudaMalloc"; return 1; }
    cudaMemcpy(j.d_idx, ji + 2 + 7 * nj, (size_t)nidx * sizeof(int), cudaMemcpyHostToDevice);
    if (cudaMalloc(&j.d_ptrs, 2 * (size_t)nz * sizeof(double *)) != cudaSuccess) { err = "cudaMalloc"; return 1; }
    return 0;
}

// Connector's flattened donor tree (miseAPlatDonorTree__): HOOK['param_int_tc'] / HOOK['param_real_tc'] as
// K_FASTC::setInterpTransfersIntra reads them (C/Com/setInterpTransfersFast.cpp:469-477 header, :563-566 and :641-654 per-join
// tables, :750-760 point lists; documented in TD 2.4.3.1).  Layout followed here:
//   sizecomIBC = pit[2], sizecomID = pit[3 + sizecomIBC], shift_graph = sizecomIBC + sizecomID + 3,
//   ech = pit[NoTransfert + shift_graph] (NoTransfert = 1: the joins inside this process),
//   nrac = pit[ech+1], nrac_inst = pit[ech+2], timelevel = pit[ech+3], tables of nrac ints from T = ech + 4 + 2 timelevel:
//   T+0 nrac: size of the donor point list   T+1: number of coefficients   T+3: IBC type (< 0: plain interpolation)
//   T+5: donor zone   T+6: address of the donor points   T+7: address of [ntype, count per type..., type per point...]
//   T+8: address of the coefficients in param_real_tc   T+10: receiver points   T+11: receiver zone   T+12: address of the
//   receiver points   T+13: variables (5 / 6)   T+14: azimuthal periodicity   T+16: conservative-flux windows.
// Scope of this library: match joins -- interpolation type 0 / 1 with ONE donor of coefficient 1 (ST/tgvMpi.py:136-139); anything
// else (IBC, chimera stencils, unsteady joins, rotation, nearmatch flux windows) is rejected loudly.  Point addresses are the
// 0-based cell addresses l of the ghosted zones.  Converts to the internal table and goes through joins_build.
int joins_build_tc(JoinSet &j, const int *pit, int n_int, const double *prt, int n_real, int nz, const int *ndimdx, const int *neqs,
                   std::string &err)
{
    auto bad = [&](const char *m) { err = m; return 1; };
    if (n_int < 6) return bad("param_int_tc too short");
    const int sizecomIBC = pit[2];
    if (sizecomIBC < 0 || 3 + sizecomIBC >= n_int) return bad("param_int_tc: bad IBC communication header");
    const int sizecomID = pit[3 + sizecomIBC];
    const int shift_graph = sizecomIBC + sizecomID + 3;
    const int NoTransfert = 1;
    if (sizecomID < 0 || NoTransfert + shift_graph >= n_int) return bad("param_int_tc: bad ID communication header");
    const int ech = pit[NoTransfert + shift_graph];
    if (ech < 0 || ech + 4 > n_int) return bad("param_int_tc: bad address of the intra-process exchange");
    const int nrac = pit[ech + 1], nrac_inst = pit[ech + 2], timelevel = pit[ech + 3];
    if (nrac < 0 || timelevel < 0) return bad("param_int_tc: negative join count");
    if (nrac_inst != 0) return bad("unsteady joins (nrac_inst > 0) are outside the hot-path scope");
    const int T = ech + 4 + timelevel * 2;
    if (T + 18 * nrac > n_int) return bad("param_int_tc: join tables run past the end of the array");
    std::vector<int> tab(2 + 7 * (size_t)nrac), idx;
    tab[0] = nrac;
    for (int irac = 0; irac < nrac; irac++) {
        const int sr = T + irac;
        if (pit[sr + nrac * 3] >= 0) return bad("IBC joins are outside the hot-path scope");
        const int NoD = pit[sr + nrac * 5], NoR = pit[sr + nrac * 11], nv = pit[sr + nrac * 13];
        const int npts = pit[sr + nrac * 10], ncoef = pit[sr + nrac];
        if (pit[sr + nrac * 14] != 0) return bad("azimuthal periodicity (rotation) is outside the hot-path scope");
        if (pit[sr + nrac * 16] != 0) return bad("conservative nearmatch flux windows are outside the hot-path scope");
        if (NoD < 0 || NoD >= nz || NoR < 0 || NoR >= nz) return bad("param_int_tc: zone number out of range");
        if (nv != neqs[NoD] || nv != neqs[NoR]) return bad("join carries a number of variables different from the zones' NEQ");
        const int pt = pit[sr + nrac * 7], pd = pit[sr + nrac * 6], pr_ = pit[sr + nrac * 12], pc = pit[sr + nrac * 8];
        if (npts < 0 || pt < 0 || pt + 1 >= n_int || pd < 0 || pd + npts > n_int || pr_ < 0 || pr_ + npts > n_int)
            return bad("param_int_tc: point list address out of range");
        const int ntype = pit[pt];
        if (ntype < 0 || pt + 1 + ntype + npts > n_int) return bad("param_int_tc: interpolation type list out of range");
        int tot = 0;
        for (int q = 0; q < ntype; q++) tot += pit[pt + 1 + q];
        if (tot != npts) return bad("param_int_tc: interpolation type counts do not add up to the receiver points");
        for (int n = 0; n < npts; n++) {
            const int ty = pit[pt + 1 + ntype + n];
            if (ty != 0 && ty != 1) return bad("only match joins (interpolation type 0 / 1, one donor) are built");
        }
        if (ncoef != npts || pc < 0 || pc + npts > n_real) return bad("match joins carry one coefficient per point");
        for (int n = 0; n < npts; n++) if (prt[pc + n] != 1.) return bad("match-join coefficient different from 1");
        for (int n = 0; n < npts; n++) {
            const int ld = pit[pd + n], lr = pit[pr_ + n];
            if (ld < 0 || ld >= ndimdx[NoD] || lr < 0 || lr >= ndimdx[NoR]) return bad("param_int_tc: cell address out of range");
        }
        int *d = tab.data() + 2 + 7 * irac;
        d[0] = 0; d[1] = NoD; d[2] = NoR; d[3] = npts; d[4] = (int)idx.size(); d[5] = (int)idx.size() + npts; d[6] = 0;
        idx.insert(idx.end(), pit + pd, pit + pd + npts);
        idx.insert(idx.end(), pit + pr_, pit + pr_ + npts);
    }
    tab[1] = (int)idx.size();
    tab.insert(tab.end(), idx.begin(), idx.end());
    return joins_build(j, tab.data(), (int)tab.size(), nz, ndimdx, neqs, err);
}

void joins_free(JoinSet &j)
{
    cudaFree(j.d_desc); cudaFree(j.d_idx); cudaFree(j.d_ptrs);
    j.d_desc = nullptr; j.d_idx = nullptr; j.d_ptrs = nullptr; j.njoin = 0;
}

static int run(JoinSet &j, double **zone_ptr, int nz, double *buf, int kind, cudaStream_t st)
{
    if (j.njoin == 0) return 0;
    bool any = false;
    for (int q = 0; q < j.njoin; q++) if (j.h_desc[JD * q] == kind) any = true;
    if (!any) return 0;
    // the zone pointers rotate every iteration: pass them through a small device table (two slots, alternating,
    // so that an in-flight launch never sees its table overwritten)
    j.flip ^= 1;
    double **tab = j.d_ptrs + j.flip * nz;
    cudaMemcpyAsync(tab, zone_ptr, nz * sizeof(double *), cudaMemcpyHostToDevice, st);
    dim3 b(128), g((j.maxpts + 127) / 128, j.njoin);
    k_join<<<g, b, 0, st>>>(j.d_desc, j.d_idx, tab, buf, kind);
    return 1;
}
int joins_apply(JoinSet &j, double **zone_ptr, int nz, cudaStream_t st) { return run(j, zone_ptr, nz, nullptr, 0, st); }
int joins_pack(JoinSet &j, double **zone_ptr, double *buf, long long nbytes, cudaStream_t st)
{
    if (nbytes < j.send_doubles * 8) return 1;
    run(j, zone_ptr, j.nz, buf, 1, st); return 0;
}
int joins_unpack(JoinSet &j, double **zone_ptr, const double *buf, long long nbytes, cudaStream_t st)
{
    if (nbytes < j.recv_doubles * 8) return 1;
    run(j, zone_ptr, j.nz, const_cast<double *>(buf), 2, st); return 0;
}
