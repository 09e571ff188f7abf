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
