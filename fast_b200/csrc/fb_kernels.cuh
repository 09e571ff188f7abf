// fb_kernels.cuh -- launchers shared between fb_kernels.cu / fb_implicit.cu and the C-ABI driver (fb_api.cu).
#pragma once
#include <atomic>
#include "fb_common.cuh"

// cudaFuncSetAttribute is per device: true the first time a call site runs on the current device (one bit per device)
static inline bool fb_first_on_device(std::atomic<unsigned long long> &done)
{
    int d = 0; cudaGetDevice(&d);
    const unsigned long long bit = 1ull << (d & 63);
    return !(done.fetch_or(bit) & bit);
}

struct BcWin {          // one BC window, ghost range already resolved (C/BC/indice_cl_sdm.for, npass = 0)
    int type, idir;     // reference tag (C/PyTree.py:30-70), 1..6
    int w[6];           // ghost cells imin..kmax (1-based interior numbering, ghosts <= 0 or > n)
};

void launch_metric(const ZoneDev &z, cudaStream_t st);
void launch_invist(const ZoneDev &z, const double *rop, int ext, cudaStream_t st);
void launch_sensor(const ZoneDev &z, const double *rop, cudaStream_t st);
bool fb_sensor_fused();
// explicit stage (init_rhs + flux balance + core3ark3) or implicit rhs (src + flux balance + core3as2)
// part: 0 the whole zone, 1 only the tiles that touch a zone face, 2 only the interior tiles (halo-exchange overlap of the
// explicit stages; implicit zones always run whole); faces: bit 0..5 = imin .. kmax faces that carry cross-rank joins
void launch_rhs(const ZoneDev &z, const double *rop, double *rop_out, const double *ro_n, const double *ro_n1,
                int nitcfg, int first_it, cudaStream_t st, int part = 0, int faces = 63, int pdl = 0);
void launch_bc(const ZoneDev &z, const BcWin *d_wins, int nwin, int maxcells, double *rop, int lrhs, cudaStream_t st);
void launch_corners(const ZoneDev &z, double *rop, const int box[6], cudaStream_t st);
