// fb_implicit.cuh -- implicit chain (LU-SGS coefficients, residual norms, hyperplane LU sweep) and match-join
// transfers; implemented in fb_implicit.cu / fb_joins.cu.
#pragma once
#include <string>
#include <vector>
#include "fb_common.cuh"

struct LuPlan {
    int box[6] = {0, 0, 0, 0, 0, 0};   // sweep box = ishift_lu (S/BC/BCzone.cpp:136-143)
    int nplanes = 0;
    int ndefault_rhs = 0;             // leading default-extrapolation windows of the rhs BC table
    int modulo_verif = 0;
    // tile-pipelined wavefront (fb_lu.cu)
    int tj = 8, tk = 8, NJ = 0, NK = 0;   // tile size in j and k (lines), tiles per direction of the (j,k) cross-section
    bool cell = false;                // thread-per-cell kernel k_lu_cell (TMA-staged) instead of k_lu_tile
    void *d_tm = nullptr;             // k_lu_cell: 9 CUtensorMap in device memory: drodm, coe, tcm, then {lower, upper} of up to 3 rop levels
    mutable const double *rop_seen[3] = {nullptr, nullptr, nullptr};
    int *d_order = nullptr;           // ticket -> tile, i+j+k order
    int *d_ctl = nullptr;             // [ticket lower, ticket upper, error flag, pad]
    double *d_mb = nullptr;           // inter-tile mailboxes: j interfaces then k interfaces
    size_t n_ifj = 0, n_ifk = 0;
    double *d_tcm = nullptr;          // cell-centred mean normals (3dfull: 9 x NDIMDX, 3dhomo: 5 x ni*nj)
    int nrelax = 1;                   // nb_relax (set before lu_plan_build); >= 2: LU-SSOR relaxations (fb_lu.cu: run_lu_ssor)
    double *d_rhs = nullptr, *d_us = nullptr;   // nb_relax >= 2: rhs of the Newton step, U-sum of the previous relaxation
    long long tstride = 0;
};
int  lu_plan_build(LuPlan &p, const ZoneDev &z, const int box[6]);
void lu_plan_free(LuPlan &p);
void lu_plan_metrics(LuPlan &p, const ZoneDev &z, cudaStream_t st);   // after the metrics exist
int  lu_plan_error(const LuPlan &p);
int  launch_lu_tiles(const ZoneDev &z, const LuPlan &p, const double *rop_ssiter, double *rop_out, int lu_match, cudaStream_t st);
// all zones of a multi-zone solver swept by one pair of launches (fb_lu.cu: k_lu_tile_b); nullptr when the zones cannot share
// a launch (different mesh family / number of equations)
struct LuBatch;
LuBatch *lu_batch_create(int nz, const ZoneDev *const *zs, const LuPlan *const *plans);
void lu_batch_free(LuBatch *b);
int  lu_batch_launch(LuBatch *b, const ZoneDev *const *zs, const LuPlan *const *plans, const double *const *rop_ssiter,
                     double *const *rop_out, int lu_match, cudaStream_t st);
void launch_coe(const ZoneDev &z, const double *rop, int first_it, cudaStream_t st);
void launch_cfl(const ZoneDev &z, const double *rop, double *cflv, cudaStream_t st);   // cflv: [max, min, sum, non-finite cells]
void launch_sa_ring(const ZoneDev &z, const double *rop, cudaStream_t st);
void launch_resid(const ZoneDev &z, double *rdm, double *cflv, int nitcfg, int nss, cudaStream_t st);
int  launch_lu(const ZoneDev &z, const LuPlan &p, const double *rop_ssiter, double *rop_out, int lu_match, cudaStream_t st);

// match joins (1 donor, coefficient 1: C/Com/setInterpTransfersFast.cpp:440-955 for interpolation type 0/1)
struct JoinSet {
    int njoin = 0;
    int *d_desc = nullptr;     // njoin x 8 ints: kind, dz, rz, npts, off_d, off_r, boff, neq
    int *d_idx = nullptr;      // donor / receiver cell addresses (0-based l)
    int maxpts = 0;
    std::vector<int> h_desc;
    long long send_doubles = 0, recv_doubles = 0;
    double **d_ptrs = nullptr; // per-zone base pointers (refreshed at every apply)
    int nz = 0;
    int flip = 0;              // slot of d_ptrs the next launch uses
};
int  joins_build(JoinSet &j, const int *ji, int n, int nz, const int *ndimdx, const int *neqs, std::string &err);
int  joins_build_tc(JoinSet &j, const int *pit, int n_int, const double *prt, int n_real, int nz, const int *ndimdx, const int *neqs,
                    std::string &err);   // Connector's param_int_tc / param_real_tc (match joins)
void joins_free(JoinSet &j);
int  joins_apply(JoinSet &j, double **zone_ptr, int nz, cudaStream_t st);          // kind 0 (local) joins
int  joins_pack(JoinSet &j, double **zone_ptr, double *buf, long long nbytes, cudaStream_t st);      // kind 1
int  joins_unpack(JoinSet &j, double **zone_ptr, const double *buf, long long nbytes, cudaStream_t st); // kind 2
