/*
 * fastb200.h -- C ABI of the B200-native FastS residual-and-update hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / numpy / CUDA types.  Each entry
 * point names the reference interface it replaces (paths relative to onera/Fast, `S/` = Fast/FastS/FastS,
 * `C/` = Fast/FastC/FastC).  The reference's CPython modules `fasts` / `fastc` bind these (see
 * INTEGRATION.md); in this repo `fast_b200/fasts.py` / `fastc.py` do so through ctypes.
 *
 * Conventions: every function returns 0 on success, a negative FB_E* code otherwise and leaves a
 * message readable through fb_last_error().  All host arrays are in the reference layout:
 * Parameter_int / Parameter_real (S/param_solver.h:80-312), rop(NDIMDX,NEQ) variable-major FP64,
 * node coordinates (NDIMDX_XYZ).  Device memory is owned by the solver object; time levels rotate on
 * the device in lock-step with the reference's pointer swap (S/Compute/computePT.cpp:675-679).
 */
#ifndef FASTB200_H
#define FASTB200_H
#ifdef __cplusplus
extern "C" {
#endif

#define FB_OK        0
#define FB_EINVAL   -1   /* bad argument / option outside the hot-path scope          */
#define FB_ECUDA    -2   /* CUDA runtime error (incl. "no device": there is no CPU fallback) */
#define FB_ESTATE   -3   /* call order violated (e.g. compute before init_metric)      */

typedef struct fb_solver fb_solver;

int         fb_version(void);
const char *fb_last_error(void);
int         fb_device_count(void);

/* Lifecycle.  Replaces the per-call pointer harvest of S/Compute/computePT.cpp:225-369: zones are
 * registered once and stay resident in HBM. */
int  fb_solver_create(int device, int nzones, fb_solver **out);
void fb_solver_destroy(fb_solver *s);

/* Parameter_int / Parameter_real of zone nd after C/PyTree.py:_buildOwnData + _BCcompact and the integer
 * part of fasts.allocate_metric (S/Metric/allocate_metric.cpp:124-176). */
int fb_zone_set_params(fb_solver *s, int nd, const int *param_int, int n_int,
                       const double *param_real, int n_real);
/* GridCoordinates/CoordinateX|Y|Z incl. ghost nodes (computePT.cpp:300-310). */
int fb_zone_set_coords(fb_solver *s, int nd, const double *x, const double *y, const double *z);
/* FastC.fastc.init_metric (C/Metric/init_metric.cpp:34-366): cp_tijk_, cp_vol_ on the device. */
int fb_init_metric(fb_solver *s);
int fb_zone_get_metric(fb_solver *s, int nd, double *ti, double *tj, double *tk, double *vol);

/* FlowSolution#Centers blocks: level 0 = Density (ro), 1 = Density_M1, 2 = Density_P1; n = NDIMDX*NEQ.
 * aux = [ViscosityEddy | TurbulentDistance] block (S/Compute/src_term.for:96). */
int fb_zone_upload_state(fb_solver *s, int nd, int level, const double *rop);
int fb_zone_download_state(fb_solver *s, int nd, int level, double *rop);
int fb_zone_upload_aux(fb_solver *s, int nd, const double *aux, int n);
int fb_zone_download_aux(fb_solver *s, int nd, double *aux, int n);
int fb_zone_download_work(fb_solver *s, int nd, double *drodm, double *coe);
/* HOOK['wiggle'] (C/PyTree.py:2201): the SENSOR switch of S/Compute/SENSOR/wiggle2001.for, kept on the device as one byte per
 * cell (bit d = lower face in direction d; the reference stores three doubles per cell); n = NDIMDX. */
int fb_zone_download_wig(fb_solver *s, int nd, unsigned char *wig, int n);

/* Match joins: flattened donor tree in this repo's documented layout (DESIGN.md "match joins"), the
 * stand-in for Connector's miseAPlatDonorTree__ tables read at C/Com/setInterpTransfersFast.cpp:469-477. */
int fb_set_joins(fb_solver *s, const int *join_int, int n_int);
/* The reference's own wire format: HOOK['param_int_tc'] / HOOK['param_real_tc'], the donor tree flattened by Connector's
 * miseAPlatDonorTree__ (S/PyTree.py:495), read field by field as K_FASTC::setInterpTransfersIntra does
 * (C/Com/setInterpTransfersFast.cpp:469-477, 641-654, 750-760; layout documented in the FastS technical document 2.4.3.1).
 * Match joins only (interpolation type 0 / 1, one donor of coefficient 1): IBC, chimera stencils, unsteady joins, azimuthal
 * periodicity and conservative nearmatch windows return FB_EINVAL.  Point lists are 0-based cell addresses of the ghosted zones. */
int fb_set_joins_tc(fb_solver *s, const int *param_int_tc, int n_int, const double *param_real_tc, int n_real);

/* fasts._computePT(zones, metrics, nitrun, nstep_deb, nstep_fin, layer_mode, nit_c, hook)
 * (S/Compute/computePT.cpp:44): runs nit_c iterations x sub-iterations nstep_deb..nstep_fin with
 * layer_mode=1 semantics (transfers + BC inside), rotates the time levels, updates *first_it. */
int fb_compute(fb_solver *s, int nitrun, int nstep_deb, int nstep_fin, int nit_c, int *first_it);
/* fasts._applyBC(zones, metrics, hook, nstep, var) (S/BC/applyBC.cpp:33) on time level `level`. */
int fb_apply_bc(fb_solver *s, int level);
/* Connector.___setInterpTransfers for match joins (S/PyTree.py:1218) on time level `level`. */
int fb_transfers(fb_solver *s, int level);
/* fasts._computePT_mut (S/Compute/LES/computePT_mut.cpp): initial mu / mu+mut into ViscosityEddy. */
int fb_compute_mut(fb_solver *s);

/* fasts.display_ss_iteration inputs (S/DISPLAY/display_ss_iteration.cpp:29): rdm(nssiter,neq,2) and
 * CFL (max,min,mean) of zone nd as gathered at S/Compute/computePT.cpp:619-645 from S/Compute/cpcfl3.for (reduced on the
 * device at the first sub-iteration of every check iteration, S/Compute/cptst3.for:113-120). */
int fb_zone_get_residuals(fb_solver *s, int nd, double *rdm, int n);
int fb_zone_get_cfl(fb_solver *s, int nd, double cfl[3]);
/* Device side of display_temporal_criteria's stopAtNan (S/PyTree.py:2070-2073): number of interior cells whose state was
 * not finite at the last CFL check (S/Compute/cpcfl3.for runs over them anyway) + residual norms (cprdu3s1) that came out
 * non-finite since; both are evaluated on the iterations the reference checks (nitrun % modulo_verif == 0). */
int fb_zone_isfinite(fb_solver *s, int nd, long long *bad);

/* End-to-end iteration on HOST buffers (what a reference-side caller holding the tree's numpy arrays does):
 * uploads the input time levels of every zone (ro; ro_M1 too where the BDF2 term of S/Compute/core3as2.for:37-46 reads it,
 * i.e. implicit, global time step, after the first iteration), zone by zone on a copy stream so that a zone's first kernels
 * overlap the upload of the next zones (pinned buffers make the copies asynchronous), runs one fb_compute
 * iteration (nit_c must be 1), downloads the new solution into the caller's ro_P1 buffers.  The caller then rotates
 * its three pointers like FastC.switchPointers__ (C/PyTree.py:2298).  ro/ro_m1/ro_p1[nd]: NDIMDX*NEQ doubles each. */
int fb_compute_host(fb_solver *s, int nitrun, int nit_c, int *first_it,
                    double **ro, double **ro_m1, double **ro_p1);

/* Instrumentation for bench.py: CUDA-event time (ms) and launch count accumulated per kernel family
 * since the last reset. names: "rhs", "lu", "bc", "transfer", "coe", "all". */
int fb_timing_enable(fb_solver *s, int on);
int fb_timing_reset(fb_solver *s);
int fb_timing_get(fb_solver *s, const char *name, double *ms, long long *launches);
int fb_sync(fb_solver *s);
/* Which kernels ran (process-wide): "rhs_tma" / "rhs_l1" = launches of the TMA-staged (k_rhs3) and of the L1-path
 * (k_rhs2) residual kernel, "lu_batch" = multi-zone LU launch pairs (one lower + one upper launch over all zones), "lu_cell" = sweeps run by the
 * TMA-staged thread-per-cell kernel k_lu_cell[_b] (the others ran the 8-byte cp.async kernel k_lu_tile[_b]).  New with this library
 * (no reference counterpart); bench.py names the kernel it reports a roofline for from it. */
int fb_counter(const char *name, long long *out);
/* Test hook, needs no device: the order in which the two-part explicit stage of the halo-exchange overlap visits a
 * gx x gy x gz tile grid (tiles on the faces flagged in `faces` -- bit 0..5 = imin, imax, jmin, jmax, kmin, kmax -- first),
 * computed by the function the kernels use; thin: bit d = the last tile of direction d is a single cell thick, so that the
 * upper face of that direction takes its two outermost tile layers (the donors are the two cell layers behind a face).
 * out: 3 ints (x, y, z) per tile in launch order; *nface: tiles of part 1. */
int fb_debug_tile_order(int gx, int gy, int gz, int faces, int thin, int *out, int *nface);
/* Device pointer of a zone's state block (uintptr) so the Python host can hand halo buffers to
 * torch.distributed/NCCL without a host copy. */
unsigned long long fb_zone_device_ptr(fb_solver *s, int nd, int level);
/* Multi-GPU halo exchange helpers: pack the donor cells of the cross-rank joins into / unpack from a
 * contiguous device buffer (FP64, nvar-major per join). */
int fb_halo_pack(fb_solver *s, int level_or_drodm, unsigned long long dev_buf, long long nbytes);
int fb_halo_unpack(fb_solver *s, int level_or_drodm, unsigned long long dev_buf, long long nbytes);
int fb_halo_sizes(fb_solver *s, long long *send_doubles, long long *recv_doubles);
/* Cross-rank match joins (one process per GPU).  The join table may contain kind 1 (pack into the send buffer) and
 * kind 2 (unpack from the receive buffer) entries; inside fb_compute / fb_transfers, at the two points where the
 * reference exchanges MPI messages (drodm before the LU sweep when lu_match=1, S/Compute/lhs.cpp:1-15; rop_CL after
 * it, S/Compute/gsdr3.cpp:362 -> C/Com/setInterpTransfersFast.cpp:962), the library packs on its stream, calls
 * cb(user, which) (which: 0 state, 1 drodm) -- the host transport moves fb_halo_buffers() send -> peers' recv, ordered
 * on fb_exchange_stream() -- and unpacks.  Without a callback cross-rank entries are ignored. */
typedef void (*fb_exchange_fn)(void *user, int which);
int fb_set_exchange_cb(fb_solver *s, fb_exchange_fn cb, void *user);
int fb_halo_buffers(fb_solver *s, unsigned long long *send_ptr, unsigned long long *recv_ptr);
/* Native transport (default of fast_b200.dist): the library itself moves the packed donor values between the ranks with
 * ncclSend / ncclRecv on its exchange stream -- no host callback on the data path.  libnccl.so.2 is bound at run time.
 * fb_nccl_unique_id: rank 0 draws the 128-byte NCCL id, the host side hands it to every rank (any out-of-band channel: the
 * reference would use MPI_Bcast, S/Mpi.py); fb_nccl_init: collective over the ranks, creates this rank's communicator on the
 * solver's device; fb_set_peers: per neighbour rank the slices [off, off + n) (in doubles) of the send and receive buffers that
 * fb_set_joins sized -- the message of C/Com/setInterpTransfersFast.cpp:962 (setInterpTransfersInter) towards that rank. */
int fb_nccl_unique_id(char id[128]);
int fb_nccl_init(fb_solver *s, int rank, int world, const char id[128]);
int fb_set_peers(fb_solver *s, int npeers, const int *rank, const long long *send_off, const long long *send_n,
                 const long long *recv_off, const long long *recv_n);
/* dtloc[1] of .Solver#dtloc (C/PyTree.py:1188-1200): residual / CFL check every modulo_verif iterations. */
int fb_set_modulo_verif(fb_solver *s, int modulo_verif);
/* cudaStream_t of the solver (uintptr) so that torch/NCCL work can be ordered against it. */
unsigned long long fb_stream(fb_solver *s);
/* cudaStream_t (uintptr) the library runs the cross-rank halo exchange on: the exchange callback must order its transfers
 * on THIS stream (pack precedes, unpack follows on it); the solver stream only waits for it where the ghost cells are
 * needed, so the exchange overlaps the interior tiles of an explicit stage. */
unsigned long long fb_exchange_stream(fb_solver *s);

#ifdef __cplusplus
}
#endif
#endif
