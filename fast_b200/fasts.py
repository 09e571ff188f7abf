"""Stand-in for the reference's CPython extension ``fasts`` (method table: FastS/FastS/fastS.cpp:26-66) on top of
the CUDA C-ABI.  Same method names and argument order as the entries FastS.PyTree calls on the hot path."""
import ctypes
import numpy

from . import _lib, VSHARE as V
from .FastC import PyTree as FastC
from .FastC.PyTree import Internal

_vp = ctypes.c_void_p


def _dp(a):
    return a.ctypes.data_as(_lib.c_dp)


def _solver(hook):
    s = hook.get('solver') if hook is not None else None
    if s is None: raise _lib.FastError("fast_b200: zones are not registered on a GPU solver (call FastS.warmup first)")
    return s


def allocate_metric(z, nssiter, verbose=0):
    """fasts.allocate_metric(zone, nssiter) (S/Metric/allocate_metric.cpp:34): the integer part (sizes, mesh family,
    NEQ...) has already been written into Parameter_int by FastS._warmup_host; returns the ``metrics`` list of the
    zone in the reference's order [ti|tj|tk|vol, ti0, venti, ti_df, rdm, ind_dm, it_lu_ssdom, degen]
    (S/param_solver.h:338-345).  The metric block itself stays in HBM; the numpy is filled on demand."""
    own = Internal.getNodeFromName1(z, '.Solver#ownData')
    pi = Internal.getNodeFromName1(own, 'Parameter_int')[1]
    nm = int(pi[V.NDIMDX_MTR]); nij = int(pi[V.NEQ_IJ]); nk = int(pi[V.NEQ_K])
    mtr = numpy.zeros((nm, 2 * nij + nk + 1), numpy.float64, order='F')
    rdm = numpy.zeros(int(nssiter) * int(pi[V.NEQ]) * 2, numpy.float64)
    ind_dm = numpy.array([1, pi[V.IJKV], 1, pi[V.IJKV + 1], 1, pi[V.IJKV + 2]], Internal.E_NpyInt)
    return [mtr, mtr, None, None, rdm, ind_dm, numpy.zeros(1, Internal.E_NpyInt), None]


def register(zones, metrics, hook, tc=None, device=None):
    """Pointer harvest of computePT.cpp:225-369 done ONCE: create the device-resident solver, upload parameters,
    coordinates, the three time levels and [ViscosityEddy|TurbulentDistance].  Device: FAST_B200_DEVICE, else
    LOCAL_RANK (one process per GPU), else 0."""
    import os
    if device is None: device = int(os.environ.get('FAST_B200_DEVICE', os.environ.get('LOCAL_RANK', '0')))
    L = _lib.lib()
    s = _vp()
    _lib.check(L.fb_solver_create(int(device), len(zones), ctypes.byref(s)))
    hook['solver'] = s
    hook['nzones'] = len(zones)
    for nd, z in enumerate(zones):
        own = Internal.getNodeFromName1(z, '.Solver#ownData')
        pi = numpy.ascontiguousarray(Internal.getNodeFromName1(own, 'Parameter_int')[1], dtype=numpy.int32)
        pr = numpy.ascontiguousarray(Internal.getNodeFromName1(own, 'Parameter_real')[1], dtype=numpy.float64)
        _lib.check(L.fb_zone_set_params(s, nd, pi.ctypes.data_as(_lib.c_ip), pi.size, _dp(pr), pr.size))
        gc = Internal.getNodeFromName1(z, Internal.__GridCoordinates__)
        xyz = [numpy.ascontiguousarray(Internal.getNodeFromName1(gc, n)[1].ravel(order='F')) for n in
               ('CoordinateX', 'CoordinateY', 'CoordinateZ')]
        _lib.check(L.fb_zone_set_coords(s, nd, _dp(xyz[0]), _dp(xyz[1]), _dp(xyz[2])))
        for lev, suf in enumerate(('', '_M1', '_P1')):
            _lib.check(L.fb_zone_upload_state(s, nd, lev, _dp(FastC.block_of(z, suf))))
        aux = FastC.aux_block_of(z)
        if aux is not None: _lib.check(L.fb_zone_upload_aux(s, nd, _dp(aux), aux.size))
    dtloc = hook['dtloc']
    _lib.check(L.fb_set_modulo_verif(s, int(dtloc[1])))
    return s


def set_joins(hook, tc):
    """The flattened donor tree, stored like the reference in HOOK['param_int_tc'] / HOOK['param_real_tc'] (S/PyTree.py:495-500):
    tc = (param_int_tc, param_real_tc) in Connector's miseAPlatDonorTree__ layout (C/Com/setInterpTransfersFast.cpp:469-477) ->
    fb_set_joins_tc; a single int32 array = this library's own table (csrc/fb_joins.cu; it also carries the pack / unpack
    entries of the cross-rank joins of fast_b200.dist)."""
    L = _lib.lib()
    if isinstance(tc, (tuple, list)) and len(tc) == 2:
        pit = numpy.ascontiguousarray(tc[0], dtype=numpy.int32); prt = numpy.ascontiguousarray(tc[1], dtype=numpy.float64)
        hook['param_int_tc'] = pit; hook['param_real_tc'] = prt
        _lib.check(L.fb_set_joins_tc(_solver(hook), pit.ctypes.data_as(_lib.c_ip), pit.size, _dp(prt), prt.size))
        return
    ji = numpy.ascontiguousarray(tc, dtype=numpy.int32)
    hook['param_int_tc'] = ji
    _lib.check(L.fb_set_joins(_solver(hook), ji.ctypes.data_as(_lib.c_ip), ji.size))


def _computePT(zones, metrics, nitrun, nstep_deb, nstep_fin, layer_mode, nit_c, hook):
    """fasts._computePT (S/Compute/computePT.cpp:44).  Returns the time spent in transfers (0.: they are on-stream)."""
    if layer_mode != 1: raise _lib.FastError("fast_b200: only layer_mode=1 (C layer) is built")
    L = _lib.lib()
    fi = ctypes.c_int(int(hook.get('FIRST_IT', 0)))
    _lib.check(L.fb_compute(_solver(hook), int(nitrun), int(nstep_deb), int(nstep_fin), int(nit_c), ctypes.byref(fi)))
    hook['FIRST_IT'] = fi.value
    return 0.


_LEVEL = {'Density': 0, 'Density_M1': 1, 'Density_P1': 2}


def _applyBC(zones, metrics, hook, nstep, var="Density"):
    """fasts._applyBC (S/BC/applyBC.cpp:33)."""
    _lib.check(_lib.lib().fb_apply_bc(_solver(hook), _LEVEL[var]))


def transfers(zones, hook, var="Density"):
    """Connector.___setInterpTransfers restricted to match joins (S/PyTree.py:1218)."""
    _lib.check(_lib.lib().fb_transfers(_solver(hook), _LEVEL[var]))


def _computePT_mut(zones, metrics, hook):
    """fasts._computePT_mut (S/Compute/LES/computePT_mut.cpp): initial mu (+mut) into ViscosityEddy."""
    _lib.check(_lib.lib().fb_compute_mut(_solver(hook)))


def allocate_ssor(zones, metrics, hook):
    """fasts.allocate_ssor (S/Compute/LU/allocate_ssor.cpp): the reference's host-side ssor work arrays of nb_relax > 1; here the
    two per-zone device arrays of the relaxations (rhs copy, U-sum) belong to the library's LU plan (csrc/fb_lu.cu: run_lu_ssor)."""
    return None


def download(zones, hook, levels=('', '_M1', '_P1'), aux=True):
    L = _lib.lib(); s = _solver(hook)
    for nd, z in enumerate(zones):
        for suf in levels:
            lev = {'': 0, '_M1': 1, '_P1': 2}[suf]
            _lib.check(L.fb_zone_download_state(s, nd, lev, _dp(FastC.block_of(z, suf))))
        if aux:
            a = FastC.aux_block_of(z)
            if a is not None: _lib.check(L.fb_zone_download_aux(s, nd, _dp(a), a.size))


def download_work(zones, hook):
    """HOOK['rhs'] / HOOK['coe'] / HOOK['wiggle'] <- HBM (per-zone offsets of gsdr3.cpp:282-290); wiggle is expanded from the
    device's one byte per cell (bit d = lower face in direction d) to the reference's three doubles per cell."""
    L = _lib.lib(); s = _solver(hook)
    for nd, z in enumerate(zones):
        pi = Internal.getNodeFromName1(Internal.getNodeFromName1(z, '.Solver#ownData'), 'Parameter_int')[1]
        ndx, neq, nco = int(pi[V.NDIMDX]), int(pi[V.NEQ]), max(int(pi[V.NEQ_COE]), 0)
        o_r, o_c, o_w = hook['work_offsets'][nd]
        rhs = hook['rhs'][o_r:o_r + neq * ndx]
        coe = hook['coe'][o_c:o_c + nco * ndx] if nco and int(pi[V.ITYPCP]) <= 1 else None
        _lib.check(L.fb_zone_download_work(s, nd, _dp(rhs), _dp(coe) if coe is not None else None))
        if int(pi[V.KFLUDOM]) == 2:
            b = numpy.zeros(ndx, numpy.uint8)
            _lib.check(L.fb_zone_download_wig(s, nd, b.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)), ndx))
            w = hook['wiggle'][o_w:o_w + 3 * ndx].reshape(3, ndx)
            for d in range(3): w[d] = (b >> d) & 1


def timing(hook, on):
    """CUDA-event timers of the library (fb_timing_*), the GPU stand-in for HOOK['TIMER_OMP']."""
    L = _lib.lib(); s = _solver(hook)
    _lib.check(L.fb_timing_enable(s, 1 if on else 0))
    if on: _lib.check(L.fb_timing_reset(s))
    hook['timing'] = bool(on); hook['timed_subit'] = 0


def timing_get(hook, name='all'):
    L = _lib.lib(); v = ctypes.c_double(0); n = ctypes.c_longlong(0)
    _lib.check(L.fb_timing_get(_solver(hook), name.encode(), ctypes.byref(v), ctypes.byref(n)))
    return v.value, int(n.value)


def get_metric(zones, metrics, hook):
    """Fill the host numpy of the metrics list from HBM (tests / post-processing)."""
    L = _lib.lib(); s = _solver(hook)
    out = []
    for nd, z in enumerate(zones):
        pi = Internal.getNodeFromName1(Internal.getNodeFromName1(z, '.Solver#ownData'), 'Parameter_int')[1]
        nm = int(pi[V.NDIMDX_MTR]); nij = int(pi[V.NEQ_IJ]); nk = int(pi[V.NEQ_K])
        ti = numpy.zeros(nm * nij); tj = numpy.zeros(nm * nij); tk = numpy.zeros(nm * max(nk, 1)); vol = numpy.zeros(nm)
        _lib.check(L.fb_zone_get_metric(s, nd, _dp(ti), _dp(tj), _dp(tk), _dp(vol)))
        out.append((ti, tj, tk[:nm * nk], vol))
    return out


def display_ss_iteration(zones, metrics, nitrun, nssiter, lft, iverb):
    """fasts.display_ss_iteration (S/DISPLAY/display_ss_iteration.cpp:29): returns {zone: {'L2','Loo','CFL'}} with
    L2 = sqrt(sum drodm^2 / ncells) and Loo = max|drodm| per sub-iteration and equation (dpssiter.for)."""
    L = _lib.lib(); s = _solver(FastC.HOOK)
    out = {}
    for nd, z in enumerate(zones):
        pi = Internal.getNodeFromName1(Internal.getNodeFromName1(z, '.Solver#ownData'), 'Parameter_int')[1]
        neq = int(pi[V.NEQ]); n = int(nssiter) * neq * 2
        rdm = metrics[nd][V.METRIC_RDM]
        if rdm.size != n: rdm = numpy.zeros(n)
        _lib.check(L.fb_zone_get_residuals(s, nd, _dp(rdm), n))
        cfl = numpy.zeros(3); _lib.check(L.fb_zone_get_cfl(s, nd, _dp(cfl)))
        own = Internal.getNodeFromName1(z, '.Solver#ownData')
        Internal.getNodeFromName1(own, 'CFL_minmaxmoy')[1][:] = [cfl[1], cfl[0], cfl[2]]
        r = rdm.reshape((int(nssiter), neq, 2), order='F')
        ncell = float(pi[V.IJKV]) * float(pi[V.IJKV + 1]) * float(pi[V.IJKV + 2])
        bad = ctypes.c_longlong(0); _lib.check(L.fb_zone_isfinite(s, nd, ctypes.byref(bad)))
        out[z[0]] = {'L2': numpy.sqrt(r[:, :, 0] / ncell), 'Loo': r[:, :, 1].copy(), 'CFL': cfl, 'nonfinite': int(bad.value)}
        if iverb == 1: print("zone %s nitrun %d: L2 %s" % (z[0], nitrun, out[z[0]]['L2'][:, 0]))
    return out
