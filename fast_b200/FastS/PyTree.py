"""Drop-in mirror of ``FastS.PyTree`` for the hot path (reference: Fast/FastS/FastS/PyTree.py).

``warmup`` (:341-545), ``_compute`` (:42-236), ``_applyBC`` (:1158), ``_fillGhostcells`` (:1162),
``display_temporal_criteria`` (:2039) keep the reference's names, arguments and tree side effects; the C
extension modules they call (``fasts``, ``fastc``) are replaced by ``fast_b200.fasts`` / ``fast_b200.fastc``,
thin ctypes layers over the CUDA C-ABI (include/fastb200.h).  There is no CPU fallback: without a GPU or
without the built library these calls raise.
"""
import numpy

from ..FastC import PyTree as FastC
from ..FastC.PyTree import Internal
from .. import VSHARE, geom

__version__ = "0.1-b200"


def _zone_arrays(z):
    own = Internal.getNodeFromName1(z, '.Solver#ownData')
    pi = Internal.getNodeFromName1(own, 'Parameter_int')[1]
    pr = Internal.getNodeFromName1(own, 'Parameter_real')[1]
    gc = Internal.getNodeFromName1(z, Internal.__GridCoordinates__)
    xyz = [Internal.getNodeFromName1(gc, n)[1] for n in ('CoordinateX', 'CoordinateY', 'CoordinateZ')]
    return own, pi, pr, xyz


def _warmup_host(t, tc=None, Padding=None, verbose=0):
    """Host half of warmup (PyTree.py:341-470): option freezing, sizes and mesh family (integer part of
    fasts.allocate_metric), BC compaction, primitive variables, contiguous time levels.  No GPU needed."""
    FastC._buildOwnData(t, Padding)
    zones = Internal.getZones(t)
    for z in zones:
        own, pi, pr, xyz = _zone_arrays(z)
        deg = geom.fill_sizes(pi, *xyz)
        if deg is not None and numpy.any(deg):
            # C/Metric/init_metric.cpp:318 -> tijk_extrap.for / dist_extrap.for rebuild the metrics (and the wall distance) of
            # degenerate ghost cells from their interior mirror; that repair is not built here, and computing on collapsed
            # cells would silently give other numbers than the reference
            raise ValueError("fast_b200: zone %s has %d degenerate cells (edges <= 2e-8, nature_geom_dom.for:36-160); the "
                             "tijk_extrap / dist_extrap repair of degenerate ghost cells is outside the hot-path scope"
                             % (z[0], int(numpy.count_nonzero(deg))))
        Internal.getNodeFromName1(own, 'type_zone')[1] = numpy.array([pi[VSHARE.ITYPZONE]], Internal.E_NpyInt)
        if verbose > 0:
            names = {0: '3D curvilinear', 1: '3D homogenous k direction with constant step',
                     2: '3D Cartesian with constant step', 3: '2D curvilinear'}
            print("typezone: %s, %s (%d, %d, %d)" % (names[int(pi[VSHARE.ITYPZONE])], z[0], pi[VSHARE.IJKV],
                                                      pi[VSHARE.IJKV + 1], pi[VSHARE.IJKV + 2]))
    FastC._BCcompact(t)
    first_it = FastC._createPrimVars(t, FastC.OMP_MODE)
    FastC._compact(t)
    FastC.FIRST_IT = first_it
    return zones


def host_views(t):
    """[(Parameter_int, Parameter_real, [ro, ro_M1, ro_P1] blocks, aux block, (x,y,z))] per zone."""
    out = []
    for z in Internal.getZones(t):
        own, pi, pr, xyz = _zone_arrays(z)
        out.append((pi, pr, [FastC.block_of(z, s) for s in ('', '_M1', '_P1')], FastC.aux_block_of(z), xyz))
    return out


def warmup(t, tc, graph=None, infos_ale=None, Adjoint=False, tmy=None, list_graph=None, Padding=None, verbose=0,
           exchange=None):
    """Perform necessary operations to prepare the tree for computation (PyTree.py:341).
    Returns (t, tc, metrics).  One process per GPU (the reference's FastS.Mpi.warmup, S/Mpi.py): `exchange(solver)` builds
    the cross-rank transport (fast_b200.dist.HaloExchange) once the join table -- with its pack / unpack entries -- is
    on the device and before the first ghost fill."""
    from .. import fasts, fastc
    if graph is not None or infos_ale is not None or Adjoint or tmy is not None:
        raise ValueError("fast_b200: MPI graphs, ALE, adjoint and statistics trees are outside the hot-path scope")
    zones = _warmup_host(t, tc, Padding, verbose)
    dtloc = Internal.getNodeFromName1(Internal.getNodeFromName1(t, '.Solver#ownData'), '.Solver#dtloc')[1]
    FastC.HOOK = FastC.createWorkArrays__(zones, dtloc, FastC.FIRST_IT)
    nssiter = int(dtloc[0])
    metrics = [fasts.allocate_metric(z, nssiter, verbose) for z in zones]          # PyTree.py:291
    fasts.register(zones, metrics, FastC.HOOK, tc)                                   # pointer harvest, once
    fastc.init_metric(zones, metrics, FastC.HOOK)                                    # PyTree.py:326
    if tc is not None: fasts.set_joins(FastC.HOOK, tc)                               # miseAPlatDonorTree__ stand-in
    if exchange is not None: FastC.HOOK['halo_exchange'] = exchange(FastC.HOOK['solver'])
    # first ghost fill on Density (PyTree.py:536) then initial viscosity (:543)
    _fillGhostcells(zones, tc, metrics, 1, ['Density'], 1, hook1=FastC.HOOK)
    fasts._computePT_mut(zones, metrics, FastC.HOOK)
    return (t, tc, metrics)


def _compute(t, metrics, nitrun, tc=None, graph=None, layer="c", NIT=1, ucData=None, vtune=False):
    """Compute NIT iterations (PyTree.py:42).  Only the reference's default ``layer="c"`` driver exists here:
    sub-iterations, match-join transfers and BC fill all run inside fasts._computePT."""
    from .. import fasts
    if graph is not None or ucData is not None:
        raise ValueError("fast_b200: MPI graphs / unsteady chimera data are outside the hot-path scope")
    if layer != "c": raise ValueError('fast_b200: only layer="c" is built (the Python layer loops in the reference)')
    zones = Internal.getZones(t)
    dtloc = FastC.HOOK['dtloc']
    nitmax = int(dtloc[0])
    FastC.HOOK['FIRST_IT'] = FastC.FIRST_IT
    fasts._computePT(zones, metrics, nitrun, 1, nitmax, 1, NIT, FastC.HOOK)          # PyTree.py:219
    # time levels rotated on the device; rotate the tree's numpys the same way (PyTree.py:226-231)
    # -- switchPointers__(zones, case, nitmax%2+2) for explicit schemes: order 2 (RK2, an even number of stages) is a no-op,
    # the last stage has written Density itself
    itypcp = int(Internal.getNodeFromName1(Internal.getNodeFromName1(zones[0], '.Solver#ownData'), 'Parameter_int')[1][VSHARE.ITYPCP])
    order = 3 if itypcp < 2 else nitmax % 2 + 2
    for _ in range(NIT): FastC.switchPointers__(zones, 1, order)
    FastC.FIRST_IT = 1
    FastC.HOOK['FIRST_IT'] = 1
    if FastC.HOOK.get('timing'):
        FastC.HOOK['timed_subit'] += NIT * (nitmax - 1 if itypcp < 2 else nitmax)
    return None


def _applyBC(t, metrics, hook1, nstep, var="Density"):
    from .. import fasts
    zones = Internal.getZones(t)
    fasts._applyBC(zones, metrics, hook1, nstep, var)
    return None


def _fillGhostcells(zones, tc, metrics, timelevel_target, vars, nstep, omp_mode=0, hook1=None, nitmax=1, rk=1,
                    exploc=0, num_passage=1, gradP=False, TBLE=False, isWireModel=False):
    """Match-join transfers then physical BCs on ``vars[0]`` (PyTree.py:1162-1260)."""
    from .. import fasts
    if hook1 is None: hook1 = FastC.HOOK
    if tc is not None: fasts.transfers(zones, hook1, vars[0])
    fasts._applyBC(zones, metrics, hook1, 1, vars[0])
    return None


def sync_tree(t, levels=('', '_M1', '_P1'), aux=True):
    """Copy the device-resident state back into the tree's numpys (what save / display / post need)."""
    from .. import fasts
    fasts.download(Internal.getZones(t), FastC.HOOK, levels, aux)


def sync_work(t):
    """Copy the device-resident work arrays into HOOK['rhs'] / HOOK['coe'] / HOOK['wiggle'] (reference layout)."""
    from .. import fasts
    fasts.download_work(Internal.getZones(t), FastC.HOOK)


def display_cpu_efficiency(t, mask_cpu=0.08, mask_cell=0.01, diag='compact', FILEOUT='listZonesSlow.dat', FILEOUT1='diagCPU.dat',
                           RECORD=None):
    """Cell updates per second (PyTree.py:2673-2806).  The reference derives, per zone, the time per cell and sub-iteration from
    HOOK['TIMER_OMP'] and stores [mean, min, max] cell updates / s in the zone's `.Solver#define/Cups` node (:2764-2772:
    1/CUPS = microseconds per (iteration x sub-iteration) per cell).  Here the clock is the library's CUDA-event timer over the
    whole sub-iteration: the zones of a solver share its kernels (one LU launch sweeps all of them), so every zone gets the
    solver-wide figure.  First call: starts the timers and returns None, like the reference when too few iterations have been
    sampled (:2692-2695); later calls return microseconds per cell and sub-iteration (the reference's RECORD value)."""
    from .. import fasts
    hook = FastC.HOOK
    if not hook.get('timing') or hook.get('timed_subit', 0) == 0:
        fasts.timing(hook, True)
        if RECORD is None: print('nombre iterations insuffisant pour diagnostic: nitrun * ss_iteration > 15')
        return None
    ms, _ = fasts.timing_get(hook, 'all')
    zones = Internal.getZones(t)
    cells = []
    for z in zones:
        pi = Internal.getNodeFromName1(Internal.getNodeFromName1(z, '.Solver#ownData'), 'Parameter_int')[1]
        cells.append(int(pi[VSHARE.IJKV]) * int(pi[VSHARE.IJKV + 1]) * int(pi[VSHARE.IJKV + 2]))
    tps_percell = ms * 1e3 / (hook['timed_subit'] * float(sum(cells)))          # microseconds per cell and sub-iteration
    tps_percell = max(tps_percell, 1.e-11)
    with open(FILEOUT1, 'w') as f1:
        for z, c in zip(zones, cells):
            perfo = numpy.empty(3, dtype=numpy.float64)
            perfo[:] = int(1.e6 / tps_percell)
            solver_def = Internal.createUniqueChild(z, '.Solver#define', 'UserDefinedData_t')
            Internal.createUniqueChild(solver_def, 'Cups', 'DataArray_t', value=perfo)
            f1.write('cpumoy/cell=  %s  dim zone=%d  %s\n' % (tps_percell, c, z[0]))
            if RECORD is None: print('cpu/cell zone=', z[0], 'moy=', tps_percell, 'dim zone=', c, 'dim tot=', sum(cells))
        f1.write('cpu moyen %%cell en microsec: %s\n' % tps_percell)
    open(FILEOUT, 'w').close()          # no zone is slower than the others: they share the kernels
    if RECORD is None: print('cpu moyen %cell en microsec: ', tps_percell)
    return tps_percell if RECORD is not None else None


def display_temporal_criteria(t, metrics, nitrun, format=None, gmres=None, verbose='firstlast', stopAtNan=True):
    """Display CFL and implicit convergence (PyTree.py:2039).  Returns the per-zone dict that the reference prints."""
    from .. import fasts
    zones = Internal.getZones(t)
    dtloc = FastC.HOOK['dtloc']
    nssiter = int(dtloc[0])
    out = fasts.display_ss_iteration(zones, metrics, nitrun, nssiter, 0, {'firstlast': 0, 'all': 1, 'none': 2}.get(verbose, 0))
    if stopAtNan:
        for zname, d in out.items():
            if d['nonfinite'] or not numpy.all(numpy.isfinite(d['L2'])):
                raise FloatingPointError("fast_b200: NaN in the state / Newton residual of zone %s (%d non-finite cells or norms seen "
                                         "by the device checks)" % (zname, d['nonfinite']))
    return out
