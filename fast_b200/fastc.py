"""Stand-in for the reference's CPython extension ``fastc`` (FastC/FastC/fastc.cpp:25-34) for the two entries the
hot path calls."""
from . import _lib
from .FastC import PyTree as FastC


def init_metric(zones, metrics, hook):
    """fastc.init_metric (C/Metric/init_metric.cpp:34): cp_tijk_ -> cp_vol_ (-> tijk_extrap_ for degenerate ghost
    cells, which the synthetic meshes never contain) on the device."""
    s = hook.get('solver')
    if s is None: raise _lib.FastError("fast_b200: init_metric before register")
    _lib.check(_lib.lib().fb_init_metric(s))


def souszones_list(zones, metrics, hook, nitrun, nstep, verbose=0):
    """fastc.souszones_list (C/Compute/souszones_list.cpp:98-213; flags :36-83), called at S/PyTree.py:122,258,403 as
    ``hook1.update(fastc.souszones_list(...))``: returns the reference's dictionary {'lexit_lu', 'lssiter_verif'}
    (souszones_list.cpp:198-213).  There is one task per zone on the GPU, so the sub-zone lists the reference also refreshes in
    HOOK have no counterpart; the C layer (fb_compute) evaluates the same two flags itself."""
    from .FastC.PyTree import Internal
    from . import VSHARE
    dtloc = hook['dtloc']
    verif = 1 if nitrun % int(dtloc[1]) == 0 else 0
    pi = Internal.getNodeFromName1(Internal.getNodeFromName1(zones[0], '.Solver#ownData'), 'Parameter_int')[1]
    itypcp = int(pi[VSHARE.ITYPCP])
    lexit = 1 if (verif and nstep == int(dtloc[0]) and itypcp != 2) else 0
    return {'lexit_lu': lexit, 'lssiter_verif': verif}
