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
    """fastc.souszones_list (C/Compute/souszones_list.cpp): one task per zone on the GPU; only the two flags matter."""
    dtloc = hook['dtloc']
    verif = 1 if nitrun % int(dtloc[1]) == 0 else 0
    itypcp = 1
    lexit = 1 if (verif and nstep == int(dtloc[0]) and itypcp != 2) else 0
    return {'lexit_lu': lexit, 'lssiter_verif': verif}
