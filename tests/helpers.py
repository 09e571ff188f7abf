"""Shared test plumbing: tree -> oracle zones, comparisons."""
import os, sys
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import oracle as orc                                    # noqa: E402  (test infrastructure)
from fast_b200.FastS import PyTree as FastS            # noqa: E402
from fast_b200.FastC import PyTree as FastC            # noqa: E402
from fast_b200 import VSHARE as V                      # noqa: E402


def oracle_from_tree(t, joins=None, fast=False):
    """Host warmup (no GPU) then one oracle Zone per tree zone."""
    FastS._warmup_host(t)
    zs = []
    for pi, pr, lev, aux, xyz in FastS.host_views(t):
        zs.append(orc.Zone(pi, pr, lev[0], lev[1], lev[2], aux, *xyz))
    o = orc.Oracle(zs, pitc=joins, fast=fast)
    o.first_it.value = FastC.FIRST_IT
    o.fill_ghost(0)            # warmup's first _fillGhostcells (S/PyTree.py:536)
    return o


def interior(pi, a, e=None):
    nd = int(pi[V.NDIMDX]); ni, nj, nk = int(pi[V.NIJK]), int(pi[V.NIJK + 1]), int(pi[V.NIJK + 2])
    nf, nfk = int(pi[V.NIJK + 3]), int(pi[V.NIJK + 4])
    neq = a.size // nd
    b = a.reshape(neq, nd)[:, :ni * nj * nk].reshape(neq, nk, nj, ni)
    return b[:, nfk:nk - nfk, nf:nj - nf, nf:ni - nf]


def relerr(a, b):
    """max over variables of max|a-b| / max|b| (the 1e-10 criterion of BASELINE.json)."""
    out = 0.
    for e in range(a.shape[0]):
        sc = max(numpy.abs(b[e]).max(), 1e-300)
        out = max(out, numpy.abs(a[e] - b[e]).max() / sc)
    return out
