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


def heun_pin(step_once, kw=None):
    """Oracle-independent pin of the explicit RK2 path (core3ark3.for:70-75, init_rhs.for:70-76: stage 1 W1 = W0 + dt R(W0),
    stage 2 W = W1 + dt/2 (R(W1) - R(W0)), i.e. Heun's W0 + dt/2 (R0 + R1)) AND of where its result lives: after one
    iteration it must be `Density` itself -- FastS._compute does not rotate the time levels when the number of stages is
    even (S/PyTree.py:226-228 -> switchPointers__(zones, case, 2) is a no-op).  R0 and R1 are measured with two independent
    forward-Euler (rk=1) steps of the same code path; `step_once(tree)` advances a tree by one iteration and returns
    (Parameter_int, Parameter_real, interior primitive block of `Density`)."""
    from fast_b200 import synth
    from fast_b200.FastC.PyTree import Internal
    kw = dict(kw or dict(n=10, scheme='ausmpred'))

    def mk(rk, prim=None):
        t = synth.case_euler_box(rk=rk, **kw)
        if prim is not None:
            sol = Internal.getNodeFromName1(Internal.getZones(t)[0], 'FlowSolution#Centers')
            g = synth.NG
            for e, n in enumerate(('Density', 'VelocityX', 'VelocityY', 'VelocityZ', 'Temperature')):
                a = Internal.getNodeFromName1(sol, n)[1]
                a[g:-g, g:-g, g:-g] = numpy.transpose(prim[e], (2, 1, 0))
        return t

    def cons(q, cv):
        ro = q[0]
        return numpy.stack([ro, ro * q[1], ro * q[2], ro * q[3], ro * (cv * q[4] + 0.5 * (q[1] ** 2 + q[2] ** 2 + q[3] ** 2))])

    t0 = mk(1)
    sol = Internal.getNodeFromName1(Internal.getZones(t0)[0], 'FlowSolution#Centers')
    g = synth.NG
    w0 = numpy.stack([numpy.transpose(Internal.getNodeFromName1(sol, n)[1][g:-g, g:-g, g:-g], (2, 1, 0)).copy()
                      for n in ('Density', 'VelocityX', 'VelocityY', 'VelocityZ', 'Temperature')])
    pi, pr, w1 = step_once(t0)                   # W1 = W0 + dt R0
    _, _, w2 = step_once(mk(1, w1))              # W1 + dt R1
    _, _, wh = step_once(mk(2))                  # the RK2 iteration, read from `Density`
    cv = float(pr[V.CVINF])
    c0, c1, c2, ch = cons(w0, cv), cons(w1, cv), cons(w2, cv), cons(wh, cv)
    heun = c0 + 0.5 * ((c1 - c0) + (c2 - c1))
    return relerr(ch, heun), relerr(ch, c1)
