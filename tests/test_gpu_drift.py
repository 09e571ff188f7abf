"""Bounded drift over many iterations (BASELINE north_star: "<= 1e-10 per step, and a bounded residual-history drift after N
iterations"; SURVEY 8c tolerance statement: residual history within 1e-6 relative in L2 norm after 100 iterations) and the
BASELINE configs[0] case at its full size."""
import numpy
import pytest
from helpers import oracle_from_tree, interior, relerr, FastS, FastC, V
from fast_b200 import synth
from fast_b200.FastC.PyTree import Internal

NIT = 100


@pytest.mark.gpu
def test_c1_full_size_parity():
    """BASELINE configs[0]: Euler 64^3 Cartesian, periodic, AUSM+ o3, explicit RK3 -- three iterations against the oracle."""
    kw = dict(n=64, scheme='ausmpred')
    t1 = synth.case_euler_box(**kw); t2 = synth.case_euler_box(**kw)
    o = oracle_from_tree(t1, fast=True)
    (t2, tc, metrics) = FastS.warmup(t2, None)
    for it in range(1, 4):
        o.compute(it, 1, 3, 1000, nthreads=o.L.fo_max_threads())
        FastS._compute(t2, metrics, it)
    FastS.sync_tree(t2)
    zo = o.zones[0]; z = Internal.getZones(t2)[0]
    a = FastC.block_of(z, ''); b = zo.lev[0]
    assert numpy.isfinite(a).all()
    assert relerr(interior(zo.pi, a), interior(zo.pi, b)) <= 1e-10
    assert numpy.abs(a - b).max() / numpy.abs(b).max() <= 1e-10


@pytest.mark.gpu
def test_explicit_state_drift_100_iterations():
    """C1-style case, 100 RK3 iterations (300 stages): the state stays within 1e-8 of the oracle's (no growth beyond round-off
    accumulation; a per-step error of 1e-10 compounding would still pass, a wrong coefficient would not)."""
    kw = dict(n=16, scheme='ausmpred', deform_amp=0.02)
    t1 = synth.case_euler_box(**kw); t2 = synth.case_euler_box(**kw)
    o = oracle_from_tree(t1)
    (t2, tc, metrics) = FastS.warmup(t2, None)
    for it in range(1, NIT + 1):
        o.compute(it, 1, 3, 1000)
        FastS._compute(t2, metrics, it)
    FastS.sync_tree(t2)
    zo = o.zones[0]; z = Internal.getZones(t2)[0]
    a = interior(zo.pi, FastC.block_of(z, '')); b = interior(zo.pi, zo.lev[0])
    assert numpy.isfinite(a).all()
    assert relerr(a, b) <= 1e-8


@pytest.mark.gpu
def test_implicit_residual_history_drift_100_iterations():
    """C3-style case (RANS-SA plate, LU-SGS, 5 Newton sub-iterations, residual norms every iteration): the L2 / Loo residual
    history of every equation and sub-iteration stays within 1e-6 relative of the oracle's over 100 iterations, the state
    within 1e-8."""
    t1, j1 = synth.case_flatplate(n=(24, 16, 6), modulo_verif=1); t2, j2 = synth.case_flatplate(n=(24, 16, 6), modulo_verif=1)
    o = oracle_from_tree(t1, joins=j1)
    (t2, tc, metrics) = FastS.warmup(t2, j2)
    nss = int(o.zones[0].pi[V.NSSITER])
    zo = o.zones[0]; z = Internal.getZones(t2)[0]
    ncell = float(zo.pi[V.IJKV]) * float(zo.pi[V.IJKV + 1]) * float(zo.pi[V.IJKV + 2])
    worst = 0.
    for it in range(1, NIT + 1):
        o.compute(it, 1, nss, 1)
        FastS._compute(t2, metrics, it)
        crit = FastS.display_temporal_criteria(t2, metrics, it)
        r = zo.rdm.reshape((nss, zo.neq, 2), order='F')
        l2 = numpy.sqrt(r[:, :, 0] / ncell)
        g = numpy.asarray(crit[z[0]]['L2'])
        assert numpy.isfinite(g).all()
        scale = numpy.maximum(numpy.abs(l2), 1e-9 * numpy.abs(l2).max())   # a component sitting at round-off is measured against the largest
        worst = max(worst, float((numpy.abs(g - l2) / scale).max()))
        scale = numpy.maximum(numpy.abs(r[:, :, 1]), 1e-9 * numpy.abs(r[:, :, 1]).max())
        worst = max(worst, float((numpy.abs(numpy.asarray(crit[z[0]]['Loo']) - r[:, :, 1]) / scale).max()))
    assert worst <= 1e-6, worst
    FastS.sync_tree(t2)
    a = interior(zo.pi, FastC.block_of(z, '')); b = interior(zo.pi, zo.lev[0])
    assert relerr(a, b) <= 1e-8
