"""GPU parity (through the C-ABI / FastS surface) against the CPU oracle: explicit RK chain."""
import numpy
import pytest
from helpers import oracle_from_tree, interior, relerr, FastS, FastC, V
from fast_b200 import synth, fasts
from fast_b200.FastC.PyTree import Internal

TOL = 1e-10   # BASELINE.json: conserved variables <= 1e-10 relative per step

CASES = {
    'euler_cart_ausm_o3': dict(n=16, scheme='ausmpred'),
    'euler_full_ausm_o3': dict(n=16, scheme='ausmpred', deform_amp=0.02),
    'lamin_cart_senseur': dict(n=16, scheme='senseur', model='NSLaminar', ReInf=2000.),
    'lamin_full_senseur': dict(n=16, scheme='senseur', model='NSLaminar', ReInf=2000., deform_amp=0.02),
    'lamin_full_roe_minmod': dict(n=12, scheme='roe', slope='minmod', model='NSLaminar', ReInf=2000., deform_amp=0.02),
    'euler_cart_roe_o1_rk2': dict(n=12, scheme='roe', slope='o1', rk=2),
    'euler_ragged_ausm': dict(n=8, n3=(19, 7, 10), scheme='ausmpred'),
}


def run_pair(kw, nit=3):
    t1 = synth.case_euler_box(**kw); t2 = synth.case_euler_box(**kw)
    o = oracle_from_tree(t1)
    (t2, tc, metrics) = FastS.warmup(t2, None)
    nss = int(o.zones[0].pi[V.NSSITER])
    for it in range(1, nit + 1):
        o.compute(it, 1, nss, 1)
        FastS._compute(t2, metrics, it)
    FastS.sync_tree(t2)
    return o, t2, metrics


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(CASES))
def test_explicit_parity(name):
    o, t2, metrics = run_pair(CASES[name])
    for zo, z in zip(o.zones, Internal.getZones(t2)):
        for lev, suf in enumerate(('', '_M1')):
            a = interior(zo.pi, FastC.block_of(z, suf)); b = interior(zo.pi, zo.lev[lev])
            assert numpy.isfinite(a).all()
            assert relerr(a, b) <= TOL, (name, suf, relerr(a, b))
        # ghost cells of the current level too (BC fill + corners)
        a = FastC.block_of(z, ''); b = zo.lev[0]
        sc = numpy.abs(b).max()
        assert numpy.abs(a - b).max() / sc <= TOL


@pytest.mark.gpu
@pytest.mark.parametrize('amp', [0., 0.03])
def test_metric_parity(amp):
    t1 = synth.case_euler_box(n=10, deform_amp=amp); t2 = synth.case_euler_box(n=10, deform_amp=amp)
    o = oracle_from_tree(t1)
    (t2, tc, metrics) = FastS.warmup(t2, None)
    ti, tj, tk, vol = fasts.get_metric(Internal.getZones(t2), metrics, FastC.HOOK)[0]
    zo = o.zones[0]
    for a, b in ((ti, zo.ti), (tj, zo.tj), (tk, zo.tk), (vol, zo.vol)):
        assert numpy.abs(a - b[:a.size]).max() <= 1e-14 * max(numpy.abs(b).max(), 1e-300)
    # closed cells: sum of outward face vectors vanishes (analytic pin)
    if amp > 0:
        pi = zo.pi; nm = int(pi[V.NDIMDX_MTR])
        nx, ny, nz = int(pi[V.NIJK_XYZ]), int(pi[V.NIJK_XYZ + 1]), int(pi[V.NIJK_XYZ + 2])
        for e in range(3):
            a3 = [t[e * nm:(e + 1) * nm].reshape((nx, ny, nz), order='F') for t in (ti, tj, tk)]
            s = (a3[0][1:, :-1, :-1] - a3[0][:-1, :-1, :-1] + a3[1][:-1, 1:, :-1] - a3[1][:-1, :-1, :-1]
                 + a3[2][:-1, :-1, 1:] - a3[2][:-1, :-1, :-1])
            assert numpy.abs(s).max() < 1e-15
