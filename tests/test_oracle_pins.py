"""CPU tests of the oracle against the pins available without the reference's (unshipped) fixtures: in-source
constants, analytic invariants and the committed regression fixtures in tests/golden/."""
import ctypes, os
import numpy
import pytest
from helpers import oracle_from_tree, interior, relerr, orc, V, ROOT
from fast_b200 import synth
from fast_b200.FastC.PyTree import Internal

_dp = ctypes.POINTER(ctypes.c_double); _ip = ctypes.POINTER(ctypes.c_int)


def _uniform(t, vals=(1., 1., 0.3, -0.2, 1.)):
    for z in Internal.getZones(t):
        sol = Internal.getNodeFromName1(z, 'FlowSolution#Centers')
        for n, v in zip(('Density', 'VelocityX', 'VelocityY', 'VelocityZ', 'Temperature'), vals):
            Internal.getNodeFromName1(sol, n)[1][...] = v


@pytest.mark.parametrize('kw', [dict(scheme='ausmpred'), dict(scheme='roe', deform_amp=0.03),
                                dict(scheme='senseur', model='NSLaminar', ReInf=500., deform_amp=0.03),
                                dict(scheme='roe', model='NSLaminar', ReInf=500., temporal='implicit', deform_amp=0.02)])
def test_free_stream_preserved(kw):
    """Uniform flow on a closed (periodic) mesh => drodm == 0 to round-off, Cartesian and curvilinear."""
    t = synth.case_euler_box(n=10, **kw); _uniform(t)
    o = oracle_from_tree(t); z = o.zones[0]
    nss = int(z.pi[V.NSSITER])
    for it in (1, 2): o.compute(it, 1, nss, 1)
    a = interior(z.pi, z.lev[0])
    for e, v in enumerate((1., 1., 0.3, -0.2, 1.)):
        assert numpy.abs(a[e] - v).max() < 5e-13


@pytest.mark.parametrize('wall', ['BCWallInviscid', 'BCWallViscous'])
def test_roe_wall_flux_correction_keeps_fluid_at_rest(wall):
    """correct_flux.for (Roe + walls): in a closed box of walls a fluid at rest must stay at rest -- the corrected wall
    flux is p*S and the pressure forces of a closed cell cancel.  Without the correction the Roe dissipation acting on
    the mirrored ghost cells is also zero here, so the pin checks signs / addressing of the two corrections: a wrong sign
    leaves +-2 p*S in the wall cells."""
    faces = ('imin', 'imax', 'jmin', 'jmax', 'kmin', 'kmax')
    t = synth.case_euler_box(n=8, scheme='roe', deform_amp=0.03, model='NSLaminar', ReInf=500., bcs={f: wall for f in faces})
    _uniform(t, (1., 0., 0., 0., 1.))
    o = oracle_from_tree(t); z = o.zones[0]
    for it in (1, 2, 3): o.compute(it, 1, 3, 1)
    a = interior(z.pi, z.lev[0])
    for e, v in enumerate((1., 0., 0., 0., 1.)):
        assert numpy.abs(a[e] - v).max() < 5e-13


def test_rk2_is_heun_and_lands_in_density():
    """RK2 (the time levels must NOT rotate after an even number of explicit stages)."""
    from helpers import heun_pin
    def step(t):
        o = oracle_from_tree(t); z = o.zones[0]
        o.compute(1, 1, int(z.pi[V.NSSITER]), 1)
        return z.pi, z.pr, interior(z.pi, z.lev[0]).copy()
    err, dist_from_stage1 = heun_pin(step)
    assert err <= 1e-12, err
    assert dist_from_stage1 > 1e3 * err          # ... and it is not the forward-Euler intermediate


@pytest.mark.parametrize('case', ['tgv', 'plate_blocks', 'multiblock', 'lamin_plate_homo', 'plate_full', 'plate_blocks_lumatch', 'plate_ssor'])
def test_threaded_oracle_is_bit_identical(case):
    """The thread decompositions of the oracle -- k slabs for the rhs; for the LU one zone per thread (full-size GPU parity tests)
    or teams of threads per zone running a (j slab, i chunk) wavefront (timed CPU baseline) -- must not change a single bit."""
    def mk():
        if case == 'tgv': return synth.case_tgv(n=20, perturb=1e-3), None
        if case == 'plate_blocks': return synth.case_flatplate(n=(48, 32, 32), nblocks=(2, 1, 1))
        if case == 'lamin_plate_homo': return synth.case_flatplate(n=(40, 20, 20), model='NSLaminar', ReInf=2.e4)   # 5 equations on 3dhomo metrics
        if case == 'plate_full': return synth.case_flatplate(n=(70, 20, 12), sheared=True)          # 3dfull SA, several i chunks
        if case == 'plate_blocks_lumatch': return synth.case_flatplate(n=(24, 16, 8), nblocks=(2, 2, 2), lu_match=1)
        if case == 'plate_ssor': return synth.case_flatplate(n=(24, 16, 6), nb_relax=2)              # relaxations stay serial
        return synth.case_multiblock(nblk=(2, 1, 2), n=10)
    res = []
    for nth, team in ((1, 0), (8, 0), (8, 1), (5, 1)):
        t, j = mk()
        o = oracle_from_tree(t, joins=j)
        nss = int(o.zones[0].pi[V.NSSITER])
        o.L.fo_set_lu_team(team)       # 1: the threads left over after one per zone sweep the zones in teams (fo_invlu_team)
        try:
            for it in (1, 2): o.compute(it, 1, nss, 1, nthreads=nth)
        finally:
            o.L.fo_set_lu_team(0)
        res.append([z.lev[0].copy() for z in o.zones])
    for r in res[1:]:
        for a, b in zip(res[0], r): assert numpy.array_equal(a, b)


def test_conservation_periodic():
    """Telescoping fluxes: total mass, momentum and energy constant to round-off in a periodic Cartesian box (Euler:
    the viscous cross terms read the edge ghost cells, which correct_coins replaces by averages -- reference behaviour)."""
    t = synth.case_euler_box(n=12, scheme='senseur')
    o = oracle_from_tree(t); z = o.zones[0]
    def totals():
        a = interior(z.pi, z.lev[0]); cv = z.pr[V.CVINF]
        ro = a[0]; e = ro * (cv * a[4] + 0.5 * (a[1] ** 2 + a[2] ** 2 + a[3] ** 2))
        return numpy.array([ro.sum(), (ro * a[1]).sum(), (ro * a[2]).sum(), (ro * a[3]).sum(), e.sum()])
    t0 = totals()
    for it in range(1, 6): o.compute(it, 1, 3, 1)
    t1 = totals()
    assert numpy.abs(t1 - t0).max() <= 1e-12 * numpy.abs(t0).max()


def test_metric_pins():
    """vol of a Cartesian cell = hx*hy*hz, |ti| = hy*hz; closed curvilinear cells: sum of face vectors = 0."""
    t = synth.case_euler_box(n=8, n3=(8, 10, 12))
    o = oracle_from_tree(t); z = o.zones[0]
    hx, hy, hz = 1. / 8, 1. / 10, 1. / 12
    assert abs(z.vol[0] - hx * hy * hz) < 1e-17 and abs(abs(z.ti[0]) - hy * hz) < 1e-16
    t = synth.case_euler_box(n=8, deform_amp=0.04)
    o = oracle_from_tree(t); z = o.zones[0]; pi = z.pi
    assert int(pi[V.ITYPZONE]) == 0
    nm = int(pi[V.NDIMDX_MTR]); nx, ny, nz = [int(pi[V.NIJK_XYZ + d]) for d in range(3)]
    for e in range(3):
        a3 = [q[e * nm:(e + 1) * nm].reshape((nx, ny, nz), order='F') for q in (z.ti, z.tj, z.tk)]
        s = (a3[0][1:, :-1, :-1] - a3[0][:-1, :-1, :-1] + a3[1][:-1, 1:, :-1] - a3[1][:-1, :-1, :-1]
             + a3[2][:-1, :-1, 1:] - a3[2][:-1, :-1, :-1])
        assert numpy.abs(s).max() < 1e-16
    assert z.vol[:nm].reshape((nx, ny, nz), order='F')[:-1, :-1, :-1].min() > 0


def test_sensor_zero_on_monotone_data():
    """wiggle2001: no double sign change => wig == 0 everywhere; a saw-tooth sets it to 1."""
    t = synth.case_euler_box(n=8, scheme='senseur')
    o = oracle_from_tree(t); z = o.zones[0]; L = o.L; pi = z.pi
    nd = int(pi[V.NDIMDX]); ni, nj, nk = int(pi[V.NIJK]), int(pi[V.NIJK + 1]), int(pi[V.NIJK + 2])
    ii = numpy.arange(nd) % ni
    rop = numpy.zeros(5 * nd)
    for e in range(5): rop[e * nd:(e + 1) * nd] = 1. + 0.01 * (e + 1) * ii          # monotone in i, constant in j,k
    ind = (ctypes.c_int * 6)(1, int(pi[V.IJKV]), 1, int(pi[V.IJKV + 1]), 1, int(pi[V.IJKV + 2]))
    wig = numpy.full(3 * nd, -1.)
    L.fo_wiggle2001(pi.ctypes.data_as(_ip), z.pr.ctypes.data_as(_dp), ind, rop.ctypes.data_as(_dp), wig.ctypes.data_as(_dp))
    assert (wig[wig >= 0] == 0).all()
    rop[:nd] = 1. + 0.01 * (ii % 2)                                                  # saw-tooth density
    L.fo_wiggle2001(pi.ctypes.data_as(_ip), z.pr.ctypes.data_as(_dp), ind, rop.ctypes.data_as(_dp), wig.ctypes.data_as(_dp))
    assert interior(pi, wig[:nd].copy())[0].min() == 1.


def test_lu_sgs_reduces_to_diagonal():
    """invlu with zero metrics (A = 0) and zero spectral radii => dW = drodm / coe5 (SURVEY 8c pin)."""
    t = synth.case_euler_box(n=6, scheme='roe', temporal='implicit', deform_amp=0.02)
    o = oracle_from_tree(t); z = o.zones[0]; L = o.L; pi = z.pi
    nd = int(pi[V.NDIMDX]); rng = numpy.random.default_rng(0)
    drodm = rng.standard_normal(5 * nd); ref = drodm.copy()
    coe = numpy.zeros(5 * nd); coe[:nd] = 0.01; coe[4 * nd:] = 1.5 + rng.random(nd)
    zero = numpy.zeros_like(z.ti); zk = numpy.zeros_like(z.tk)
    box = (ctypes.c_int * 6)(1, int(pi[V.IJKV]), 1, int(pi[V.IJKV + 1]), 1, int(pi[V.IJKV + 2]))
    out = z.lev[0].copy()
    L.fo_invlu(pi.ctypes.data_as(_ip), z.pr.ctypes.data_as(_dp), box, z.lev[0].ctypes.data_as(_dp), out.ctypes.data_as(_dp),
               drodm.ctypes.data_as(_dp), coe.ctypes.data_as(_dp), zero.ctypes.data_as(_dp), zero.ctypes.data_as(_dp), zk.ctypes.data_as(_dp))
    a = interior(pi, drodm); b = interior(pi, ref) / interior(pi, coe[4 * nd:].copy())[0][None]
    assert numpy.abs(a - b).max() < 1e-14
    # Newton update of density is rho + dW1
    assert numpy.abs(interior(pi, out)[0] - (interior(pi, z.lev[0])[0] + a[0])).max() < 1e-14


@pytest.mark.parametrize('case', ['lamin_full', 'sa_homo'])
def test_lussor_relaxations(case):
    """nb_relax >= 2 (LU/invlussor*.for) as symmetric Gauss-Seidel on (D + L + U) dW = rhs, pinned three ways: the first
    relaxation IS LU-SGS (invlu.for nb_relax = 1 path, to round-off); every further relaxation lowers the linear residual
    max |rhs - (D + L + U) dW| (operators of the sweeps, fo_lu_residual); and many relaxations reach the exact solve."""
    if case == 'lamin_full': t = synth.case_euler_box(n=8, scheme='roe', model='NSLaminar', ReInf=1000., temporal='implicit', deform_amp=0.02)
    else: t, _ = synth.case_flatplate(n=(12, 10, 6))
    o = oracle_from_tree(t); z = o.zones[0]; L = o.L
    nss = int(z.pi[V.NSSITER])
    o.compute(1, 1, 1, 1)                      # one sub-iteration: coe, rhs and metrics of a real Newton step are in place
    L.fo_lu_residual.restype = ctypes.c_double
    nd = int(z.pi[V.NDIMDX]); neq = z.neq
    box = (ctypes.c_int * 6)(1, int(z.pi[V.IJKV]), 1, int(z.pi[V.IJKV + 1]), 1, int(z.pi[V.IJKV + 2]))
    rng = numpy.random.default_rng(3)
    rhs = 1e-3 * rng.standard_normal(neq * nd)
    state = z.lev[0].copy()

    def solve(nrelax):
        pi = z.pi.copy(); pi[V.NB_RELAX] = nrelax
        dw = rhs.copy(); out = state.copy()
        L.fo_invlu(pi.ctypes.data_as(_ip), z.pr.ctypes.data_as(_dp), box, state.ctypes.data_as(_dp), out.ctypes.data_as(_dp),
                   dw.ctypes.data_as(_dp), z.coe.ctypes.data_as(_dp), z.ti.ctypes.data_as(_dp), z.tj.ctypes.data_as(_dp), z.tk.ctypes.data_as(_dp))
        res = L.fo_lu_residual(pi.ctypes.data_as(_ip), z.pr.ctypes.data_as(_dp), box, state.ctypes.data_as(_dp), rhs.ctypes.data_as(_dp),
                               dw.ctypes.data_as(_dp), z.coe.ctypes.data_as(_dp), z.ti.ctypes.data_as(_dp), z.tj.ctypes.data_as(_dp), z.tk.ctypes.data_as(_dp))
        return dw, out, res
    dw1, out1, r1 = solve(1)
    dw2, out2, r2 = solve(2); dw4, out4, r4 = solve(4); dw30, out30, r30 = solve(30)
    scale = numpy.abs(rhs).max()
    assert r2 < r1 and r4 < r2 and r30 < 1e-9 * scale, (r1, r2, r4, r30)
    assert numpy.isfinite(out4).all()
    # Newton update of the last relaxation: rho_new = rho + dW1 on the interior
    a = interior(z.pi, out4)[0]; b = interior(z.pi, state)[0] + interior(z.pi, dw4)[0]
    assert numpy.abs(a - b).max() < 1e-14


def test_lussor_first_relaxation_is_lu_sgs():
    """One relaxation of the nb_relax >= 2 sweeps ((D+U) dW = rhs + L-sum) reproduces invlu_l / invlu_u ((D+U) dW = D dW*),
    Newton update included, on an interior box and with the state aliased to the output (plane-by-plane update)."""
    t = synth.case_euler_box(n=8, scheme='roe', model='NSLaminar', ReInf=1000., temporal='implicit', deform_amp=0.02)
    o = oracle_from_tree(t); z = o.zones[0]; L = o.L
    o.compute(1, 1, 1, 1)
    nd = int(z.pi[V.NDIMDX]); neq = z.neq
    box = (ctypes.c_int * 6)(1, int(z.pi[V.IJKV]), 1, int(z.pi[V.IJKV + 1]), 1, int(z.pi[V.IJKV + 2]))
    rhs = 1e-3 * numpy.random.default_rng(4).standard_normal(neq * nd)
    for alias in (False, True):
        outs = []
        for relax in (0, 1):
            dw = rhs.copy(); st = z.lev[0].copy(); out = st if alias else st.copy()
            args = [z.pi.ctypes.data_as(_ip), z.pr.ctypes.data_as(_dp), box, st.ctypes.data_as(_dp), out.ctypes.data_as(_dp), dw.ctypes.data_as(_dp),
                    z.coe.ctypes.data_as(_dp), z.ti.ctypes.data_as(_dp), z.tj.ctypes.data_as(_dp), z.tk.ctypes.data_as(_dp)]
            if relax: L.fo_invlu_relax(*args, ctypes.c_int(1))
            else: L.fo_invlu(*args)
            outs.append((interior(z.pi, dw).copy(), interior(z.pi, out).copy()))
        assert relerr(outs[1][0], outs[0][0]) < 1e-12 and relerr(outs[1][1], outs[0][1]) < 1e-13


def test_newton_converges_and_bdf_first_iteration():
    """Implicit dual time stepping: the Newton residual drops monotonically over the sub-iterations."""
    t = synth.case_euler_box(n=10, scheme='roe', temporal='implicit')
    o = oracle_from_tree(t); z = o.zones[0]
    nss = int(z.pi[V.NSSITER])
    o.compute(1, 1, nss, 1)
    r = z.rdm.reshape((nss, z.neq, 2), order='F')[:, 0, 0]
    assert (numpy.diff(r) < 0).all() and r[-1] < 1e-3 * r[0]


def test_openmp_slabs_equal_serial():
    """The timed CPU baseline (k slabs / zone-parallel LU) gives the serial result (same per-cell accumulation order)."""
    ts = [synth.case_euler_box(n=10, scheme='senseur', model='NSLaminar', ReInf=500., deform_amp=0.02) for _ in range(2)]
    o1 = oracle_from_tree(ts[0]); o2 = oracle_from_tree(ts[1])
    for it in (1, 2):
        o1.compute(it, 1, 3, 1, nthreads=1); o2.compute(it, 1, 3, 1, nthreads=3)
    assert numpy.array_equal(interior(o1.zones[0].pi, o1.zones[0].lev[0]), interior(o2.zones[0].pi, o2.zones[0].lev[0]))


GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'oracle_regression.npz')


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="fixtures not generated")
def test_golden_regression():
    """The committed fixtures (tests/golden/make_golden.py) are reproduced bit-for-bit by the strict-FP oracle build."""
    import sys; sys.path.insert(0, os.path.join(ROOT, "tests", "golden")); import make_golden
    g = numpy.load(GOLDEN)
    for name in make_golden.CASES:
        a = make_golden.run(name)
        assert numpy.allclose(a, g[name], rtol=1e-13, atol=1e-15), name
