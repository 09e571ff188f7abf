"""GPU parity against the CPU oracle: implicit chain (coe, BDF residual, LU-SGS wavefront, Newton update), SA,
wall / farfield BCs, match joins."""
import numpy
import pytest
from helpers import oracle_from_tree, interior, relerr, FastS, FastC, V
from fast_b200 import synth
from fast_b200.FastC.PyTree import Internal

TOL = 1e-10


def _box(**kw):
    return lambda: (synth.case_euler_box(**kw), None)


CASES = {
    'euler_cart_roe_bdf': _box(n=12, scheme='roe', temporal='implicit'),
    'lamin_full_roe_bdf': _box(n=12, scheme='roe', model='NSLaminar', ReInf=1000., deform_amp=0.02, temporal='implicit'),
    'lamin_cart_ausm_dtloc': _box(n=12, scheme='ausmpred', model='NSLaminar', ReInf=1000., temporal='implicit',
                                  extra_numz={'time_step_nature': 'local', 'cfl': 4.}),
    'lamin_full_senseur_minmod': _box(n=10, scheme='senseur', model='NSLaminar', ReInf=500., deform_amp=0.02, temporal='implicit'),
    'sa_plate_homo': lambda: synth.case_flatplate(n=(24, 16, 6)),
    'sa_plate_full': lambda: synth.case_flatplate(n=(20, 12, 6), sheared=True),
    'lamin_roe_walls_bdf': _box(n=10, scheme='roe', model='NSLaminar', ReInf=1000., deform_amp=0.02, temporal='implicit',
                               bcs={'imin': 'BCWallInviscid', 'imax': 'BCFarfield', 'jmin': 'BCWallViscous', 'jmax': 'BCFarfield',
                                    'kmin': 'BCFarfield', 'kmax': 'BCWall'}),
    'sa_plate_roe': lambda: synth.case_flatplate(n=(16, 12, 4), scheme='roe'),
    'sa_plate_blocks': lambda: synth.case_flatplate(n=(24, 16, 8), nblocks=(2, 2, 2)),
    # larger than one LU tile (16 x 8 lines, 16 x 6 for 3dfull SA) in both j and k: inter-tile mailboxes, ragged last tiles
    'sa_plate_homo_tiles': lambda: synth.case_flatplate(n=(20, 20, 20)),
    'sa_plate_full_tiles': lambda: synth.case_flatplate(n=(18, 20, 16), sheared=True),
    'lamin_full_roe_tiles': _box(n=8, n3=(14, 22, 20), scheme='roe', model='NSLaminar', ReInf=1000., deform_amp=0.02, temporal='implicit'),
    'sa_plate_medium': lambda: synth.case_flatplate(n=(64, 40, 24)),      # several rhs CTAs and LU tiles in every direction
    'multiblock_roe_lumatch': lambda: synth.case_multiblock(nblk=(2, 2, 1), n=8),
    # several zones with several LU tiles each: the multi-zone launch (k_lu_tile_b) interleaves the tiles of all zones
    'sa_plate_blocks_tiles': lambda: synth.case_flatplate(n=(24, 40, 40), nblocks=(1, 2, 2)),
    'multiblock_roe_lumatch_tiles': lambda: synth.case_multiblock(nblk=(2, 1, 2), n=18),
    'sa_plate_blocks_lumatch': lambda: synth.case_flatplate(n=(24, 16, 8), nblocks=(2, 2, 2), lu_match=1),   # invlu_overlap_*_SA
    # nb_relax >= 2: LU-SSOR relaxations (invlussor*.for) around the same sweep kernels
    'sa_plate_ssor2': lambda: synth.case_flatplate(n=(24, 16, 6), nb_relax=2),
    'sa_plate_ssor2_tiles': lambda: synth.case_flatplate(n=(20, 20, 20), nb_relax=2),
    'lamin_full_roe_ssor3': _box(n=12, scheme='roe', model='NSLaminar', ReInf=1000., deform_amp=0.02, temporal='implicit',
                                 extra_numz={'nb_relax': 3}),
    'multiblock_roe_lumatch_ssor2': lambda: synth.case_multiblock(nblk=(2, 1, 1), n=10, extra_numz={'nb_relax': 2}),
    # the same joins in the reference's wire format: Connector's param_int_tc / param_real_tc through fb_set_joins_tc
    'sa_plate_blocks_tc': lambda: synth.case_flatplate(n=(24, 16, 8), nblocks=(2, 2, 2), tc_format=True),
    'multiblock_roe_lumatch_tc': lambda: synth.case_multiblock(nblk=(2, 2, 1), n=8, tc_format=True),
    'multiblock_explicit': lambda: synth.case_multiblock(nblk=(2, 1, 2), n=8, temporal='explicit', scheme='senseur'),
    # the 5-equation kernels on a 3dhomo zone (shared metric tile of k_rhs3, k_lu_cell<1,5,*>): the plate without the SA equation
    'lamin_plate_homo': lambda: synth.case_flatplate(n=(24, 16, 6), model='NSLaminar', ReInf=2.e4),
    'lamin_plate_homo_tiles': lambda: synth.case_flatplate(n=(40, 20, 20), model='NSLaminar', ReInf=2.e4),
    'lamin_plate_homo_rk3': lambda: synth.case_flatplate(n=(24, 16, 6), model='NSLaminar', ReInf=2.e4, temporal='explicit', scheme='senseur'),
}


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(CASES))
def test_implicit_parity(name):
    t1, j1 = CASES[name](); t2, j2 = CASES[name]()
    o = oracle_from_tree(t1, joins=j1)
    (t2, tc, metrics) = FastS.warmup(t2, j2)
    nss = int(o.zones[0].pi[V.NSSITER])
    for it in range(1, 4):
        o.compute(it, 1, nss, 1)
        FastS._compute(t2, metrics, it)
    FastS.sync_tree(t2)
    crit = FastS.display_temporal_criteria(t2, metrics, 3)
    for zo, z in zip(o.zones, Internal.getZones(t2)):
        for lev, suf in enumerate(('', '_M1')):
            a = interior(zo.pi, FastC.block_of(z, suf)); b = interior(zo.pi, zo.lev[lev])
            assert numpy.isfinite(a).all()
            assert relerr(a, b) <= TOL, (name, z[0], suf, relerr(a, b))
        # CFL statistics of the check iteration (cpcfl3.for via cptst3.for:113-120; max / min exact up to the last bit of the
        # fused arithmetic, the mean is an atomic sum in arbitrary order)
        cfl = crit[z[0]]['CFL']
        assert cfl[0] > 0. and numpy.allclose(cfl[:2], zo.cfl[:2], rtol=1e-10, atol=0.), (name, cfl, zo.cfl)
        assert numpy.allclose(cfl[2], zo.cfl[2], rtol=1e-8, atol=0.), (name, cfl, zo.cfl)
        own = Internal.getNodeFromName1(z, '.Solver#ownData')
        assert numpy.array_equal(Internal.getNodeFromName1(own, 'CFL_minmaxmoy')[1], [cfl[1], cfl[0], cfl[2]])
        assert crit[z[0]]['nonfinite'] == 0
        if int(zo.pi[V.ITYPCP]) <= 1:
            r = zo.rdm.reshape((nss, zo.neq, 2), order='F')
            ncell = float(zo.pi[V.IJKV]) * float(zo.pi[V.IJKV + 1]) * float(zo.pi[V.IJKV + 2])
            l2 = numpy.sqrt(r[:, :, 0] / ncell)
            assert numpy.allclose(crit[z[0]]['L2'], l2, rtol=1e-8, atol=1e-300)
            assert numpy.allclose(crit[z[0]]['Loo'], r[:, :, 1], rtol=1e-8, atol=1e-300)


@pytest.mark.gpu
def test_device_isfinite_check_raises():
    """stopAtNan (S/PyTree.py:2070-2073): a NaN planted in the state is seen by the device-side checks of the next check
    iteration and display_temporal_criteria raises."""
    t, j = synth.case_flatplate(n=(16, 12, 4))
    (t, tc, metrics) = FastS.warmup(t, j)
    FastS._compute(t, metrics, 1)
    assert FastS.display_temporal_criteria(t, metrics, 1)[Internal.getZones(t)[0][0]]['nonfinite'] == 0
    z = Internal.getZones(t)[0]
    FastS.sync_tree(t)
    blk = FastC.block_of(z, '')
    pi = Internal.getNodeFromName1(Internal.getNodeFromName1(z, '.Solver#ownData'), 'Parameter_int')[1]
    interior(pi, blk)[4, 1, 5, 7] = numpy.nan
    from fast_b200 import _lib
    import ctypes
    _lib.check(_lib.lib().fb_zone_upload_state(FastC.HOOK['solver'], 0, 0, blk.ctypes.data_as(_lib.c_dp)))
    FastS._compute(t, metrics, 2)
    with pytest.raises(FloatingPointError): FastS.display_temporal_criteria(t, metrics, 2)


@pytest.mark.gpu
def test_reference_join_tables_match_private_ones_and_reject_out_of_scope():
    """fb_set_joins_tc: the reference layout gives bit-identical results to the library's own table, HOOK holds the two arrays
    like S/PyTree.py:495-500, and content outside the match-join scope is refused (never silently ignored)."""
    from fast_b200 import _lib
    res = []
    for fmt in (False, True):
        t, j = synth.case_multiblock(nblk=(2, 1, 2), n=8, tc_format=fmt)
        (t, tc, metrics) = FastS.warmup(t, j)
        if fmt: assert FastC.HOOK['param_int_tc'].dtype == numpy.int32 and FastC.HOOK['param_real_tc'].dtype == numpy.float64
        for it in (1, 2): FastS._compute(t, metrics, it)
        FastS.sync_tree(t)
        res.append([FastC.block_of(z, '').copy() for z in Internal.getZones(t)])
    for a, b in zip(*res): assert numpy.array_equal(a, b)
    L = _lib.lib(); s = FastC.HOOK['solver']
    pit, prt = [x.copy() for x in j]
    def rc(pi_, pr_): return L.fb_set_joins_tc(s, pi_.ctypes.data_as(_lib.c_ip), pi_.size, pr_.ctypes.data_as(_lib.c_dp), pr_.size)
    ech = int(pit[1 + int(pit[2]) + int(pit[3 + int(pit[2])]) + 3]); nrac = int(pit[ech + 1]); T = ech + 4
    bad = pit.copy(); bad[T + 3 * nrac] = 3                      # an IBC join
    assert rc(bad, prt) == -1 and b'IBC' in L.fb_last_error()
    bad = pit.copy(); bad[T + 14 * nrac] = 1                     # azimuthal periodicity
    assert rc(bad, prt) == -1 and b'rotation' in L.fb_last_error()
    bad = pit.copy(); bad[int(pit[T + 7 * nrac]) + 2] = 2        # a chimera stencil (interpolation type 2)
    assert rc(bad, prt) == -1 and b'match joins' in L.fb_last_error()
    badr = prt.copy(); badr[5] = 0.5
    assert rc(pit, badr) == -1 and b'coefficient' in L.fb_last_error()
    bad = pit.copy(); bad[ech + 2] = 1                            # unsteady joins
    assert rc(bad, prt) == -1
    assert rc(pit, prt) == 0
