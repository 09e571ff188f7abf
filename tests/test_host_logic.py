"""CPU tests of the host-side mirror of the reference's Python surface (numb/numz -> Parameter_int/real, BC
compaction, mesh-family detection, pointer rotation) and of the C-ABI library's exports."""
import ctypes, os, re
import numpy
import pytest
from helpers import FastS, FastC, V, ROOT
from fast_b200 import synth, geom, _lib
from fast_b200.FastC.PyTree import Internal


def _pi_pr(t, n=0):
    z = Internal.getZones(t)[n]
    own = Internal.getNodeFromName1(z, '.Solver#ownData')
    return Internal.getNodeFromName1(own, 'Parameter_int')[1], Internal.getNodeFromName1(own, 'Parameter_real')[1]


def test_param_tables_explicit():
    t = synth.case_euler_box(n=8, scheme='senseur', model='NSLaminar', ReInf=1600.)
    FastS._warmup_host(t)
    pi, pr = _pi_pr(t)
    assert pi[V.NIJK:V.NIJK + 5].tolist() == [12, 12, 12, 2, 2] and pi[V.IJKV:V.IJKV + 3].tolist() == [8, 8, 8]
    assert pi[V.NDIMDX] == 12 ** 3 and pi[V.NDIMDX_XYZ] == 13 ** 3
    assert pi[V.ITYPZONE] == 2 and pi[V.NDIMDX_MTR] == 1 and pi[V.NEQ_IJ] == 1 and pi[V.NEQ_K] == 1
    assert pi[V.IFLOW] == 2 and pi[V.KFLUDOM] == 2 and pi[V.SLOPE] == 2 and pi[V.ITYPCP] == 2
    assert pi[V.NEQ] == 5 and pi[V.NEQ_COE] == 1 and pi[V.NSSITER] == 3 and pi[V.RK] == 3
    assert pr[V.GAMMA] == 1.4 and abs(pr[V.PRANDT] - 0.72) < 1e-15 and pr[V.WIG_AMP] == 1. and pr[V.PSIROE] == 0.1
    assert abs(pr[V.VINF] - 1.) < 1e-15 and abs(pr[V.XMUL0] - 1. / 1600.) < 1e-18
    dt = Internal.getNodeFromName1(Internal.getNodeFromName1(t, '.Solver#ownData'), '.Solver#dtloc')[1]
    assert dt[0] == 3 and dt[6] == 3
    assert FastC.FIRST_IT == 0     # no _M1 fields in the input tree => BDF1 / first iteration


def test_param_tables_implicit_sa_and_bcs():
    t, j = synth.case_flatplate(n=(12, 8, 4))
    FastS._warmup_host(t)
    pi, pr = _pi_pr(t)
    assert pi[V.ITYPZONE] == 1 and pi[V.NEQ_IJ] == 2 and pi[V.NEQ_K] == 1 and pi[V.NIJK_MTR:V.NIJK_MTR + 5].tolist() == [1, 17, 0, 2, 0]
    assert pi[V.IFLOW] == 3 and pi[V.NEQ] == 6 and pi[V.NEQ_COE] == 6 and pi[V.ITYPCP] == 1 and pi[V.DTLOC] == 1
    assert pi[V.NSSITER] == 6 and pr[V.CFL] == 7.
    pt = pi[V.PT_BC]; nb = pi[pt]
    assert nb == 6
    recs = {}
    for b in range(nb):
        pb = pi[pt + 1 + b]
        recs[int(pi[pb + V.BC_IDIR])] = (int(pi[pb + V.BC_TYPE]), pi[pb + 2:pb + 8].tolist())
    assert recs[3] == (6, [1, 12, 1, 1, 1, 4])        # jmin viscous wall, window in interior cell numbering
    assert recs[1] == (1, [1, 1, 1, 8, 1, 4]) and recs[2] == (1, [12, 12, 1, 8, 1, 4]) and recs[4][0] == 1
    assert recs[5][0] == 11 and recs[6] == (11, [1, 12, 1, 8, 4, 4])
    assert pi[V.NB_RELAX] == 1 and pi[V.LU_MATCH] == 0
    # nb_relax / lu_match travel in the reference's slots (S/param_solver.h: NB_RELAX 71, LU_MATCH 75)
    t, j = synth.case_flatplate(n=(12, 8, 4), nb_relax=3, lu_match=1)
    FastS._warmup_host(t)
    pi, pr = _pi_pr(t)
    assert pi[V.NB_RELAX] == 3 and pi[V.LU_MATCH] == 1 and V.NB_RELAX == 71 and V.LU_MATCH == 75


def test_mesh_family_detection():
    for amp, kw, want in ((0., {}, 2), (0.02, {}, 0)):
        t = synth.case_euler_box(n=6, deform_amp=amp)
        FastS._warmup_host(t)
        assert _pi_pr(t)[0][V.ITYPZONE] == want
    t, _ = synth.case_flatplate(n=(8, 6, 4), sheared=True); FastS._warmup_host(t)
    assert _pi_pr(t)[0][V.ITYPZONE] == 0
    # stretched in x,y but constant k planes => 3dhomo; degenerate (collapsed) cells are flagged
    t, _ = synth.case_flatplate(n=(8, 6, 4))
    z = Internal.getZones(t)[0]; gc = Internal.getNodeFromName1(z, 'GridCoordinates')
    x, y, zz = [Internal.getNodeFromName1(gc, n)[1] for n in ('CoordinateX', 'CoordinateY', 'CoordinateZ')]
    assert geom.nature_geom_dom(x, y, zz, 2, 2)[0] == 1
    x2 = x.copy(); x2[1, :, :] = x2[0, :, :]
    y2 = y.copy(); y2[1, :, :] = y2[0, :, :]
    z2 = zz.copy(); z2[1, :, :] = z2[0, :, :]
    assert geom.degenerate_cells(x2, y2, z2)[0].all() and not geom.degenerate_cells(x, y, zz).any()


def test_switch_pointers_and_compact_layout():
    t = synth.case_euler_box(n=6)
    FastS._warmup_host(t)
    z = Internal.getZones(t)[0]; pi, _ = _pi_pr(t)
    blk = FastC.block_of(z, '')
    sol = Internal.getNodeFromName1(z, 'FlowSolution#Centers')
    # VelocityX = Density + NDIMDX (C/PyTree.py:2262-2274)
    d = Internal.getNodeFromName1(sol, 'Density')[1]; u = Internal.getNodeFromName1(sol, 'VelocityX')[1]
    assert u.ctypes.data - d.ctypes.data == 8 * int(pi[V.NDIMDX]) and blk.size == 5 * int(pi[V.NDIMDX])
    ids = [id(FastC.block_of(z, s)) for s in ('', '_M1', '_P1')]
    FastC.switchPointers__([z], 1, 3)
    assert [id(FastC.block_of(z, s)) for s in ('', '_M1', '_P1')] == [ids[2], ids[0], ids[1]]


def test_out_of_scope_options_fail_loudly():
    t = synth.case_euler_box(n=6)
    FastC._setNum2Zones(t, {'scheme': 'ausmpred', 'slope': 'o5'})
    with pytest.raises(ValueError): FastS._warmup_host(t)
    t = synth.case_euler_box(n=6); FastC._setNum2Zones(t, {'scheme': 'ausmpred', 'nb_relax': 0, 'time_step': 1e-3})
    with pytest.raises(ValueError): FastS._warmup_host(t)
    t = synth.case_euler_box(n=6); FastC._setNum2Base(t, {'temporal_scheme': 'explicit_local'})
    with pytest.raises(ValueError): FastS._warmup_host(t)


def test_degenerate_cells_are_refused():
    """nature_geom_dom.for flags collapsed cells; the reference repairs their metrics (tijk_extrap.for), this library refuses
    them instead of computing on them silently."""
    t = synth.case_euler_box(n=6)
    z = Internal.getZones(t)[0]
    gc = Internal.getNodeFromName1(z, 'GridCoordinates')
    for n in ('CoordinateX', 'CoordinateY', 'CoordinateZ'):
        a = Internal.getNodeFromName1(gc, n)[1]
        a[0, :, :] = a[1, :, :]                       # collapse the outer ghost layer in i
    with pytest.raises(ValueError, match='degenerate'): FastS._warmup_host(t)


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'fastb200.h')).read()
    names = sorted(set(re.findall(r'\b(fb_[a-z_0-9]+)\s*\(', hdr)))
    assert len(names) >= 30
    L = _lib.lib()
    for n in names: assert hasattr(L, n), n
    assert set(_lib.SYMBOLS) == set(names)


def test_overlap_tile_order_is_a_face_first_bijection():
    """Halo-exchange overlap (csrc/fb_tiles.cuh): part 1 of an explicit stage must cover exactly the tiles on the flagged faces,
    part 2 the rest, every tile once -- for every face mask, including degenerate tile grids."""
    import ctypes, itertools
    L = _lib.lib()
    for gx, gy, gz in itertools.product((1, 2, 3, 5), (1, 2, 4), (1, 2, 3)):
        n = gx * gy * gz
        g = (gx, gy, gz)
        for faces in range(64):
            for thin in (0, 1, 2, 4, 7):
                out = (ctypes.c_int * (3 * n))(); nf = ctypes.c_int(0)
                assert L.fb_debug_tile_order(gx, gy, gz, faces, thin, out, ctypes.byref(nf)) == 0
                tiles = [(out[3 * c], out[3 * c + 1], out[3 * c + 2]) for c in range(n)]
                assert len(set(tiles)) == n and all(0 <= x < gx and 0 <= y < gy and 0 <= z < gz for x, y, z in tiles)
                # layers of part 1 behind every face: 1, or 2 on an upper face whose last tile is a single cell thick (ADVICE r1:
                # the second donor layer then lives in the penultimate tile), never more than the direction holds
                lo = [(faces >> (2 * d)) & 1 for d in range(3)]
                hi = [min(((faces >> (2 * d + 1)) & 1) * (2 if (thin >> d) & 1 else 1), g[d] - lo[d]) for d in range(3)]

                def on_face(t):
                    return any(t[d] < lo[d] or t[d] >= g[d] - hi[d] for d in range(3))
                assert all(on_face(t) for t in tiles[:nf.value]) and not any(on_face(t) for t in tiles[nf.value:]), (gx, gy, gz, faces, thin)
    assert L.fb_debug_tile_order(0, 1, 1, 0, 0, None, None) != 0


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly (never route through the oracle)."""
    L = _lib.lib()
    if L.fb_device_count() > 0: pytest.skip("GPU present")
    s = ctypes.c_void_p()
    assert L.fb_solver_create(0, 1, ctypes.byref(s)) == -2 and b'no CPU fallback' in L.fb_last_error()
    t = synth.case_euler_box(n=6)
    with pytest.raises(_lib.FastError): FastS.warmup(t, None)
    import glob
    for f in glob.glob(os.path.join(ROOT, 'fast_b200', '**', '*.py'), recursive=True) + \
            glob.glob(os.path.join(ROOT, 'fast_b200', 'csrc', '*.cu*')):
        assert 'oracle' not in open(f).read().lower(), f + " must not reference the oracle"


def test_bench_reference_arm_contract():
    """bench.py --impl reference: one JSON line with the contract keys, the oracle timed on the requested mesh (tiny here)."""
    import json, subprocess, sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--n', '12,12,8', '--steps', '2', '--warmup', '1',
                        '--ref-budget', '0.5'], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d['impl'] == 'reference' and d['unit'] == 'Mcell.it/s' and d['value'] > 0 and d['higher_is_better'] is True
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'Mcell.it/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'configs[4]' in d['config']['workload'] and d['config']['steps_timed'] >= 2
