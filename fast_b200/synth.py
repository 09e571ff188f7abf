"""Seeded synthetic cases of BASELINE.json: trees are produced directly in the layout
FastS.warmup expects (ghost layers included, FlowSolution#Centers in conservative or primitive variables,
ZoneBC windows in ghosted node numbering, ReferenceState, GoverningEquations).  No CGNS file, no Cassiopee."""
import math
import numpy
from . import Internal

NG = 2  # ghost layers


def adim1(MInf=0.5, alphaZ=0., alphaY=0., ReInf=1.e8, MutSMuInf=0.2, Gamma=1.4, Pr=0.72):
    """Non-dimensional reference state with RoInf = UInf = TInf = 1 (the convention of Cassiopee's KCore.Adim.adim1,
    restated from its documented definition): p = 1/(gamma M^2), cv = 1/(gamma (gamma-1) M^2), mu(TInf) = 1/Re."""
    ro, u, T = 1., 1., 1.
    p = 1. / (Gamma * MInf * MInf)
    cv = 1. / (Gamma * (Gamma - 1.) * MInf * MInf)
    az, ay = math.radians(alphaZ), math.radians(alphaY)
    ux = u * math.cos(az) * math.cos(ay); uy = u * math.sin(az) * math.cos(ay); uz = u * math.sin(ay)
    roe = p / (Gamma - 1.) + 0.5 * ro * u * u
    Ts = 1.; Cs = 110.4 / 288.15
    mus = 1. / ReInf          # Sutherland reference viscosity at Ts = TInf
    nutilde = _nutilde_from_mut(MutSMuInf) * mus / ro
    return dict(Density=ro, MomentumX=ro * ux, MomentumY=ro * uy, MomentumZ=ro * uz, EnergyStagnationDensity=roe,
                Pressure=p, Temperature=T, Cv=cv, Mach=MInf, Reynolds=ReInf, Gamma=Gamma,
                TurbulentSANuTildeDensity=ro * nutilde, Mus=mus, Cs=Cs, Ts=Ts, Pr=Pr)


def _nutilde_from_mut(r):
    """chi such that chi * fv1(chi) = mut/mu (SA), Newton iterations."""
    chi = max(r, 1e-3)
    for _ in range(50):
        f = chi ** 4 / (chi ** 3 + 357.911) - r
        df = (4 * chi ** 3 * (chi ** 3 + 357.911) - chi ** 4 * 3 * chi ** 2) / (chi ** 3 + 357.911) ** 2
        chi -= f / df
    return chi


def add_state(node, state, model):
    rs = Internal.createUniqueChild(node, 'ReferenceState', 'ReferenceState_t')
    for k, v in state.items(): Internal.createUniqueChild(rs, k, 'DataArray_t', float(v))
    fe = Internal.createUniqueChild(node, 'FlowEquationSet', 'FlowEquationSet_t')
    Internal.createUniqueChild(fe, 'GoverningEquations', 'GoverningEquations_t', model)


def _axis(n, L, o, stretch=None):
    """n interior cells on [o, o+L], NG ghost cells each side (mapping simply continued)."""
    s = (numpy.arange(-NG, n + NG + 1, dtype=numpy.float64)) / n
    if stretch is not None: s = stretch(s)
    return o + L * s


def make_block(name, n, L=(1., 1., 1.), origin=(0., 0., 0.), deform=None, stretch=(None, None, None),
               index_offset=(0, 0, 0), nglob=None):
    """Structured zone of n=(ni,nj,nk) interior cells + NG ghost layers; `deform(X,Y,Z)->(X,Y,Z)` is applied to the
    analytic node positions (ghost nodes included, so periodic / matching geometry stays consistent).
    With nglob the block is the window [index_offset, index_offset+n) of a global nglob-cell mesh."""
    nglob = nglob or n
    ax = []
    for d in range(3):
        full = _axis(nglob[d], L[d], origin[d], stretch[d])
        a = index_offset[d]
        ax.append(full[a:a + n[d] + 2 * NG + 1])
    X, Y, Z = numpy.meshgrid(ax[0], ax[1], ax[2], indexing='ij')
    if deform is not None: X, Y, Z = deform(X, Y, Z)
    z = Internal.newZone(name, n[0] + 2 * NG + 1, n[1] + 2 * NG + 1, n[2] + 2 * NG + 1)
    gc = Internal.createChild(z, 'GridCoordinates', 'GridCoordinates_t')
    for nm, a in zip(('CoordinateX', 'CoordinateY', 'CoordinateZ'), (X, Y, Z)):
        Internal.createChild(gc, nm, 'DataArray_t', numpy.asfortranarray(a))
    return z


def centers(z):
    gc = Internal.getNodeFromName1(z, 'GridCoordinates')
    out = []
    for nm in ('CoordinateX', 'CoordinateY', 'CoordinateZ'):
        a = Internal.getNodeFromName1(gc, nm)[1]
        c = 0.125 * (a[:-1, :-1, :-1] + a[1:, :-1, :-1] + a[:-1, 1:, :-1] + a[1:, 1:, :-1]
                     + a[:-1, :-1, 1:] + a[1:, :-1, 1:] + a[:-1, 1:, 1:] + a[1:, 1:, 1:])
        out.append(c)
    return out


def set_fields(z, ro, u, v, w, T, nutilde=None, dist=None):
    sol = Internal.createUniqueChild(z, 'FlowSolution#Centers', 'FlowSolution_t')
    Internal.createUniqueChild(sol, 'GridLocation', 'GridLocation_t', 'CellCenter')
    for nm, a in zip(('Density', 'VelocityX', 'VelocityY', 'VelocityZ', 'Temperature'), (ro, u, v, w, T)):
        Internal.createUniqueChild(sol, nm, 'DataArray_t', numpy.asfortranarray(a, dtype=numpy.float64))
    if nutilde is not None:
        Internal.createUniqueChild(sol, 'TurbulentSANuTilde', 'DataArray_t', numpy.asfortranarray(nutilde))
    if dist is not None:
        Internal.createUniqueChild(sol, 'TurbulentDistance', 'DataArray_t', numpy.asfortranarray(dist))


def add_bc(z, name, btype, face):
    """face in imin,imax,jmin,jmax,kmin,kmax; window = whole face in ghosted node numbering (1-based)."""
    ni, nj, nk = z[1][0, 0], z[1][1, 0], z[1][2, 0]
    r = [[1, ni], [1, nj], [1, nk]]
    d = 'ijk'.index(face[0])
    n = (ni, nj, nk)[d]
    r[d] = [1, 1] if face[1:] == 'min' else [n, n]
    zbc = Internal.createUniqueChild(z, 'ZoneBC', 'ZoneBC_t')
    bc = Internal.createChild(zbc, name, 'BC_t', btype)
    Internal.createChild(bc, 'PointRange', 'IndexRange_t', numpy.array(r, dtype=Internal.E_NpyInt, order='F'))
    return bc


FACES = ('imin', 'imax', 'jmin', 'jmax', 'kmin', 'kmax')


def new_tree(zones, state, model):
    t = Internal.newCGNSTree()
    b = Internal.newCGNSBase('Base', parent=t)
    b[2] += zones
    add_state(b, state, model)
    return t


def vortex_state(xc, yc, zc, state, rng, amp=1e-3, L=1.):
    """Uniform stream + Lamb-Oseen-like vortex + seeded 1e-3 relative noise."""
    gam = state['Gamma']; cv = state['Cv']
    u0 = state['MomentumX'] / state['Density']; v0 = state['MomentumY'] / state['Density']
    w0 = state['MomentumZ'] / state['Density']
    r2 = ((xc - 0.5 * L) ** 2 + (yc - 0.5 * L) ** 2) / (0.1 * L) ** 2
    g = 0.2 * numpy.exp(0.5 * (1. - r2))
    u = u0 - g * (yc - 0.5 * L) / (0.1 * L); v = v0 + g * (xc - 0.5 * L) / (0.1 * L)
    w = w0 + 0.05 * numpy.sin(2 * math.pi * zc / L)
    T = state['Temperature'] * (1. - (gam - 1.) * g * g * 0.5 * state['Mach'] ** 2)
    ro = state['Density'] * (T / state['Temperature']) ** (1. / (gam - 1.))
    def noise(a): return a * (1. + amp * rng.standard_normal(a.shape))
    return noise(ro), noise(u) , noise(v), w + amp * rng.standard_normal(w.shape), noise(T)


def case_euler_box(n=64, scheme='ausmpred', deform_amp=0., seed=1234, model='Euler', rk=3, slope='o3',
                   MInf=0.4, ReInf=1.e8, time_step=None, bcs=None, temporal='explicit', extra_numz=None,
                   extra_numb=None, n3=None):
    """C1-style case: n^3 box, periodic (BCautoperiod) unless `bcs` maps faces to BC types."""
    from .FastC import PyTree as FastC
    n3 = n3 or (n, n, n)
    rng = numpy.random.default_rng(seed)
    state = adim1(MInf=MInf, ReInf=ReInf)
    deform = None
    if deform_amp > 0.:
        def deform(X, Y, Z):
            tp = 2 * math.pi
            return (X + deform_amp * numpy.sin(tp * Y) * numpy.sin(tp * Z),
                    Y + deform_amp * numpy.sin(tp * Z) * numpy.sin(tp * X),
                    Z + deform_amp * numpy.sin(tp * X) * numpy.sin(tp * Y))
    z = make_block('cart', n3, deform=deform)
    xc, yc, zc = centers(z)
    set_fields(z, *vortex_state(xc, yc, zc, state, rng))
    for f in FACES:
        bt = 'BCautoperiod' if bcs is None else bcs[f]
        add_bc(z, f, bt, f)
    t = new_tree([z], state, model)
    h = 1. / n
    numb = {'temporal_scheme': temporal, 'rk': rk, 'modulo_verif': 1}
    if temporal == 'implicit': numb['ss_iteration'] = 3
    if extra_numb: numb.update(extra_numb)
    numz = {'scheme': scheme, 'slope': slope, 'time_step': time_step if time_step is not None else 1e-3 * h * 100}
    if extra_numz: numz.update(extra_numz)
    FastC._setNum2Zones(t, numz); FastC._setNum2Base(t, numb)
    return t


def _tgv_fields(z, state, perturb=0., seed=7):
    x, y, zz = centers(z)
    gam = state['Gamma']; p0 = state['Pressure']; ro0 = 1.; U0 = 1.
    u = U0 * numpy.sin(x) * numpy.cos(y) * numpy.cos(zz)
    v = -U0 * numpy.cos(x) * numpy.sin(y) * numpy.cos(zz)
    w = numpy.zeros_like(u)
    p = p0 + ro0 * U0 * U0 / 16. * (numpy.cos(2 * x) + numpy.cos(2 * y)) * (numpy.cos(2 * zz) + 2.)
    ro = p / p0 * ro0
    T = p / (ro * state['Cv'] * (gam - 1.))
    if perturb > 0.:
        # the exact TGV field is symmetric: differences across its extrema are pure round-off and the SENSOR switch (a sign
        # test) becomes a coin toss there -- parity tests break the symmetry, the benchmark keeps the exact field
        rng = numpy.random.default_rng(seed)
        ro = ro * (1. + perturb * rng.standard_normal(ro.shape)); T = T * (1. + perturb * rng.standard_normal(T.shape))
        u = u + perturb * U0 * rng.standard_normal(u.shape); v = v + perturb * U0 * rng.standard_normal(v.shape)
        w = w + perturb * U0 * rng.standard_normal(w.shape)
    set_fields(z, ro, u, v, w, T)


def case_tgv(n=256, seed=7, scheme='senseur', Re=1600., MInf=0.1, time_step=0.002, perturb=0.):
    """C2: laminar Taylor-Green vortex on [0,2pi]^3, fields as in FastS/test/tgvMpi.py:54-58 (restated)."""
    from .FastC import PyTree as FastC
    state = adim1(MInf=MInf, ReInf=Re)
    L = 2 * math.pi
    z = make_block('tgv', (n, n, n), L=(L, L, L))
    _tgv_fields(z, state, perturb, seed)
    for f in FACES: add_bc(z, f, 'BCautoperiod', f)
    t = new_tree([z], state, 'NSLaminar')
    FastC._setNum2Zones(t, {'scheme': scheme, 'slope': 'o3', 'time_step': time_step * (256. / n)})
    FastC._setNum2Base(t, {'temporal_scheme': 'explicit', 'rk': 3, 'modulo_verif': 50})
    return t


def case_tgv_blocks(n=256, nblk=(2, 1, 1), only=None, scheme='senseur', Re=1600., MInf=0.1, time_step=0.002):
    """Weak-scaling family of C2: nblk blocks of n^3 cells, one Taylor-Green period per block (domain 2 pi nblk), periodic
    everywhere: BCautoperiod where a direction has one block, match joins otherwise (the hand-built donor tree of
    FastS/test/tgvMpi.py:125-195).  `only` = global zone ids to build (the others belong to other ranks).
    Returns (tree with the `only` zones, layout {(bi,bj,bk): zone id})."""
    from .FastC import PyTree as FastC
    state = adim1(MInf=MInf, ReInf=Re)
    L = 2 * math.pi
    n3 = (n, n, n) if isinstance(n, int) else tuple(n)
    ng = tuple(n3[d] * nblk[d] for d in range(3))
    layout, zones = {}, []
    for bk in range(nblk[2]):
        for bj in range(nblk[1]):
            for bi in range(nblk[0]):
                zid = len(layout); layout[(bi, bj, bk)] = zid
                if only is not None and zid not in only: continue
                z = make_block('tgv_%d_%d_%d' % (bi, bj, bk), n3, L=tuple(L * b for b in nblk),
                               index_offset=(bi * n3[0], bj * n3[1], bk * n3[2]), nglob=ng)
                _tgv_fields(z, state)
                for d, f in enumerate('ijk'):
                    if nblk[d] == 1:
                        add_bc(z, f + 'min', 'BCautoperiod', f + 'min'); add_bc(z, f + 'max', 'BCautoperiod', f + 'max')
                zones.append(z)
    t = new_tree(zones, state, 'NSLaminar')
    FastC._setNum2Zones(t, {'scheme': scheme, 'slope': 'o3', 'time_step': time_step * (256. / max(n3))})
    FastC._setNum2Base(t, {'temporal_scheme': 'explicit', 'rk': 3, 'modulo_verif': 50})
    return t, layout


def case_flatplate(n=(64, 48, 8), seed=2025, MInf=0.2, ReInf=25.e6, ss_iteration=5, cfl=7., sheared=False,
                   scheme='ausmpred', modulo_verif=1, nblocks=(1, 1, 1), only=None, return_layout=False, lu_match=0, nb_relax=1,
                   tc_format=False, model='NSTurbulent', temporal='implicit'):
    """C3 / C5-style RANS-SA flat plate: x geometric stretch, y tanh clustering toward the wall (jmin = viscous wall),
    z uniform => typezone 1 (3dhomo); `sheared=True` adds a smooth 3-D shear => typezone 0 (3dfull).
    Farfield on imin/imax/jmax, autoperiodic in k.  model='NSLaminar' / 'Euler' run the same mesh and state without the SA
    equation (the 3dhomo instantiations of the 5-equation kernels); temporal='explicit' switches to RK3 with a global time
    step.  nblocks splits the plate into match-joined blocks; `only` (a set of
    global block ids, k-major numbering) builds just those blocks -- the per-rank share of a sharded case, every block
    drawing from its own random stream -- and `return_layout` returns (tree, layout) for dist.local_case."""
    from .FastC import PyTree as FastC
    rng = numpy.random.default_rng(seed)
    state = adim1(MInf=MInf, ReInf=ReInf, MutSMuInf=15.)
    Lx, H, Lz = 2., 1., 0.25
    ax, by = 2.0, 3.2
    sx = lambda s: (numpy.exp(ax * s) - 1.) / (math.exp(ax) - 1.)
    sy = lambda s: 1. + numpy.tanh(by * (s - 1.)) / math.tanh(by)
    deform = None
    if sheared:
        def deform(X, Y, Z):
            return X + 0.02 * numpy.sin(6. * Z) * Y, Y * (1. + 0.03 * numpy.cos(5. * Z) * numpy.sin(3. * X)), Z + 0.01 * numpy.sin(4. * X) * Y
    zones, layout = [], {}
    nb = nblocks
    nloc = [n[d] // nb[d] for d in range(3)]
    for bk in range(nb[2]):
        for bj in range(nb[1]):
            for bi in range(nb[0]):
                zid = len(layout)
                if only is not None:
                    layout[(bi, bj, bk)] = zid
                    if zid not in only: continue
                    rng = numpy.random.default_rng(seed + 7919 * zid)
                name = 'plate' if nb == (1, 1, 1) else 'plate_%d_%d_%d' % (bi, bj, bk)
                z = make_block(name, tuple(nloc), L=(Lx, H, Lz), stretch=(sx, sy, None), deform=deform,
                               index_offset=(bi * nloc[0], bj * nloc[1], bk * nloc[2]), nglob=n)
                xc, yc, zc = centers(z)
                d = numpy.abs(yc) + 1e-12
                delta = 0.05
                prof = numpy.tanh(numpy.abs(yc) / delta)
                u = prof * (1. + 1e-3 * rng.standard_normal(xc.shape))
                v = 1e-3 * rng.standard_normal(xc.shape) * prof
                w = 1e-3 * rng.standard_normal(xc.shape) * prof
                ro = numpy.ones_like(xc) * (1. + 1e-4 * rng.standard_normal(xc.shape))
                T = numpy.ones_like(xc) * (1. + 0.2 * (state['Gamma'] - 1.) * MInf ** 2 * (1. - prof ** 2))
                nut = state['TurbulentSANuTildeDensity'] * (0.1 + 3. * prof * numpy.exp(-numpy.abs(yc) / (4 * delta))) \
                    * (1. + 1e-2 * rng.random(xc.shape))
                sa = model == 'NSTurbulent'
                set_fields(z, ro, u, v, w, T, nutilde=nut if sa else None, dist=d if sa else None)
                if only is None: layout[(bi, bj, bk)] = len(zones)
                zones.append(z)
                if bi == 0: add_bc(z, 'imin', 'BCFarfield', 'imin')
                if bi == nb[0] - 1: add_bc(z, 'imax', 'BCFarfield', 'imax')
                if bj == 0: add_bc(z, 'jmin', 'BCWallViscous', 'jmin')
                if bj == nb[1] - 1: add_bc(z, 'jmax', 'BCFarfield', 'jmax')
                if nb[2] == 1:
                    add_bc(z, 'kmin', 'BCautoperiod', 'kmin'); add_bc(z, 'kmax', 'BCautoperiod', 'kmax')
    t = new_tree(zones, state, model)
    if temporal == 'implicit':
        FastC._setNum2Zones(t, {'scheme': scheme, 'slope': 'o3', 'time_step_nature': 'local', 'cfl': cfl, 'nb_relax': nb_relax, 'lu_match': lu_match})
        FastC._setNum2Base(t, {'temporal_scheme': 'implicit', 'ss_iteration': ss_iteration, 'modulo_verif': modulo_verif})
    else:
        FastC._setNum2Zones(t, {'scheme': scheme, 'slope': 'o3', 'time_step': 2.e-5})
        FastC._setNum2Base(t, {'temporal_scheme': 'explicit', 'modulo_verif': modulo_verif})
    if return_layout: return t, layout
    if tc_format: return t, build_joins_tc(layout, nb, tuple(nloc), periodic=(False, False, nb[2] > 1), neq=6)   # reference wire format
    joins = build_joins(layout, nb, tuple(nloc), periodic=(False, False, nb[2] > 1))
    return t, joins


def case_multiblock(nblk=(2, 2, 2), n=16, seed=99, scheme='roe', model='NSLaminar', deform_amp=0.015, lu_match=1,
                    ss_iteration=3, ReInf=5000., MInf=0.3, time_step=None, temporal='implicit', modulo_verif=1,
                    only=None, return_layout=False, extra_numz=None, tc_format=False):
    """C4-style case: nblk blocks of n^3 cells, smooth sinusoidal deformation (3dfull), match joins inside, farfield
    outside, Roe o3, implicit LU-SGS with lu_match."""
    from .FastC import PyTree as FastC
    rng = numpy.random.default_rng(seed)
    state = adim1(MInf=MInf, ReInf=ReInf)
    ng = tuple(n * b for b in nblk)
    L = tuple(float(b) for b in nblk)
    deform = None
    if deform_amp > 0.:
        def deform(X, Y, Z):
            tp = math.pi
            return (X + deform_amp * numpy.sin(tp * Y) * numpy.sin(tp * Z), Y + deform_amp * numpy.sin(tp * Z) * numpy.sin(tp * X),
                    Z + deform_amp * numpy.sin(tp * X) * numpy.sin(tp * Y))
    zones, layout = [], {}
    for bk in range(nblk[2]):
        for bj in range(nblk[1]):
            for bi in range(nblk[0]):
                zid = len(layout)
                brng = numpy.random.default_rng(seed + 7919 * zid) if only is not None else rng   # per-block stream when sharded
                if only is not None and zid not in only:
                    layout[(bi, bj, bk)] = zid
                    continue
                z = make_block('blk_%d_%d_%d' % (bi, bj, bk), (n, n, n), L=L, deform=deform,
                               index_offset=(bi * n, bj * n, bk * n), nglob=ng)
                xc, yc, zc = centers(z)
                set_fields(z, *vortex_state(xc, yc, zc, state, brng, L=L[0]))
                for d, f in enumerate(('i', 'j', 'k')):
                    b = (bi, bj, bk)[d]
                    if b == 0: add_bc(z, f + 'min', 'BCFarfield', f + 'min')
                    if b == nblk[d] - 1: add_bc(z, f + 'max', 'BCFarfield', f + 'max')
                layout[(bi, bj, bk)] = zid
                zones.append(z)
    t = new_tree(zones, state, model)
    h = 1. / n
    numz = {'scheme': scheme, 'slope': 'o3', 'time_step': time_step if time_step is not None else 0.1 * h, 'psiroe': 0.1}
    numb = {'temporal_scheme': temporal, 'modulo_verif': modulo_verif}
    if temporal == 'implicit':
        numz['lu_match'] = lu_match; numb['ss_iteration'] = ss_iteration
    if extra_numz: numz.update(extra_numz)
    FastC._setNum2Zones(t, numz); FastC._setNum2Base(t, numb)
    if return_layout: return t, layout
    if tc_format: return t, build_joins_tc(layout, nblk, (n, n, n), neq=5)        # reference wire format
    return t, build_joins(layout, nblk, (n, n, n))


def join_records(layout, nblk, nloc, periodic=(False, False, False)):
    """Global, deterministically ordered list of match joins (donor zone, receiver zone, donor cells, receiver cells):
    for every pair of face-adjacent blocks the NG ghost layers of the receiver (interior tangential extent) are copied
    from the matching interior cells of the donor."""
    ni, nj, nk = [nloc[d] + 2 * NG for d in range(3)]

    def addr(i, j, k):   # 1-based interior numbering -> 0-based cell address
        return (i + NG - 1) + (j + NG - 1) * ni + (k + NG - 1) * ni * nj
    recs = []
    for (bi, bj, bk), rz in sorted(layout.items(), key=lambda kv: kv[1]):
        for d in range(3):
            for side in (-1, 1):
                nb = [bi, bj, bk]; nb[d] += side
                if nb[d] < 0 or nb[d] >= nblk[d]:
                    if not periodic[d]: continue
                    nb[d] %= nblk[d]
                dz = layout[tuple(nb)]
                rng = [numpy.arange(1, nloc[0] + 1), numpy.arange(1, nloc[1] + 1), numpy.arange(1, nloc[2] + 1)]
                g = numpy.arange(1 - NG, 1) if side < 0 else numpy.arange(nloc[d] + 1, nloc[d] + NG + 1)
                src = g + nloc[d] if side < 0 else g - nloc[d]
                r = list(rng); r[d] = g
                s = list(rng); s[d] = src
                R = numpy.meshgrid(*r, indexing='ij'); S = numpy.meshgrid(*s, indexing='ij')
                recs.append((dz, rz, addr(S[0], S[1], S[2]).ravel(order='F').astype(numpy.int32),
                             addr(R[0], R[1], R[2]).ravel(order='F').astype(numpy.int32)))
    return recs


def build_joins(layout, nblk, nloc, periodic=(False, False, False), rank_of=None, rank=0, neq=5):
    """Flattened match-join table (csrc/fb_joins.cu layout).
    Without rank_of: every join is local (kind 0); returns the table, or None when the case has no join.
    With rank_of (global zone id -> rank): the table of `rank`, zones renumbered locally in global-id order; a join whose
    donor and receiver live on different ranks is split into a pack entry (kind 1) on the donor's rank and an unpack
    entry (kind 2) on the receiver's rank.  The send (receive) buffer is the concatenation, peer by peer in rank order,
    of the joins to (from) that peer in global join order, each join var-major [neq][npts] FP64 -- so both sides agree
    on the offsets without talking.  Returns (table, peers) with peers = {peer: (send_off, send_n, recv_off, recv_n)}
    in doubles."""
    recs = join_records(layout, nblk, nloc, periodic)
    if rank_of is None:
        if not recs: return None
        desc, idx, off = [], [], 0
        for dz, rz, s, r in recs:
            desc += [0, dz, rz, s.size, off, off + s.size, 0]
            idx += [s, r]; off += 2 * s.size
        return numpy.concatenate([numpy.array([len(recs), off], numpy.int32), numpy.array(desc, numpy.int32)] + idx)
    mine = sorted(z for z, r in rank_of.items() if r == rank)
    lz = {z: n for n, z in enumerate(mine)}
    send_n, recv_n = {}, {}
    for dz, rz, s, r in recs:
        rd, rr = rank_of[dz], rank_of[rz]
        if rd == rr: continue
        if rd == rank: send_n[rr] = send_n.get(rr, 0) + s.size * neq
        if rr == rank: recv_n[rd] = recv_n.get(rd, 0) + s.size * neq
    peers, so, ro = {}, 0, 0
    for p in sorted(set(send_n) | set(recv_n)):
        peers[p] = (so, send_n.get(p, 0), ro, recv_n.get(p, 0))
        so += send_n.get(p, 0); ro += recv_n.get(p, 0)
    cur_s = {p: v[0] for p, v in peers.items()}; cur_r = {p: v[2] for p, v in peers.items()}
    desc, idx, off, nj = [], [], 0, 0
    for dz, rz, s, r in recs:
        rd, rr = rank_of[dz], rank_of[rz]
        if rd == rank and rr == rank:
            desc += [0, lz[dz], lz[rz], s.size, off, off + s.size, 0]; idx += [s, r]; off += 2 * s.size; nj += 1
        elif rd == rank:
            desc += [1, lz[dz], -1, s.size, off, -1, cur_s[rr]]; idx += [s]; off += s.size; cur_s[rr] += s.size * neq; nj += 1
        elif rr == rank:
            desc += [2, -1, lz[rz], s.size, -1, off, cur_r[rd]]; idx += [r]; off += r.size; cur_r[rd] += r.size * neq; nj += 1
    table = numpy.concatenate([numpy.array([nj, off], numpy.int32), numpy.array(desc, numpy.int32).reshape(-1)] + idx) \
        if nj else numpy.array([0, 0], numpy.int32)
    return table.astype(numpy.int32), peers


def build_joins_tc(layout, nblk, nloc, periodic=(False, False, False), neq=5):
    """The same match joins in the REFERENCE wire format: (param_int_tc, param_real_tc) as Connector's miseAPlatDonorTree__ lays
    them out and K_FASTC::setInterpTransfersIntra reads them (C/Com/setInterpTransfersFast.cpp:469-477, 641-654, 750-760;
    technical document 2.4.3.1): header [-, NbP2P, sizecomIBC = 0, sizecomID = 1, NbcomID = 0, Ptr_P2P_1], then the intra-process
    exchange [dest, nrac, nrac_inst = 0, timelevel = 0, 18 tables of nrac ints, 27 ints per join (local time stepping levels)],
    then per join the interpolation type block [ntype = 1, npts, type (= 1) per point], the donor and the receiver point lists.
    Coefficients (1.0 per point, ST/tgvMpi.py:136-139) go to param_real_tc.  Returns None when the case has no join."""
    recs = join_records(layout, nblk, nloc, periodic)
    if not recs: return None
    nrac = len(recs)
    head = [0, 1, 0, 1, 0]                      # pit[2] = sizecomIBC, pit[3] = sizecomID, one entry (NbcomID = 0)
    ech = len(head) + 1                          # pit[NoTransfert + shift_graph] = pit[5] -> first P2P block
    T = ech + 4
    end_tables = T + 18 * nrac + 27 * nrac
    tables = numpy.zeros((18, nrac), numpy.int64)
    blocks, pos, cpos, coefs = [], end_tables, 0, []
    for q, (dz, rz, sidx, ridx) in enumerate(recs):
        npts = int(sidx.size)
        tables[0, q] = npts; tables[1, q] = npts; tables[2, q] = 1; tables[3, q] = -1; tables[4, q] = 0
        tables[5, q] = dz; tables[9, q] = 0; tables[10, q] = npts; tables[11, q] = rz; tables[13, q] = neq
        tables[14, q] = 0; tables[15, q] = -1; tables[16, q] = 0
        tblock = numpy.concatenate([[1, npts], numpy.ones(npts, numpy.int64)])
        tables[7, q] = pos; pos += tblock.size
        tables[6, q] = pos; pos += npts
        tables[12, q] = pos; pos += npts
        tables[8, q] = cpos; cpos += npts
        blocks += [tblock, sidx.astype(numpy.int64), ridx.astype(numpy.int64)]
        coefs.append(numpy.ones(npts))
    lev = numpy.zeros((nrac, 27), numpy.int64); lev[:, 24] = 1; lev[:, 25] = 1
    pit = numpy.concatenate([numpy.array(head + [ech], numpy.int64), numpy.array([0, nrac, 0, 0], numpy.int64), tables.ravel(), lev.ravel()] + blocks)
    assert pit.size == pos
    return pit.astype(numpy.int32), numpy.concatenate(coefs)
