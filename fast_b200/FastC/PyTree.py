"""Host-side mirror of ``FastC.PyTree`` for the hot path (reference: Fast/FastC/FastC/PyTree.py).

Same function names, argument meaning and tree/array layout as the reference for the entries the
residual-and-update path needs: the ``numb``/``numz`` dictionaries are written under
``.Solver#define`` (PyTree.py:109-140), frozen into the flat ``Parameter_int`` / ``Parameter_real``
tables (:920-2047), the BC windows appended (:2788-3046) and the fields compacted into one
contiguous SoA block per time level (:2236-2274).  Everything outside the hot path (ALE, LBM, IBM,
chimera, unstructured) is rejected loudly rather than silently ignored.
"""
import math
import numpy

try:  # real Cassiopee when present (never in this container)
    import Converter.Internal as Internal  # pragma: no cover
except Exception:  # noqa
    from .. import Internal
from .. import VSHARE

MX_SYNCHRO = 1000
MX_SSZONE = 10
MX_OMP_SIZE_INT = -1
OMP_MODE = 0
FIRST_IT = 0
HOOK = None

# PyTree.py:30-70
tagBCDict = {
    "BCExtrapolate": 0, "BCDegenerateLine": 0, "BCFarfield": 1, "BCInflowSupersonic": 2,
    "BCWallInviscid": 3, "BCSymmetryPlane": 3, "BCWall": 4, "FamilySpecified": 5, "BCWallViscous": 6,
    "BCWallViscous_isot_fich": 7, "BC_fich": 8, "Nearmatch": 9, "BCOutflow": 10, "BCautoperiod": 11,
    "BCWallViscous_transition": 12, "BCInflow": 13, "BCExtrapolateRANS": 14, "BCPeriodic": 15,
    "BCOutpres": 16, "BCInj1": 17, "BCRacinf": 18, "BCInflowLund": 19, "BCInjMFR": 20, "BCOutMFR": 21,
    "BCOverlap": 22, "BCWallModel": 30, "BCWallExchange": 31, "BCWallViscousIsothermal": 32,
}
# BC families with a CUDA ghost-fill kernel
BC_IN_SCOPE = (0, 1, 3, 4, 6, 11, 22)

keys4Base = {'temporal_scheme': ['explicit', 'implicit'], 'ss_iteration': 0, 'rk': 0, 'modulo_verif': 0,
             'omp_mode': 0, 'restart_fields': 0, 'exp_local': 0, 'time_begin_ale': 1}
keys4Zone = {'scheme': ['ausmpred', 'senseur', 'roe', 'roe_min', 'roe_kap', 'roe_nul'],
             'slope': ['o1', 'o3', 'minmod'], 'time_step': 1, 'time_step_nature': ['global', 'local'],
             'cfl': 1, 'nb_relax': 0, 'lu_match': 0, 'implicit_solver': ['lussor'], 'epsi_newton': 1,
             'psiroe': 1, 'ausmDamping': 1, 'senseurType': 0, 'ransmodel': ['SA'], 'ratiom': 1,
             'prandtltb': 1, 'cache_blocking_I': 0, 'cache_blocking_J': 0, 'cache_blocking_K': 0,
             'shiftvar': 0, 'io_thread': 0, 'model': 2, 'motion': ['none'], 'extract_res': 0}


def tagBC(bcname):
    tag = tagBCDict.get(bcname, -1)
    if tag == -1: print("Warning: Fast: unknown BC type %s." % bcname)
    return tag


def checkKeys(d, keys):
    """PyTree.py:2449 -- unknown or out-of-scope keys are an error here (no silent fallback)."""
    for c in d[2]:
        k = c[0]
        if k not in keys:
            raise ValueError("fast_b200: key '%s' is outside the hot-path scope (known: %s)" % (k, sorted(keys)))
        allowed = keys[k]
        if isinstance(allowed, list):
            v = Internal.getValue(c)
            if v not in allowed:
                raise ValueError("fast_b200: value '%s' of key '%s' is outside the hot-path scope %s" % (v, k, allowed))


# PyTree.py:109-140
def _setNum2Zones(a, num):
    """Set numeric data dictionary in zones."""
    for z in Internal.getZones(a):
        cont = Internal.createUniqueChild(z, '.Solver#define', 'UserDefinedData_t')
        for k in num: Internal.createUniqueChild(cont, k, 'DataArray_t', num[k])
    return None


def setNum2Zones(a, num):
    ap = Internal.copyRef(a); _setNum2Zones(ap, num); return ap


def _setNum2Base(a, num):
    """Set numeric data dictionary in bases."""
    cont = Internal.createUniqueChild(a, '.Solver#define', 'UserDefinedData_t')
    for k in num: Internal.createUniqueChild(cont, k, 'DataArray_t', num[k])
    return None


def setNum2Base(a, num):
    ap = Internal.copyRef(a); _setNum2Base(ap, num); return ap


_STATE_NAMES = ['Density', 'MomentumX', 'MomentumY', 'MomentumZ', 'EnergyStagnationDensity', 'Pressure',
                'Temperature', 'Cv', 'Mach', 'Reynolds', 'Cs_', 'Gamma', 'TurbulentEnergyKineticDensity',
                'TurbulentDissipationRateDensity', 'TurbulentSANuTildeDensity', 'Mus', 'Cs', 'Ts', 'Pr']


def getState(ref):
    """Same 19-slot list as Cassiopee's ``C.getState`` (used at PyTree.py:1959-1977)."""
    out = []
    for n in _STATE_NAMES:
        c = Internal.getNodeFromName1(ref, n)
        out.append(float(Internal.getValue(c)) if c is not None and c[1] is not None else 0.)
    return out


def _get(d, key, default):
    if d is None: return default
    a = Internal.getNodeFromName1(d, key)
    return default if a is None else Internal.getValue(a)


def _buildOwnData(t, Padding=None):
    """numb/numz -> .Solver#ownData/{Parameter_int, Parameter_real, .Solver#dtloc} (PyTree.py:920-2047)."""
    bases = Internal.getBases(t)
    zones = Internal.getZones(t)
    d = Internal.getNodeFromName1(t, '.Solver#define')
    temporal_scheme, ss_iteration, rk, modulo_verif, restart_fields, exploc, ompmode = \
        "implicit", 30, 3, 200, 1, 0, 0
    if d is not None:
        checkKeys(d, keys4Base)
        ompmode = min(max(_get(d, 'omp_mode', 0), 0), 1)
        temporal_scheme = _get(d, 'temporal_scheme', temporal_scheme)
        ss_iteration = _get(d, 'ss_iteration', ss_iteration)
        modulo_verif = _get(d, 'modulo_verif', modulo_verif)
        restart_fields = _get(d, 'restart_fields', restart_fields)
        rk = _get(d, 'rk', rk)
        if temporal_scheme == "implicit": rk = 3
        exploc = _get(d, 'exp_local', 0)
        if exploc != 0: raise ValueError("fast_b200: exp_local (local explicit time stepping) is out of scope")
    if temporal_scheme == "explicit": nssiter = {1: 1, 2: 2, 3: 3}[int(rk)]
    elif temporal_scheme == "implicit": nssiter = int(ss_iteration) + 1
    else: raise ValueError("fast_b200: temporal_scheme %s is out of scope" % temporal_scheme)
    o = Internal.createUniqueChild(t, '.Solver#ownData', 'UserDefinedData_t')
    # .Solver#dtloc header (PyTree.py:1188-1200); the OpenMP task table that follows in the reference
    # has no meaning on the GPU and is left empty.
    dt = numpy.zeros(13, Internal.E_NpyInt)
    dt[0] = nssiter; dt[1] = int(modulo_verif); dt[2] = restart_fields - 1; dt[6] = rk
    dt[8] = ompmode; dt[9] = 1; dt[11] = 13
    Internal.createUniqueChild(o, '.Solver#dtloc', 'DataArray_t', dt)

    ndom = 0
    for b in bases:
        for z in Internal.getNodesFromType1(b, 'Zone_t'):
            a = Internal.getNodeFromName2(z, 'GoverningEquations')
            if a is None: a = Internal.getNodeFromName2(b, 'GoverningEquations')
            model = Internal.getValue(a) if a is not None else 'Euler'
            ref = Internal.getNodeFromName1(z, 'ReferenceState')
            if ref is None: ref = Internal.getNodeFromName1(b, 'ReferenceState')
            adim = getState(ref) if ref is not None else None
            prandtltb = adim[18] if adim is not None else 1.
            dz = Internal.getNodeFromName1(z, '.Solver#define')
            if dz is not None: checkKeys(dz, keys4Zone)
            scheme = _get(dz, 'scheme', 'ausmpred'); slope = _get(dz, 'slope', 'o3')
            dtnature = _get(dz, 'time_step_nature', 'global'); dtc = _get(dz, 'time_step', -0.000001)
            cfl = _get(dz, 'cfl', 1.); epsi_newton = _get(dz, 'epsi_newton', 0.1)
            psiroe = _get(dz, 'psiroe', 0.1); ausmdamping = _get(dz, 'ausmDamping', 1.)
            nbr_relax = _get(dz, 'nb_relax', 1); lu_match = _get(dz, 'lu_match', 0)
            senseurtype = _get(dz, 'senseurType', 1); ratiom = _get(dz, 'ratiom', 10000.)
            prandtltb = _get(dz, 'prandtltb', prandtltb)
            cb = [_get(dz, 'cache_blocking_I', 2048), _get(dz, 'cache_blocking_J', 2), _get(dz, 'cache_blocking_K', 7)]
            shiftvar = _get(dz, 'shiftvar', 0); io_th = _get(dz, 'io_thread', 0)
            extract_res = _get(dz, 'extract_res', 0)
            if int(nbr_relax) < 1: raise ValueError("fast_b200: nb_relax must be >= 1")
            iflow = {'Euler': 1, 'euler': 1, 'NSLaminar': 2, 'nslaminar': 2, 'NSTurbulent': 3, 'nsspalart': 3}.get(model)
            if iflow is None: raise ValueError("fast_b200: model %s is out of scope" % model)
            itypcp = 2 if temporal_scheme == "explicit" else 1
            islope = {'o1': 1, 'o3': 2, 'minmod': 3}[slope]
            kfludom = {'ausmpred': 1, 'senseur': 2, 'roe': 5, 'roe_kap': 5, 'roe_min': 5, 'roe_nul': 5}[scheme]
            if scheme == 'roe_kap': islope = 2
            elif scheme == 'roe_min': islope = 3
            elif scheme == 'roe_nul': islope = 1
            dtloc = {'global': 0, 'local': 1}[dtnature]

            pi = numpy.empty(VSHARE.PARAM_INT_SIZE, Internal.E_NpyInt)
            pi[:] = -1000; pi[0:25] = -1
            pi[25] = 0; pi[26] = 0; pi[27] = iflow; pi[28] = 0; pi[29] = itypcp
            pi[30] = 100; pi[31] = 10; pi[32] = 10
            pi[33] = kfludom; pi[34] = 0; pi[35] = 0; pi[36:45] = -1
            pi[45] = io_th; pi[46] = dtloc
            pi[47] = 0; pi[48] = 1; pi[49] = 2; pi[50] = 0; pi[51] = 0
            pi[52] = rk; pi[53] = 1; pi[54] = 0; pi[55] = 0; pi[56] = 1; pi[57] = 0; pi[58] = nssiter
            pi[59:62] = cb; pi[62] = 0; pi[63] = 1; pi[64] = islope; pi[65] = 10; pi[66] = shiftvar
            pi[67] = extract_res; pi[68] = 0; pi[69] = 0; pi[70] = 0
            pi[71] = nbr_relax; pi[72] = 0; pi[73] = 0; pi[74] = 0; pi[75] = lu_match
            pi[76:83] = 0; pi[83] = 0; pi[84] = 0; pi[85] = senseurtype; pi[86] = -1; pi[87] = 1; pi[88] = 0
            pi[89:129] = 0
            pi[VSHARE.IBC_PT_FLUX] = -1
            pi[VSHARE.SA_LOW_RE] = 0; pi[VSHARE.SA_ROT_CORR] = 0
            pi[VSHARE.SA_DIST] = 1 if iflow == 3 else 0
            pi[VSHARE.NONZ] = ndom
            Internal.createUniqueChild(Internal.createUniqueChild(z, '.Solver#ownData', 'UserDefinedData_t'),
                                       'Parameter_int', 'DataArray_t', pi)
            oz = Internal.getNodeFromName1(z, '.Solver#ownData')

            pr = numpy.zeros(VSHARE.PARAM_REAL_SIZE, numpy.float64)
            if dtc < 0:
                if dtloc == 0:
                    print('Warning: Fast: time_step_nature global requires a time_step value. time_step set to default value (0.0001).')
                dtc = 0.0001
            pr[0] = dtc
            if adim is not None:
                pr[1] = adim[11]; pr[2] = adim[7]; pr[3] = adim[0]; pr[4] = adim[5]
                pr[5] = math.sqrt(adim[1] ** 2 + adim[2] ** 2 + adim[3] ** 2) / adim[0]
                pr[6] = adim[6]; pr[7] = adim[8]; pr[8] = adim[14]; pr[9] = adim[9]; pr[10] = adim[18]
                pr[11] = adim[15]; pr[12] = adim[17]; pr[13] = adim[16]
                pr[19] = adim[1] / adim[0]; pr[20] = adim[2] / adim[0]; pr[21] = adim[3] / adim[0]
            pr[14] = epsi_newton; pr[15] = cfl; pr[16] = 0.01; pr[17] = 0.2; pr[18] = ratiom; pr[22] = 0.
            pr[34] = psiroe; pr[35] = 0.; pr[36] = 0.; pr[37] = 1.e-2; pr[38] = prandtltb; pr[39] = 1.e-2
            pr[72] = ausmdamping
            Internal.createUniqueChild(oz, 'Parameter_real', 'DataArray_t', pr)
            Internal.createUniqueChild(oz, 'CFL_minmaxmoy', 'DataArray_t', numpy.zeros(3))
            Internal.createUniqueChild(oz, 'type_zone', 'DataArray_t', 0)
            Internal.createUniqueChild(oz, 'model', 'DataArray_t', model)
            Internal.createUniqueChild(oz, 'temporal_scheme', 'DataArray_t', temporal_scheme)
            Internal.createUniqueChild(oz, 'exp_local', 'DataArray_t', 0)
            Internal.createUniqueChild(oz, 'rk', 'DataArray_t', int(rk))
            ndom += 1
    return None


PRIM = ['Density', 'VelocityX', 'VelocityY', 'VelocityZ', 'Temperature']
CONS = ['Density', 'MomentumX', 'MomentumY', 'MomentumZ', 'EnergyStagnationDensity']


def _createPrimVars(t, omp_mode=0, rmConsVars=True):
    """Conservative -> primitive + the three time levels (PyTree.py:222-420, _createVarsFast).
    Returns FIRST_IT (0 when the tree held no ``_M1`` fields => BDF1 on the first step, :497-516)."""
    first_it = 1
    for z in Internal.getZones(t):
        sol = Internal.getNodeFromName1(z, Internal.__FlowSolutionCenters__)
        if sol is None: raise ValueError("fast_b200: zone %s has no FlowSolution#Centers" % z[0])
        own = Internal.getNodeFromName1(z, '.Solver#ownData')
        pi = Internal.getNodeFromName1(own, 'Parameter_int')[1]
        sa = pi[VSHARE.IFLOW] == 3
        pr = Internal.getNodeFromName1(own, 'Parameter_real')[1]
        if Internal.getNodeFromName1(sol, 'VelocityX') is None:
            ro = Internal.getNodeFromName1(sol, 'Density')[1]
            mom = [Internal.getNodeFromName1(sol, n)[1] for n in CONS[1:4]]
            roe = Internal.getNodeFromName1(sol, CONS[4])[1]
            vel = [m / ro for m in mom]
            T = (roe / ro - 0.5 * (vel[0] ** 2 + vel[1] ** 2 + vel[2] ** 2)) / pr[VSHARE.CVINF]
            for n, v in zip(PRIM[1:], vel + [T]): Internal.createUniqueChild(sol, n, 'DataArray_t', numpy.asfortranarray(v))
            if sa:
                rn = Internal.getNodeFromName1(sol, 'TurbulentSANuTildeDensity')[1]
                Internal.createUniqueChild(sol, 'TurbulentSANuTilde', 'DataArray_t', numpy.asfortranarray(rn / ro))
            if rmConsVars:
                for n in CONS[1:] + ['TurbulentSANuTildeDensity']: Internal._rmNodesByName(sol, n)
        names = PRIM + (['TurbulentSANuTilde'] if sa else [])
        for n in names:
            base = Internal.getNodeFromName1(sol, n)
            for suf in ('_M1', '_P1'):
                if Internal.getNodeFromName1(sol, n + suf) is None:
                    if suf == '_M1': first_it = 0
                    Internal.createChild(sol, n + suf, 'DataArray_t', numpy.array(base[1], order='F', copy=True))
        shape = Internal.getNodeFromName1(sol, 'Density')[1].shape
        if pi[VSHARE.IFLOW] >= 2 and Internal.getNodeFromName1(sol, 'ViscosityEddy') is None:
            Internal.createChild(sol, 'ViscosityEddy', 'DataArray_t', numpy.zeros(shape, order='F'))
        if sa and Internal.getNodeFromName1(sol, 'TurbulentDistance') is None:
            raise ValueError("fast_b200: NSTurbulent zone %s needs centers:TurbulentDistance" % z[0])
    return first_it


def _compact(t, containers=None, fields=None, mode=None, init=True, ific=2, dtloc=None):
    """Re-lay every time level as one contiguous variable-major block rop(NDIMDX,NEQ) (PyTree.py:2236-2274):
    after this ``Density`` is the base address of the block and ``VelocityX = Density + NDIMDX``."""
    for z in Internal.getZones(t):
        sol = Internal.getNodeFromName1(z, Internal.__FlowSolutionCenters__)
        own = Internal.getNodeFromName1(z, '.Solver#ownData')
        pi = Internal.getNodeFromName1(own, 'Parameter_int')[1]
        sa = pi[VSHARE.IFLOW] == 3
        names = PRIM + (['TurbulentSANuTilde'] if sa else [])
        shape = Internal.getNodeFromName1(sol, 'Density')[1].shape
        ncell = int(numpy.prod(shape)); ndimdx = ncell + int(pi[VSHARE.SHIFTVAR])
        for suf in ('', '_M1', '_P1'):
            blk = numpy.zeros(ndimdx * len(names), numpy.float64)
            for e, n in enumerate(names):
                node = Internal.getNodeFromName1(sol, n + suf)
                view = blk[e * ndimdx:e * ndimdx + ncell].reshape(shape, order='F')
                view[...] = node[1]
                node[1] = view
        aux = []
        if pi[VSHARE.IFLOW] >= 2: aux.append('ViscosityEddy')
        if sa: aux.append('TurbulentDistance')
        if aux:
            blk = numpy.zeros(ndimdx * len(aux), numpy.float64)
            for e, n in enumerate(aux):
                node = Internal.getNodeFromName1(sol, n)
                view = blk[e * ndimdx:e * ndimdx + ncell].reshape(shape, order='F')
                view[...] = node[1]
                node[1] = view
    return None


def block_of(z, suffix=''):
    """The contiguous rop(NDIMDX*NEQ) block behind ``Density<suffix>`` (what computePT.cpp:228-345 harvests)."""
    sol = Internal.getNodeFromName1(z, Internal.__FlowSolutionCenters__)
    a = Internal.getNodeFromName1(sol, 'Density' + suffix)[1]
    base = a
    while base.base is not None: base = base.base
    return base


def aux_block_of(z):
    sol = Internal.getNodeFromName1(z, Internal.__FlowSolutionCenters__)
    n = Internal.getNodeFromName1(sol, 'ViscosityEddy')
    if n is None: return None
    base = n[1]
    while base.base is not None: base = base.base
    return base


# FastC/getRange.cpp:57-125
def getRange(in_bc, pi):
    f = [0] * 6
    nf, nfk = int(pi[VSHARE.NIJK + 3]), int(pi[VSHARE.NIJK + 4])
    ijkv = [int(pi[VSHARE.IJKV + i]) for i in range(3)]
    fic = [nf, nf, nfk]
    for d in range(3):
        if in_bc[2 * d] == in_bc[2 * d + 1]:
            v = in_bc[2 * d] - 1 if in_bc[2 * d] > 1 else in_bc[2 * d]
            v = max(1, v - 2 * fic[d])
            f[2 * d] = v; f[2 * d + 1] = v
            for o in range(3):
                if o == d: continue
                f[2 * o] = max(in_bc[2 * o] - fic[o], 1)
                f[2 * o + 1] = min(in_bc[2 * o + 1] - fic[o] - 1, ijkv[o])
            break
    return f


def getDir(ijkv, f):
    for d in range(3):
        if f[2 * d] == f[2 * d + 1] and f[2 * d] == 1: return 2 * d + 1
        if f[2 * d] == f[2 * d + 1] and f[2 * d] == ijkv[d]: return 2 * d + 2
    return 0


def _BCcompact(t):
    """Append the BC table to Parameter_int / Parameter_real (PyTree.py:2788-3046):
    ``[nb_bc, ptr_int x nb_bc, ptr_real x nb_bc, {type, idir, imin..kmax, nbdata, sizes..} ...]``."""
    for z in Internal.getZones(t):
        own = Internal.getNodeFromName1(z, '.Solver#ownData')
        npi = Internal.getNodeFromName1(own, 'Parameter_int'); npr = Internal.getNodeFromName1(own, 'Parameter_real')
        pi = npi[1]; pr = npr[1]
        bcs = Internal.getNodesFromType2(z, 'BC_t')
        nb = len(bcs)
        c_int = pi.size; c_real = pr.size
        size_real = 0
        for bc in bcs:
            for dta in Internal.getNodesFromType3(bc, 'DataArray_t'): size_real += dta[1].size
        size_int = 1 + 2 * nb + 9 * nb + sum(len(Internal.getNodesFromType3(bc, 'DataArray_t')) for bc in bcs)
        pin = numpy.zeros(c_int + size_int, Internal.E_NpyInt); pin[:c_int] = pi
        prn = numpy.zeros(c_real + size_real, numpy.float64); prn[:c_real] = pr
        pt = c_int
        pin[VSHARE.PT_BC] = pt; pin[pt] = nb
        sint = 1 + 2 * nb; sreal = 0
        ijkv = [int(pin[VSHARE.IJKV + i]) for i in range(3)]
        for i, bc in enumerate(bcs, start=1):
            pin[pt + i] = sint + pt; pin[pt + i + nb] = sreal + c_real
            pt_bc = pin[pt + i]
            btype = Internal.getValue(bc)
            tag = tagBC(btype)
            if tag not in BC_IN_SCOPE:
                raise ValueError("fast_b200: BC %s (type %d) of zone %s is outside the hot-path scope" % (btype, tag, z[0]))
            pin[pt_bc] = tag
            rng = Internal.getNodeFromType1(bc, 'IndexRange_t')[1]
            in_bc = [int(rng[0][0]), int(rng[0][1]), int(rng[1][0]), int(rng[1][1]), int(rng[2][0]), int(rng[2][1])]
            fen = getRange(in_bc, pin)
            pin[pt_bc + 2:pt_bc + 8] = fen
            pin[pt_bc + 1] = getDir(ijkv, fen)
            datas = Internal.getNodesFromType3(bc, 'DataArray_t')
            pin[pt_bc + 8] = len(datas)
            ctot = 0
            for j, dta in enumerate(datas, start=1):
                c = dta[1].size; ctot += c
                pin[pt_bc + 8 + j] = ctot
                deb = c_real + sreal
                prn[deb:deb + c] = dta[1].ravel(order='F')
                dta[1] = prn[deb:deb + c]
                sreal += c
            sint += 9 + len(datas)
        npi[1] = pin; npr[1] = prn
    return None


# PyTree.py:2298-2321
def switchPointers__(zones, case, order=3):
    if order == 2: return None
    for z in zones:
        sol = Internal.getNodeFromName1(z, Internal.__FlowSolutionCenters__)
        own = Internal.getNodeFromName1(z, '.Solver#ownData')
        model = Internal.getValue(Internal.getNodeFromName1(own, 'model'))
        names = list(PRIM)
        if model in ('nsspalart', 'NSTurbulent'): names.append('TurbulentSANuTilde')
        for v in names:
            ca = Internal.getNodeFromName1(sol, v); caM1 = Internal.getNodeFromName1(sol, v + '_M1')
            caP1 = Internal.getNodeFromName1(sol, v + '_P1')
            if case == 1:
                ta = caM1[1]; caM1[1] = ca[1]; ca[1] = caP1[1]; caP1[1] = ta
            elif case == 2:
                ta = caP1[1]; caP1[1] = ca[1]; ca[1] = caM1[1]; caM1[1] = ta
    return None


def createWorkArrays__(zones, dtloc, FIRST_IT):
    """HOOK dictionary (PyTree.py:2064-2230): the work arrays `rhs` (drodm), `coe` and `wiggle` are host numpys of the
    reference's sizes, all zones concatenated with the per-zone offsets of gsdr3.cpp:282-290 (neq*NDIMDX, neq_coe*NDIMDX,
    3*NDIMDX per zone).  On the GPU the live copies stay in HBM, owned by the solver object; FastS.sync_work(t) fills these
    host arrays on demand (they are never read back by the solver).  numpy.zeros does not commit the pages before they are
    touched, so unused work arrays cost nothing."""
    neq_max = 0; ndimt = 0; ndimcoe = 0; ndimwig = 0
    offsets = []
    for z in zones:
        pi = Internal.getNodeFromName1(Internal.getNodeFromName1(z, '.Solver#ownData'), 'Parameter_int')[1]
        neq_max = max(neq_max, int(pi[VSHARE.NEQ]))
        offsets.append((ndimt, ndimcoe, ndimwig))
        ndimt += int(pi[VSHARE.NEQ]) * int(pi[VSHARE.NDIMDX])
        ndimcoe += max(int(pi[VSHARE.NEQ_COE]), 0) * int(pi[VSHARE.NDIMDX])
        if pi[VSHARE.KFLUDOM] == 2: ndimwig += 3 * int(pi[VSHARE.NDIMDX])
    return {'rhs': numpy.zeros(ndimt, numpy.float64), 'coe': numpy.zeros(ndimcoe, numpy.float64),
            'wiggle': numpy.zeros(ndimwig, numpy.float64), 'work_offsets': offsets,
            'dtloc': dtloc, 'FIRST_IT': FIRST_IT, 'neq_max': neq_max,
            'MX_SSZONE': MX_SSZONE, 'MX_SYNCHRO': MX_SYNCHRO, 'MX_OMP_SIZE_INT': MX_OMP_SIZE_INT, 'param_int_tc': None,
            'param_real_tc': None, 'mpi': 0, 'lssiter_loc': 0, 'solver': None, 'timing': False, 'timed_subit': 0}
