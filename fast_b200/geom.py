"""Host-side integer part of ``fasts.allocate_metric`` (reference: FastS/Metric/allocate_metric.cpp:34-481
and FastC/Metric/nature_geom_dom.for): zone sizes, ghost counts, mesh-family detection and the metric
strides written into Parameter_int.  Pure numpy -- this is control logic, the metric arithmetic itself
runs in CUDA (csrc/fb_kernels.cu: k_tijk, k_vol)."""
import numpy
from . import VSHARE as V

SOUSZERO = 2.e-8


def degenerate_cells(x, y, z):
    """nature_geom_dom.for:36-160 -- a cell is flagged when its four edges along i, or j, or k are all <= 2e-8."""
    def elen(a0, a1):
        return numpy.sqrt((x[a0] - x[a1]) ** 2 + (y[a0] - y[a1]) ** 2 + (z[a0] - z[a1]) ** 2)
    s = slice(None, -1); p = slice(1, None)
    deg = numpy.zeros(tuple(n - 1 for n in x.shape), dtype=bool)
    # i edges at (j,k),(j+1,k),(j+1,k+1),(j,k+1)
    di = [elen((s, a, b), (p, a, b)) for a, b in ((s, s), (p, s), (p, p), (s, p))]
    dj = [elen((a, s, b), (a, p, b)) for a, b in ((s, s), (p, s), (p, p), (s, p))]
    dk = [elen((a, b, s), (a, b, p)) for a, b in ((s, s), (p, s), (p, p), (s, p))]
    for grp in (di, dj, dk):
        deg |= numpy.logical_and.reduce([g <= SOUSZERO for g in grp])
    return deg


def nature_geom_dom(x, y, z, nfic, nfic_k):
    """Returns (type_zone, neq_ij, neq_k, nijk_mtr[5], ndimdx_mtr, degen) -- nature_geom_dom.for:162-460.
    x,y,z: Fortran-ordered node arrays (ni,nj,nk) including ghost nodes."""
    ni, nj, nk = x.shape
    deg = degenerate_cells(x, y, z)
    ok = ~deg
    if nk == 2:
        return 3, 2, 0, [1, ni, 0, nfic, nfic_k], ni * nj, deg
    s = slice(None, -1); p = slice(1, None)
    e = SOUSZERO

    def plane_const(a, axis, f_other):
        # reference value: node (1,1,k) etc. = first interior node of the two other directions
        idx = [f_other[0], f_other[1], f_other[2]]
        idx[axis] = slice(None)
        ref = a[tuple(idx)]
        shp = [1, 1, 1]; shp[axis] = -1
        ref = ref.reshape(shp)
        d = numpy.abs(ref - a)          # node-wise deviation from the plane reference
        lo = [s, s, s]; hi = [p, p, p]
        lo_n = list(lo); hi_n = list(hi)
        lo_n[axis] = s; hi_n[axis] = s    # nodes (.,.,k): corner (i,j) and (i+1,j+1)
        lo_p = list(lo); hi_p = list(hi)
        lo_p[axis] = p; hi_p[axis] = p    # same at k+1
        bad = (d[tuple(lo_n)] >= e) | (d[tuple(hi_n)] >= e) | (d[tuple(lo_p)] >= e) | (d[tuple(hi_p)] >= e)
        return not numpy.any(bad & ok)

    def step_const(a, axis, first):
        i0 = list(first); i1 = list(first); i1[axis] += 1
        cible = a[tuple(i1)] - a[tuple(i0)]
        lo = [s, s, s]; hi = [s, s, s]; hi[axis] = p
        return not numpy.any((numpy.abs((a[tuple(hi)] - a[tuple(lo)]) - cible) >= e) & ok)

    first = (nfic, nfic, nfic_k)
    zcart = plane_const(z, 2, first); zcte = step_const(z, 2, first)
    ycart = plane_const(y, 1, first); ycte = step_const(y, 1, first)
    xcart = plane_const(x, 0, first); xcte = step_const(x, 0, first)
    if xcart and xcte and ycart and ycte and zcart and zcte:
        return 2, 1, 1, [0, 0, 0, 0, 0], 1, deg
    if zcart and zcte:
        def kinv(a):
            ref = a[:, :, nfic_k:nfic_k + 1]
            d = numpy.abs(ref - a)
            bad = (d[s, s, s] >= e) | (d[p, p, s] >= e) | (d[s, s, p] >= e) | (d[p, p, p] >= e)
            return not numpy.any(bad & ok)
        if kinv(x) and kinv(y):
            return 1, 2, 1, [1, ni, 0, nfic, 0], ni * nj, deg
    return 0, 3, 3, [1, ni, ni * nj, nfic, nfic_k], ni * nj * nk, deg


def fill_sizes(pi, x, y, z):
    """allocate_metric.cpp:96-176: NIJK_XYZ, NIJK, IJKV, NDIMDX*, ITYPZONE, NEQ, NEQ_COE, NIJK_MTR."""
    ni, nj, nk = x.shape
    if pi[V.SLOPE] in (5, 7): raise ValueError("fast_b200: o5 slopes (3 ghost layers) are out of scope")
    nfic = 2
    pi[V.NIJK_XYZ:V.NIJK_XYZ + 5] = [ni, nj, nk, nfic, 0 if nk == 2 else nfic]
    pi[V.NIJK:V.NIJK + 5] = [ni - 1, nj - 1, nk - 1, nfic, 0 if nk == 2 else nfic]
    pi[V.IJKV] = pi[V.NIJK] - 2 * pi[V.NIJK + 3]
    pi[V.IJKV + 1] = pi[V.NIJK + 1] - 2 * pi[V.NIJK + 3]
    pi[V.IJKV + 2] = pi[V.NIJK + 2] - 2 * pi[V.NIJK + 4]
    pi[V.NDIMDX] = (ni - 1) * (nj - 1) * (nk - 1) + pi[V.SHIFTVAR]
    pi[V.NDIMDX_XYZ] = ni * nj * nk
    typ, neq_ij, neq_k, nijk_mtr, ndimdx_mtr, deg = nature_geom_dom(x, y, z, nfic, int(pi[V.NIJK_XYZ + 4]))
    if typ == 3: raise ValueError("fast_b200: 2-D zones (typezone 3) are not built yet")
    pi[V.NIJK_MTR:V.NIJK_MTR + 5] = nijk_mtr
    pi[V.NDIMDX_MTR] = ndimdx_mtr; pi[V.ITYPZONE] = typ; pi[V.NEQ_IJ] = neq_ij; pi[V.NEQ_K] = neq_k
    coe_shift = 0 if pi[V.ITYPCP] <= 1 else 1
    if pi[V.IFLOW] == 3 and pi[V.ILES] == 0: pi[V.NEQ_COE] = 6 - coe_shift * 5
    else: pi[V.NEQ_COE] = 5 - coe_shift * 4
    pi[V.NEQ] = 6 if pi[V.IFLOW] == 3 else 5
    pi[V.ITYPVENT] = 2; pi[V.NEQ_VENT] = 3; pi[V.NDIMDX_VENT] = 1
    pi[V.NIJK_VENT:V.NIJK_VENT + 5] = 0
    pi[V.PT_VOL] = 0
    pi[V.MXSSDOM_LU] = 1
    return deg
