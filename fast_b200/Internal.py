"""Minimal stand-in for Cassiopee's ``Converter.Internal`` (external to the reference).

Only what the hot-path host code needs: a CGNS/Python node is the 4-list
``[name, value, children, type]`` and the helpers below walk it.  When the real
``Converter.Internal`` is importable, ``fast_b200.FastC.PyTree`` uses it instead.
"""
import numpy

E_NpyInt = numpy.int32
__FlowSolutionNodes__ = 'FlowSolution'
__FlowSolutionCenters__ = 'FlowSolution#Centers'
__GridCoordinates__ = 'GridCoordinates'


def _isnode(n):
    return isinstance(n, list) and len(n) == 4 and isinstance(n[0], str)


def getNodeFromName1(node, name):
    for c in node[2]:
        if c[0] == name: return c
    return None


def getNodeFromName2(node, name):
    for c in node[2]:
        if c[0] == name: return c
    for c in node[2]:
        r = getNodeFromName1(c, name)
        if r is not None: return r
    return None


def getNodeFromName(node, name):
    if node[0] == name: return node
    for c in node[2]:
        r = getNodeFromName(c, name)
        if r is not None: return r
    return None


def getNodesFromName(node, name, out=None):
    if out is None: out = []
    if node[0] == name: out.append(node)
    for c in node[2]: getNodesFromName(c, name, out)
    return out


def getNodesFromType1(node, ntype):
    return [c for c in node[2] if c[3] == ntype]


def getNodesFromType2(node, ntype):
    out = [c for c in node[2] if c[3] == ntype]
    for c in node[2]: out += [d for d in c[2] if d[3] == ntype]
    return out


def getNodesFromType3(node, ntype):
    out = getNodesFromType2(node, ntype)
    for c in node[2]:
        for d in c[2]: out += [e for e in d[2] if e[3] == ntype]
    return out


def getNodesFromType(node, ntype, out=None):
    if out is None: out = []
    if node[3] == ntype: out.append(node)
    for c in node[2]: getNodesFromType(c, ntype, out)
    return out


def getNodeFromType1(node, ntype):
    r = getNodesFromType1(node, ntype)
    return r[0] if r else None


def getNodeFromPath(node, path):
    cur = node
    for p in path.split('/'):
        if p == '' or p == cur[0] and cur is node: continue
        cur = getNodeFromName1(cur, p)
        if cur is None: return None
    return cur


def getBases(t):
    return getNodesFromType1(t, 'CGNSBase_t')


def getZones(t):
    if t[3] == 'Zone_t': return [t]
    if t[3] == 'CGNSBase_t': return getNodesFromType1(t, 'Zone_t')
    return getNodesFromType2(t, 'Zone_t')


def getValue(node):
    v = node[1]
    if v is None: return None
    if isinstance(v, numpy.ndarray):
        if v.dtype.kind in ('S', 'U'):
            return v.tobytes().decode().strip() if v.dtype.kind == 'S' else ''.join(v.ravel().tolist())
        if v.size == 1: return v.ravel()[0].item()
        return v
    return v


def _tovalue(value):
    if value is None or isinstance(value, numpy.ndarray): return value
    if isinstance(value, str): return numpy.frombuffer(value.encode(), dtype='S1').copy()
    if isinstance(value, bool): return numpy.array([int(value)], dtype=E_NpyInt)
    if isinstance(value, int): return numpy.array([value], dtype=E_NpyInt)
    if isinstance(value, float): return numpy.array([value], dtype=numpy.float64)
    if isinstance(value, (list, tuple)):
        if len(value) and all(isinstance(x, (int, numpy.integer)) for x in value):
            return numpy.array(value, dtype=E_NpyInt)
        return numpy.array(value, dtype=numpy.float64)
    return value


def setValue(node, value):
    node[1] = _tovalue(value)


def createNode(name, ntype, value=None, children=None, parent=None):
    n = [name, _tovalue(value), children if children is not None else [], ntype]
    if parent is not None: parent[2].append(n)
    return n


def createChild(node, name, ntype, value=None, children=None):
    return createNode(name, ntype, value, children, parent=node)


def createUniqueChild(node, name, ntype, value=None, children=None):
    c = getNodeFromName1(node, name)
    if c is None: return createChild(node, name, ntype, value, children)
    c[3] = ntype
    c[1] = _tovalue(value)
    if children is not None: c[2] = children
    return c


def _rmNodesByName(node, name):
    node[2] = [c for c in node[2] if c[0] != name]
    for c in node[2]: _rmNodesByName(c, name)


def copyRef(node):
    return [node[0], node[1], [copyRef(c) for c in node[2]], node[3]]


def getZoneDim(z):
    d = z[1]
    return ['Structured', int(d[0, 0]), int(d[1, 0]), int(d[2, 0]), 3]


def newCGNSTree():
    return ['CGNSTree', None, [createNode('CGNSLibraryVersion', 'CGNSLibraryVersion_t', 3.1)], 'CGNSTree_t']


def newCGNSBase(name='Base', cellDim=3, physDim=3, parent=None):
    return createNode(name, 'CGNSBase_t', numpy.array([cellDim, physDim], dtype=E_NpyInt), parent=parent)


def newZone(name, ni, nj, nk, parent=None):
    """Structured zone of (ni,nj,nk) nodes (ghost layers included, as FastS expects)."""
    dims = numpy.array([[ni, ni - 1, 0], [nj, nj - 1, 0], [nk, nk - 1, 0]], dtype=E_NpyInt, order='F')
    z = createNode(name, 'Zone_t', dims, parent=parent)
    createChild(z, 'ZoneType', 'ZoneType_t', 'Structured')
    return z
