"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this)."""
import ctypes, os, subprocess
import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double); _ip = ctypes.POINTER(ctypes.c_int)


class fo_zone(ctypes.Structure):
    _fields_ = [('pi', _ip), ('pr', _dp), ('ro', _dp), ('ro_m1', _dp), ('ro_p1', _dp), ('xmut', _dp),
                ('x', _dp), ('y', _dp), ('z', _dp), ('ti', _dp), ('tj', _dp), ('tk', _dp), ('vol', _dp),
                ('drodm', _dp), ('drodm2', _dp), ('coe', _dp), ('wig', _dp), ('rdm', _dp), ('cfl', ctypes.c_double * 3)]


def build(force=False):
    out = os.path.join(_HERE, '_build', 'liboracle.so')
    if force or not os.path.exists(out) or not os.path.exists(os.path.join(_HERE, '_build', 'liboracle_fast.so')):
        subprocess.check_call(['make', '-C', _HERE], stdout=subprocess.DEVNULL)
    return out


_libs = {}


def lib(fast=False):
    if fast not in _libs:
        build()
        L = ctypes.CDLL(os.path.join(_HERE, '_build', 'liboracle_fast.so' if fast else 'liboracle.so'))
        L.fo_max_threads.restype = ctypes.c_int
        _libs[fast] = L
    return _libs[fast]


def _p(a, t=_dp):
    return None if a is None else a.ctypes.data_as(t)


class Zone:
    """Host arrays of one zone in the reference layout + the fo_zone view handed to C."""
    def __init__(self, pi, pr, ro, ro_m1, ro_p1, xmut, x, y, z):
        from fast_b200 import VSHARE as V
        self.pi = numpy.ascontiguousarray(pi, dtype=numpy.int32); self.pr = numpy.ascontiguousarray(pr, dtype=numpy.float64)
        nd = int(pi[V.NDIMDX]); neq = int(pi[V.NEQ]); nm = int(pi[V.NDIMDX_MTR])
        self.lev = [numpy.array(a, dtype=numpy.float64).ravel().copy() for a in (ro, ro_m1, ro_p1)]
        self.xmut = numpy.zeros(2 * nd) if xmut is None else numpy.array(xmut, dtype=numpy.float64).ravel().copy()
        if self.xmut.size < 2 * nd: self.xmut = numpy.concatenate([self.xmut, numpy.zeros(2 * nd - self.xmut.size)])
        self.x, self.y, self.z = [numpy.ascontiguousarray(a.ravel(order='F')) for a in (x, y, z)]
        self.ti = numpy.zeros(nm * int(pi[V.NEQ_IJ])); self.tj = numpy.zeros(nm * int(pi[V.NEQ_IJ]))
        self.tk = numpy.zeros(nm * max(int(pi[V.NEQ_K]), 1)); self.vol = numpy.zeros(nm)
        self.drodm = numpy.zeros(nd * neq); self.coe = numpy.zeros(nd * max(int(pi[V.NEQ_COE]), 1))
        self.wig = numpy.zeros(nd * 3 if int(pi[V.KFLUDOM]) == 2 else 3)     # SENSOR switch (wiggle2001.for)
        nblk = 1
        self.rdm = numpy.zeros(int(pi[V.NSSITER]) * neq * 2 * nblk)
        self.neq = neq; self.nd = nd

    def fill(self, s):
        s.pi = _p(self.pi, _ip); s.pr = _p(self.pr)
        s.ro, s.ro_m1, s.ro_p1 = _p(self.lev[0]), _p(self.lev[1]), _p(self.lev[2])
        s.xmut = _p(self.xmut); s.x, s.y, s.z = _p(self.x), _p(self.y), _p(self.z)
        s.ti, s.tj, s.tk, s.vol = _p(self.ti), _p(self.tj), _p(self.tk), _p(self.vol)
        s.drodm = _p(self.drodm); s.drodm2 = None; s.coe = _p(self.coe); s.wig = _p(self.wig); s.rdm = _p(self.rdm)

    def metrics(self, L):
        L.fo_cp_tijk(_p(self.pi, _ip), _p(self.x), _p(self.y), _p(self.z), _p(self.ti), _p(self.tj), _p(self.tk))
        L.fo_cp_vol(_p(self.pi, _ip), _p(self.x), _p(self.y), _p(self.z), _p(self.ti), _p(self.tj), _p(self.tk), _p(self.vol))


def tc_to_table(pit, prt):
    """Connector's flattened donor tree (param_int_tc, param_real_tc) -> the match-join table fo_transfers applies, reading the
    fields exactly where K_FASTC::setInterpTransfersIntra reads them (C/Com/setInterpTransfersFast.cpp:469-477 header, :641-654
    zones / variables, :750-760 point lists).  Match joins only: type 0 / 1, coefficient 1."""
    pit = numpy.asarray(pit); prt = numpy.asarray(prt)
    sizecomIBC = int(pit[2]); sizecomID = int(pit[3 + sizecomIBC]); shift_graph = sizecomIBC + sizecomID + 3
    ech = int(pit[1 + shift_graph])                                  # NoTransfert = 1
    nrac, nrac_inst, timelevel = int(pit[ech + 1]), int(pit[ech + 2]), int(pit[ech + 3])
    assert nrac_inst == 0
    desc, idx, off = [], [], 0
    for irac in range(nrac):
        sr = ech + 4 + timelevel * 2 + irac
        assert pit[sr + nrac * 3] < 0 and pit[sr + nrac * 14] == 0
        NoD, NoR, npts = int(pit[sr + nrac * 5]), int(pit[sr + nrac * 11]), int(pit[sr + nrac * 10])
        pt = int(pit[sr + nrac * 7]); ntype = int(pit[pt]); types = pit[pt + 1 + ntype: pt + 1 + ntype + npts]
        assert numpy.all((types == 0) | (types == 1))
        pc = int(pit[sr + nrac * 8]); assert numpy.all(prt[pc:pc + npts] == 1.)
        pd, pr = int(pit[sr + nrac * 6]), int(pit[sr + nrac * 12])
        desc += [0, NoD, NoR, npts, off, off + npts, 0]; idx += [pit[pd:pd + npts], pit[pr:pr + npts]]; off += 2 * npts
    return numpy.concatenate([numpy.array([nrac, off], numpy.int32), numpy.array(desc, numpy.int32)] + idx).astype(numpy.int32)


class Oracle:
    """Runs fo_compute on a list of Zone objects, rotating the python-side level list like the reference."""
    def __init__(self, zones, pitc=None, prtc=None, fast=False):
        self.L = lib(fast); self.zones = zones
        self.arr = (fo_zone * len(zones))()
        if isinstance(pitc, (tuple, list)) and len(pitc) == 2: pitc = tc_to_table(*pitc)     # reference wire format
        self.pitc = None if pitc is None else numpy.ascontiguousarray(pitc, dtype=numpy.int32)
        self.prtc = None if prtc is None else numpy.ascontiguousarray(prtc, dtype=numpy.float64)
        for z in zones: z.metrics(self.L)
        self.first_it = ctypes.c_int(0)

    def fill_ghost(self, level=0):
        for z, s in zip(self.zones, self.arr): z.fill(s)
        self.L.fo_fill_ghost(self.arr, len(self.zones), level, _p(self.pitc, _ip), _p(self.prtc))

    def compute(self, nitrun, nit, nssiter, modulo_verif, first_it=None, nthreads=1):
        if first_it is not None: self.first_it.value = first_it
        for z, s in zip(self.zones, self.arr): z.fill(s)
        self.L.fo_compute(self.arr, len(self.zones), nitrun, nit, nssiter, modulo_verif, ctypes.byref(self.first_it),
                          _p(self.pitc, _ip), _p(self.prtc), nthreads)
        from fast_b200 import VSHARE as V
        for z, s in zip(self.zones, self.arr): z.cfl = numpy.array([s.cfl[0], s.cfl[1], s.cfl[2]])   # max, min, mean
        if int(self.zones[0].pi[V.ITYPCP]) == 2 and nssiter % 2 == 0: return self.first_it.value   # RK2: no rotation (fo_driver.c)
        for z in self.zones:
            for _ in range(nit): z.lev = [z.lev[2], z.lev[0], z.lev[1]]   # ro<-ro_p1, ro_m1<-ro, ro_p1<-old ro_m1
        return self.first_it.value
