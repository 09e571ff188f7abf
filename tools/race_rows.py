"""Locate a nondeterministic event: per iteration, zone, variable, k plane and j row a wrap-around checksum of the state row.
Usage: race_rows.py workload nit out.npz [n]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy
import bench
from fast_b200.FastS import PyTree as FastS
from fast_b200.FastC import PyTree as FastC
from fast_b200.FastC.PyTree import Internal
from fast_b200 import VSHARE as V
w, nit, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
n = tuple(int(x) for x in sys.argv[4].split(',')) if len(sys.argv) > 4 else None
t, j, desc = bench.make_case(w, n=n)
(t, tc, metrics) = FastS.warmup(t, j)
zones = Internal.getZones(t)
res = []
for it in range(1, nit + 1):
    FastS._compute(t, metrics, it)
    FastS.sync_tree(t, levels=('',), aux=False)
    rows = []
    for z in zones:
        pi = Internal.getNodeFromName1(Internal.getNodeFromName1(z, '.Solver#ownData'), 'Parameter_int')[1]
        ni, nj, nk, neq = int(pi[V.NIJK]), int(pi[V.NIJK + 1]), int(pi[V.NIJK + 2]), int(pi[V.NEQ])
        a = FastC.block_of(z, '').view(numpy.uint64)[:neq * ni * nj * nk].reshape(neq, nk, nj, ni)
        rows.append(a.sum(axis=-1, dtype=numpy.uint64))
    res.append(numpy.stack(rows))
numpy.savez_compressed(out, rows=numpy.stack(res))
print('done')
