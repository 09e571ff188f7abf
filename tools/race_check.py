"""Determinism / race check of the LU kernels: run NIT iterations of a workload and dump, per iteration and zone, a hash of the
state.  Two runs (same kernel twice, or the two LU kernels -- their arithmetic is ordered identically) must give identical files."""
import sys, os, hashlib, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy
import bench
from fast_b200.FastS import PyTree as FastS
from fast_b200.FastC import PyTree as FastC
from fast_b200.FastC.PyTree import Internal
w, nit, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
n = None
if len(sys.argv) > 4: n = tuple(int(x) for x in sys.argv[4].split(','))
t, j, desc = bench.make_case(w, n=n)
(t, tc, metrics) = FastS.warmup(t, j)
res = []
stride = int(os.environ.get('STRIDE', '1'))
for it in range(1, nit + 1):
    FastS._compute(t, metrics, it)
    if it % stride: continue
    FastS.sync_tree(t, levels=('',), aux=False)
    res.append([hashlib.md5(FastC.block_of(z, '').tobytes()).hexdigest() for z in Internal.getZones(t)])
json.dump(res, open(out, 'w'))
print(w, 'done', len(res))
