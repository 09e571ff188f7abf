import sys, os, numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import oracle_from_tree, interior, relerr, FastS, FastC, V
from fast_b200 import synth
from fast_b200.FastC.PyTree import Internal
n = int(sys.argv[1]); nit = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t1 = synth.case_tgv(n=n); t2 = synth.case_tgv(n=n)
o = oracle_from_tree(t1)
(t2, tc, metrics) = FastS.warmup(t2, None)
for it in range(1, nit + 1):
    o.compute(it, 1, 3, 1); FastS._compute(t2, metrics, it)
FastS.sync_tree(t2)
zo = o.zones[0]; z = Internal.getZones(t2)[0]
a = interior(zo.pi, FastC.block_of(z, '')); b = interior(zo.pi, zo.lev[0])
d = numpy.abs(a - b)
print(os.environ.get('FAST_B200_RHS'), os.environ.get('FAST_B200_SENSOR'), 'n', n, 'relerr', relerr(a, b))
for e in range(5):
    idx = numpy.unravel_index(numpy.argmax(d[e]), d[e].shape)
    print('  var', e, 'max', d[e].max(), 'at (k,j,i)', idx, 'count>1e-9', int((d[e] > 1e-9).sum()))
bad = numpy.argwhere(d[0] > 1e-9)
if len(bad): print('  bad k range', bad[:,0].min(), bad[:,0].max(), 'j', bad[:,1].min(), bad[:,1].max(), 'i', bad[:,2].min(), bad[:,2].max(), 'distinct i', sorted(set(bad[:,2]))[:20], 'distinct j', sorted(set(bad[:,1]))[:20], 'distinct k', sorted(set(bad[:,0]))[:20])
