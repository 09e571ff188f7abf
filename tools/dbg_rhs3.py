import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_b200 import synth
from fast_b200.FastS import PyTree as FastS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
t = synth.case_tgv(n=n)
(t, tc, metrics) = FastS.warmup(t, None)
for it in (1, 2):
    FastS._compute(t, metrics, it)
FastS.sync_tree(t)
print("ok", n)
