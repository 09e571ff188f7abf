#!/usr/bin/env python
"""rhs / LU time per iteration of a laminar 3dhomo plate (the 5-equation kernels with the shared metric tile): A/B of library builds
through FAST_B200_LIB.  python tools/homo_lam_time.py [ni nj nk]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_b200 import synth, fasts
from fast_b200.FastS import PyTree as FastS
from fast_b200.FastC import PyTree as FastC
n = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (256, 128, 64)
t, j = synth.case_flatplate(n=n, model='NSLaminar', ReInf=2.e4, modulo_verif=1000)
(t, tc, metrics) = FastS.warmup(t, j)
for it in range(1, 4): FastS._compute(t, metrics, it)
fasts.timing(FastC.HOOK, True)
N = 10
for it in range(4, 4 + N): FastS._compute(t, metrics, it)
from fast_b200 import _lib
_lib.check(_lib.lib().fb_sync(FastC.HOOK['solver']))
out = {k: fasts.timing_get(FastC.HOOK, k) for k in ('all', 'rhs', 'lu')}
cells = n[0] * n[1] * n[2]
print(os.path.basename(os.environ.get('FAST_B200_LIB', 'default')), n, 'ms/iteration %.3f' % (out['all'][0] / N),
      'rhs ms/iteration %.3f' % (out['rhs'][0] / N), 'lu ms/iteration %.3f' % (out['lu'][0] / N),
      'Mcell.it/s %.1f' % (cells * N / (out['all'][0] * 1e-3) / 1e6))
