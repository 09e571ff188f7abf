"""fast_b200 -- B200-native (sm_100a, FP64 CUDA) implementation of FastS's residual-and-update hot path
behind the reference's Python surface.  ``install_aliases()`` registers FastS / FastC / Fast so that
Fast/valid-style scripts (``import FastS.PyTree as FastS``) run unchanged."""
import sys

__all__ = ['install_aliases']


def install_aliases():
    from . import FastC, FastS, Fast
    from .FastC import PyTree as _c
    from .FastS import PyTree as _s
    from .Fast import PyTree as _f
    sys.modules.setdefault('FastC', FastC); sys.modules.setdefault('FastC.PyTree', _c)
    sys.modules.setdefault('FastS', FastS); sys.modules.setdefault('FastS.PyTree', _s)
    sys.modules.setdefault('Fast', Fast); sys.modules.setdefault('Fast.PyTree', _f)
