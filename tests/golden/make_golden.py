"""Generates tests/golden/oracle_regression.npz: interior state after 2 iterations of seeded small cases, computed
by the CPU ORACLE (strict-FP build).  These are regression fixtures of this repo's oracle -- the reference ships no
fixtures for the path (PARITY UNPINNED) -- used to detect accidental changes of the oracle and, on the
GPU box, as a second comparison target for the CUDA path.
    python tests/golden/make_golden.py"""
import os, sys
import numpy
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
from helpers import oracle_from_tree, interior, V   # noqa: E402
from fast_b200 import synth                           # noqa: E402

CASES = {
    'c1_euler_cart_ausm_rk3': lambda: (synth.case_euler_box(n=12, scheme='ausmpred'), None),
    'c2_lamin_cart_senseur_rk3': lambda: (synth.case_euler_box(n=12, scheme='senseur', model='NSLaminar', ReInf=1600.), None),
    'c3_sa_plate_homo_lu': lambda: synth.case_flatplate(n=(16, 12, 4)),
    'c4_blocks_roe_lu_match': lambda: synth.case_multiblock(nblk=(2, 1, 1), n=8),
    'c5_lamin_roe_wall_flux': lambda: (synth.case_euler_box(n=10, scheme='roe', model='NSLaminar', ReInf=2000., deform_amp=0.02,
                                                            bcs={'imin': 'BCWallInviscid', 'imax': 'BCFarfield', 'jmin': 'BCWallViscous',
                                                                 'jmax': 'BCFarfield', 'kmin': 'BCFarfield', 'kmax': 'BCWall'}), None),
    'c6_sa_plate_ssor2': lambda: synth.case_flatplate(n=(16, 12, 4), nb_relax=2),
    'c7_lamin_full_roe_ssor3': lambda: (synth.case_euler_box(n=10, scheme='roe', model='NSLaminar', ReInf=1000., deform_amp=0.02,
                                                              temporal='implicit', extra_numz={'nb_relax': 3}), None),
}


def run(name, nit=2):
    t, j = CASES[name]()
    o = oracle_from_tree(t, joins=j)
    nss = int(o.zones[0].pi[V.NSSITER])
    for it in range(1, nit + 1): o.compute(it, 1, nss, 1)
    return numpy.concatenate([interior(z.pi, z.lev[0]).ravel() for z in o.zones])


if __name__ == '__main__':
    numpy.savez_compressed(os.path.join(HERE, 'oracle_regression.npz'), **{k: run(k) for k in CASES})
    print('written', {k: run(k).size for k in CASES})
