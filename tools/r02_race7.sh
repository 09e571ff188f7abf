#!/bin/bash
mkdir -p gpurun_out
for r in 1 2 3; do timeout 600 python tools/race_rows.py c5 ${NIT:-10} gpurun_out/rows$r.npz > /dev/null 2>&1; done
python - <<'PY'
import numpy
R=[numpy.load('gpurun_out/rows%d.npz'%r)['rows'] for r in (1,2,3)]
for a,b,name in ((0,1,'1v2'),(0,2,'1v3'),(1,2,'2v3')):
    d = R[a]!=R[b]
    its = numpy.nonzero(d.reshape(d.shape[0],-1).any(axis=1))[0]
    print(name, 'iterations with differences:', its.tolist())
    if len(its):
        it=its[0]; idx=numpy.argwhere(d[it])
        print('  first differing iteration', it+1, ': rows differing', len(idx))
        for z in sorted(set(idx[:,0])):
            sub=idx[idx[:,0]==z]
            print('   zone',z,'vars',sorted(set(sub[:,1])),'k range',sub[:,2].min(),sub[:,2].max(),'j range',sub[:,3].min(),sub[:,3].max(),'rows',len(sub))
            print('      sample (var,k,j):', sub[:12,1:].tolist())
PY
rm -f gpurun_out/rows*.npz
