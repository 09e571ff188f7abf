#!/bin/bash
mkdir -p gpurun_out
W=${W:-c5}; NIT=${NIT:-12}
timeout 600 python tools/race_check.py $W $NIT gpurun_out/race_cell1.json $N 2>&1 | tail -1
timeout 600 python tools/race_check.py $W $NIT gpurun_out/race_cell2.json $N 2>&1 | tail -1
FAST_B200_LU=tile timeout 600 python tools/race_check.py $W $NIT gpurun_out/race_tile.json $N 2>&1 | tail -1
python - <<'PY'
import json
a,b,c=[json.load(open('gpurun_out/race_%s.json'%n)) for n in ('cell1','cell2','tile')]
for name,x,y in (('cell vs cell',a,b),('cell vs tile',a,c)):
    bad=[(it,z) for it,(r,s) in enumerate(zip(x,y)) for z,(p,q) in enumerate(zip(r,s)) if p!=q]
    print(name, 'mismatching (iteration, zone):', bad[:20], 'of', len(x)*len(x[0]))
PY
