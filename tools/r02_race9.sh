#!/bin/bash
mkdir -p gpurun_out
export STRIDE=1
for m in 1 2; do
for r in 1 2 3; do FAST_B200_FAN=$m timeout 600 python tools/race_check.py c5 10 gpurun_out/race_m${m}_$r.json > /dev/null 2>&1; done
python - $m <<'PY'
import json,sys
m=sys.argv[1]
R=[json.load(open('gpurun_out/race_m%s_%d.json'%(m,r))) for r in (1,2,3)]
for a in range(3):
    for b in range(a+1,3):
        bad=[(it,z) for it,(r,s) in enumerate(zip(R[a],R[b])) for z,(p,q) in enumerate(zip(r,s)) if p!=q]
        print('fan mode',m, a+1,'v',b+1, bad[:6], len(bad))
PY
done
