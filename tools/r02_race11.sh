#!/bin/bash
mkdir -p gpurun_out
export STRIDE=1
for r in 1 2 3 4; do FAST_B200_RHS=2 timeout 600 python tools/race_check.py c5 10 gpurun_out/race_k$r.json > /dev/null 2>&1; done
python - <<'PY'
import json
R=[json.load(open('gpurun_out/race_k%d.json'%r)) for r in (1,2,3,4)]
for a in range(4):
    for b in range(a+1,4):
        bad=[(it,z) for it,(r,s) in enumerate(zip(R[a],R[b])) for z,(p,q) in enumerate(zip(r,s)) if p!=q]
        print('fan + k_rhs2', a+1,'v',b+1, bad[:6], len(bad))
PY
