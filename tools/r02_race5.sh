#!/bin/bash
mkdir -p gpurun_out
cmp() { python - "$@" <<'PY'
import json,sys
a,b=[json.load(open(f)) for f in sys.argv[2:4]]
bad=[(it,z) for it,(r,s) in enumerate(zip(a,b)) for z,(p,q) in enumerate(zip(r,s)) if p!=q]
print(sys.argv[1], 'mismatching (iteration, zone):', bad[:8], len(bad), 'of', len(a)*len(a[0]))
PY
}
N=128,128,32; NIT=150
for r in 1 2 3 4 5 6; do FAST_B200_LU=plane timeout 900 python tools/race_check.py c5 $NIT gpurun_out/race_p$r.json $N > /dev/null 2>&1; done
for r in 2 3 4 5 6; do cmp "plane 1v$r" gpurun_out/race_p1.json gpurun_out/race_p$r.json; done
