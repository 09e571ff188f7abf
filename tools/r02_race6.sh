#!/bin/bash
mkdir -p gpurun_out
cmp() { python - "$@" <<'PY'
import json,sys
a,b=[json.load(open(f)) for f in sys.argv[2:4]]
bad=[(it,z) for it,(r,s) in enumerate(zip(a,b)) for z,(p,q) in enumerate(zip(r,s)) if p!=q]
print(sys.argv[1], 'first mismatching (sample, zone):', bad[:6], len(bad), 'of', len(a)*len(a[0]))
PY
}
export STRIDE=20
N=${N:-64,64,16}; NIT=${NIT:-2000}
for r in 1 2 3 4 5 6; do timeout 300 python tools/race_check.py c5 $NIT gpurun_out/race_a$r.json $N > /dev/null 2>&1; done
for r in 2 3 4 5 6; do cmp "fan 1v$r" gpurun_out/race_a1.json gpurun_out/race_a$r.json; done
for r in 1 2 3 4 5 6; do FAST_B200_FAN=0 timeout 300 python tools/race_check.py c5 $NIT gpurun_out/race_b$r.json $N > /dev/null 2>&1; done
for r in 1 2 3 4 5 6; do cmp "nofan a1v$r" gpurun_out/race_a1.json gpurun_out/race_b$r.json; done
