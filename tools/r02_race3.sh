#!/bin/bash
mkdir -p gpurun_out
cmp() { python - "$@" <<'PY'
import json,sys
a,b=[json.load(open(f)) for f in sys.argv[2:4]]
bad=[(it,z) for it,(r,s) in enumerate(zip(a,b)) for z,(p,q) in enumerate(zip(r,s)) if p!=q]
print(sys.argv[1], 'mismatching (iteration, zone):', bad[:12], len(bad), 'of', len(a)*len(a[0]))
PY
}
W=${W:-c5}; N=${N:-128,128,32}; NIT=${NIT:-100}; R=${R:-4}
for r in $(seq 1 $R); do timeout 600 python tools/race_check.py $W $NIT gpurun_out/race_s$r.json $N > /dev/null 2>&1; done
for r in $(seq 2 $R); do cmp "$W $N 1v$r" gpurun_out/race_s1.json gpurun_out/race_s$r.json; done
