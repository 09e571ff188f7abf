#!/bin/bash
mkdir -p gpurun_out
cmp() { python - "$@" <<'PY'
import json,sys
a,b=[json.load(open(f)) for f in sys.argv[2:4]]
bad=[(it,z) for it,(r,s) in enumerate(zip(a,b)) for z,(p,q) in enumerate(zip(r,s)) if p!=q]
print(sys.argv[1], 'mismatching (iteration, zone):', bad[:12], len(bad), 'of', len(a)*len(a[0]))
PY
}
N=128,128,32; NIT=100
for r in 1 2 3 4; do FAST_B200_LU=tile timeout 600 python tools/race_check.py c5 $NIT gpurun_out/race_t$r.json $N > /dev/null 2>&1; done
for r in 2 3 4; do cmp "tile 1v$r" gpurun_out/race_t1.json gpurun_out/race_t$r.json; done
for r in 1 2 3 4; do FAST_B200_FAN=0 timeout 600 python tools/race_check.py c5 $NIT gpurun_out/race_f$r.json $N > /dev/null 2>&1; done
for r in 1 2 3 4; do cmp "cell nofan t1v$r" gpurun_out/race_t1.json gpurun_out/race_f$r.json; done
