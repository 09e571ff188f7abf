#!/bin/bash
mkdir -p gpurun_out
cmp() { python - "$@" <<'PY'
import json,sys
a,b=[json.load(open(f)) for f in sys.argv[2:4]]
bad=[(it,z) for it,(r,s) in enumerate(zip(a,b)) for z,(p,q) in enumerate(zip(r,s)) if p!=q]
print(sys.argv[1], 'mismatching (iteration, zone):', bad[:16], len(bad), 'of', len(a)*len(a[0]))
PY
}
# small bricks, many iterations
for r in 1 2 3; do timeout 600 python tools/race_check.py c5 60 gpurun_out/race_s$r.json 128,128,32 > /dev/null 2>&1; done
cmp "small 1v2" gpurun_out/race_s1.json gpurun_out/race_s2.json; cmp "small 1v3" gpurun_out/race_s1.json gpurun_out/race_s3.json
# full size without the stream fan-out
for r in 1 2; do FAST_B200_FAN=0 timeout 600 python tools/race_check.py c5 12 gpurun_out/race_nf$r.json > /dev/null 2>&1; done
cmp "full nofan 1v2" gpurun_out/race_nf1.json gpurun_out/race_nf2.json
