#!/bin/bash
# round 2, second pass: whole GPU suite, determinism of the default configuration, quick bench lines
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x --durations=8) > gpurun_out/r02_pytest_gpu.log 2>&1; tail -16 gpurun_out/r02_pytest_gpu.log
export STRIDE=1
for r in 1 2 3; do timeout 600 python tools/race_check.py c5 8 gpurun_out/race_d$r.json > /dev/null 2>&1; done
python - <<'PY'
import json
R=[json.load(open('gpurun_out/race_d%d.json'%r)) for r in (1,2,3)]
for a in range(3):
    for b in range(a+1,3):
        bad=[(it,z) for it,(r,s) in enumerate(zip(R[a],R[b])) for z,(p,q) in enumerate(zip(r,s)) if p!=q]
        print('default config determinism', a+1,'v',b+1, bad[:6], len(bad))
PY
for w in c5 multiblock plate; do
  timeout 400 python bench.py --steps 10 --warmup 3 --workload $w --no-secondary --no-cpu --no-e2e > gpurun_out/r02_p2_$w.json 2> gpurun_out/r02_p2_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_p2_$w.json').read().strip().splitlines()[-1])
    print('$w', round(d['value'],1), round(d['ms_per_step'],2), d['roofline']['whole_subiteration']['frac'], json.dumps(d['roofline']['kernels']), d['roofline']['step_share'])
except Exception as e: print('$w ERR', e); print(open('gpurun_out/r02_p2_$w.err').read()[-1500:])
PY
done

