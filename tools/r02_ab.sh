#!/bin/bash
# round 2: A/B of rhs kernel configurations (variant builds of the library, FAST_B200_LIB selects one); the GPU suite runs on the default build first
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x) > gpurun_out/ab_pytest.log 2>&1; tail -4 gpurun_out/ab_pytest.log
run() {   # variant, workloads...
  v=$1; shift
  lib=fast_b200/libfastb200${v:+_$v}.so
  for w in "$@"; do
    FAST_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --workload $w --no-secondary --no-cpu --no-e2e > gpurun_out/ab_${v:-base}_$w.json 2> gpurun_out/ab_${v:-base}_$w.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_${v:-base}_$w.json').read().strip().splitlines()[-1])
    k=d['roofline']['kernels']
    print('${v:-base}', '$w', round(d['value'],1), round(d['ms_per_step'],2), 'rhs/zone ms', round(k['k_rhs3']['avg_launch_ms_per_zone'],4), d['clocks']['sm_mhz'])
except Exception as e: print('${v:-base} $w ERR', e); print(open('gpurun_out/ab_${v:-base}_$w.err').read()[-800:])
PY
  done
}
run "" c5 plate multiblock
run slots3 c5 plate multiblock
run "" c5
export STRIDE=1
for r in 1 2 3; do timeout 600 python tools/race_check.py c5 8 gpurun_out/race_e$r.json > /dev/null 2>&1; done
python - <<'PY'
import json
R=[json.load(open('gpurun_out/race_e%d.json'%r)) for r in (1,2,3)]
for a in range(3):
    for b in range(a+1,3):
        bad=[(it,z) for it,(r,s) in enumerate(zip(R[a],R[b])) for z,(p,q) in enumerate(zip(r,s)) if p!=q]
        print('default config determinism', a+1,'v',b+1, bad[:6], len(bad))
PY
