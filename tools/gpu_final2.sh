#!/bin/bash
# end-of-round check of the committed build: parity tests, smoke, the four bench lines, ncu of the SA rhs kernel
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log | head -1
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for w in tgv256 plate multiblock c5; do
  timeout 500 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/final_bench_$w.json 2> gpurun_out/final_bench_$w.err
done
KERN=k_rhs3 WORK=plate SKIP=4 COUNT=2 NAME=final_prof_rhs3_plate bash tools/gpu_ncu_one.sh
WORK=plate NAME=final_launches_plate bash tools/gpu_launchlist.sh
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/final_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[25:], d.get('value'), d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
