#!/bin/bash
# full evidence pass for profiles/: parity tests, smoke, bench lines (with cpu_baseline), launch lists, ncu --set full of the top kernels
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log | head -1
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for w in tgv256 plate multiblock c5; do
  timeout 500 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/final_bench_$w.json 2> gpurun_out/final_bench_$w.err
done
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
for w in tgv256 plate multiblock; do WORK=$w NAME=final_launches_$w bash tools/gpu_launchlist.sh; done
KERN=k_rhs3 WORK=tgv256 SKIP=4 COUNT=3 NAME=final_prof_rhs3_tgv bash tools/gpu_ncu_one.sh
KERN=k_lu_tile_b WORK=multiblock SKIP=2 COUNT=2 NAME=final_prof_lub_multiblock bash tools/gpu_ncu_one.sh
KERN=k_lu_tile WORK=plate SKIP=8 COUNT=2 NAME=final_prof_lu_plate bash tools/gpu_ncu_one.sh
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/final_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[25:], d.get('value'), d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
