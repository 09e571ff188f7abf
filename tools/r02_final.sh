#!/bin/bash
# round 2 evidence pass: GPU tests, smoke, the driver's two bench arms as the driver runs them, racecheck of the LU mailboxes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | sed -n 2p
(time timeout 1500 python -m pytest tests -m gpu -q --durations=6) > gpurun_out/r02_final_pytest_gpu.log 2>&1; tail -12 gpurun_out/r02_final_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
(time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5) > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err; tail -4 gpurun_out/r02_final_bench_reference.err
(time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; tail -4 gpurun_out/r02_final_bench.err
python - <<'PY'
import json
try:
    r=json.loads(open('gpurun_out/r02_final_bench_reference.json').read().strip().splitlines()[-1])
    print('reference arm', r['value'], r['config']['steps_timed'], r['cpu_baseline']['cores'], r['config']['workload'][:60])
    d=json.loads(open('gpurun_out/r02_final_bench.json').read().strip().splitlines()[-1])
    print('c5', d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['whole_subiteration']['frac'], 'traffic', d['roofline']['traffic'])
    print(' e2e', d['e2e']); print(' cpu', d['cpu_baseline']); print(' launches', d['gpu_launches'], d['clocks'], d['config']['timed_region_s'])
    print(json.dumps(d['roofline']['kernels']))
    for s in d['config'].get('secondary', []): print(s['workload'][:40], round(s['value'],1), round(s['ms_per_step'],2), s['roofline']['kernel'], round(s['roofline']['frac'],3), round(s['roofline']['whole_subiteration']['frac'],3), round(s['e2e']['value'],1), s.get('cpu_baseline',{}).get('value'), s['steps'], round(s['timed_region_s'],2))
except Exception as e: print('ERR', e)
PY
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_implicit.py -q -x -k "sa_plate_homo_tiles or lamin_full_roe_tiles" > gpurun_out/r02_racecheck_lu.log 2>&1; grep -E "RACECHECK SUMMARY|passed|failed|Error" gpurun_out/r02_racecheck_lu.log | head -5
