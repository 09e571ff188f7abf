#!/bin/bash
# round 2, first GPU pass: parity suite (incl. the full-size BASELINE configs), smoke, default bench (c5 + secondaries)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | sed -n 2p
(time timeout 1500 python -m pytest tests -m gpu -q -x --durations=12) > gpurun_out/r02_pytest_gpu.log 2>&1; tail -25 gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
(time timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -3 gpurun_out/r02_bench_default.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
    print('c5', d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['whole_subiteration']['frac'], d['e2e'], d['cpu_baseline'])
    print(json.dumps(d['roofline']['kernels']))
    for s in d['config'].get('secondary', []): print(s['workload'][:40], s['value'], s['ms_per_step'], s['roofline']['kernel'], s['roofline']['frac'], s['roofline']['whole_subiteration']['frac'], s['e2e']['value'], s.get('cpu_baseline',{}).get('value'))
except Exception as e: print('ERR', e)
PY
