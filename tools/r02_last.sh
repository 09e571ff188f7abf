#!/bin/bash
# round 2, closing GPU call: the GPU suite and the default bench line of the last build
mkdir -p gpurun_out
(time timeout 200 python -m pytest tests -m gpu -q -x) > gpurun_out/r02_last_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02_last_pytest_gpu.log
(time timeout 150 python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r02_last_bench.json 2> gpurun_out/r02_last_bench.err; tail -3 gpurun_out/r02_last_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_last_bench.json').read().strip().splitlines()[-1])
    print('c5', d['value'], d['ms_per_step'], json.dumps(d['roofline']['kernels']), 'e2e', d['e2e']['value'])
    for s in d['config'].get('secondary', []): print(s['workload'][:40], round(s['value'],1), round(s['ms_per_step'],2), json.dumps(s['roofline']['kernels']), round(s['e2e']['value'],1))
except Exception as e: print('ERR', e)
PY
