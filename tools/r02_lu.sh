#!/bin/bash
# LU kernel A/B: parity (implicit cases + full-size C3/C4/C5), then c5 / plate / multiblock bench lines with the thread-per-cell
# kernel (default) and with the four-lanes-per-line kernel (FAST_B200_LU=tile)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_implicit.py tests/test_gpu_drift.py tests/test_zz_gpu_golden.py -q -x) > gpurun_out/r02_lu_pytest.log 2>&1; tail -5 gpurun_out/r02_lu_pytest.log
(timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -x -k "c3 or c4 or c5") > gpurun_out/r02_lu_pytest_full.log 2>&1; tail -5 gpurun_out/r02_lu_pytest_full.log
for w in c5 plate multiblock; do
  for v in cell tile; do
    FAST_B200_LU=$v timeout 400 python bench.py --steps 10 --warmup 3 --workload $w --no-secondary --no-cpu --no-e2e > gpurun_out/r02_lu_${w}_$v.json 2> gpurun_out/r02_lu_${w}_$v.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_lu_${w}_$v.json').read().strip().splitlines()[-1])
    print('$w $v', round(d['value'],1), round(d['ms_per_step'],2), d['roofline']['whole_subiteration']['frac'], json.dumps(d['roofline']['kernels']))
except Exception as e: print('$w $v ERR', e); print(open('gpurun_out/r02_lu_${w}_$v.err').read()[-1500:])
PY
  done
done
