#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
for w in tgv256 plate multiblock; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload $w > gpurun_out/b3_$w.json 2> gpurun_out/b3_$w.err
  FAST_B200_RHS=2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload $w > gpurun_out/b2_$w.json 2> gpurun_out/b2_$w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b[23]_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, 'value %.1f ms/step %.3f kern %s avg_ms %.3f share %s'%(d['value'],d['ms_per_step'],r['kernel'],r['avg_launch_ms'],{k:round(v,3) for k,v in r['step_share'].items()}))
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
