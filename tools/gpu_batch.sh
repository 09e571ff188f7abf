#!/bin/bash
# multi-zone LU launch (k_lu_tile_b): parity tests, A/B on the multi-block workloads, one-GPU share of C5
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log | head -1
show() { python -c "
import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);r=d['roofline'];print('$1',round(d['value'],1),round(d['ms_per_step'],2),r['kernel'],round(r['avg_launch_ms'],3),round(r['frac'],3),round(r['whole_subiteration']['frac'],3),{k:round(v,3) for k,v in r['step_share'].items()})"; }
for b in 1 0; do
  FAST_B200_LU_BATCH=$b timeout 300 python bench.py --steps 5 --warmup 3 --workload multiblock --no-cpu > gpurun_out/batch${b}_multiblock.json 2> gpurun_out/batch${b}_multiblock.err; show gpurun_out/batch${b}_multiblock.json
done
for b in 1 0; do
  FAST_B200_LU_BATCH=$b timeout 400 python bench.py --steps 4 --warmup 3 --workload c5 --no-cpu > gpurun_out/batch${b}_c5.json 2> gpurun_out/batch${b}_c5.err; show gpurun_out/batch${b}_c5.json
done
timeout 300 python bench.py --steps 5 --warmup 3 --workload plate --no-cpu > gpurun_out/batch_plate.json 2> gpurun_out/batch_plate.err; show gpurun_out/batch_plate.json
