#!/bin/bash
# quick regression check: bench lines of the workloads named in $W (default: all four), no CPU baseline
mkdir -p gpurun_out
show() { python -c "
import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);r=d['roofline'];print('$1',round(d['value'],1),round(d['ms_per_step'],2),r['kernel'],round(r['avg_launch_ms'],3),round(r['frac'],3),round(r['whole_subiteration']['frac'],3),{k:round(v*d['ms_per_step'],2) for k,v in r['step_share'].items()})"; }
for w in ${W:-tgv256 plate multiblock c5}; do
  timeout 400 python bench.py --steps ${STEPS:-5} --warmup 3 --workload $w --no-cpu > gpurun_out/quick_$w.json 2> gpurun_out/quick_$w.err; show gpurun_out/quick_$w.json
done
