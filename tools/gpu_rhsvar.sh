#!/bin/bash
# A/B of the explicit rhs kernel builds (make VAR=v1 EXTRA=-DRHS_VAR=1, VAR=v2 EXTRA=-DRHS_VAR=2) on the TGV workload
mkdir -p gpurun_out
for v in "" _v1 _v2; do
  FAST_B200_LIB=$PWD/fast_b200/libfastb200$v.so timeout 300 python bench.py --steps 5 --warmup 3 --workload tgv256 --no-cpu > gpurun_out/rhsvar$v.json 2> gpurun_out/rhsvar$v.err
  python -c "
import json;d=json.loads(open('gpurun_out/rhsvar$v.json').read().strip().splitlines()[-1]);r=d['roofline'];print('var[$v]',round(d['value'],1),round(d['ms_per_step'],2),r['kernel'],round(r['avg_launch_ms'],3))"
done
