#!/bin/bash
# A/B of a variant build of the library (make -C fast_b200/csrc VAR=name ...): parity of the SA cases + bench lines
# usage: gpu_libvar.sh name "workloads"
V=$1; mkdir -p gpurun_out
FAST_B200_LIB=$PWD/fast_b200/libfastb200_$V.so timeout 600 python -m pytest tests -m gpu -x -q -k "sa_ or smoke or c1" 2>&1 | tail -1
for lib in "" _$V; do
  for w in ${2:-plate c5}; do
    FAST_B200_LIB=$PWD/fast_b200/libfastb200$lib.so timeout 400 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu --no-e2e > gpurun_out/libvar${lib}_$w.json 2> gpurun_out/libvar${lib}_$w.err
    python -c "
import json;d=json.loads(open('gpurun_out/libvar${lib}_$w.json').read().strip().splitlines()[-1]);r=d['roofline'];print('lib[$lib] $w',round(d['value'],1),round(d['ms_per_step'],2),{k:round(v*d['ms_per_step'],2) for k,v in r['step_share'].items()})"
  done
done
