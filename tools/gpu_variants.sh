#!/bin/bash
# A/B runs of library variants (fast_b200/csrc/Makefile VAR=...) on tgv256 / plate / multiblock; LU tile timing debug
mkdir -p gpurun_out
for v in "" $VARIANTS; do
  lib=$PWD/fast_b200/libfastb200${v:+_$v}.so
  for w in ${WORKLOADS:-tgv256 plate}; do
    FAST_B200_LIB=$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload $w > gpurun_out/v_${v:-base}_$w.json 2> gpurun_out/v_${v:-base}_$w.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/v_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; sh=r['step_share']; ms=d['ms_per_step']; ns=d['config']['sub_iterations_per_step']
        print('%-40s value %8.1f ms/step %7.3f | per sub-iteration: rhs %.3f lu %.3f bc %.3f coe %.3f ms'%(f[13:],d['value'],ms,ms*sh['rhs']/ns,ms*sh['lu']/ns,ms*sh['bc']/ns,ms*sh['coe']/ns))
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-300:])
PY
if [ -n "$LUDBG" ]; then FAST_B200_LU_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --workload plate > /dev/null 2> gpurun_out/ludbg.txt; grep LUDBG gpurun_out/ludbg.txt | head -150; fi
