#!/bin/bash
# ncu --set full of one kernel: KERN=regex WORK=workload SKIP=n NAME=out
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KERN -s ${SKIP:-4} -c ${COUNT:-2} -f -o gpurun_out/$NAME python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-secondary --workload $WORK > gpurun_out/$NAME.log 2>&1
ls -la gpurun_out/$NAME.ncu-rep
