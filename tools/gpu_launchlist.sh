#!/bin/bash
# ncu launch list of one workload: WORK=tgv256 NAME=launches_x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/$NAME.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-secondary --workload $WORK > gpurun_out/$NAME.log 2>&1
