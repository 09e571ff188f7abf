#!/bin/bash
# round-2 profiling evidence: launch list of the default workload + ncu --set full of its two dominant kernels
mkdir -p gpurun_out
WORK=c5 NAME=r02_launches_c5 bash tools/gpu_launchlist.sh
KERN=k_lu_cell_b WORK=c5 SKIP=2 COUNT=2 NAME=r02_ncu_lu_cell_c5 bash tools/gpu_ncu_one.sh
KERN=k_rhs3 WORK=c5 SKIP=9 COUNT=1 NAME=r02_ncu_rhs3_c5 bash tools/gpu_ncu_one.sh
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_c5.csv
