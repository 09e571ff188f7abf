#!/bin/bash
# round 2, last GPU call: the evidence pass, then the profiles of the final build (launch list + ncu --set full of the dominant kernels)
bash tools/r02_final.sh
KERN=k_rhs3 WORK=c5 SKIP=9 COUNT=1 NAME=r02_ncu_rhs3_c5 bash tools/gpu_ncu_one.sh
KERN=k_rhs3 WORK=multiblock SKIP=9 COUNT=1 NAME=r02_ncu_rhs3_c4 bash tools/gpu_ncu_one.sh
WORK=c5 NAME=r02_launches_c5 bash tools/gpu_launchlist.sh
KERN=k_lu_cell_b WORK=c5 SKIP=2 COUNT=2 NAME=r02_ncu_lu_cell_c5 bash tools/gpu_ncu_one.sh
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_c5.csv
