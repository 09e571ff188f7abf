#!/bin/bash
mkdir -p gpurun_out
KERN=k_rhs3 WORK=tgv256 SKIP=4 COUNT=2 NAME=prof_rhs3_tgv bash tools/gpu_ncu_one.sh
KERN=k_rhs3 WORK=plate SKIP=4 COUNT=2 NAME=prof_rhs3_plate bash tools/gpu_ncu_one.sh
KERN=k_lu_tile WORK=plate SKIP=9 COUNT=2 NAME=prof_lu_tile_upper_plate bash tools/gpu_ncu_one.sh
KERN="k_sensor|k_visc|k_bc|k_coe" WORK=tgv256 SKIP=4 COUNT=6 NAME=prof_small_tgv bash tools/gpu_ncu_one.sh
