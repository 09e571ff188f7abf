#!/bin/bash
# ncu --set full captures of the dominant kernels (one GPU); reports come back in gpurun_out/
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rhs2 -s 6 -c 2 -f -o gpurun_out/prof_rhs2_tgv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full_tgv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rhs2 -s 6 -c 2 -f -o gpurun_out/prof_rhs2_plate python bench.py --steps 2 --warmup 1 --no-cpu --workload plate > gpurun_out/ncu_full_plate_rhs.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lu_tile -s 12 -c 2 -f -o gpurun_out/prof_lu_plate python bench.py --steps 2 --warmup 1 --no-cpu --workload plate > gpurun_out/ncu_full_plate_lu.log 2>&1
ls -la gpurun_out/*.ncu-rep
