#!/bin/bash
# per-role cycle counters of k_lu_cell on the C3 plate (single zone: run_lu_tiles reports them)
mkdir -p gpurun_out
FAST_B200_LIB=$PWD/fast_b200/libfastb200_clk.so FAST_B200_LU_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 1 --workload plate --no-secondary --no-cpu --no-e2e > gpurun_out/r02_ludbg.json 2> gpurun_out/r02_ludbg.err
grep LUDBG gpurun_out/r02_ludbg.err | head -60
