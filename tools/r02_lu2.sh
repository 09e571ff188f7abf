#!/bin/bash
mkdir -p gpurun_out
for v in tile cell; do
FAST_B200_LU=$v timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -s -k "c5" > gpurun_out/r02_lu_full_$v.log 2>&1; grep -E "per-variable|passed|failed" gpurun_out/r02_lu_full_$v.log | cut -c1-200
done
