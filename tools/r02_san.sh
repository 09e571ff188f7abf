#!/bin/bash
mkdir -p gpurun_out
for k in euler_cart_roe_bdf sa_plate_homo-; do
timeout 600 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_gpu_implicit.py -q -x -k "$k" > gpurun_out/r02_san_$k.log 2>&1
grep -E "=========|passed|failed" gpurun_out/r02_san_$k.log | head -40
done
