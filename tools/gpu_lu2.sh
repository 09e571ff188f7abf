#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log | cut -c1-220
WORKLOADS="plate multiblock" LUDBG=1 bash tools/gpu_variants.sh
