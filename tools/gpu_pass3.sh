#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log | cut -c1-200
WORKLOADS="${WORKLOADS:-tgv256 plate multiblock}" bash tools/gpu_variants.sh
