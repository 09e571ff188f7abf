#!/bin/bash
mkdir -p gpurun_out
for n in 32 64; do timeout 120 python tools/dbg_rhs3.py $n 2>&1 | tail -1; done
bash tools/gpu_pass2.sh
