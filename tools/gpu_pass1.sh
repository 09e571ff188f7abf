#!/bin/bash
# one GPU pass: parity tests, three bench workloads, ncu launch lists (for profiles/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tgv256.json 2> gpurun_out/bench_tgv256.err
timeout 300 python bench.py --steps 5 --warmup 3 --workload plate > gpurun_out/bench_plate.json 2> gpurun_out/bench_plate.err
timeout 300 python bench.py --steps 5 --warmup 3 --workload multiblock > gpurun_out/bench_multiblock.json 2> gpurun_out/bench_multiblock.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tgv256.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_tgv.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_plate.csv python bench.py --steps 2 --warmup 1 --no-cpu --workload plate > gpurun_out/ncu_plate.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_*.json
