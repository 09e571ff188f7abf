#!/bin/bash
# multi-GPU check: N = $1 ranks on one box (bench.py weak-scaling family + NCCL parity test)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 1500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err | cut -c1-300
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-600
(timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -x -q) 2>&1 | tail -5
