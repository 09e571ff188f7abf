#!/bin/bash
# halo-exchange overlap A/B at N ranks: NCCL parity test, bench with and without overlap
N=${1:-2}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -x -q) 2>&1 | tail -3
for ov in 1 0; do
FAST_B200_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$ov bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/ov_$ov.err | tee gpurun_out/bench_n${N}_ov$ov.json | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('overlap=$ov',d['value'],d['ms_per_step'],d['roofline']['step_share'])"
done
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('N=1',d['value'],d['ms_per_step'])"
