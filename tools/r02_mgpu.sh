#!/bin/bash
# 2-GPU pass: the NCCL parity cases (the driver's 1-GPU box skips them) + weak-scaling member of the default bench
mkdir -p gpurun_out
N=${N:-2}
(time timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q -x) > gpurun_out/r02_pytest_nccl_${N}gpu.log 2>&1; tail -6 gpurun_out/r02_pytest_nccl_${N}gpu.log
FAST_B200_EXCHANGE=torch timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tests/mgpu_parity.py c5 2>&1 | grep MGPU
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/r02_bench_c5_${N}gpu.json 2> gpurun_out/r02_bench_c5_${N}gpu.err
tail -2 gpurun_out/r02_bench_c5_${N}gpu.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_c5_${N}gpu.json').read().strip().splitlines()[-1])
    print('c5 ${N} GPUs', round(d['value'],1), round(d['ms_per_step'],2), d['roofline']['whole_subiteration']['frac'], d['config'].get('halo'), d['roofline']['step_share'])
except Exception as e: print('ERR', e)
PY
