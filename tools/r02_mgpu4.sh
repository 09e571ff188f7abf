#!/bin/bash
mkdir -p gpurun_out
N=${N:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29657 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/r02_bench_c5_${N}gpu.json 2> gpurun_out/r02_bench_c5_${N}gpu.err
tail -2 gpurun_out/r02_bench_c5_${N}gpu.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_c5_${N}gpu.json').read().strip().splitlines()[-1])
    print('c5 ${N} GPUs', round(d['value'],1), round(d['ms_per_step'],2), d['roofline']['whole_subiteration']['frac'], d['config'].get('halo'), d['roofline']['step_share'], d['config']['workload'])
except Exception as e: print('ERR', e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29658 tests/mgpu_parity.py c5 2>&1 | grep MGPU
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29659 tests/mgpu_parity.py tgv 2>&1 | grep MGPU
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29660 bench.py --impl reference --gpus $N --steps 2 --warmup 1 --ref-budget 20 2>/dev/null | cut -c1-200
