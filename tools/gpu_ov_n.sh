#!/bin/bash
# overlap A/B of the default multi-GPU bench at N ranks (no pytest)
N=${1:-4}; mkdir -p gpurun_out
for ov in 1 0; do
FAST_B200_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$ov bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e 2> gpurun_out/ovn_$ov.err | tee gpurun_out/bench_n${N}_ov$ov.json | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('N=$N overlap=$ov',round(d['value'],1),round(d['ms_per_step'],3),{k:round(v*d['ms_per_step'],2) for k,v in d['roofline']['step_share'].items()})"
done
