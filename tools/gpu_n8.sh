#!/bin/bash
# N-GPU lines: C5 (BASELINE configs[4]; 8 GPUs = the 1024x1024x256 plate in 64 blocks) and the default TGV family
N=${1:-8}
mkdir -p gpurun_out
run() { timeout ${T:-400} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err
  python -c "
import json;d=json.loads(open('gpurun_out/$2.json').read().strip().splitlines()[-1]);r=d['roofline'];print('$2',d['n_gpus'],round(d['value'],1),round(d['ms_per_step'],2),r['kernel'],{k:round(v*d['ms_per_step'],2) for k,v in r['step_share'].items()})" || tail -5 gpurun_out/$2.err; }
free -g | head -2
run 29711 bench_c5_n$N --workload c5 --steps 4 --warmup 2 --no-cpu --no-e2e
run 29712 bench_tgv_n$N --steps 10 --warmup 3 --no-cpu --no-e2e
