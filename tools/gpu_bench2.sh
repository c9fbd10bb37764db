#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/bench_n$NG.log 2> gpurun_out/bench_n$NG.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_n$NG.log').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'], d['index_parity'])
" || grep -v "^\s*$" gpurun_out/bench_n$NG.err | head -20 | cut -c1-300
