#!/bin/bash
# multi-GPU: NCCL parity test + scaling bench at N = 1 .. all GPUs
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -5 gpurun_out/pytest_multi.log
for N in 1 2 4 8; do
  if [ $N -le $NG ]; then
    if [ $N -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
    fi
    echo "bench N=$N rc=$?"
    python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/bench_n$N.log').read().strip().splitlines()[-1])
    print('N=$N', 'value', d['value'], 'ms', d['ms_per_step'], 'search', d['kernel_ms']['search'], 'e2e', d['e2e'] and d['e2e']['value'], d['clocks'])
except Exception as e:
    print('parse fail', e); print(open('gpurun_out/bench_n$N.err').read()[-1500:])
"
  fi
done
