#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi.log | cut -c1-300
for N in 1 2 4 8; do
  if [ $N -le $NG ]; then
    if [ $N -eq 1 ]; then
      timeout 400 python bench.py --gpus 1 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err
    else
      timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
    fi
    echo "bench N=$N rc=$?"
    python -c "
import json
try:
    d=json.loads(open('gpurun_out/bench_n$N.log').read().strip().splitlines()[-1])
    print('N=$N', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'search', round(d['kernel_ms']['search'],3), 'e2e', d['e2e'] and round(d['e2e']['value']), d['clocks']['sm_mhz'], d['index_parity'] and d['index_parity']['beyond_tf32_tolerance'])
except Exception as e:
    print('parse fail', e); print(open('gpurun_out/bench_n$N.err').read()[-800:])
"
  fi
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29566 tools/sweep.py multi > gpurun_out/sweep_multi_$NG.jsonl 2> gpurun_out/sweep_multi_$NG.err; echo "sweep multi $NG rc=$?"
grep "^{" gpurun_out/sweep_multi_$NG.jsonl | cut -c1-330; tail -3 gpurun_out/sweep_multi_$NG.err | cut -c1-300
