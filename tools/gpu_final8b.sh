#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 tests/dist_gpu_check.py > gpurun_out/dist_check_$NG.log 2> gpurun_out/dist_check_$NG.err; echo "dist check rc=$?"
grep -E "dist check|exchange" gpurun_out/dist_check_$NG.log; tail -3 gpurun_out/dist_check_$NG.err | cut -c1-300
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29566 tools/sweep.py multi > gpurun_out/sweep_multi_$NG.jsonl 2> gpurun_out/sweep_multi_$NG.err; echo "sweep multi $NG rc=$?"
grep "^{" gpurun_out/sweep_multi_$NG.jsonl | cut -c1-260; tail -3 gpurun_out/sweep_multi_$NG.err | cut -c1-300
