#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
if [ "$1" == "single" ]; then
  timeout 1200 python tools/sweep.py single > gpurun_out/sweep_single.jsonl 2> gpurun_out/sweep_single.err; echo "single rc=$?"
  python -c "
import json
for l in open('gpurun_out/sweep_single.jsonl'):
    d=json.loads(l); print(d['config'],d['init'],'N',d['N'],'K',d['K'],'ms',round(d['ms_fwd_bwd'],3),'Mlat/s',round(d['latents_per_s']/1e6,3),'TF',round(d.get('search_tflops',0),1))
"; tail -3 gpurun_out/sweep_single.err
else
  for N in 2 4 8; do
    if [ $N -le $NG ]; then
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29566 tools/sweep.py multi > gpurun_out/sweep_multi_$N.jsonl 2> gpurun_out/sweep_multi_$N.err; echo "multi $N rc=$?"
      cat gpurun_out/sweep_multi_$N.jsonl; tail -3 gpurun_out/sweep_multi_$N.err
    fi
  done
fi
