#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "graph" > gpurun_out/pytest_graph.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_graph.log | cut -c1-250
timeout 600 python tools/sweep.py single > gpurun_out/sweep_single.jsonl 2> gpurun_out/sweep_single.err; echo "single rc=$?"
python -c "
import json
for l in open('gpurun_out/sweep_single.jsonl'):
    d=json.loads(l); print(d['config'],d['init'],'N',d['N'],'K',d['K'],'ms',round(d['ms_fwd_bwd'],3),'graph',d.get('ms_fwd_bwd_cuda_graph'),'Mlat/s',round(d['latents_per_s']/1e6,3),'TF',round(d.get('search_tflops',0),1))
"; tail -3 gpurun_out/sweep_single.err
