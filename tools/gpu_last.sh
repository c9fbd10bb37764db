#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
PCMD="python tools/probe.py one tf32_refine 1048576 256 65536 normal"
$PCMD > gpurun_out/probe_refine_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:search_tf32 -s 1 -c 1 -o gpurun_out/search_top2 $PCMD > gpurun_out/ncu_top2.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/probe_refine_plain.log | cut -c1-300
