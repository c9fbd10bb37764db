#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python tests/harness_shelgon_step.py > gpurun_out/shelgon.jsonl 2> gpurun_out/shelgon.err; echo "shelgon rc=$?"; cat gpurun_out/shelgon.jsonl | cut -c1-1200; tail -5 gpurun_out/shelgon.err
