#!/bin/bash
# single-GPU validation + evidence: gpu tests, smoke, bench, ncu launch list, ncu --set full of the three main kernels
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.log
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-cublas"
$CMD > gpurun_out/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/bench_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"search_tf32|quantize_kernel|segmented_backward" -s 9 -c 3 -o gpurun_out/final_kernels $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
python -c "
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','kernel_ms','clocks','gpu_launches','index_parity']}); print(d['roofline']); print(d['e2e']); print(d['cpu_baseline'])
"
