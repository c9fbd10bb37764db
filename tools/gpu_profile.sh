#!/bin/bash
# peaks + ncu evidence.  One ncu family per call (launch list + one --set full capture).
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
python tools/peaks.py > gpurun_out/peaks.json 2> gpurun_out/peaks.err; cat gpurun_out/peaks.json
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
PCMD="python tools/probe.py one tf32 262144 256 65536 normal"
$PCMD > gpurun_out/probe_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:search_tf32 -s 1 -c 1 -o gpurun_out/search_tf32 $PCMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
