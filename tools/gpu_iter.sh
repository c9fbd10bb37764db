#!/bin/bash
# quick iteration: probe (default config only) -> gpu tests -> bench
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python tools/probe.py quick > gpurun_out/probe.log 2>&1; echo "probe rc=$?"
grep -E "tf32|FAILED|TIMEOUT" gpurun_out/probe.log | python -c "
import sys,json
for l in sys.stdin:
    try:
        i=l.index('} {')+2; d=json.loads(l[i:l.rindex('}')+1]); print(d['mode'],d['N'],d['D'],d['K'],d['init'],'mism',round(d['mismatch_vs_fp64'],5),'bad',d['beyond_tol'],'ms',round(d['ms'],3),'TF',round(d['tflops'],1))
    except Exception as e: print(l.strip()[:300])
"
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','kernel_ms','clocks']}); print(d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'])
"
