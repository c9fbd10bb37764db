#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "refine" > gpurun_out/pytest_refine.log 2>&1; echo "pytest rc=$?"; grep -E "refine\[|passed|failed|Error|assert" gpurun_out/pytest_refine.log | head -20 | cut -c1-250
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
import kindergarten_vq_vae_b200 as kvq
F = kvq.functional
dev = 'cuda:0'
g = torch.Generator(device=dev).manual_seed(69)
N, D, K = 1 << 20, 256, 65536
z = torch.randn(N, D, device=dev, generator=g); E = torch.randn(K, D, device=dev, generator=g)
ws = F.workspace(N, D, K, dev)
for mode in ('tf32', 'tf32_refine'):
    for _ in range(2): F.search(z, E, mode=mode, ws=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): idx, _ = F.search(z, E, mode=mode, ws=ws)
    e1.record(); torch.cuda.synchronize()
    print(mode, 'ms', e0.elapsed_time(e1) / 5)
PY
