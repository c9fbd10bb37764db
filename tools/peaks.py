"""Measure library GEMM peaks on this GPU the way MEASURED_PEAKS.json was measured (torch.matmul 8192^3, best of 10
and a 4 s back-to-back loop), for bf16 and tf32.  tf32 is the denominator of the search kernel's roofline."""
import json, sys, time
import torch
dev = "cuda:0"
out = {}
for name, dtype, tf32 in (("bf16", torch.bfloat16, False), ("tf32", torch.float32, True), ("fp32", torch.float32, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    n = 8192
    a = torch.randn(n, n, device=dev, dtype=dtype); b = torch.randn(n, n, device=dev, dtype=dtype)
    for _ in range(3): a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = 2 * n ** 3 / best / 1e9
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(10, int((0.5 if name == "fp32" else 4.0) * 1e3 / best))
    e0.record()
    for _ in range(reps): a @ b
    e1.record(); torch.cuda.synchronize()
    sus = 2 * n ** 3 * reps / e0.elapsed_time(e1) / 1e9
    out[name] = dict(burst_tflops=burst, sustained_tflops=sus, reps=reps)
print(json.dumps(out))
