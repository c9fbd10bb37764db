"""Developer probe for the GPU box: correctness + timing of the search kernels against a torch fp64 argmin.
Each case runs in its own subprocess under a timeout so that a faulting kernel cannot take the rest down.

    python tools/probe.py            # all cases
    python tools/probe.py one <mode> <N> <D> <K> <init>
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(mode, N, D, K, init):
    import torch
    import kindergarten_vq_vae_b200 as kvq
    F = kvq.functional
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(69)
    z = torch.randn(N, D, device=dev, generator=g)
    if init == "default":
        E = (torch.rand(K, D, device=dev, generator=g) * 2 - 1) / K
    else:
        E = torch.randn(K, D, device=dev, generator=g)
    ws = F.workspace(N, D, K, dev)
    idx, _ = F.search(z, E, mode=mode, ws=ws)
    torch.cuda.synchronize()
    # truth on a row sample in fp64
    rows = torch.arange(0, N, max(1, N // 4096), device=dev)[:4096]
    zs = z[rows].double()
    Ed = E.double()
    e2 = (Ed * Ed).sum(1)
    best = torch.empty(rows.numel(), dtype=torch.int64, device=dev)
    gap_chosen = torch.empty(rows.numel(), dtype=torch.float64, device=dev)
    for s in range(0, rows.numel(), 512):
        d = e2 - 2.0 * zs[s:s + 512] @ Ed.t()
        best[s:s + 512] = d.argmin(1)
        gap_chosen[s:s + 512] = d.gather(1, idx[rows[s:s + 512], None]).squeeze(1) - d.min(1).values
    mism = (best != idx[rows]).float().mean().item()
    tol = 2.0 ** -9 * zs.norm(dim=1) * Ed.norm(dim=1).max()
    bad = int((gap_chosen > tol + 1e-6 * (zs * zs).sum(1)).sum())
    # timing
    for _ in range(2):
        F.search(z, E, mode=mode, ws=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        F.search(z, E, mode=mode, ws=ws)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps(dict(mode=mode, N=N, D=D, K=K, init=init, mismatch_vs_fp64=mism, beyond_tol=bad,
                          max_gap=float(gap_chosen.max()), ms=ms, tflops=2.0 * N * K * D / ms / 1e9,
                          idx_min=int(idx.min()), idx_max=int(idx.max()))),
          flush=True)


CASES = [
    ("fp32", 4096, 64, 512, "normal"),
    ("tf32", 128, 32, 256, "normal"),
    ("tf32", 4096, 64, 512, "normal"),
    ("tf32", 4096, 256, 1024, "normal"),
    ("tf32", 8192, 768, 512, "normal"),
    ("tf32", 1 << 16, 256, 8192, "normal"),
    ("fp32", 1 << 16, 256, 8192, "normal"),
    ("tf32", 1 << 20, 256, 8192, "normal"),
    ("tf32", 1 << 20, 256, 65536, "normal"),
    ("tf32", 1 << 18, 256, 8192, "default"),
    ("fp32", 1 << 18, 256, 8192, "default"),
]

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6])
        sys.exit(0)
    configs = [dict(KVQ_TF32_CTA_GROUP="2"), dict(KVQ_TF32_CTA_GROUP="1"), dict(KVQ_TF32_CTA_GROUP="2", KVQ_TMA_ROUND_TF32="0")]
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        configs = configs[:1]
    for ci, cfg in enumerate(configs):
        for c in CASES:
            if ci > 0 and (c[0] != "tf32" or c[1] < (1 << 16)):
                continue
            env = dict(os.environ, **cfg)
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, __file__, "one", *map(str, c)], env=env, capture_output=True,
                                   text=True, timeout=180)
                tail = (r.stdout.strip().splitlines() or [""])[-1]
                if r.returncode != 0:
                    tail = f"FAILED rc={r.returncode} case={c} :: " + (r.stdout + r.stderr)[-1500:]
            except subprocess.TimeoutExpired:
                tail = f"TIMEOUT case={c}"
            print(json.dumps(cfg), tail, f"[{time.time() - t0:.1f}s]", flush=True)
