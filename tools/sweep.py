"""Throughput of every BASELINE.json configuration (synthetic data, seeded).  One JSON line per measurement.

    python tools/sweep.py single                      # C1, C2 (VQ layer part), C3 sweep on one GPU
    torchrun --nproc-per-node G tools/sweep.py multi  # C4 (K-sharded, K = 2^20) and C5 (batch-sharded) on G GPUs
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kindergarten_vq_vae_b200 as kvq  # noqa: E402
from kindergarten_vq_vae_b200 import _lib  # noqa: E402
import ctypes  # noqa: E402

BETA = 0.25


def timed(fn, warm=3, iters=10, sync=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if sync:
        sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def make(dev, N, D, K, init, seed=69):
    g = torch.Generator(device=dev).manual_seed(seed)
    z = torch.randn(N, D, device=dev, generator=g)
    gz = torch.randn(N, D, device=dev, generator=g)
    if init == "default":
        E = (torch.rand(K, D, device=dev, generator=g) * 2 - 1) / K
    else:
        E = torch.randn(K, D, device=dev, generator=g) + 0.1 * torch.randn(K, D, device=dev, generator=g)
    return z, gz, E


def profile(lib, fn, iters=5):
    lib.kvq_profile_enable(1)
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    lib.kvq_profile_enable(0)
    ms = (ctypes.c_double * 6)(); cnt = (ctypes.c_int * 6)()
    lib.kvq_profile_collect(ms, cnt, 6)
    return {t: (ms[i] / cnt[i] if cnt[i] else None) for i, t in enumerate(_lib.PROF_TAGS)}


def single():
    dev = torch.device("cuda:0")
    lib = _lib.load()
    cases = [("C1", 64, 64, 768, 512), ("C2-B512", 512, 12, 768, 512), ("C2-B2048", 2048, 12, 768, 512)]
    cases += [(f"C3-K{K}", 16384, 64, 256, K) for K in (8192, 16384, 32768, 65536)]
    for name, B, S, D, K in cases:
        for init in ("default", "data"):
            N = B * S
            z, gz, E = make(dev, N, D, K, init)
            z3 = z.view(B, S, D).requires_grad_(True)
            g3 = gz.view(B, S, D)
            vq = kvq.VectorQuantizer(K, D, BETA, vq_codebook_init_values=E, min_encodings=False).to(dev)   # default search mode
            one = torch.ones((), device=dev)

            def step():
                z3.grad = None; vq.embedding.weight.grad = None
                loss, z_q, perp, _, idx = vq.forward(z3, dev)
                torch.autograd.backward([loss, z_q], [one, g3])
            ms = timed(step, iters=10 if N >= (1 << 18) else 50)
            prof = profile(lib, step)
            out = dict(config=name, N=N, D=D, K=K, init=init, search=vq.search, ms_fwd_bwd=ms, latents_per_s=N / ms * 1e3, kernel_ms=prof)
            if prof["search"]:
                out["search_tflops"] = 2.0 * N * K * D / prof["search"] / 1e9
            if N <= 32768:   # small shapes are launch-bound: the same step as ONE CUDA-graph launch
                side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    step()
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                z3.grad = None; vq.embedding.weight.grad = None
                with torch.cuda.graph(graph):
                    step()
                out["ms_fwd_bwd_cuda_graph"] = timed(graph.replay, iters=200)
            print(json.dumps(out), flush=True)
            del vq, z, gz, E, z3, g3


def multi():
    import torch.distributed as dist
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    lib = _lib.load()
    one = torch.ones((), device=dev)

    def barrier():
        dist.barrier(); torch.cuda.synchronize()

    # ---- C4: codebook-sharded, K = 2^20, latents replicated ----
    D, K = 256, 1 << 20
    for N, exchange in ((1 << 18, "nccl"), (1 << 20, "nccl"), (1 << 18, "nvlink"), (1 << 20, "nvlink")):
        g = torch.Generator(device=dev).manual_seed(69)                 # same z on every rank
        z = torch.randn(N, D, device=dev, generator=g); gz = torch.randn(N, D, device=dev, generator=g)
        per = (K + world - 1) // world
        ge = torch.Generator(device=dev).manual_seed(1000 + rank)       # this rank's codebook rows
        vq = kvq.CodebookShardedVectorQuantizer(K, D, BETA, exchange=exchange).to(dev)
        with torch.no_grad():
            vq.embedding.weight.copy_(torch.randn(per, D, device=dev, generator=ge))
        z3 = z.view(N // 64, 64, D).requires_grad_(True); g3 = gz.view_as(z3)

        def step():
            z3.grad = None; vq.embedding.weight.grad = None
            loss, z_q, perp, _, idx = vq.forward(z3, dev)
            torch.autograd.backward([loss, z_q], [one, g3])
        ms = timed(step, warm=2, iters=5, sync=barrier)
        t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        prof = profile(lib, step, iters=2)
        if rank == 0:
            ms = float(t)
            print(json.dumps(dict(config="C4-kshard", exchange=exchange, gpus=world, N=N, D=D, K=K, ms_fwd_bwd=ms, latents_per_s=N / ms * 1e3,
                                  search_ms=prof["search"], search_tflops_per_gpu=2.0 * N * per * D / prof["search"] / 1e9,
                                  aggregate_tflops=2.0 * N * K * D / ms / 1e9)), flush=True)
        del vq, z, gz, z3, g3
        torch.cuda.empty_cache()

    # ---- C5: batch-sharded, K = 65536, weak (N = G * 2^20) and strong (N = 2^20) scaling ----
    K = 65536
    for mode, n_local, exchange in (("weak", 1 << 20, "nccl"), ("strong", (1 << 20) // world, "nccl"),
                                    ("weak", 1 << 20, "nvlink"), ("strong", (1 << 20) // world, "nvlink")):
        g = torch.Generator(device=dev).manual_seed(69 + rank)
        z = torch.randn(n_local, D, device=dev, generator=g); gz = torch.randn(n_local, D, device=dev, generator=g)
        ge = torch.Generator(device=dev).manual_seed(7)
        E = torch.randn(K, D, device=dev, generator=ge)
        vq = kvq.BatchShardedVectorQuantizer(K, D, BETA, vq_codebook_init_values=E, exchange=exchange).to(dev)
        z3 = z.view(n_local // 64, 64, D).requires_grad_(True); g3 = gz.view_as(z3)

        def step():
            z3.grad = None; vq.embedding.weight.grad = None
            loss, z_q, perp, _, idx = vq.forward(z3, dev)
            torch.autograd.backward([loss, z_q], [one, g3])
        ms = timed(step, warm=3, iters=10, sync=barrier)
        t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms = float(t)
            print(json.dumps(dict(config=f"C5-dp-{mode}", exchange=exchange, gpus=world, N_total=n_local * world, D=D, K=K, ms_fwd_bwd=ms,
                                  latents_per_s=n_local * world / ms * 1e3)), flush=True)
        del vq, z, gz, E, z3, g3
        torch.cuda.empty_cache()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    (single if sys.argv[1] == "single" else multi)()
