#!/usr/bin/env python
"""bench.py -- VQ latents/sec (forward + backward) at K=65536, D=256 on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl kvq|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic latents: the drop-in VectorQuantizer's forward
(code norms, fused tf32 distance+argmin, gather + straight-through + loss + histogram, loss/perplexity) and its
backward (dz fused with the bucketed scatter-add into dE).  Workload: BASELINE.json's metric configuration --
2^20 latents x 256 dims against a 65536-code book per GPU (weak scaling: batch-sharded latents, replicated
codebook, codebook gradient + histogram + loss partial all-reduced over NCCL).

One JSON line is printed by rank 0 (see README / DESIGN.md for the field meanings).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "VQ latents/sec (fwd+bwd) at K=65536,D=256"
UNIT = "latents/s"
N_PER_GPU = 1 << 20
D = 256
K = 65536
BETA = 0.25
SEED = 69  # the reference's DS_GEN_SEED (common/consts.py:3)
CPU_SAMPLE_ROWS = 4096
PARITY_ROWS = 8192          # rows of the measured batch checked against the oracle's reference-order fp32 argmin
REF_CUDA_ROWS = 16384       # "same box" bar: the reference's own torch ops on the B200 at the largest N that fits

# dram__bytes_read.sum + dram__bytes_write.sum per launch at the headline shape, from the `ncu --set full` captures
# under profiles/ (file named per entry); None = not captured for this build.
NCU_TRAFFIC = {
    "search": (3.912e9, "profiles/r02/ncu_full_kernels.summary.csv"),         # search_tf32_kernel<2, TOP2>
    "quantize": (2.343e9, "profiles/r02/ncu_full_kernels.summary.csv"),       # quantize_refine_kernel<2>
    "bwd_segmented": (3.290e9, "profiles/r02/ncu_full_kernels.summary.csv"),  # segmented_kernel<2, false>
}


def peaks():
    """Roofline denominators: driver-measured numbers if present, else the profiling guide's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                    bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML, else nvidia-smi)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _reasons(self, mask: int):
        n = self._nvml
        table = [("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"),
                 ("hw_power_brake", "nvmlClocksThrottleReasonHwPowerBrakeSlowdown")]
        for name, attr in table:
            bit = getattr(n, attr, None)
            if bit is not None and mask & bit:
                self.reasons.add(name)

    def _loop(self):
        n = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self._reasons(mask)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self._nvml is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        if not self.samples:
            return self._smi_once()
        return dict(sm_mhz=statistics.median(self.samples), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(self.samples))

    def _smi_once(self):
        import subprocess
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            a, b = [float(x) for x in out.strip().split(",")]
            return dict(sm_mhz=a, sm_max_mhz=b, reasons=[], samples=1, note="single idle nvidia-smi sample")
        except Exception as exc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0, note=f"no clock source: {exc}")


def bind_to_gpu_numa_node(index: int):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host buffers allocated
    afterwards are node-local (one process per GPU: without this every rank's staging memory may sit on socket 0
    and all host<->device copies cross the inter-socket link).  Measurement plumbing; returns a description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # 00000000:1B:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return f"gpu {index} ({bus}): no NUMA affinity reported"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return f"gpu {index} ({bus}): node {node} has no allowed CPUs"
        os.sched_setaffinity(0, allowed)
        return f"gpu {index} ({bus}) -> NUMA node {node}, {len(allowed)} CPUs"
    except Exception as exc:                      # no NVML / sysfs: leave the default placement
        return f"not bound ({exc})"


# ---------------------------------------------------------------------------------------------------------
def synth_device(torch, dev, n_rows, seed):
    """Seeded synthetic inputs: z, g_zq ~ N(0,1); data-scale codebook = K latents + noise
    (mirrors the reference's k-means init with minit='points', vq_codebook_init_weights.py:85)."""
    gen = torch.Generator(device=dev).manual_seed(seed)
    z = torch.randn(n_rows, D, device=dev, generator=gen)
    gz = torch.randn(n_rows, D, device=dev, generator=gen)
    gen_e = torch.Generator(device=dev).manual_seed(SEED)       # same codebook on every rank
    base = torch.randn(K, D, device=dev, generator=gen_e)
    E = base + 0.1 * torch.randn(K, D, device=dev, generator=gen_e)
    return z, gz, E


def cpu_reference_rate(torch, steps, warmup, rows):
    """Literal CPU port of the reference step (oracle.vq_oracle.literal_step_cpu) on a bounded row sample."""
    from oracle import vq_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(SEED)
    z = torch.randn(rows // 64, 64, D, generator=gen)
    gz = torch.randn(rows // 64, 64, D, generator=gen)
    E = torch.randn(K, D, generator=gen) + 0.1 * torch.randn(K, D, generator=gen)
    for _ in range(warmup):
        O.literal_step_cpu(z, E, BETA, gz)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.literal_step_cpu(z, E, BETA, gz)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return rows / dt, dt, torch.get_num_threads()


def reference_on_cuda(torch, dev):
    """The reference layer's own PyTorch ops (oracle.vq_oracle.LiteralVectorQuantizer = VectorQuantizer.py:55-93 restated,
    host-built one-hot and all) forward + backward on the GPU, fp32 and with allow_tf32, at the largest N whose dense
    N x K temporaries fit comfortably.  A reported bar, not part of any timed kvq region."""
    from oracle import vq_oracle as O
    rows = REF_CUDA_ROWS
    gen = torch.Generator(device=dev).manual_seed(SEED)
    z = torch.randn(rows // 64, 64, D, device=dev, generator=gen).requires_grad_(True)
    gz = torch.randn(rows // 64, 64, D, device=dev, generator=gen)
    E = torch.randn(K, D, device=dev, generator=gen)
    one = torch.ones((), device=dev)
    out = {"latents_per_step": rows, "K": K, "D": D,
           "what": "oracle.vq_oracle.LiteralVectorQuantizer (the reference's tensor expressions) on cuda, fwd + bwd"}
    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        for name, flag in (("fp32", False), ("allow_tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = flag
            vq = O.LiteralVectorQuantizer(K, D, BETA, vq_codebook_init_values=E).to(dev)

            def it():
                z.grad = None
                vq.embedding.weight.grad = None
                loss, zq, *_ = vq.forward(z, dev)
                torch.autograd.backward([loss, zq], [one, gz])
            for _ in range(2):
                it()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                it()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            out[name] = {"ms_per_step": ms, "latents_per_s": rows / (ms * 1e-3)}
            del vq
    except Exception as exc:                      # e.g. out of memory on a smaller part: report, do not fail the bench
        out["error"] = f"{type(exc).__name__}: {exc}"[:300]
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
        torch.cuda.empty_cache()
    return out


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = CPU_SAMPLE_ROWS
    rate, dt, threads = cpu_reference_rate(torch, args.steps, args.warmup, rows)
    sample = (f"{rows} latents x D={D} against the full K={K} codebook per step (the reference's dense N x K fp32 "
              f"temporaries are {rows * K * 4 / 2**30:.2f} GiB each; cost is linear in N)")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"VQ fwd+bwd, K={K}, D={D}, CPU port of models/shelgon3/VectorQuantizer.py on a "
                               f"{rows}-latent sample per step", "K": K, "D": D, "latents_per_step": rows},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
def run_kvq(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- kvq has no CPU path (use --impl reference for the CPU arm)")
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "single process: default placement"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    import kindergarten_vq_vae_b200 as kvq
    from kindergarten_vq_vae_b200 import _lib
    F = kvq.functional
    lib = _lib.load()

    n_rows = N_PER_GPU
    z, gz, E = synth_device(torch, dev, n_rows, SEED + 1 + rank)
    z3 = z.view(n_rows // 64, 64, D).requires_grad_(True)     # (B=16384, S=64, D): the module's (B,S,D) input
    gz3 = gz.view_as(z3)
    one = torch.ones((), device=dev)
    if world > 1:
        vq = kvq.BatchShardedVectorQuantizer(K, D, BETA, vq_codebook_init_values=E).to(dev)
    else:
        vq = kvq.VectorQuantizer(K, D, BETA, vq_codebook_init_values=E, min_encodings=False).to(dev)
    search_mode = vq.search          # the module's DEFAULT: "auto" (= tf32 top-2 search + exact re-evaluation of the pair)
    last = {}

    def step(module=None):
        m = vq if module is None else module
        z3.grad = None
        m.embedding.weight.grad = None
        loss, z_q, perp, _, idx = m.forward(z3, dev)
        torch.autograd.backward([loss, z_q], [one, gz3])
        last["idx"] = idx
        return loss, perp

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()

    # ---- timed region: exactly K steps, CUDA events on the launching stream, max over ranks -------------
    lib.kvq_profile_enable(1)
    launches0 = lib.kvq_launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the barrier comes LAST before the start event: anything a rank does between the barrier and its first launch
    # (starting the clock sampler took a rank-dependent 5-15 ms) is charged to the other ranks, which wait for it in the
    # first collective, and the max over ranks would report that skew as step time
    sync_all()
    e0.record()
    for _ in range(args.steps):
        loss, perp = step()
    e1.record()
    sync_all()
    clocks = sampler.stop()
    total_ms = e0.elapsed_time(e1)
    launches = lib.kvq_launch_count() - launches0
    lib.kvq_profile_enable(0)
    ms = (ctypes.c_double * 6)()
    cnt = (ctypes.c_int * 6)()
    _lib.check(lib.kvq_profile_collect(ms, cnt, 6), "kvq_profile_collect")
    prof = {tag: (ms[i] / cnt[i] if cnt[i] else None) for i, tag in enumerate(_lib.PROF_TAGS)}
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = n_rows * world / (ms_per_step * 1e-3)

    # ---- where the multi-GPU step time goes: per-rank spread (the headline is the MAX over ranks) and the bare cost of
    #      the one large collective (all-reduce of the 64 MiB codebook gradient) ----------------------------------------
    dp_breakdown = None
    if world > 1:
        mine = torch.tensor([total_ms / args.steps, prof["search"] or 0.0], device=dev, dtype=torch.float64)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        buf = torch.empty(K, D, device=dev)
        for _ in range(2):
            dist.all_reduce(buf)
        torch.cuda.synchronize(); dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(5):
            dist.all_reduce(buf)
        a1.record(); torch.cuda.synchronize()
        dp_breakdown = {"ms_per_step_by_rank": [float(x[0]) for x in allr], "search_ms_by_rank": [float(x[1]) for x in allr],
                        "allreduce_dE_ms": a0.elapsed_time(a1) / 5, "allreduce_bytes": K * D * 4,
                        "note": "ms_per_step is the max over ranks: the spread between GPUs of the power-capped search kernel "
                                "and the unoverlapped dE all-reduce are the two terms that separate N-GPU from 1-GPU step time"}
        del buf

    idx_timed = last["idx"].reshape(-1).clone()       # the indices the LAST TIMED step produced (default search mode)
    dE_timed = vq.embedding.weight.grad.detach().clone()

    # ---- end to end with host buffers (H2D + compute + D2H inside the timed region) --------------------
    # Measured right after the headline.  The call is bound by the box's host <-> device copy path, which varies between
    # boxes and between moments on one box (traced calls, tools/e2e_trace_stats.py: the inbound copies alone end after
    # 46 ms on one box and after 54 ms on another, with compute never more than 0.5 ms behind them), so the bare-copy
    # ceiling is measured directly after it.
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(torch, dist, F, vq, z, gz, E, dev, world, args)
        e2e["bare_copy_ceiling"] = measure_copy_ceiling(torch, dist, dev, world, e2e["h2d_bytes_per_step"],
                                                        e2e["d2h_bytes_per_step"], n_rows)

    # ---- result check on the measured configuration: sampled rows of the timed step's own output against
    #      (a) the oracle = the reference's fp32 evaluation order on the CPU, (b) an fp64 argmin on the device
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import vq_oracle as O              # checker only, outside every timed region
        with torch.no_grad():
            rows = torch.arange(0, n_rows, n_rows // PARITY_ROWS, device=dev)[:PARITY_ROWS]
            zs_cpu, E_cpu, ours = z[rows].cpu(), E.cpu(), idx_timed[rows].cpu()
            t0 = time.perf_counter()
            ref = O.forward_fp32(zs_cpu.view(-1, 64, D), E_cpu, BETA, row_chunk=1024)
            par = O.index_parity(ours, ref.idx, zs_cpu, E_cpu)
            oracle_s = time.perf_counter() - t0
            zs = z[rows].double(); Ed = E.double()
            d = (Ed * Ed).sum(1) - 2.0 * zs @ Ed.t()
            best = d.argmin(1)
            chosen = d.gather(1, idx_timed[rows, None]).squeeze(1)
            gap = chosen - d.min(1).values
            tol = 2.0 ** -9 * zs.norm(dim=1) * Ed.norm(dim=1).max()
            parity = {"search": search_mode, "rows_checked": int(rows.numel()),
                      "vs_oracle_reference_order_fp32": {"raw_mismatch": par.raw_mismatch, "raw_rate": par.raw_rate,
                                                         "unexcused": par.unexcused,
                                                         "max_gap_over_tolerance": par.max_gap_over_tol,
                                                         "oracle_seconds": oracle_s},
                      "index_mismatch_vs_fp64": int((best.cpu() != ours).sum()),
                      "reference_order_fp32_mismatch_vs_fp64": int((best.cpu() != ref.idx.reshape(-1)).sum()),
                      "beyond_tf32_tolerance_vs_fp64": int((gap > tol).sum()),
                      "tolerance": "oracle.vq_oracle.tf32_tolerance: 2^-9 |z_i| max_k|E_k| + 4 ulp32(d) on the fp64 "
                                   "squared-distance gap (DESIGN.md section 3)"}
            del d, zs, Ed

    # ---- batch-sharded run: the all-reduced dE of the timed step against an fp64 sum of every rank's contribution
    dp_check = None
    if world > 1 and not args.no_parity:
        with torch.no_grad():
            codes = torch.arange(0, K, K // 64, device=dev)[:64]
            c2 = 2.0 * BETA / (float(n_rows) * world * D)              # g_loss = 1, normalised by the GLOBAL N * D
            contrib = torch.zeros(codes.numel(), D, dtype=torch.float64, device=dev)
            for j, k in enumerate(codes.tolist()):
                rows_k = (idx_timed == k).nonzero().flatten()
                if rows_k.numel():
                    contrib[j] = (E[k].double()[None, :] - z[rows_k].double()).sum(0) * c2
            dist.all_reduce(contrib, op=dist.ReduceOp.SUM)
            err = float((dE_timed[codes].double() - contrib).abs().max())
            scale = float(contrib.abs().max())
            dp_check = {"codes_checked": int(codes.numel()), "max_abs_err": err, "max_abs_ref": scale,
                        "rel_err": err / max(scale, 1e-30), "tolerance_rel": 1e-5,
                        "what": "all-reduced dE rows of the last timed step vs an fp64 sum of all ranks' contributions"}
            assert dp_check["rel_err"] <= 1e-5, dp_check

    # ---- plain tf32 search (no exact re-evaluation) timed beside the default, same inputs, outside the headline
    refine = None
    if rank == 0 and world == 1 and not args.no_side:
        vq_t = kvq.VectorQuantizer(K, D, BETA, vq_codebook_init_values=E, search="tf32", min_encodings=False).to(dev)
        for _ in range(2):
            step(vq_t)
        torch.cuda.synchronize()
        ab = {"tf32": [], "auto": []}
        for _rep in range(3):                 # interleaved blocks of 4 steps: both modes see the same thermal / power state
            for name, module in (("tf32", vq_t), ("auto", vq)):
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                r0.record()
                for _ in range(4):
                    step(module)
                r1.record(); torch.cuda.synchronize()
                ab[name].append(r0.elapsed_time(r1) / 4)
                if name == "tf32":
                    idx_t = last["idx"].reshape(-1).clone()
        t_ms, a_ms = statistics.mean(ab["tf32"]), statistics.mean(ab["auto"])
        refine = {"ms_per_step_search_tf32": t_ms, "ms_per_step_search_auto": a_ms,
                  "overhead_of_the_default_mode_pct": 100.0 * (a_ms - t_ms) / t_ms, "blocks_ms": ab,
                  "rows_changed_by_the_exact_pass": int((idx_t != idx_timed).sum()),
                  "note": "interleaved A/B outside the headline (3 x [4 steps search='tf32', 4 steps search='auto']): "
                          "'auto' adds top-2 tracking in the search epilogue and the exact float64 re-evaluation fused into "
                          "the gather kernel; the headline above times 'auto'"}
        del vq_t

    # ---- BASELINE.json configs[3]: K = 2^20 codebook sharded over the ranks, latents replicated ---------------------
    kshard = None
    if world > 1 and not args.no_kshard:
        del vq
        torch.cuda.empty_cache()
        kshard = measure_kshard(torch, dist, kvq, lib, _lib, dev, rank, world, args)

    # ---- rooflines -----------------------------------------------------------------------------------------
    pk = peaks()
    sustained = args.steps * ms_per_step > 1000.0
    tf32_peak = (pk["bf16_tflops_sustained"] if sustained else pk["bf16_tflops"]) / 2.0
    # library tf32 GEMM on this GPU, measured live (torch.matmul 8192^3, best of 5): context for the search kernel
    cublas_tf32 = None
    if rank == 0 and not args.no_cublas:
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(8192, 8192, device=dev); b = torch.randn(8192, 8192, device=dev)
        for _ in range(2):
            a @ b
        best = 1e9
        for _ in range(5):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(); a @ b; c1.record(); torch.cuda.synchronize()
            best = min(best, c0.elapsed_time(c1))
        cublas_tf32 = 2 * 8192 ** 3 / best / 1e9
        torch.backends.cuda.matmul.allow_tf32 = prev
        del a, b
    def hbm_roof(tag, algo_bytes, note=None):
        """HBM-bound kernel: `achieved` = algorithmic bytes / CUDA-event time; `frac_dram` uses the bytes ncu measured."""
        t = prof[tag] * 1e-3
        ach = algo_bytes / t / 1e9
        traffic, src = NCU_TRAFFIC.get(tag, (None, None)) if headline_shape else (None, None)
        r = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
             "algorithmic_bytes_per_launch": algo_bytes, "ms_per_launch": prof[tag], "traffic": traffic,
             "traffic_source": src}
        if traffic:
            r["dram_gbs"] = traffic / t / 1e9
            r["frac_dram"] = traffic / t / 1e9 / pk["hbm_gbs"]
        if note:
            r["note"] = note
        return r

    headline_shape = (n_rows == 1 << 20 and K == 65536)
    roof = None
    if prof["search"]:
        flops = 2.0 * n_rows * K * D
        ach = flops / (prof["search"] * 1e-3) / 1e12
        roof = {"kernel": "search_tf32_kernel<2, TOP2> (tcgen05 distance + argmin, keeps the two best codes)"
                          if search_mode in ("auto", "tf32_refine") else "search_tf32_kernel<2> (tcgen05 distance + argmin)",
                "bound": "tensor", "achieved": ach,
                "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak,
                "traffic": NCU_TRAFFIC["search"][0] if headline_shape else None,
                "traffic_source": NCU_TRAFFIC["search"][1] if headline_shape else None,
                "cublas_tf32_tflops_live": cublas_tf32,
                "peak_source": f"{pk['source']} bf16 dense {'sustained' if sustained else 'burst'} / 2 "
                               "(tf32 runs at half the bf16 rate; tf32 itself is not in MEASURED_PEAKS.json)",
                "algorithmic_flops_per_launch": flops, "ms_per_launch": prof["search"]}
    others = {}
    if prof["quantize"]:
        # z read + codebook row read + z_q write + idx; the default mode also reads the 8-byte top-2 word
        b = 3 * 4.0 * n_rows * D + 8.0 * n_rows + 4.0 * K + (8.0 * n_rows if search_mode in ("auto", "tf32_refine") else 0.0)
        r = hbm_roof("quantize", b, "the 4ND bytes of gathered codebook rows are L2 hits, so `achieved` (algorithmic "
                                    "bytes) can exceed the HBM peak; `frac_dram` (measured DRAM bytes / time) is the "
                                    "roofline fraction to read")
        if "frac_dram" in r:       # report the measured-traffic fraction as THE fraction, the algorithmic one as context
            r["frac_algorithmic"], r["frac"] = r["frac"], r["frac_dram"]
        others["quantize"] = r
    if prof["bwd_segmented"]:
        b = 3 * 4.0 * n_rows * D + 8.0 * n_rows + 4.0 * K * D
        others["bwd_segmented"] = hbm_roof("bwd_segmented", b)
    if prof["bwd_bucket"]:
        others["bwd_sort"] = {"ms_per_launch_group": prof["bwd_bucket"],
                              "what": "dE memset + histogram scan + stable 2-pass radix sort of (row, code) pairs"}

    # ---- the reference's own torch ops on this B200 ("same box" bar): fp32 and allow_tf32 ---------------
    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_refcuda:
        ref_cuda = reference_on_cuda(torch, dev)

    # ---- CPU baseline beside it (rank 0, single-GPU run only) ---------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, dt, threads = cpu_reference_rate(torch, 20, 5, CPU_SAMPLE_ROWS)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{CPU_SAMPLE_ROWS} latents x D={D} vs K={K}, literal port of the reference step "
                         f"(oracle.vq_oracle.literal_step_cpu), 5 warm-ups + 20 timed iterations, {dt:.2f} s each"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32 (fp32 accumulate; fp32 everywhere outside the distance GEMM; float64 for the exact top-2 re-evaluation)",
            "data": "synthetic",
            "config": {"workload": f"VQ fwd+bwd, N={n_rows} latents per GPU, D={D}, K={K}, beta={BETA}",
                       "latents_per_gpu": n_rows, "D": D, "K": K,
                       "parallelism": "single GPU" if world == 1 else f"batch-sharded x{world}, codebook replicated, "
                                      "all-reduce of dE + histogram + loss partial (NCCL)",
                       "l2": "inputs exceed L2 (z and g_zq are 1 GiB each per step; L2 is 126 MB), no flush needed",
                       "codebook_init": "data-scale: N(0,1) rows + 0.1 noise"},
            "roofline": roof, "roofline_other_kernels": others, "kernel_ms": prof,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "kshard": kshard, "dp_dE_check": dp_check, "dp_breakdown": dp_breakdown, "host_placement": numa, "clocks": clocks, "loss": float(loss.detach()), "perplexity": float(perp),
            "search_mode": search_mode, "index_parity": parity, "plain_tf32_beside": refine, "reference_cuda": ref_cuda,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


KSHARD_K = 1 << 20


def measure_copy_ceiling(torch, dist, dev, world, h2d_bytes, d2h_bytes, n_rows):
    """Bare pinned-memory copies of exactly the bytes one end-to-end step moves, both directions concurrently on two
    streams, all ranks at once: the ceiling the host side of this box puts on the end-to-end number."""
    nh, nd = h2d_bytes // 4, d2h_bytes // 4
    src_h = torch.empty(nh, dtype=torch.float32, pin_memory=True); dst_d = torch.empty(nh, dtype=torch.float32, device=dev)
    src_d = torch.empty(nd, dtype=torch.float32, device=dev); dst_h = torch.empty(nd, dtype=torch.float32, pin_memory=True)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def once():
        with torch.cuda.stream(s_in):
            dst_d.copy_(src_h, non_blocking=True)
        with torch.cuda.stream(s_out):
            dst_h.copy_(src_d, non_blocking=True)
        s_in.synchronize(); s_out.synchronize()
    once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        once()
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) / 3
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return {"ms_per_step": dt * 1e3, "latents_per_s_ceiling": n_rows * world / dt,
            "h2d_gbs_per_gpu": h2d_bytes / dt / 1e9, "d2h_gbs_per_gpu": d2h_bytes / dt / 1e9,
            "what": "cudaMemcpyAsync of the step's H2D and D2H bytes from/to pinned memory on two streams, no compute, "
                    "all ranks concurrently (max over ranks)"}


def measure_kshard(torch, dist, kvq, lib, _lib, dev, rank, world, args):
    """BASELINE.json configs[3]: K = 2^20 codes (D = 256) sharded over the ranks, N = 2^20 latents replicated; forward +
    backward of CodebookShardedVectorQuantizer with its default exchange (fused NVLink peer-memory argmin + peer gather when
    symmetric memory is available).  Index parity on sampled rows against oracle.kshard_merge."""
    Kt, n_rows = KSHARD_K, N_PER_GPU
    gen = torch.Generator(device=dev).manual_seed(SEED + 7)         # same latents on every rank
    z = torch.randn(n_rows, D, device=dev, generator=gen)
    gz = torch.randn(n_rows, D, device=dev, generator=gen)
    vq = kvq.CodebookShardedVectorQuantizer(Kt, D, BETA).to(dev)

    def shard_values(r):
        g = torch.Generator(device=dev).manual_seed(SEED + 100 + r)
        return torch.randn(vq.k_per, D, device=dev, generator=g) * 1.005
    with torch.no_grad():
        vq.embedding.weight.copy_(shard_values(rank))
    z3 = z.view(n_rows // 64, 64, D).requires_grad_(True)
    gz3 = gz.view_as(z3)
    one = torch.ones((), device=dev)
    last = {}

    def step():
        z3.grad = None
        vq.embedding.weight.grad = None
        loss, z_q, perp, _, idx = vq.forward(z3, dev)
        torch.autograd.backward([loss, z_q], [one, gz3])
        last["idx"], last["loss"], last["perp"] = idx, loss, perp
    for _ in range(2):
        step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    lib.kvq_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.kshard_steps):
        step()
    e1.record()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    lib.kvq_profile_enable(0)
    ms = (ctypes.c_double * 6)(); cnt = (ctypes.c_int * 6)()
    _lib.check(lib.kvq_profile_collect(ms, cnt, 6), "kvq_profile_collect")
    search_ms = ms[1] / cnt[1] if cnt[1] else None
    t = torch.tensor([e0.elapsed_time(e1) / args.kshard_steps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item())
    out = {"workload": f"codebook-sharded VQ fwd+bwd, K={Kt} over {world} ranks ({vq.k_per} codes each), N={n_rows} latents "
                       f"replicated, D={D}",
           "K": Kt, "N": n_rows, "D": D, "ranks": world, "exchange": vq.exchange, "steps": args.kshard_steps,
           "ms_per_step": ms_step, "latents_per_s": n_rows / (ms_step * 1e-3),
           "aggregate_tflops_whole_step": 2.0 * n_rows * Kt * D / (ms_step * 1e-3) / 1e12,
           "search_ms_per_rank": search_ms,
           "search_tflops_per_rank": (2.0 * n_rows * vq.k_per * D / (search_ms * 1e-3) / 1e12) if search_ms else None,
           "loss": float(last["loss"].detach()), "perplexity": float(last["perp"])}
    if rank == 0 and not args.no_parity:
        from oracle import vq_oracle as O              # checker only
        with torch.no_grad():
            rows = torch.arange(0, n_rows, n_rows // 1024, device=dev)[:1024]
            E_full = torch.cat([shard_values(r) for r in range(world)])[:Kt].cpu()
            zs = z[rows].cpu()
            ours = last["idx"].reshape(-1)[rows].cpu()
            t0 = time.perf_counter()
            ref = O.kshard_merge(zs, E_full, world)
            par = O.index_parity(ours, ref, zs, E_full)
            out["index_parity_vs_oracle_kshard_merge"] = {"rows_checked": int(rows.numel()), "raw_mismatch": par.raw_mismatch,
                                                          "unexcused": par.unexcused,
                                                          "max_gap_over_tolerance": par.max_gap_over_tol,
                                                          "oracle_seconds": time.perf_counter() - t0}
            assert par.unexcused == 0, par
    return out


def measure_e2e(torch, dist, F, vq, z, gz, E, dev, world, args):
    """Same metric through the public host-buffer call: every step copies z and g_zq (and the codebook) from
    pinned host memory, runs forward + backward, and copies z_q, idx, dz, dE, loss, perplexity back."""
    n_rows = z.shape[0]
    steps = max(2, min(args.steps, 5))
    zh = torch.empty(n_rows, D, pin_memory=True).copy_(z)
    gh = torch.empty(n_rows, D, pin_memory=True).copy_(gz)
    Eh = torch.empty(K, D, pin_memory=True).copy_(E)
    h2d = (2 * n_rows * D + K * D) * 4 + 4
    d2h = (2 * n_rows * D + K * D) * 4 + 8 * n_rows + 8
    torch.cuda.synchronize()
    if world == 1:
        out = None
        for _ in range(2):
            out = F.forward_backward_host(zh, Eh, gh, 1.0, BETA, mode="auto", rows_per_chunk=args.chunk_rows, out=out)
        t0 = time.perf_counter()
        for _ in range(steps):
            out = F.forward_backward_host(zh, Eh, gh, 1.0, BETA, mode="auto", rows_per_chunk=args.chunk_rows, out=out)
        dt = (time.perf_counter() - t0) / steps
        api = "kvq_forward_backward_host (C ABI, pinned host buffers, chunked copy/compute overlap)"
        F._lib.load().kvq_host_release()
    else:
        out = None
        n_global = n_rows * world
        for _ in range(2):
            out = F.forward_backward_host_sharded(zh, Eh, gh, 1.0, BETA, n_global, mode="auto",
                                                  rows_per_chunk=args.chunk_rows, out=out)
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = F.forward_backward_host_sharded(zh, Eh, gh, 1.0, BETA, n_global, mode="auto",
                                                  rows_per_chunk=args.chunk_rows, out=out)
        dist.barrier()
        dt = (time.perf_counter() - t0) / steps
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        api = ("kvq_forward_backward_host_sharded per rank (C ABI, pinned host buffers, chunked copy/compute overlap) "
               "+ NCCL all-reduce of dE, histogram, loss partial")
        F._lib.load().kvq_host_release()
    return {"value": n_rows * world / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": dt * 1e3, "steps": steps, "api": api}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kvq", choices=["kvq", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cublas", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-side", action="store_true")
    ap.add_argument("--no-refcuda", action="store_true")
    ap.add_argument("--no-kshard", action="store_true")
    ap.add_argument("--kshard-steps", type=int, default=4)
    ap.add_argument("--chunk-rows", type=int, default=0)  # 0 = the library's default: one wave of the search kernel per chunk
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "kvq":
        args.warmup = 3          # timing rule: at least three warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_kvq(args)


if __name__ == "__main__":
    main()
