"""Several traced host-buffer calls at the headline shape; per call: total, when the inbound copies ended, the largest
lag of compute behind the inbound copies and of the return copies behind compute (which stage made a slow call slow).
    KVQ_PIPE_TRACE=1 python tools/e2e_trace_stats.py [calls] 2> trace.txt ; python tools/e2e_trace_stats.py --parse trace.txt"""
import os
import re
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse(path):
    calls, cur = [], None
    for line in open(path):
        m = re.search(r"kvq host pipeline: (\d+) chunks, codebook landed ([\d.]+) ms, all done ([\d.]+) ms", line)
        if m:
            cur = {"total": float(m.group(3)), "rows": []}
            calls.append(cur)
            continue
        m = re.search(r"chunk\s+(\d+) rows\s+(\d+)\s+z in\s+([\d.]+)\s+g in\s+([\d.]+)\s+computed\s+([\d.]+)\s+returned\s+([\d.]+)", line)
        if m and cur is not None:
            cur["rows"].append(tuple(float(x) for x in m.groups()[2:]))
    for i, c in enumerate(calls):
        r = c["rows"]
        lag_c = max(x[2] - x[1] for x in r)
        lag_r = max(x[3] - x[2] for x in r)
        steps = [b[1] - a[1] for a, b in zip(r[1:-1], r[2:])]
        print(f"call {i}: total {c['total']:.2f} ms, inbound done {r[-1][1]:.2f}, max compute lag {lag_c:.2f}, "
              f"max return lag {lag_r:.2f}, inbound chunk period min/max {min(steps):.3f}/{max(steps):.3f}, last returned {r[-1][3]:.2f}")


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--parse":
        return parse(sys.argv[2])
    import torch
    from kindergarten_vq_vae_b200 import functional as F
    calls = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    N, D, K = 1 << 20, 256, 65536
    g = torch.Generator().manual_seed(69)
    zh = torch.empty(N, D, pin_memory=True).normal_(generator=g)
    gh = torch.empty(N, D, pin_memory=True).normal_(generator=g)
    Eh = torch.empty(K, D, pin_memory=True).normal_(generator=g)
    out = None
    for _ in range(calls):
        t0 = time.perf_counter()
        out = F.forward_backward_host(zh, Eh, gh, 1.0, 0.25, mode="auto", rows_per_chunk=0, out=out)
        print(f"wall {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)


if __name__ == "__main__":
    main()
