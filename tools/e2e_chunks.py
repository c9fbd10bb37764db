"""End-to-end host-buffer call (kvq_forward_backward_host) at the headline shape for several chunk sizes, and, with
--trace, one traced call per size (KVQ_PIPE_TRACE=1: per-chunk timeline on stderr).   python tools/e2e_chunks.py [--trace] [rows ...]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kindergarten_vq_vae_b200 import functional as F  # noqa: E402

N, D, K, BETA = 1 << 20, 256, 65536, 0.25


def main():
    trace = "--trace" in sys.argv
    rows = [int(a) for a in sys.argv[1:] if a != "--trace"] or [18944, 37888, 75776, 131072]
    g = torch.Generator().manual_seed(69)
    zh = torch.empty(N, D, pin_memory=True).normal_(generator=g)
    gh = torch.empty(N, D, pin_memory=True).normal_(generator=g)
    Eh = torch.empty(K, D, pin_memory=True).normal_(generator=g)
    out = None
    for r in rows:
        for _ in range(2):
            out = F.forward_backward_host(zh, Eh, gh, 1.0, BETA, mode="auto", rows_per_chunk=r, out=out)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            out = F.forward_backward_host(zh, Eh, gh, 1.0, BETA, mode="auto", rows_per_chunk=r, out=out)
            ts.append((time.perf_counter() - t0) * 1e3)
        if trace:
            os.environ["KVQ_PIPE_TRACE"] = "1"
            print(f"--- trace, rows_per_chunk {r}", file=sys.stderr, flush=True)
            F.forward_backward_host(zh, Eh, gh, 1.0, BETA, mode="auto", rows_per_chunk=r, out=out)
            os.environ["KVQ_PIPE_TRACE"] = "0"
        print(f"rows_per_chunk {r:7d}: {min(ts):.2f} / {sorted(ts)[2]:.2f} / {max(ts):.2f} ms (min / median / max)", flush=True)


if __name__ == "__main__":
    main()
