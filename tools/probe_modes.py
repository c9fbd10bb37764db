"""Same-box, interleaved A/B of the search kernel alone (CUDA events around the launch, library profiler) for the plain
tensor-core search ("tf32") and the default mode's top-2 search ("auto"), N = 2^20, D = 256, K = 8192 ... 65536.
    python tools/probe_modes.py [K ...]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kindergarten_vq_vae_b200 as kvq  # noqa: E402
from kindergarten_vq_vae_b200 import _lib  # noqa: E402
from tools.sweep import make, profile  # noqa: E402


def main():
    Ks = [int(a) for a in sys.argv[1:]] or [8192, 65536]
    dev = torch.device("cuda:0")
    lib = _lib.load()
    N, D = 1 << 20, 256
    for K in Ks:
        z, gz, E = make(dev, N, D, K, "data")
        z3 = z.view(N // 64, 64, D)
        mods = {m: kvq.VectorQuantizer(K, D, 0.25, vq_codebook_init_values=E, min_encodings=False, search=m).to(dev)
                for m in ("tf32", "auto")}
        mods["auto_nofilter"] = mods["auto"]            # KVQ_TOP2_FILTER=0: runner-up tracking without the tau band
        res = {m: [] for m in mods}
        with torch.no_grad():
            for m, vq in mods.items():
                for _ in range(3):
                    vq.forward(z3, dev)
            for rep in range(4):
                for m, vq in mods.items():
                    os.environ["KVQ_TOP2_FILTER"] = "0" if m == "auto_nofilter" else "1"
                    p = profile(lib, lambda: vq.forward(z3, dev), iters=5 if K > 16384 else 20)
                    res[m].append(round(p["search"], 4))
        fl = 2.0 * N * K * D
        print(json.dumps({"K": K, "search_ms": res, "tflops_best": {m: round(fl / min(v) / 1e9, 1) for m, v in res.items()},
                          "top2_over_plain_pct": round(100 * (min(res["auto"]) / min(res["tf32"]) - 1), 2)}), flush=True)
        del mods, z, gz, E, z3
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
