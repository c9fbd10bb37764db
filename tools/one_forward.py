"""A few forward passes at N = 2^20, D = 256 for profiling one kernel under ncu.   python tools/one_forward.py K mode [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kindergarten_vq_vae_b200 as kvq  # noqa: E402
from tools.sweep import make  # noqa: E402

K, mode = int(sys.argv[1]), sys.argv[2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
N, D = 1 << 20, 256
z, gz, E = make(dev, N, D, K, "data")
vq = kvq.VectorQuantizer(K, D, 0.25, vq_codebook_init_values=E, min_encodings=False, search=mode).to(dev)
with torch.no_grad():
    for _ in range(reps):
        vq.forward(z.view(N // 64, 64, D), dev)
torch.cuda.synchronize()
print("done")
