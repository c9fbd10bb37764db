"""A few forward+backward steps of the layer at BASELINE config 1 (N=4096, D=768, K=512), for `ncu` launch lists:
    ncu --metrics gpu__time_duration.sum --csv --log-file gpurun_out/c1_launches.csv python tools/probe_small.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kindergarten_vq_vae_b200 as k  # noqa: E402

dev = "cuda:0"
g = torch.Generator().manual_seed(1)
B, S, D, K = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (64, 64, 768, 512)))
z = torch.randn(B, S, D, generator=g).to(dev).requires_grad_(True)
gz = torch.randn(B, S, D, generator=g).to(dev)
vq = k.VectorQuantizer(K, D, 0.25, vq_codebook_init_values=torch.randn(K, D, generator=g), min_encodings=False).to(dev)
one = torch.ones((), device=dev)
for _ in range(3):
    z.grad = None
    vq.embedding.weight.grad = None
    loss, z_q, perp, _, idx = vq.forward(z, dev)
    torch.autograd.backward([loss, z_q], [one, gz])
torch.cuda.synchronize()
print("ok", float(loss))
