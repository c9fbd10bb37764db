"""A few forward + backward passes of the GumbelQuantizer at the reference's largest batch (for an ncu launch list)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kindergarten_vq_vae_b200 as kvq  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(69)
K = 512
gq = kvq.GumbelQuantizer(768, K, 768, temperature=1.0, kl_div_scale=5e-4, straight_through=False).to(dev)
z = torch.randn(2048, 12, 768, device=dev, generator=g).requires_grad_(True)
gz = torch.randn(2048, 12, 768, device=dev, generator=g)
one = torch.ones((), device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    z.grad = None
    gq.zero_grad(set_to_none=True)
    z_q, diff, ind = gq(z, True, seed=3)
    torch.autograd.backward([z_q, diff], [gz, one])
torch.cuda.synchronize()
print("done")
