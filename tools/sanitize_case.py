"""Small ragged forward+backward through every kernel family, meant to run under compute-sanitizer (memcheck)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kindergarten_vq_vae_b200 as kvq

dev = "cuda:0"
g = torch.Generator().manual_seed(1)
for (B, S, D, K, search) in [(3, 111, 96, 300, "tf32"), (3, 111, 96, 300, "fp32"), (2, 70, 256, 1000, "tf32"), (1, 45, 768, 130, "tf32"),
                            (1, 200, 36, 77, "fp32")]:
    z = torch.randn(B, S, D, generator=g).to(dev).requires_grad_(True)
    E = torch.randn(K, D, generator=g)
    gz = torch.randn(B, S, D, generator=g).to(dev)
    vq = kvq.VectorQuantizer(K, D, 0.25, vq_codebook_init_values=E, search=search, min_encodings=True).to(dev)
    loss, z_q, perp, onehot, idx = vq.forward(z, dev)
    (loss * 1.5 + (z_q * gz).sum()).backward()
    torch.cuda.synchronize()
    print(B, S, D, K, search, float(loss.detach()), float(perp), int(idx.max()), flush=True)
ids = torch.randint(0, 50, (64, 12), generator=g).to(dev)
print(kvq.seq_acc(ids, ids.roll(1, 0))[0].item())
print(kvq.replace_pct_rand_values(ids, 0.3, 0, 100, seed=1).sum().item(), kvq.change_percentage_of_elements(ids, 1, 0.5, 0, 9, seed=2).sum().item())
c, l = kvq.kmeans2(torch.randn(3000, 64, device=dev), 17, iter=3, seed=0)
print(c.shape, int(l.max()))
t = kvq.analysis.code_usage_by_token(ids, torch.randint(0, 9, (64, 12, 1), generator=g).to(dev), 50, 9)
print(int(t.sum()))
torch.cuda.synchronize()
print("sanitize case done")
