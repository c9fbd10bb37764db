"""Generate the golden fixtures in this directory from the UNMODIFIED reference class.

Run in the build container only (the reference is mounted read-only at /root/reference and does not
exist on the GPU box):

    python tests/golden/make_golden.py

It imports `VectorQuantizer` from /root/reference/models/shelgon3/VectorQuantizer.py, runs forward and
autograd backward on seeded CPU inputs and stores inputs + outputs as .npz.  The committed .npz files
are what tests compare against; this script is committed so that they can be regenerated and audited.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/models/shelgon3"


def load_reference():
    sys.path.insert(0, REF)
    from VectorQuantizer import VectorQuantizer  # noqa: E402  (the reference's own class)
    return VectorQuantizer


def make_inputs(name, B, S, D, K, init, seed):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(B, S, D, generator=g)
    if init == "default":          # VectorQuantizer.py:29
        E = (torch.rand(K, D, generator=g) * 2 - 1) / K
    elif init == "normal":
        E = torch.randn(K, D, generator=g)
    elif init == "points":         # mirrors kmeans2(minit='points'), vq_codebook_init_weights.py:85
        zf = z.view(-1, D)
        pick = torch.randperm(zf.shape[0], generator=g)[:K]
        E = zf[pick] + 0.1 * torch.randn(K, D, generator=g)
    elif init == "ties":           # duplicated codes: argmin must return the lowest index
        E = torch.randn(K, D, generator=g)
        E[K // 2:] = E[:K - K // 2]
    else:
        raise ValueError(init)
    gz = torch.randn(B, S, D, generator=g)
    return z.contiguous(), E.contiguous(), gz.contiguous()


def run_reference(VQ, z, E, gz, beta, w):
    K, D = E.shape
    vq = VQ(n_e=K, e_dim=D, beta=beta, vq_codebook_init_values=E.clone())
    zin = z.clone().requires_grad_(True)
    loss, z_q, perp, onehot, idx = vq.forward(zin, "cpu")
    total = loss * w + (z_q * gz).sum()
    total.backward()
    return dict(loss=loss.detach(), z_q=z_q.detach(), perplexity=perp.detach(), idx=idx,
                onehot_rowsum=onehot.sum(1), onehot_argmax=onehot.argmax(1),
                dz=zin.grad, dE=vq.embedding.weight.grad)


CASES = [
    # name,        B,  S,   D,   K,  init,      beta, w(loss weight), seed
    ("small",      4, 16,  64,  32, "normal",   0.25, 1.0,  69),
    ("default",    8, 12, 128,  64, "default",  0.69, 0.5,  70),
    ("bert",       8, 12, 768,  64, "points",   0.10, 2.0,  71),
    ("ties",       2, 32,  32,  16, "ties",     0.25, 1.0,  72),
    ("wide",       2, 64, 256, 512, "normal",   0.25, 1.0,  73),
]


def main():
    VQ = load_reference()
    torch.set_num_threads(1)  # one thread: summation order independent of the host's core count
    for name, B, S, D, K, init, beta, w, seed in CASES:
        z, E, gz = make_inputs(name, B, S, D, K, init, seed)
        out = run_reference(VQ, z, E, gz, beta, w)
        np.savez_compressed(
            os.path.join(HERE, f"vq_{name}.npz"),
            z=z.numpy(), E=E.numpy(), gz=gz.numpy(), beta=np.float64(beta), w=np.float64(w),
            loss=out["loss"].numpy(), z_q=out["z_q"].numpy(), perplexity=out["perplexity"].numpy(),
            idx=out["idx"].numpy(), dz=out["dz"].numpy(), dE=out["dE"].numpy(),
            onehot_rowsum=out["onehot_rowsum"].numpy(), onehot_argmax=out["onehot_argmax"].numpy())
        print(name, "loss", float(out["loss"]), "perp", float(out["perplexity"]))

    # BASELINE config 1 (B=64,S=64,D=768,K=512): inputs are regenerated from the seed (12 MB of latents is
    # too much to commit); the fixture stores the outputs that identify the result plus an input digest.
    for init in ("default", "points"):
        z, E, gz = make_inputs("c1", 64, 64, 768, 512, init, 69)
        out = run_reference(VQ, z, E, gz, 0.25, 1.0)
        digest = hashlib.sha256(z.numpy().tobytes() + E.numpy().tobytes() + gz.numpy().tobytes()).hexdigest()
        np.savez_compressed(
            os.path.join(HERE, f"vq_c1_{init}.npz"),
            input_sha256=np.array(digest), beta=np.float64(0.25), w=np.float64(1.0),
            loss=out["loss"].numpy(), perplexity=out["perplexity"].numpy(),
            idx=out["idx"].numpy().astype(np.int16),
            z_q_sum=np.float64(out["z_q"].double().sum()), z_q_abs=np.float64(out["z_q"].double().abs().sum()),
            dz_abs=np.float64(out["dz"].double().abs().sum()), dE_abs=np.float64(out["dE"].double().abs().sum()),
            dE_rows=out["dE"].double().abs().sum(1).numpy())
        print("c1", init, "loss", float(out["loss"]), "perp", float(out["perplexity"]))

    # seq_acc (common/metrics.py:8-36)
    sys.path.insert(0, "/root/reference")
    from common.metrics import seq_acc
    g = torch.Generator().manual_seed(74)
    a = torch.randint(0, 6, (16, 12), generator=g)
    b = torch.randint(0, 6, (16, 12), generator=g)
    acc, per = seq_acc(a, b)
    np.savez_compressed(os.path.join(HERE, "seq_acc.npz"), a=a.numpy(), b=b.numpy(), acc=acc.numpy(), per=per.numpy())


def gumbel_goldens():
    """GumbelQuantizer (models/shelgon3/GumbelQuantizer.py): the unmodified class, with torch's RNG seeded right before
    the forward so that the Gumbel sample F.gumbel_softmax draws can be reproduced and stored next to the outputs."""
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from GumbelQuantizer import GumbelQuantizer
    from oracle.vq_oracle import gumbel_noise_like_reference
    torch.set_num_threads(1)
    cases = [  # name, B, S, C, K, D, tau, kld_scale, straight_through, is_training, seed
        ("soft", 4, 12, 64, 32, 64, 0.9, 5e-4, False, True, 81),
        ("hard", 4, 12, 64, 32, 64, 1.0, 5e-4, True, True, 82),
        ("eval", 2, 12, 96, 9, 32, 0.5, 1e-3, False, False, 83),      # eval forces hard=True; K=9 as in the analyses
    ]
    for name, B, S, C, K, D, tau, kld, st, train, seed in cases:
        g = torch.Generator().manual_seed(seed)
        z = torch.randn(B, S, C, generator=g)
        gz = torch.randn(B, S, D, generator=g)
        torch.manual_seed(seed)
        gq = GumbelQuantizer(enc_out_size=C, n_embed=K, embedding_dim=D, temperature=tau, kl_div_scale=kld, straight_through=st)
        zin = z.clone().requires_grad_(True)
        noise = gumbel_noise_like_reference((B, K, S), seed + 1000)
        torch.manual_seed(seed + 1000)
        z_q, diff, ind = gq.forward(zin, train)
        (diff * 3.0 + (z_q * gz).sum()).backward()
        np.savez_compressed(
            os.path.join(HERE, f"gumbel_{name}.npz"),
            z=z.numpy(), gz=gz.numpy(), W=gq.proj.weight.detach()[:, :, 0].numpy(), b=gq.proj.bias.detach().numpy(),
            E=gq.embed.weight.detach().numpy(), noise=noise.numpy(), tau=np.float64(tau), kld_scale=np.float64(kld),
            hard=np.bool_(st if train else True), w=np.float64(3.0),
            z_q=z_q.detach().numpy(), diff=diff.detach().numpy(), ind=ind.numpy(), dz=zin.grad.numpy(),
            dW=gq.proj.weight.grad[:, :, 0].numpy(), db=gq.proj.bias.grad.numpy(), dE=gq.embed.weight.grad.numpy())
        print("gumbel", name, "diff", float(diff), "unique codes", int(ind.unique().numel()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gumbel":
        gumbel_goldens()
    else:
        main()
        gumbel_goldens()
