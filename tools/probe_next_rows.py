"""Timings of the scope table's "next" rows at the reference's own sizes, against the HBM / tensor
figures that bound them.  One JSON line per measurement.   python tools/probe_next_rows.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kindergarten_vq_vae_b200 as kvq  # noqa: E402
from kindergarten_vq_vae_b200.recon import recon_loss  # noqa: E402

HBM_GBS = 6535.7     # MEASURED_PEAKS.json


def timed(fn, warm=3, iters=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(69)
    # ---- (f)2 fused reconstruction loss, Trainer.py:94-101: B = 2048 sentences x 12 tokens, BERT vocabulary
    B, S, V = 2048, 12, 30522
    logits = torch.randn(B, S, V, device=dev, generator=g).requires_grad_(True)
    ids = torch.randint(0, V, (B, S), device=dev, generator=g)
    one = torch.ones((), device=dev)
    with torch.no_grad():
        ms_f = timed(lambda: recon_loss(logits, ids))

    def fb():
        logits.grad = None
        loss, *_ = recon_loss(logits, ids)
        loss.backward(one)
    ms_fb = timed(fb)
    bytes_f, bytes_b = 4.0 * B * S * V, 8.0 * B * S * V
    print(json.dumps({"row": "(f)2 recon loss", "rows": B * S, "V": V, "fwd_ms": ms_f, "fwd_gbs": bytes_f / ms_f / 1e6,
                      "fwd_frac_hbm": bytes_f / ms_f / 1e6 / HBM_GBS, "bwd_ms": ms_fb - ms_f,
                      "bwd_gbs": bytes_b / (ms_fb - ms_f) / 1e6, "bwd_frac_hbm": bytes_b / (ms_fb - ms_f) / 1e6 / HBM_GBS}), flush=True)
    del logits
    torch.cuda.empty_cache()
    # ---- (f)1 k-means codebook init, vq_codebook_init_weights.py:79-101: BERT latents of 8192 sentences x 12 tokens, K = 512
    N, D, K = 8192 * 12, 768, 512
    data = torch.randn(N, D, device=dev, generator=g)
    for search in ("auto", "fp32", "tf32"):
        ms = timed(lambda: kvq.kmeans2(data, K, iter=10, seed=1, search=search), warm=1, iters=3)
        print(json.dumps({"row": "(f)1 kmeans2", "N": N, "D": D, "K": K, "iter": 10, "search": search, "ms": ms,
                          "ms_per_iteration": ms / 10, "hbm_floor_ms_per_iteration": 2 * 4.0 * N * D / HBM_GBS / 1e6}), flush=True)
    # ---- (f)3 code usage analysis: (token id x code) table over a corpus of 1M sentences x 12 tokens
    from kindergarten_vq_vae_b200 import analysis
    tok = torch.randint(0, V, (1 << 20, 12), device=dev, generator=g)
    codes = torch.randint(0, K, (1 << 20, 12, 1), device=dev, generator=g)
    ms = timed(lambda: analysis.code_usage_by_token(tok, codes, V, K), warm=2, iters=10)
    print(json.dumps({"row": "(f)3 code usage table", "tokens": tok.numel(), "vocab": V, "n_e": K, "ms": ms,
                      "tokens_per_s": tok.numel() / ms * 1e3}), flush=True)
    # ---- (f)4 Gumbel quantiser, GumbelQuantizer.py:43-83: B = 2048 x 12 BERT latents, K = 512 codes of 768 dims
    gq = kvq.GumbelQuantizer(768, K, 768, temperature=1.0, kl_div_scale=5e-4, straight_through=False).to(dev)
    z = torch.randn(2048, 12, 768, device=dev, generator=g).requires_grad_(True)
    gz = torch.randn(2048, 12, 768, device=dev, generator=g)

    def gfb():
        z.grad = None
        gq.zero_grad(set_to_none=True)
        z_q, diff, ind = gq(z, True, seed=3)
        torch.autograd.backward([z_q, diff], [gz, one])
    with torch.no_grad():
        ms_gf = timed(lambda: gq(z, True, seed=3))
    ms_gfb = timed(gfb)
    Nn = 2048 * 12
    flops_f = 2.0 * Nn * K * 768 * 2            # logits + mix
    flops_b = 2.0 * Nn * K * 768 * 5            # dy, dz, dW, dE (+ the transposes' traffic)
    # the reference's own tensor expressions (GumbelQuantizer.py:43-83) with torch on the same GPU
    import torch.nn.functional as Fn
    proj, embed = gq.proj, gq.embed

    def ref_forward():
        zt = z.permute(0, 2, 1)
        logits = proj(zt)
        soft = Fn.gumbel_softmax(logits, tau=1.0, dim=1, hard=False)
        z_q = torch.einsum("b n s, n d -> b d s", soft, embed.weight)
        qy = Fn.softmax(logits, dim=1)
        diff = 5e-4 * torch.sum(qy * torch.log(qy * K + 1e-10), dim=1).mean()
        ind = soft.argmax(dim=1)
        return z_q.permute(0, 2, 1), diff, ind

    def ref_fb():
        z.grad = None
        gq.zero_grad(set_to_none=True)
        z_q, diff, ind = ref_forward()
        torch.autograd.backward([z_q, diff], [gz, one])
    ref = {}
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        with torch.no_grad():
            f = timed(ref_forward)
        ref["allow_tf32" if tf32 else "fp32"] = {"fwd_ms": f, "fwd_bwd_ms": timed(ref_fb)}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    print(json.dumps({"row": "(f)4 GumbelQuantizer soft", "N": Nn, "K": K, "fwd_ms": ms_gf, "fwd_tflops": flops_f / ms_gf / 1e9,
                      "fwd_bwd_ms": ms_gfb, "fwd_bwd_tflops": (flops_f + flops_b) / ms_gfb / 1e9,
                      "reference_torch_ops_same_gpu": ref}), flush=True)


if __name__ == "__main__":
    main()
