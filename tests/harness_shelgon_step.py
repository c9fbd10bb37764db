"""BASELINE config 2: a full shelgon3 train step on one B200 -- random-init BERT-base encoder -> VQ (K=512, D=768)
-> BERT LM-head decoder with cross-attention on z_q, on dSentences-shaped synthetic batches (12 tokens, vocab 30522).

`Shelgon.forward` (models/shelgon3/Shelgon.py:50-73) and `Trainer.step` (models/shelgon3/Trainer.py:65-124) are
restated here because the reference's own classes do not construct at HEAD and need network
access for pretrained weights.  The same model is stepped twice from identical seeds: once with the kvq
VectorQuantizer, once with a literal PyTorch restatement of the reference layer; losses, indices and step times are
compared.  Test harness (it uses the oracle as the A/B partner), not product code.
"""
import json
import os
import sys
import time

import torch
import torch.nn.functional as Fn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VOCAB, SEQ, K, D, BETA = 30522, 12, 512, 768, 0.25


class Shelgon(torch.nn.Module):
    def __init__(self, vq):
        super().__init__()
        from transformers import BertConfig, BertLMHeadModel, BertModel
        self.encoder = BertModel(BertConfig(), add_pooling_layer=False)
        self.decoder = BertLMHeadModel(BertConfig(is_decoder=True, add_cross_attention=True))
        self.vector_quantizer = vq

    def forward(self, input_ids, attention_mask, device, is_training):
        embeds = self.encoder(input_ids, attention_mask=attention_mask).last_hidden_state          # Shelgon.py:52
        assert embeds.shape[-1] == self.vector_quantizer.e_dim                                      # :54
        assert type(self.vector_quantizer).__name__.endswith("VectorQuantizer")                     # :57
        vq_loss, z_q, perplexity, _, idx = self.vector_quantizer.forward(embeds, device)            # :58
        logits = self.decoder(encoder_hidden_states=z_q, input_ids=input_ids, attention_mask=attention_mask).logits  # :71
        return vq_loss, perplexity, idx, logits


def train_step(model, opt, input_ids, mask, dev, vq_weight=1.0, fused_recon=False):
    """Trainer.py:87-115.  fused_recon: kvq.recon_loss (one pass over the logits) instead of the reference expressions."""
    loss_vq, perp, idx, logits = model.forward(input_ids, mask, dev, True)
    if fused_recon:
        from kindergarten_vq_vae_b200 import recon_loss
        loss_recon, recon_ids, _acc, _per = recon_loss(logits, input_ids)
    else:
        target = Fn.one_hot(input_ids, VOCAB).reshape(-1, VOCAB).float()
        loss_recon = Fn.kl_div(Fn.log_softmax(logits.reshape(-1, VOCAB), dim=-1), target, reduction="batchmean")  # :94-98
        recon_ids = torch.argmax(torch.softmax(logits, dim=-1), dim=-1)                              # :100
    loss_vq *= vq_weight                                                                             # :104 (in place)
    loss_full = loss_recon + loss_vq
    opt.zero_grad()
    loss_full.backward()
    opt.step()
    return loss_recon.detach(), loss_vq.detach(), perp.detach(), idx, recon_ids


def build(kind, dev, seed):
    import kindergarten_vq_vae_b200 as kvq
    from oracle import vq_oracle as O
    torch.manual_seed(seed)
    init = torch.randn(K, D) * 0.7          # BERT last_hidden_state is layer-normed: O(1) entries
    vq = (kvq.VectorQuantizer(K, D, BETA, vq_codebook_init_values=init) if kind == "kvq"
          else O.LiteralVectorQuantizer(K, D, BETA, vq_codebook_init_values=init))
    torch.manual_seed(seed)
    model = Shelgon(vq).to(dev)
    model.train()
    for m in model.modules():               # deterministic A/B: no dropout
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    opt = torch.optim.Adam(model.parameters(), lr=1e-5, weight_decay=0.0, amsgrad=False)             # main.py:91
    return model, opt


def main():
    dev = torch.device("cuda:0")
    from kindergarten_vq_vae_b200 import seq_acc
    results = []
    for B in (512, 2048):
        g = torch.Generator().manual_seed(69)
        batches = [torch.randint(0, VOCAB, (B, SEQ), generator=g).to(dev) for _ in range(4)]
        mask = torch.ones(B, SEQ, dtype=torch.long, device=dev)
        rec = {}
        for kind in ("kvq", "literal"):
            model, opt = build(kind, dev, 123)
            losses = []
            for i in range(8):
                out = train_step(model, opt, batches[i % 4], mask, dev, fused_recon=(kind == "kvq"))
                losses.append((float(out[0]), float(out[1]), float(out[2])))
                if i == 0:
                    first_idx = out[3].clone()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(10):
                out = train_step(model, opt, batches[i % 4], mask, dev, fused_recon=(kind == "kvq"))
            e1.record(); torch.cuda.synchronize()
            ms_step = e0.elapsed_time(e1) / 10
            # the VQ layer alone inside this model: forward + backward on the encoder's latents
            with torch.no_grad():
                z = model.encoder(batches[0], attention_mask=mask).last_hidden_state
            z = z.detach().requires_grad_(True)
            gz = torch.randn_like(z)
            def vq_only():
                z.grad = None; model.vector_quantizer.embedding.weight.grad = None
                l, zq, *_ = model.vector_quantizer.forward(z, dev)
                torch.autograd.backward([l, zq], [torch.ones((), device=dev), gz])
            for _ in range(5):
                vq_only()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(50):
                vq_only()
            e1.record(); torch.cuda.synchronize()
            acc = seq_acc(out[4], batches[9 % 4])[0]
            rec[kind] = dict(ms_step=ms_step, ms_vq_fwd_bwd=e0.elapsed_time(e1) / 50, losses=losses, idx=first_idx.cpu(),
                             acc=float(acc))
            del model, opt
            torch.cuda.empty_cache()
        a, b = rec["kvq"], rec["literal"]
        same_idx = float((a["idx"] == b["idx"]).float().mean())
        drift = max(abs(x[0] - y[0]) / max(abs(y[0]), 1e-9) for x, y in zip(a["losses"], b["losses"]))
        drift_vq = max(abs(x[1] - y[1]) / max(abs(y[1]), 1e-9) for x, y in zip(a["losses"], b["losses"]))
        line = dict(config="C2-shelgon-step", B=B, S=SEQ, K=K, D=D, ms_step_kvq=a["ms_step"], ms_step_literal=b["ms_step"],
                    sentences_per_s_kvq=B / a["ms_step"] * 1e3, sentences_per_s_literal=B / b["ms_step"] * 1e3,
                    ms_vq_fwd_bwd_kvq=a["ms_vq_fwd_bwd"], ms_vq_fwd_bwd_literal=b["ms_vq_fwd_bwd"],
                    first_step_index_agreement=same_idx, max_rel_diff_loss_recon_8_steps=drift,
                    max_rel_diff_loss_vq_8_steps=drift_vq, loss_first=a["losses"][0], loss_first_literal=b["losses"][0],
                    loss_last=a["losses"][-1], loss_last_literal=b["losses"][-1])
        print(json.dumps(line), flush=True)
        results.append(line)


if __name__ == "__main__":
    main()
