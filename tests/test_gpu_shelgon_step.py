"""BASELINE config 2 as a test: a full shelgon3 train step on one B200 -- random-init BERT-base encoder -> VQ (K=512,
D=768) -> BERT LM-head decoder with cross-attention on z_q, dSentences-shaped synthetic batches (12 tokens, vocab 30522).
`Shelgon.forward` (models/shelgon3/Shelgon.py:50-73) and `Trainer.step` (models/shelgon3/Trainer.py:65-124) are restated
in tests/harness_shelgon_step.py (the reference's own classes do not construct at HEAD).

The same model is stepped from identical seeds with (a) the kvq VectorQuantizer + the fused kvq reconstruction loss and
(b) a literal PyTorch restatement of the reference layer + the reference's loss expressions."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_shelgon_train_steps_track_the_reference_layer():
    pytest.importorskip("transformers")
    import harness_shelgon_step as H
    dev = torch.device("cuda:0")
    B, steps = 128, 4
    g = torch.Generator().manual_seed(69)
    batches = [torch.randint(0, H.VOCAB, (B, H.SEQ), generator=g).to(dev) for _ in range(steps)]
    mask = torch.ones(B, H.SEQ, dtype=torch.long, device=dev)
    rec = {}
    for kind in ("kvq", "literal"):
        model, opt = H.build(kind, dev, 123)
        out_steps = []
        for i in range(steps):
            loss_recon, loss_vq, perp, idx, recon_ids = H.train_step(model, opt, batches[i], mask, dev, vq_weight=0.7,
                                                                     fused_recon=(kind == "kvq"))
            out_steps.append((float(loss_recon), float(loss_vq), float(perp), idx.clone(), recon_ids.clone()))
        rec[kind] = out_steps
        del model, opt
        torch.cuda.empty_cache()
    a, b = rec["kvq"], rec["literal"]
    # first step: identical weights, so the two layers see the same latents
    agree = float((a[0][3] == b[0][3]).float().mean())
    print(f"config 2: first-step code agreement {agree:.6f}; losses kvq {a[0][:3]} literal {b[0][:3]}")
    assert agree >= 0.999
    assert tuple(a[0][3].shape) == (B, H.SEQ, 1) and a[0][3].dtype == torch.int64
    assert abs(a[0][0] - b[0][0]) <= 1e-5 * abs(b[0][0])          # reconstruction loss (fused kernel vs kl_div chain)
    assert abs(a[0][1] - b[0][1]) <= 1e-4 * abs(b[0][1])          # VQ loss
    assert abs(a[0][2] - b[0][2]) <= 1e-3 * abs(b[0][2])          # perplexity
    assert float((a[0][4] == b[0][4]).float().mean()) >= 0.999    # reconstructed token ids
    # the following steps went through Adam with each path's own gradients: trajectories must stay together
    for i in range(1, steps):
        assert abs(a[i][0] - b[i][0]) <= 2e-3 * abs(b[i][0]), (i, a[i][0], b[i][0])
        assert abs(a[i][1] - b[i][1]) <= 2e-2 * abs(b[i][1]), (i, a[i][1], b[i][1])
