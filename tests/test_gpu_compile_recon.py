"""(1) The layer under torch.compile: the reference wraps its model in `model.compile()` (models/shelgon3/main.py:83); the
layer is registered as dispatcher operators with fake implementations, so it traces into ONE graph (fullgraph=True).
(2) The fused reconstruction loss (models/shelgon3/Trainer.py:94-101) against the reference's own tensor expressions."""
import pytest
import torch
import torch.nn.functional as Fn

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _kvq():
    import kindergarten_vq_vae_b200 as k
    return k


class _Toy(torch.nn.Module):
    """encoder stand-in -> VQ -> decoder stand-in, consuming the layer's outputs the way Shelgon.forward / Trainer.step do."""

    def __init__(self, vq, d):
        super().__init__()
        self.pre = torch.nn.Linear(d, d)
        self.vector_quantizer = vq
        self.post = torch.nn.Linear(d, d)

    def forward(self, x):
        h = self.pre(x)
        loss_vq, z_q, perplexity, _min_encodings, idx = self.vector_quantizer.forward(h, x.device)   # Shelgon.py:58
        y = self.post(z_q)
        loss_vq *= 0.5                                                                                # Trainer.py:104, in place
        return (y ** 2).mean() + loss_vq, perplexity, idx


@pytest.mark.parametrize("backend", ["aot_eager", "inductor"])
def test_layer_compiles_fullgraph_and_matches_eager(backend):
    k = _kvq()
    torch.manual_seed(0)
    K, D = 96, 64
    model = _Toy(k.VectorQuantizer(K, D, 0.25, vq_codebook_init_values=torch.randn(K, D)), D).to(DEV)
    x = torch.randn(8, 12, D, device=DEV)
    total_e, perp_e, idx_e = model(x)
    total_e.backward()
    grads_e = {n: p.grad.clone() for n, p in model.named_parameters()}
    model.zero_grad()
    torch._dynamo.reset()
    compiled = torch.compile(model, fullgraph=True, backend=backend)      # fullgraph: any graph break is an error
    total_c, perp_c, idx_c = compiled(x)
    total_c.backward()
    torch.cuda.synchronize()
    assert torch.equal(idx_c, idx_e)
    assert abs(float(total_c) - float(total_e)) <= 1e-5 * abs(float(total_e))
    assert abs(float(perp_c) - float(perp_e)) <= 1e-6 * float(perp_e)
    for n, p in model.named_parameters():
        assert torch.allclose(p.grad, grads_e[n], rtol=1e-4, atol=1e-6), n
    # no_grad / eval through the compiled module (Trainer.py:365)
    with torch.no_grad():
        total_n, _, idx_n = compiled(x)
    assert torch.equal(idx_n, idx_e)


def _reference_recon(logits, ids, V):
    """Trainer.py:94-101 verbatim, plus common/metrics.py:8-36."""
    loss = Fn.kl_div(input=Fn.log_softmax(logits.reshape(-1, V), dim=-1),
                     target=Fn.one_hot(ids, V).reshape(-1, V).float(), reduction="batchmean")
    recon = torch.argmax(torch.softmax(logits, dim=-1), dim=-1)
    mask = (recon - ids) == 0
    return loss, recon, mask.sum() / recon.numel(), torch.mean(mask.float(), dim=-1)


@pytest.mark.parametrize("B,S,V,scale", [(16, 12, 30522, 1.0), (64, 12, 30522, 8.0), (3, 5, 17, 2.0), (2, 7, 1000, 30.0)])
def test_fused_recon_loss_matches_the_reference_expressions(B, S, V, scale):
    k = _kvq()
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + S)
    logits = (torch.randn(B, S, V, device=DEV, generator=g) * scale).requires_grad_(True)
    ids = torch.randint(0, V, (B, S), device=DEV, generator=g)
    # make a good part of the positions "correct" so the accuracy is not trivially zero
    with torch.no_grad():
        hit = torch.rand(B, S, device=DEV, generator=g) < 0.4
        logits.view(-1, V)[hit.view(-1).nonzero().flatten(), ids.view(-1)[hit.view(-1)]] += 20.0 * scale
    loss_r, recon_r, acc_r, per_r = _reference_recon(logits, ids, V)
    (loss_r * 1.7).backward()
    grad_r = logits.grad.clone()
    logits.grad = None
    loss, recon, acc, per = k.recon_loss(logits, ids)
    loss_scaled = loss * 1.7
    loss_scaled.backward()
    torch.cuda.synchronize()
    assert torch.equal(recon, recon_r)
    # torch forms sum / numel as sum * (1 / numel) on CUDA and as a true division on the CPU: allow that last-bit difference
    assert abs(float(acc) - float(acc_r)) <= 1.2e-7 and float((per - per_r).abs().max()) <= 1.2e-7
    assert abs(float(acc) - float(((recon_r - ids) == 0).sum()) / recon_r.numel()) <= 6e-8
    assert abs(float(loss) - float(loss_r)) <= 2e-6 * abs(float(loss_r)) + 1e-7
    assert float((logits.grad - grad_r).abs().max()) <= 1e-5 * float(grad_r.abs().max()) + 1e-10
    # bitwise reproducible (fixed-order reduction of the row losses)
    loss2, *_ = k.recon_loss(logits.detach(), ids)
    assert float(loss2) == float(loss)


@pytest.mark.parametrize("V", [4098, 30522, 1001])
def test_recon_argmax_ties_nan_and_infinities(V):
    """arg-max semantics of torch.argmax on the grouped / vectorised path: first column on ties, the first NaN wins
    outright, -inf logits contribute nothing, +inf is the maximum."""
    k = _kvq()
    g = torch.Generator(device=DEV).manual_seed(V)
    logits = torch.randn(6, 2, V, device=DEV, generator=g)
    x = logits.view(-1, V)
    x[0, 700] = 50.0; x[0, 3000 % V] = 50.0; x[0, 5] = 50.0            # three equal maxima: column 5
    x[1, 900] = float("nan"); x[1, 17] = float("nan")                   # NaN is the maximum, first one: column 17
    x[2, :] = float("-inf"); x[2, 123] = 1.5                            # one finite logit
    x[3, 600:700] = float("-inf")                                       # a stretch of -inf inside a normal row
    x[4, 42] = float("inf")                                             # +inf
    x[5, V - 1] = 60.0                                                  # maximum in the very last column (tail loop)
    ids = torch.randint(0, V, (6, 2), device=DEV, generator=g)
    loss, recon, acc, per = k.recon_loss(logits, ids)
    want = torch.argmax(logits, dim=-1)
    assert torch.equal(recon, want), (recon.flatten().tolist(), want.flatten().tolist())
    rows = [0, 2, 3, 5] + list(range(6, 12))                            # rows whose log-softmax is finite
    ref = -torch.log_softmax(x[rows].double(), dim=-1).gather(1, ids.view(-1, 1)[rows]).squeeze(1)
    _, _, _, _, lse = torch.ops.kvq.recon_loss_forward(logits, ids)
    got = lse[rows].double() - x[rows].double().gather(1, ids.view(-1, 1)[rows]).squeeze(1)
    fin = torch.isfinite(ref)
    assert torch.allclose(got[fin], ref[fin], rtol=1e-5, atol=1e-5)


def test_recon_loss_inplace_scaling_and_compile():
    k = _kvq()
    V = 1000
    logits = torch.randn(4, 12, V, device=DEV, requires_grad=True)
    ids = torch.randint(0, V, (4, 12), device=DEV)

    def step(lg):
        loss, recon, acc, per = k.recon_loss(lg, ids)
        loss *= 3.0                                           # Trainer.py:103 multiplies in place
        return loss, recon, acc

    l_e, r_e, a_e = step(logits)
    l_e.backward()
    g_e = logits.grad.clone(); logits.grad = None
    torch._dynamo.reset()
    l_c, r_c, a_c = torch.compile(step, fullgraph=True, backend="aot_eager")(logits)
    l_c.backward()
    assert torch.equal(r_c, r_e) and float(a_c) == float(a_e)   # same kernel both times
    assert abs(float(l_c) - float(l_e)) <= 1e-6 * float(l_e)
    assert torch.allclose(logits.grad, g_e, rtol=1e-5, atol=1e-9)
