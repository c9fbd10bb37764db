"""CPU oracle for the hard vector-quantisation bottleneck.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the algorithm of the reference's
``models/shelgon3/VectorQuantizer.py:31-93`` (forward) and of the autograd
backward PyTorch derives from it.  It exists so that the CUDA path can be
*checked*; nothing in the product package may import it.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` use it.

Pinning.  The reference ships no golden vectors and no tests, and
all of its arithmetic lives in un-vendored, unpinned ``torch``
(``requirements.txt:3``).  The pin is therefore: the reference class itself,
imported unmodified from ``/root/reference`` in the build container and executed
on the installed torch 2.11.0 (CPU, fp32), with its inputs and outputs committed
under ``tests/golden/`` by ``tests/golden/make_golden.py``.
``tests/test_oracle.py`` checks every function below against those fixtures
(bit-exact for indices, z_q, loss; closed-form gradients to 1e-6 relative).

The arithmetic is torch-CPU fp32 on purpose: the reference *is* torch fp32, and
details such as ``argmin`` tie-breaking (lowest index) and the evaluation order
``(|z|^2 + |E|^2) - 2 z.E`` decide which code wins when distances are nearly
tied.  ``truth_fp64`` is the independent float64 evaluation used to judge those
near-ties.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

__all__ = [
    "ForwardResult", "forward_fp32", "backward_closed_form", "truth_fp64",
    "index_parity", "IndexParity", "seq_acc", "perplexity_from_counts",
    "onehot", "kshard_merge", "dp_merge",
]


@dataclass
class ForwardResult:
    loss: torch.Tensor          # 0-d fp32
    z_q: torch.Tensor           # (B,S,D) fp32, value fl(z + fl(q - z))
    perplexity: torch.Tensor    # 0-d fp32
    idx: torch.Tensor           # (B,S,1) int64
    counts: torch.Tensor        # (K,) int64 code usage


def _rows(z: torch.Tensor, D: int) -> torch.Tensor:
    # VectorQuantizer.py:55 -- z.view(-1, e_dim); requires contiguity like the reference.
    return z.view(-1, D)


def distances_fp32(zf: torch.Tensor, E: torch.Tensor) -> torch.Tensor:
    """VectorQuantizer.py:59-61, evaluation order (A + B) - C, all fp32."""
    a = torch.sum(zf ** 2, dim=1, keepdim=True)
    b = torch.sum(E ** 2, dim=1)
    c = 2 * torch.matmul(zf, E.t())
    return a + b - c


def forward_fp32(z: torch.Tensor, E: torch.Tensor, beta: float, row_chunk: int = 8192) -> ForwardResult:
    """Forward of the reference layer, without the dense N x K one-hot.

    Follows VectorQuantizer.py:52-93.  The one-hot product of line 72 is a
    bit-exact gather (products by 1.0 / 0.0), line 84's column mean of the
    one-hot is bincount / N, so neither N x K temporary is needed; rows are
    processed in chunks so that large N fits in host memory.
    """
    assert z.dtype == torch.float32 and E.dtype == torch.float32
    K, D = E.shape
    zf = _rows(z, D)
    N = zf.shape[0]
    idx = torch.empty(N, dtype=torch.int64)
    for s in range(0, N, row_chunk):
        d = distances_fp32(zf[s:s + row_chunk], E)
        idx[s:s + row_chunk] = torch.argmin(d, dim=1)            # :65, ties -> lowest index
    q = E[idx].view(z.shape)                                      # :72 (gather == one-hot GEMM)
    m = torch.mean((q - z) ** 2)                                  # :76
    loss = m + beta * torch.mean((q - z) ** 2)                    # :76-77 (value; grads in backward_closed_form)
    z_q = z + (q - z)                                             # :80
    counts = torch.bincount(idx, minlength=K)
    return ForwardResult(loss=loss, z_q=z_q, perplexity=perplexity_from_counts(counts, N),
                         idx=idx.view(*z.shape[:-1], 1), counts=counts)


def perplexity_from_counts(counts: torch.Tensor, N: int) -> torch.Tensor:
    """VectorQuantizer.py:84-85 with e_mean = bincount / N (fp32)."""
    e_mean = counts.to(torch.float32) / float(N)
    return torch.exp(-torch.sum(e_mean * torch.log(e_mean + 1e-10)))


def onehot(idx: torch.Tensor, K: int) -> torch.Tensor:
    """VectorQuantizer.py:67-68, the dense (N,K) fp32 `min_encodings`."""
    flat = idx.reshape(-1, 1)
    out = torch.zeros(flat.shape[0], K)
    out.scatter_(1, flat, 1)
    return out


def backward_closed_form(z, E, idx, beta, g_zq=None, g_loss=None, n_global=None):
    """What autograd produces for VectorQuantizer.py:72-80.

    dz    = g_zq + g_loss * 2 (z - q) / (N D)             (commitment term, weight 1)
    dE[k] = g_loss * beta * 2 / (N D) * sum_{i: idx_i = k} (q_i - z_i)   (dense, exact zeros elsewhere)
    Accumulated in float64 so that it can referee fp32 implementations.
    """
    K, D = E.shape
    zf = _rows(z, D).double()
    flat = idx.reshape(-1)
    N = zf.shape[0]
    nd = float((n_global if n_global is not None else N) * D)
    q = E.double()[flat]
    gl = 0.0 if g_loss is None else float(g_loss)
    dz = gl * 2.0 * (zf - q) / nd
    if g_zq is not None:
        dz = dz + _rows(g_zq, D).double()
    dE = torch.zeros(K, D, dtype=torch.float64)
    dE.index_add_(0, flat, gl * beta * 2.0 * (q - zf) / nd)
    return dz.view(z.shape), dE


def truth_fp64(z: torch.Tensor, E: torch.Tensor, row_chunk: int = 4096):
    """Float64 squared distances -> (idx64, dmin64, second-best gap).  No |z|^2 term needed for argmin,
    but it is kept so that `dmin` is the true squared distance used in the tolerance."""
    K, D = E.shape
    zf = _rows(z, D).double()
    Ed = E.double()
    e2 = (Ed * Ed).sum(1)
    N = zf.shape[0]
    idx = torch.empty(N, dtype=torch.int64)
    dmin = torch.empty(N, dtype=torch.float64)
    gap = torch.empty(N, dtype=torch.float64)
    for s in range(0, N, row_chunk):
        zc = zf[s:s + row_chunk]
        d = (zc * zc).sum(1, keepdim=True) + e2 - 2.0 * (zc @ Ed.t())
        if K >= 2:
            two = torch.topk(d, 2, dim=1, largest=False)
            dmin[s:s + row_chunk] = two.values[:, 0]
            gap[s:s + row_chunk] = two.values[:, 1] - two.values[:, 0]
        else:
            dmin[s:s + row_chunk] = d[:, 0]
            gap[s:s + row_chunk] = math.inf
        idx[s:s + row_chunk] = torch.argmin(d, dim=1)
    return idx, dmin, gap


def dist_fp64_at(z: torch.Tensor, E: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """Exact (float64) squared distance of every row to the code `idx` names."""
    D = E.shape[1]
    diff = _rows(z, D).double() - E.double()[idx.reshape(-1)]
    return (diff * diff).sum(1)


@dataclass
class IndexParity:
    n: int
    raw_mismatch: int        # rows where ours != reference
    unexcused: int           # mismatching rows whose fp64 gap exceeds the tolerance
    max_gap_over_tol: float  # worst ratio gap / tol over mismatching rows (<= 1 means excused)

    @property
    def raw_rate(self) -> float:
        return self.raw_mismatch / max(self.n, 1)


def tf32_tolerance(z: torch.Tensor, E: torch.Tensor, idx_a: torch.Tensor, idx_b: torch.Tensor) -> torch.Tensor:
    """Per-row bound on |d64(i, a) - d64(i, b)| below which a tf32 search and the fp32 reference may
    legitimately disagree.

      tau_i = 2^-9 * |z_i| * max_k |E_k|          tf32 operand rounding of both dot products
                                                   (2 products x 2 operands x 2^-11, Cauchy-Schwarz)
            + 4 * ulp32(d_ref_i)                   the reference rounds distances of size |z|^2+|E|^2 to fp32
    """
    D = E.shape[1]
    zn = _rows(z, D).double().norm(dim=1)
    en = E.double().norm(dim=1).max()
    d_a = dist_fp64_at(z, E, idx_a)
    d_b = dist_fp64_at(z, E, idx_b)
    dmag = torch.maximum(d_a, d_b).clamp_min(1e-30)
    ulp = torch.pow(2.0, torch.floor(torch.log2(dmag)) - 23)
    return 2.0 ** -9 * zn * en + 4.0 * ulp


def index_parity(idx_ours: torch.Tensor, idx_ref: torch.Tensor, z: torch.Tensor, E: torch.Tensor,
                 exact_fp32: bool = False) -> IndexParity:
    """Compare code indices with the reference's; excuse only near-ties (see `tf32_tolerance`).
    With `exact_fp32` the tf32 operand term is dropped (fp32 CUDA-core search): only the reference's own
    fp32 rounding of the distances (a few ulp of d, and of the dot-product accumulation order) is excused."""
    a = idx_ours.reshape(-1).cpu()
    b = idx_ref.reshape(-1).cpu()
    bad = (a != b).nonzero().reshape(-1)
    if bad.numel() == 0:
        return IndexParity(a.numel(), 0, 0, 0.0)
    D = E.shape[1]
    zb = _rows(z, D)[bad]
    gap = (dist_fp64_at(zb, E, a[bad]) - dist_fp64_at(zb, E, b[bad])).abs()
    tol = tf32_tolerance(zb, E, a[bad], b[bad])
    if exact_fp32:
        dmag = torch.maximum(dist_fp64_at(zb, E, a[bad]), dist_fp64_at(zb, E, b[bad])).clamp_min(1e-30)
        ulp = torch.pow(2.0, torch.floor(torch.log2(dmag)) - 23)
        tol = (4.0 + math.sqrt(D)) * ulp
    ratio = gap / tol
    return IndexParity(a.numel(), int(bad.numel()), int((ratio > 1.0).sum()), float(ratio.max()))


def seq_acc(inp: torch.Tensor, target: torch.Tensor):
    """common/metrics.py:8-36 -- token accuracy over the batch and per sentence."""
    assert inp.shape == target.shape, "input and target shapes must match"
    assert not inp.is_floating_point() and not target.is_floating_point()
    same = (inp - target) == 0
    return same.sum() / inp.numel(), torch.mean(same.float(), dim=-1)


# ---- multi-GPU restatements (new capability; the reference is single device) ----------------------

def kshard_merge(z: torch.Tensor, E: torch.Tensor, shards: int) -> torch.Tensor:
    """K-sharded search: each shard reports (score, global index); the global winner is the lowest score,
    ties to the lowest index.  Scores are |E_k|^2 - 2 z.E_k in fp32 (the row-constant |z|^2 dropped)."""
    K, D = E.shape
    zf = _rows(z, D)
    best_s = torch.full((zf.shape[0],), math.inf)
    best_i = torch.zeros(zf.shape[0], dtype=torch.int64)
    per = (K + shards - 1) // shards
    for r in range(shards):
        Es = E[r * per:(r + 1) * per]
        if Es.shape[0] == 0:
            continue
        s = torch.sum(Es ** 2, dim=1) - 2 * torch.matmul(zf, Es.t())
        v, i = torch.min(s, dim=1)
        i = i + r * per
        take = (v < best_s) | ((v == best_s) & (i < best_i))
        best_s = torch.where(take, v, best_s)
        best_i = torch.where(take, i, best_i)
    return best_i


def dp_merge(parts, beta: float, N_global: int, D: int, K: int):
    """Batch-sharded VQ: combine per-rank (sum of squared residuals, counts) into global loss / perplexity."""
    sq = sum(float(p[0]) for p in parts)
    counts = sum(p[1] for p in parts)
    m = np.float32(sq / (N_global * D))
    loss = np.float32(m + np.float32(beta) * m)
    return torch.tensor(loss), perplexity_from_counts(counts, N_global)


# ---- literal CPU port of the reference step (the `cpu_baseline` / `--impl reference` arm of bench.py) -------

def literal_step_cpu(z: torch.Tensor, E: torch.Tensor, beta: float, g_zq: torch.Tensor):
    """One forward + backward exactly as the reference executes it, dense N x K temporaries included
    (VectorQuantizer.py:55-85 followed by autograd): distance matrix, argmin, host-built one-hot, one-hot GEMM
    lookup, two mean-squared terms, straight-through, perplexity; backward with g_zq upstream and unit weight on
    the loss.  This is what the reference costs on CPU -- `forward_fp32` above is a lighter restatement used for
    checking, this one is used for timing.  Returns (loss, perplexity, idx, dz, dE)."""
    K, D = E.shape
    zr = z.detach().clone().requires_grad_(True)
    W = E.detach().clone().requires_grad_(True)
    flat = zr.view(-1, D)
    d = torch.sum(flat ** 2, dim=1, keepdim=True) + torch.sum(W ** 2, dim=1) - 2 * torch.matmul(flat, W.t())
    nearest = torch.argmin(d, dim=1).unsqueeze(1)
    hot = torch.zeros(nearest.shape[0], K)
    hot.scatter_(1, nearest, 1)
    q = torch.matmul(hot, W).view(zr.shape)
    loss = torch.mean((q.detach() - zr) ** 2) + beta * torch.mean((q - zr.detach()) ** 2)
    out = zr + (q - zr).detach()
    usage = torch.mean(hot, dim=0)
    perplexity = torch.exp(-torch.sum(usage * torch.log(usage + 1e-10)))
    torch.autograd.backward([loss, out], [torch.ones(()), g_zq])
    return loss.detach(), perplexity.detach(), nearest.view(*z.shape[:-1], 1), zr.grad, W.grad


class LiteralVectorQuantizer(torch.nn.Module):
    """The reference layer restated as a module (same ops as `literal_step_cpu`, any device): the A/B partner of
    the kvq module in tests/harness_shelgon_step.py on machines where /root/reference is not mounted.  TEST INFRASTRUCTURE."""

    def __init__(self, n_e, e_dim, beta, vq_codebook_init_values=None):
        super().__init__()
        self.n_e, self.e_dim, self.beta = n_e, e_dim, beta
        self.embedding = torch.nn.Embedding(n_e, e_dim)
        if vq_codebook_init_values is not None:
            self.embedding.weight.data.copy_(vq_codebook_init_values)
        else:
            self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)

    def forward(self, z, device):
        W = self.embedding.weight
        flat = z.view(-1, self.e_dim)
        d = torch.sum(flat ** 2, dim=1, keepdim=True) + torch.sum(W ** 2, dim=1) - 2 * torch.matmul(flat, W.t())
        nearest = torch.argmin(d, dim=1).unsqueeze(1)
        hot = torch.zeros(nearest.shape[0], self.n_e).to(device)
        hot.scatter_(1, nearest, 1)
        q = torch.matmul(hot, W).view(z.shape)
        loss = torch.mean((q.detach() - z) ** 2) + self.beta * torch.mean((q - z.detach()) ** 2)
        out = z + (q - z).detach()
        usage = torch.mean(hot, dim=0)
        perplexity = torch.exp(-torch.sum(usage * torch.log(usage + 1e-10)))
        return loss, out, perplexity, hot, nearest.reshape(z.shape[0], z.shape[1], 1)


# ---- GumbelQuantizer (models/shelgon3/GumbelQuantizer.py:43-83), the alternative VQ_MODE -----------------------------

def gumbel_noise_like_reference(shape_bks, seed: int) -> torch.Tensor:
    """The Gumbel(0,1) sample F.gumbel_softmax draws for logits of shape (B, K, S) right after torch.manual_seed(seed)
    (torch/nn/functional.py: `-torch.empty_like(logits).exponential_().log()`), returned in (B, S, K) layout."""
    torch.manual_seed(seed)
    g = -torch.empty(shape_bks, dtype=torch.float32).exponential_().log()
    return g.permute(0, 2, 1).contiguous()


def gumbel_forward(z: torch.Tensor, W: torch.Tensor, b: torch.Tensor, E: torch.Tensor, noise: torch.Tensor, tau: float,
                   kld_scale: float, hard: bool):
    """GumbelQuantizer.forward restated on (B, S, .) row-major tensors with the Gumbel sample given explicitly.

      z (B,S,C), W (K,C) = proj.weight[:, :, 0], b (K) = proj.bias, E (K,D) = embed.weight, noise (B,S,K)
      logits = z W^T + b                                   :55  (1x1 Conv1d over the channel axis)
      y_soft = softmax((logits + noise) / tau) over K      :57  (F.gumbel_softmax)
      y      = one_hot(argmax y_soft) - sg(y_soft) + y_soft  if hard else y_soft
      z_q    = y E                                         :64  (einsum 'b n s, n d -> b d s')
      qy     = softmax(logits);  diff = kld_scale * mean_{b,s} sum_k qy log(qy K + 1e-10)      :68-71
      ind    = argmax_k y                                  :74
    Differentiable torch code: autograd of this function is the backward oracle."""
    K = W.shape[0]
    logits = torch.matmul(z, W.t()) + b
    gum = (logits + noise) / tau
    y_soft = gum.softmax(dim=-1)
    if hard:
        index = y_soft.max(dim=-1, keepdim=True)[1]
        y_hard = torch.zeros_like(logits).scatter_(-1, index, 1.0)
        y = y_hard - y_soft.detach() + y_soft
    else:
        y = y_soft
    z_q = torch.matmul(y, E)
    qy = torch.softmax(logits, dim=-1)
    diff = kld_scale * torch.sum(qy * torch.log(qy * K + 1e-10), dim=-1).mean()
    ind = y.argmax(dim=-1)
    return z_q, diff, ind
