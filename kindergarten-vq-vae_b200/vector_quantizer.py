"""Drop-in replacement for the reference's `VectorQuantizer` (models/shelgon3/VectorQuantizer.py:8-93).

Same class name (the caller dispatches on `type(vq).__name__ == "VectorQuantizer"`, models/shelgon3/Shelgon.py:57),
same constructor, same attributes (`n_e`, `e_dim`, `beta`, `embedding: nn.Embedding` -> state-dict key
`embedding.weight`), same `forward(z, device)` signature and the same 5-tuple of outputs.  The arithmetic runs in
libkvq's sm_100a kernels through the C ABI; there is no PyTorch or CPU fallback.

Single intentional deviation: the dense (N, K) one-hot `min_encodings` (4*N*K bytes; the only caller discards it,
Shelgon.py:58) is materialised only when it is small or explicitly requested, otherwise the slot holds None.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import functional as F

# materialise `min_encodings` automatically only below this many bytes
ONEHOT_AUTO_BYTES = 64 << 20


# ---- the layer as two dispatcher-registered operators -------------------------------------------------------------
# `torch.library.custom_op` (with fake / meta implementations and a registered autograd formula) instead of a bare
# autograd.Function: the reference wraps the whole model in `model.compile()` (models/shelgon3/main.py:83), and an
# opaque custom op is traced as ONE graph node -- no graph break around the layer, `fullgraph=True` works.
# Each op is a thin call into the C ABI; nothing here computes.

@torch.library.custom_op("kvq::vq_forward", mutates_args=())
def _vq_forward_op(z: Tensor, E: Tensor, beta: float, mode: str) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    # loss and perplexity come back as fresh 0-d tensors (not views): the training loop multiplies the loss in place
    # (models/shelgon3/Trainer.py:104)
    return F.vq_forward(z, E, beta, mode=mode)


@_vq_forward_op.register_fake
def _(z, E, beta, mode):
    N, K = z.shape[0], E.shape[0]
    return (z.new_empty(()), torch.empty_like(z), z.new_empty(()), z.new_empty((N,), dtype=torch.int64),
            z.new_empty((K,), dtype=torch.int32))


@torch.library.custom_op("kvq::vq_backward", mutates_args=())
def _vq_backward_op(z: Tensor, E: Tensor, idx: Tensor, hist: Tensor, g_zq: Optional[Tensor], g_loss: Tensor, beta: float,
                    need_dz: bool, need_dE: bool) -> Tuple[Tensor, Tensor]:
    if g_zq is not None:
        g_zq = g_zq.contiguous()
        if g_zq.dtype != torch.float32:
            g_zq = g_zq.float()
    g_loss = g_loss.detach().to(torch.float32).contiguous()
    dz, dE = F.vq_backward(z, E, idx, hist, beta, g_zq=g_zq, g_loss=g_loss, need_dz=need_dz, need_dE=need_dE)
    # an operator cannot return None: an unrequested gradient comes back as an empty tensor
    return (dz if dz is not None else z.new_empty(0)), (dE if dE is not None else E.new_empty(0))


@_vq_backward_op.register_fake
def _(z, E, idx, hist, g_zq, g_loss, beta, need_dz, need_dE):
    return (torch.empty_like(z) if need_dz else z.new_empty(0)), (torch.empty_like(E) if need_dE else E.new_empty(0))


@torch.library.custom_op("kvq::onehot", mutates_args=())
def _onehot_op(idx: Tensor, n_e: int) -> Tensor:
    return F.onehot(idx, n_e)


@_onehot_op.register_fake
def _(idx, n_e):
    return idx.new_empty((idx.numel(), n_e), dtype=torch.float32)


def _vq_setup_context(ctx, inputs, output):
    z, E, beta, _mode = inputs
    _loss, _z_q, _perp, idx, hist = output
    ctx.save_for_backward(z, E, idx, hist)
    ctx.beta = beta
    ctx.set_materialize_grads(False)


def _vq_backward(ctx, g_loss, g_zq, g_perp, g_idx, g_hist):
    """dz = g_zq + g_loss 2 (z - q)/(N D);  dE[k] = g_loss beta 2/(N D) sum_{i in k} (q_i - z_i)."""
    z, E, idx, hist = ctx.saved_tensors
    need_dz, need_dE = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    if g_loss is None and g_zq is None:
        return None, None, None, None
    if g_loss is None:
        # loss unused: dz = g_zq exactly, dE = 0
        return (g_zq if need_dz else None), (torch.zeros_like(E) if need_dE else None), None, None
    dz, dE = _vq_backward_op(z, E, idx, hist, g_zq, g_loss, ctx.beta, need_dz, need_dE)
    return (dz if need_dz else None), (dE if need_dE else None), None, None


_vq_forward_op.register_autograd(_vq_backward, setup_context=_vq_setup_context)


class _VQFunction(torch.autograd.Function):
    """The same forward / backward as the dispatcher operators above, as a plain autograd.Function: used in eager mode,
    where the operator dispatch costs ~0.1 ms per step -- as much as the whole layer at the reference's own shapes."""

    @staticmethod
    def forward(ctx, z: Tensor, E: Tensor, beta: float, mode: str):
        loss, z_q, perplexity, idx, hist = F.vq_forward(z, E, beta, mode=mode)   # loss / perplexity: fresh 0-d tensors
        ctx.save_for_backward(z, E, idx, hist)
        ctx.beta = beta
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(perplexity, idx, hist)
        return loss, z_q, perplexity, idx, hist

    @staticmethod
    def backward(ctx, g_loss, g_zq, g_perp, g_idx, g_hist):
        z, E, idx, hist = ctx.saved_tensors
        need_dz, need_dE = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if g_loss is None and g_zq is None:
            return None, None, None, None
        if g_loss is None:
            return (g_zq if need_dz else None), (torch.zeros_like(E) if need_dE else None), None, None
        if g_zq is not None:
            g_zq = g_zq.contiguous()
            if g_zq.dtype != torch.float32:
                g_zq = g_zq.float()
        g_loss = g_loss.detach().to(torch.float32).contiguous()
        dz, dE = F.vq_backward(z, E, idx, hist, ctx.beta, g_zq=g_zq, g_loss=g_loss, need_dz=need_dz, need_dE=need_dE)
        return dz, dE, None, None


class _EmptyBatchFunction(torch.autograd.Function):
    """A batch with no latents.  The reference then takes means over zero elements (VectorQuantizer.py:76-85): loss and
    perplexity are NaN, every other output is empty, and autograd gives the codebook an all-zero gradient."""

    @staticmethod
    def forward(ctx, z: Tensor, E: Tensor):
        ctx.save_for_backward(E)
        ctx.set_materialize_grads(False)
        nan = torch.full((), float("nan"), dtype=torch.float32, device=z.device)
        idx = torch.empty(0, dtype=torch.int64, device=z.device)
        ctx.mark_non_differentiable(idx)
        return nan, torch.empty_like(z), nan.clone().detach(), idx

    @staticmethod
    def backward(ctx, g_loss, g_zq, g_perp, g_idx):
        (E,) = ctx.saved_tensors
        return (g_zq if ctx.needs_input_grad[0] else None), (torch.zeros_like(E) if ctx.needs_input_grad[1] else None)


class VectorQuantizer(nn.Module):
    """
    Discretization bottleneck part of the VQ-VAE (B200-native).

    Inputs:
    - n_e : number of embeddings
    - e_dim : dimension of embedding
    - beta : weight of the codebook term  beta * ||z_q - sg[z]||^2  (the reference's placement,
             VectorQuantizer.py:76-77: the commitment term has weight 1)
    - vq_codebook_init_values : optional (n_e, e_dim) initial codebook (e.g. from k-means)

    Extra keyword-only options (not in the reference):
    - search : "auto" | "tf32" | "fp32" | "tf32_refine" -- precision of the nearest-code search ("tf32_refine":
               tensor-core search for the two best codes, exact float64 re-evaluation of that pair)
    - min_encodings : "auto" | True | False -- when to materialise the dense one-hot output
    """

    def __init__(self, n_e, e_dim, beta, vq_codebook_init_values: Tensor = None, *, search: str = "auto",
                 min_encodings="auto"):
        super(VectorQuantizer, self).__init__()
        self.n_e = n_e
        self.e_dim = e_dim
        self.beta = beta
        if search not in ("auto", "tf32", "fp32", "tf32_refine"):
            raise ValueError(f"search must be auto|tf32|fp32|tf32_refine, got {search!r}")
        self.search = search
        self.return_min_encodings = min_encodings

        self.embedding = nn.Embedding(self.n_e, self.e_dim)
        if vq_codebook_init_values is not None:
            self.embedding.weight.data.copy_(vq_codebook_init_values)
        else:
            self.embedding.weight.data.uniform_(-1.0 / self.n_e, 1.0 / self.n_e)

    def forward(self, z: torch.Tensor, device=None):
        """
        z (continuous) -> z_q (discrete);  z.shape = (batch, seq_len, channel), channel == e_dim.

        Returns (loss, z_q, perplexity, min_encodings, min_encoding_indices) like the reference:
        loss 0-d fp32 (differentiable), z_q like z (gradient passes straight through to z),
        perplexity 0-d fp32, min_encodings (N, n_e) fp32 one-hot or None, indices (batch, seq_len, 1) int64.
        `device` is accepted for signature compatibility and ignored (outputs live where z lives).
        """
        if z.dim() != 3:
            raise RuntimeError(f"z must be (batch, seq_len, e_dim), got shape {tuple(z.shape)}")
        batch_size, seq_len, _ = z.shape
        z_flattened = z.view((-1, self.e_dim))            # VectorQuantizer.py:55 (raises on non-contiguous z)
        weight = self.embedding.weight
        if not z.is_cuda or not weight.is_cuda:
            raise RuntimeError("VectorQuantizer (kvq) runs on CUDA only: there is no CPU fallback "
                               f"(z on {z.device}, codebook on {weight.device})")
        N = z_flattened.shape[0]
        if N == 0:
            loss, z_q, perplexity, idx = _EmptyBatchFunction.apply(z_flattened, weight)
        elif torch.compiler.is_compiling():      # traced (model.compile(), main.py:83): one opaque dispatcher node
            loss, z_q, perplexity, idx, _hist = _vq_forward_op(z_flattened, weight, float(self.beta), self.search)
        else:                                  # eager: same C-ABI calls without the operator-dispatch overhead
            loss, z_q, perplexity, idx, _hist = _VQFunction.apply(z_flattened, weight, float(self.beta), self.search)
        z_q = z_q.view(z.shape)

        want = self.return_min_encodings
        if want == "auto":
            want = N * self.n_e * 4 <= ONEHOT_AUTO_BYTES
        if not want:
            min_encodings = None
        elif N == 0:
            min_encodings = z.new_zeros((0, self.n_e), dtype=torch.float32)
        else:                                                                    # VectorQuantizer.py:67-68
            min_encodings = _onehot_op(idx, self.n_e) if torch.compiler.is_compiling() else F.onehot(idx, self.n_e)

        min_encoding_indices = idx.reshape((batch_size, seq_len, 1))  # VectorQuantizer.py:90
        return loss, z_q, perplexity, min_encodings, min_encoding_indices
