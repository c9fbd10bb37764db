"""Drop-in replacement for the reference's `VectorQuantizer` (models/shelgon3/VectorQuantizer.py:8-93).

Same class name (the caller dispatches on `type(vq).__name__ == "VectorQuantizer"`, models/shelgon3/Shelgon.py:57),
same constructor, same attributes (`n_e`, `e_dim`, `beta`, `embedding: nn.Embedding` -> state-dict key
`embedding.weight`), same `forward(z, device)` signature and the same 5-tuple of outputs.  The arithmetic runs in
libkvq's sm_100a kernels through the C ABI; there is no PyTorch or CPU fallback.

Single intentional deviation: the dense (N, K) one-hot `min_encodings` (4*N*K bytes; the only caller discards it,
Shelgon.py:58) is materialised only when it is small or explicitly requested, otherwise the slot holds None.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import functional as F

# materialise `min_encodings` automatically only below this many bytes
ONEHOT_AUTO_BYTES = 64 << 20


class _VQFunction(torch.autograd.Function):
    """forward(z_flat, E) -> (loss, z_q, perplexity, idx, hist); backward per SURVEY.md section 3.3."""

    @staticmethod
    def forward(ctx, z: Tensor, E: Tensor, beta: float, mode: str, ws: Optional[Tensor]):
        loss, z_q, perplexity, idx, hist = F.vq_forward(z, E, beta, mode=mode, ws=ws)
        # fresh 0-d tensors (not views of the 2-float result buffer): the training loop multiplies the loss in
        # place (models/shelgon3/Trainer.py:104)
        loss, perplexity = loss.clone(), perplexity.clone()
        ctx.save_for_backward(z, E, idx, hist)
        ctx.beta = beta
        ctx.ws = ws
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(perplexity, idx, hist)
        return loss, z_q, perplexity, idx, hist

    @staticmethod
    def backward(ctx, g_loss, g_zq, g_perp, g_idx, g_hist):
        z, E, idx, hist = ctx.saved_tensors
        need_dz, need_dE = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if g_loss is None and g_zq is None:
            return None, None, None, None, None
        if g_zq is not None:
            g_zq = g_zq.contiguous()
            if g_zq.dtype != torch.float32:
                g_zq = g_zq.float()
        if g_loss is not None:
            g_loss = g_loss.detach().to(torch.float32).contiguous()
        if g_loss is None:
            # loss unused: dz = g_zq exactly, dE = 0
            dz = g_zq if need_dz else None
            dE = torch.zeros_like(E) if need_dE else None
            return dz, dE, None, None, None
        dz, dE = F.vq_backward(z, E, idx, hist, ctx.beta, g_zq=g_zq, g_loss=g_loss, need_dz=need_dz,
                               need_dE=need_dE, ws=ctx.ws)
        return dz, dE, None, None, None


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
        self._ws = None

    def _workspace(self, N: int, device) -> Tensor:
        need = F._lib.load().kvq_workspace_bytes(N, self.e_dim, self.n_e)
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    @torch.compiler.disable
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
        ws = self._workspace(N, z.device)
        loss, z_q, perplexity, idx, _hist = _VQFunction.apply(z_flattened, weight, float(self.beta), self.search, ws)
        z_q = z_q.view(z.shape)

        want = self.return_min_encodings
        if want == "auto":
            want = N * self.n_e * 4 <= ONEHOT_AUTO_BYTES
        min_encodings = F.onehot(idx, self.n_e) if want else None    # VectorQuantizer.py:67-68

        min_encoding_indices = idx.reshape((batch_size, seq_len, 1))  # VectorQuantizer.py:90
        return loss, z_q, perplexity, min_encodings, min_encoding_indices
