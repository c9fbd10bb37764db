"""Drop-in for the reference's `GumbelQuantizer` (models/shelgon3/GumbelQuantizer.py:15-83), the alternative VQ_MODE
(`models/shelgon3/main.py:68-73`; `Shelgon.forward` dispatches on the class name, models/shelgon3/Shelgon.py:60).

Same class name, constructor, attributes (`e_dim`, `n_embed`, `straight_through`, `temperature`, `kld_scale`,
`proj: nn.Conv1d(enc_out_size, n_embed, 1)`, `embed: nn.Embedding`) -> same state-dict keys; same
`forward(z, is_training) -> (z_q, diff, ind)`.  The arithmetic runs in libkvq: the dense contractions on the tcgen05
tf32 kernel (kvq_gemm_nt), everything per-row in the Gumbel row kernels.  CUDA only, no fallback.

Randomness: F.gumbel_softmax draws its sample from torch's global generator; here a counter-based device generator
keyed by a seed drawn from torch's CPU generator (so `torch.manual_seed` still makes runs repeatable).  Parity with the
reference is therefore distributional, and exact when the sample is passed in (`noise=`), which the tests do.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib
from ._lib import check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _up32(n: int) -> int:
    return (n + 31) // 32 * 32


def _gemm_nt(A: Tensor, B: Tensor, M: int, n: int, Kc: int, ldc: int, bias: Optional[Tensor] = None) -> Tensor:
    """C (M x ldc) = A (M x Kc) B^T (n x Kc) + bias on the tensor cores (tf32, fp32 accumulate)."""
    C = torch.empty(M, ldc, dtype=torch.float32, device=A.device)
    lib = _lib.load()
    nbytes = lib.kvq_gemm_nt_workspace_bytes(M, n, Kc, ldc)       # > 0: few output tiles, long contraction -> split
    ws = torch.empty(nbytes, dtype=torch.uint8, device=A.device) if nbytes else None
    check(lib.kvq_gemm_nt(A.data_ptr(), B.data_ptr(), M, n, Kc, C.data_ptr(), ldc,
                          None if bias is None else bias.data_ptr(), 1.0, None if ws is None else ws.data_ptr(), nbytes,
                          _stream()), "kvq_gemm_nt")
    return C


def _transpose_pad(src: Tensor, R: int, C: int, lds: int, ldd: int) -> Tensor:
    """(C x ldd) transpose of the (R x C) matrix inside src (leading dimension lds); columns [R, ldd) zero."""
    dst = torch.empty(C, ldd, dtype=torch.float32, device=src.device)
    check(_lib.load().kvq_transpose_pad(src.data_ptr(), R, C, lds, dst.data_ptr(), ldd, _stream()), "kvq_transpose_pad")
    return dst


class _GumbelFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z: Tensor, W: Tensor, b: Tensor, E: Tensor, tau: float, kld_scale: float, hard: bool,
                noise: Optional[Tensor], seed: int):
        lib = _lib.load()
        B_, S, C = z.shape
        K, D = E.shape
        N, Kp = B_ * S, _up32(K)
        X = z.reshape(N, C).contiguous()
        Wm = W.reshape(K, C).contiguous()                   # Conv1d weight (K, C, 1)
        with torch.cuda.device(z.device):
            logits = _gemm_nt(X, Wm, N, K, C, Kp, bias=b.contiguous())                  # GumbelQuantizer.py:55
            y = torch.empty(N, Kp, dtype=torch.float32, device=z.device)
            ind = torch.empty(N, dtype=torch.int64, device=z.device)
            diff = torch.empty((), dtype=torch.float32, device=z.device)
            kl_row = torch.empty(N, dtype=torch.float32, device=z.device)
            check(lib.kvq_gumbel_rows_forward(logits.data_ptr(), None if noise is None else noise.data_ptr(), seed, N, K, Kp,
                                              float(tau), float(kld_scale), int(hard), y.data_ptr(), ind.data_ptr(),
                                              diff.data_ptr(), kl_row.data_ptr(), _stream()), "kvq_gumbel_rows_forward")
            if hard:    # every weight but the arg-max one is an exact zero: the einsum of :64 is a scaled gather
                z_q = torch.empty(N, D, dtype=torch.float32, device=z.device)
                check(lib.kvq_gumbel_hard_gather(y.data_ptr(), ind.data_ptr(), E.data_ptr(), N, D, Kp, z_q.data_ptr(),
                                                 _stream()), "kvq_gumbel_hard_gather")
            else:
                ET = _transpose_pad(E, K, D, D, Kp)                                     # (D x Kp)
                z_q = _gemm_nt(y, ET, N, D, Kp, D)                                      # :64
        ctx.save_for_backward(X, Wm, E, logits, y, noise if noise is not None else X.new_empty(0))
        ctx.meta = (B_, S, C, K, D, N, Kp, float(tau), float(kld_scale), noise is not None, seed)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(ind)
        return z_q.view(B_, S, D), diff, ind.view(B_, S)

    @staticmethod
    def backward(ctx, g_zq, g_diff, _g_ind):
        X, Wm, E, logits, y, noise = ctx.saved_tensors
        B_, S, C, K, D, N, Kp, tau, kld_scale, has_noise, seed = ctx.meta
        if g_zq is None and g_diff is None:
            return (None,) * 9
        lib = _lib.load()
        Np = _up32(N)
        with torch.cuda.device(X.device):
            dy = None
            if g_zq is not None:
                G = g_zq.reshape(N, D).contiguous().float()
                dy = _gemm_nt(G, E, N, K, D, Kp)                                        # d y = g_zq E^T
            gd = None if g_diff is None else g_diff.detach().to(torch.float32).contiguous()
            dL = torch.empty(N, Kp, dtype=torch.float32, device=X.device)
            check(lib.kvq_gumbel_rows_backward(logits.data_ptr(), noise.data_ptr() if has_noise else None, seed,
                                               None if dy is None else dy.data_ptr(), None if gd is None else gd.data_ptr(),
                                               N, K, Kp, tau, kld_scale, dL.data_ptr(), _stream()), "kvq_gumbel_rows_backward")
            dz = dW = db = dE = None
            if ctx.needs_input_grad[0]:
                WT = _transpose_pad(Wm, K, C, C, Kp)                                    # (C x Kp)
                dz = _gemm_nt(dL, WT, N, C, Kp, C).view(B_, S, C)                       # dL W
            if ctx.needs_input_grad[1]:
                dLT = _transpose_pad(dL, N, K, Kp, Np)                                  # (K x Np)
                XT = _transpose_pad(X, N, C, C, Np)                                     # (C x Np)
                dW = _gemm_nt(dLT, XT, K, C, Np, C).view(K, C, 1)                       # dL^T z
            if ctx.needs_input_grad[2]:
                db = torch.empty(K, dtype=torch.float32, device=X.device)
                cws = torch.empty(lib.kvq_colsum_workspace_bytes(N, K), dtype=torch.uint8, device=X.device)
                check(lib.kvq_colsum(dL.data_ptr(), N, K, Kp, db.data_ptr(), cws.data_ptr(), cws.numel(), _stream()),
                      "kvq_colsum")
            if ctx.needs_input_grad[3]:
                if g_zq is None:
                    dE = torch.zeros_like(E)
                else:
                    yT = _transpose_pad(y, N, K, Kp, Np)                                # (K x Np)
                    GT = _transpose_pad(G, N, D, D, Np)                                 # (D x Np)
                    dE = _gemm_nt(yT, GT, K, D, Np, D)                                  # y^T g_zq
        return dz, dW, db, dE, None, None, None, None, None


class GumbelQuantizer(nn.Module):
    """
    Gumbel Softmax trick quantizer (B200-native).
    Categorical Reparameterization with Gumbel-Softmax, Jang et al. 2016, https://arxiv.org/abs/1611.01144
    """

    def __init__(self, enc_out_size, n_embed, embedding_dim, temperature: float, kl_div_scale: float,
                 straight_through: bool):
        super().__init__()
        self.e_dim = embedding_dim
        self.n_embed = n_embed
        self.straight_through = straight_through
        self.temperature = temperature
        self.kld_scale = kl_div_scale
        self.proj = nn.Conv1d(enc_out_size, n_embed, 1)
        self.embed = nn.Embedding(n_embed, embedding_dim)

    @torch.compiler.disable
    def forward(self, z: Tensor, is_training: bool, noise: Optional[Tensor] = None, seed: Optional[int] = None):
        """z (batch, seq_len, enc_out_size) -> (z_q (batch, seq_len, e_dim), diff 0-d, ind (batch, seq_len) int64).
        `noise`: optional explicit Gumbel(0,1) sample of shape (batch, seq_len, n_embed); `seed`: optional seed of the
        device generator used otherwise."""
        if z.dim() != 3:
            raise RuntimeError(f"z must be (batch, seq_len, enc_out_size), got shape {tuple(z.shape)}")
        W, b, E = self.proj.weight, self.proj.bias, self.embed.weight
        if not z.is_cuda or not W.is_cuda or not E.is_cuda:
            raise RuntimeError("GumbelQuantizer (kvq) runs on CUDA only: there is no CPU fallback")
        C, D = z.shape[-1], self.e_dim
        if C != W.shape[1]:
            raise RuntimeError(f"z has {C} channels but proj expects {W.shape[1]}")
        if C % 32 or D % 32:
            raise RuntimeError(f"GumbelQuantizer (kvq) needs enc_out_size and embedding_dim to be multiples of 32 "
                               f"(tcgen05 tf32 contraction granule); got {C} and {D}")
        if z.dtype != torch.float32:
            raise RuntimeError(f"z must be float32, got {z.dtype}")
        hard = self.straight_through if is_training else True      # eval must quantise (:52)
        if noise is not None:
            if tuple(noise.shape) != (z.shape[0], z.shape[1], self.n_embed):
                raise RuntimeError("noise must have shape (batch, seq_len, n_embed)")
            noise = noise.to(device=z.device, dtype=torch.float32).contiguous()
        if seed is None:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
        return _GumbelFn.apply(z.contiguous(), W, b, E.contiguous(), float(self.temperature), float(self.kld_scale),
                               bool(hard), noise, int(seed) & 0xFFFFFFFFFFFFFFFF)
