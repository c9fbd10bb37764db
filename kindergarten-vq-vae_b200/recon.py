"""Fused reconstruction loss of the shelgon3 train step (reference: models/shelgon3/Trainer.py:94-101).

    loss_recon = kl_div(log_softmax(logits.reshape(-1, V), -1), one_hot(input_ids, V).reshape(-1, V).float(), "batchmean")
    recon_ids  = argmax(softmax(logits, -1), -1)
    acc        = seq_acc(recon_ids, input_ids)                      (common/metrics.py:8-36)

One kernel pass over the (B*S) x V logits produces all of it -- no dense one-hot, no softmax tensor -- and the backward is
one more pass (dlogits = g (softmax - one_hot) / (B*S)).  Registered as dispatcher operators so that a compiled training
step traces them as single nodes.  CUDA tensors only: there is no CPU fallback.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import check


def _check(logits: Tensor, input_ids: Tensor) -> Tuple[int, int, int]:
    if not logits.is_cuda or not input_ids.is_cuda:
        raise RuntimeError("recon_loss (kvq) runs on CUDA tensors only: there is no CPU fallback")
    if logits.dim() != 3 or input_ids.dim() != 2 or logits.shape[:2] != input_ids.shape:
        raise RuntimeError(f"recon_loss expects logits (B, S, V) and input_ids (B, S); got {tuple(logits.shape)} and "
                           f"{tuple(input_ids.shape)}")
    if logits.dtype != torch.float32:
        raise RuntimeError(f"logits must be float32, got {logits.dtype}")
    if input_ids.is_floating_point():
        raise RuntimeError("input_ids must be an integer tensor")
    B, S, V = logits.shape
    return B, S, V


@torch.library.custom_op("kvq::recon_loss_forward", mutates_args=())
def _recon_forward_op(logits: Tensor, input_ids: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    B, S, V = _check(logits, input_ids)
    x = logits.contiguous()
    ids = input_ids.to(torch.int64).contiguous()
    dev = x.device
    lib = _lib.load()
    scal = torch.empty(2, dtype=torch.float32, device=dev)
    recon = torch.empty(B, S, dtype=torch.int64, device=dev)
    per = torch.empty(B, dtype=torch.float32, device=dev)
    lse = torch.empty(B * S, dtype=torch.float32, device=dev)
    ws = torch.empty(lib.kvq_recon_workspace_bytes(B, S), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.kvq_recon_loss_forward(x.data_ptr(), ids.data_ptr(), B, S, V, scal.data_ptr(), recon.data_ptr(),
                                         scal.data_ptr() + 4, per.data_ptr(), lse.data_ptr(), ws.data_ptr(), ws.numel(),
                                         torch.cuda.current_stream().cuda_stream), "kvq_recon_loss_forward")
    # fresh 0-d tensors: the training loop scales the loss in place (Trainer.py:103)
    return scal[0].clone(), recon, scal[1].clone(), per, lse


@_recon_forward_op.register_fake
def _(logits, input_ids):
    B, S, _V = logits.shape
    return (logits.new_empty(()), logits.new_empty((B, S), dtype=torch.int64), logits.new_empty(()),
            logits.new_empty((B,)), logits.new_empty((B * S,)))


@torch.library.custom_op("kvq::recon_loss_backward", mutates_args=())
def _recon_backward_op(logits: Tensor, input_ids: Tensor, row_lse: Tensor, g_loss: Tensor) -> Tensor:
    B, S, V = logits.shape
    x = logits.contiguous()
    ids = input_ids.to(torch.int64).contiguous()
    g = g_loss.detach().to(torch.float32).contiguous()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(_lib.load().kvq_recon_loss_backward(x.data_ptr(), ids.data_ptr(), row_lse.data_ptr(), g.data_ptr(), B, S, V,
                                                  out.data_ptr(), torch.cuda.current_stream().cuda_stream),
              "kvq_recon_loss_backward")
    return out


@_recon_backward_op.register_fake
def _(logits, input_ids, row_lse, g_loss):
    return torch.empty_like(logits, memory_format=torch.contiguous_format)


def _setup(ctx, inputs, output):
    logits, input_ids = inputs
    ctx.save_for_backward(logits, input_ids, output[4])
    ctx.set_materialize_grads(False)


def _backward(ctx, g_loss, g_ids, g_acc, g_per, g_lse):
    if g_loss is None or not ctx.needs_input_grad[0]:
        return None, None
    logits, input_ids, lse = ctx.saved_tensors
    return _recon_backward_op(logits, input_ids, lse, g_loss), None


_recon_forward_op.register_autograd(_backward, setup_context=_setup)


def recon_loss(logits_recon: Tensor, input_ids: Tensor):
    """(loss_recon, recon_ids, acc, acc_per_sentence) of Trainer.py:94-101 from logits (B, S, V) and token ids (B, S).

    loss_recon is a differentiable 0-d tensor (gradient flows to the logits), recon_ids is (B, S) int64 -- the arg-max of
    the logits, which is the arg-max of their softmax except on ties that fp32 softmax rounding itself creates --, acc and
    acc_per_sentence are what common.metrics.seq_acc(recon_ids, input_ids) returns."""
    loss, recon_ids, acc, per, _lse = _recon_forward_op(logits_recon, input_ids)
    return loss, recon_ids, acc, per
