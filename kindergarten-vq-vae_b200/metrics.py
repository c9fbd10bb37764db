"""Device-side `seq_acc` (reference: common/metrics.py:8-36)."""
from __future__ import annotations

import torch
from torch import Tensor

from . import _lib
from ._lib import check


def seq_acc(input: Tensor, target: Tensor):
    """Token accuracy over the whole batch and per sentence; same asserts and return pair as the reference."""
    assert input.shape == target.shape, "input and target shapes must match"
    assert not input.is_floating_point(), "input tensor must be integer type, not floating point"
    assert not target.is_floating_point(), "target tensor must be integer type, not floating point"
    if not input.is_cuda or not target.is_cuda:
        raise RuntimeError("seq_acc (kvq) runs on CUDA tensors only: there is no CPU fallback")
    a = input.to(torch.int64).contiguous()
    b = target.to(torch.int64).contiguous()
    S = a.shape[-1] if a.dim() > 0 else 1
    B = a.numel() // max(S, 1)
    acc = torch.empty((), dtype=torch.float32, device=a.device)
    per = torch.empty(a.shape[:-1], dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(_lib.load().kvq_seq_acc(a.data_ptr(), b.data_ptr(), B, S, acc.data_ptr(), per.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream), "kvq_seq_acc")
    return acc, per
