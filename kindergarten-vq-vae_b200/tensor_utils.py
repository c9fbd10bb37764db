"""Device-side token-id corruption helpers (reference: common/tensor_utils.py:13-87).

The reference draws its randomness from the host RNG (`torch.randperm`, `random.shuffle`); here a seeded
counter-based generator runs on the device, so results are reproducible per seed but not bit-identical to the
reference's stream.  What is preserved exactly: the number of corrupted positions / slices, the value range,
the early return of the *same object* when percentage ~ 0, and the CUDA-only contract.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import Tensor

from . import _lib
from ._lib import check

_counter = [0]


def _seed(seed: Optional[int]) -> int:
    if seed is not None:
        return int(seed) & 0xFFFFFFFFFFFFFFFF
    _counter[0] += 1
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _counter[0]) & 0xFFFFFFFFFFFFFFFF


def _require_integer(tensor: Tensor, who: str) -> None:
    # the kernels work on int64 token ids; a float tensor would be silently truncated by the round trip
    if tensor.is_floating_point() or tensor.is_complex() or tensor.dtype == torch.bool:
        raise RuntimeError(f"{who} (kvq) corrupts integer token-id tensors; got dtype {tensor.dtype}")


def replace_pct_rand_values(tensor: Tensor, percentage: float, rand_int_low: int, rand_int_high: int,
                            seed: Optional[int] = None) -> Tensor:
    """Replace exactly int(numel*percentage) randomly chosen elements by uniform ints in [low, high)."""
    if math.isclose(percentage, 0):
        return tensor
    if tensor.get_device() < 0:
        raise RuntimeError("replace_pct_rand_values expects a CUDA tensor (as the reference does)")
    _require_integer(tensor, "replace_pct_rand_values")
    src = tensor.to(torch.int64).contiguous()
    out = torch.empty_like(src)
    with torch.cuda.device(src.device):
        check(_lib.load().kvq_replace_pct_rand_values(src.data_ptr(), src.numel(), float(percentage), rand_int_low,
                                                      rand_int_high, _seed(seed), out.data_ptr(),
                                                      torch.cuda.current_stream().cuda_stream),
              "kvq_replace_pct_rand_values")
    return out.to(tensor.dtype)


def change_percentage_of_elements(tensor: Tensor, dim, percentage, min, max, seed: Optional[int] = None) -> Tensor:
    """Overwrite int(size(dim)*percentage) random slices along `dim` (0 or 1) of a 2-D tensor, each with one
    random int in [min, max)."""
    if math.isclose(percentage, 0):
        return tensor
    if dim not in (0, 1):
        raise ValueError("Unsupported dimension")
    if tensor.get_device() < 0:
        raise RuntimeError("change_percentage_of_elements expects a CUDA tensor (as the reference does)")
    if tensor.dim() != 2:
        raise RuntimeError("change_percentage_of_elements expects a 2-D tensor")
    _require_integer(tensor, "change_percentage_of_elements")
    src = tensor.to(torch.int64).contiguous()
    out = torch.empty_like(src)
    R, C = src.shape
    with torch.cuda.device(src.device):
        check(_lib.load().kvq_change_percentage_of_elements(src.data_ptr(), R, C, dim, float(percentage), min, max,
                                                            _seed(seed), out.data_ptr(),
                                                            torch.cuda.current_stream().cuda_stream),
              "kvq_change_percentage_of_elements")
    return out.to(tensor.dtype)
