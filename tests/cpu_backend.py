"""CPU stand-in for the CUDA functional module, built on the oracle.  TEST INFRASTRUCTURE: tests/test_sharded_gloo.py
monkeypatches it over `sharded.F` so that the collective orchestration of the sharded layers can be exercised under gloo
without a GPU.  The product has no such seam."""
import ctypes

import numpy as np
import torch

from oracle import vq_oracle as O
from kindergarten_vq_vae_b200 import _lib


def _pack(scores: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()   # host helper of libkvq: same packing as the device code
    out = np.empty(scores.numel(), dtype=np.int64)
    s, i = scores.numpy(), index.numpy()
    for n in range(out.size):
        out[n] = lib.kvq_pack_key(ctypes.c_float(float(s[n])), ctypes.c_uint32(int(i[n])))
    return torch.from_numpy(out)


def search(z, E, *, mode="auto", k_offset=0, want_idx=True, want_keys=False, keys=None, keys_accumulate=False, ws=None):
    s = torch.sum(E ** 2, dim=1) - 2 * torch.matmul(z, E.t())
    i = torch.argmin(s, dim=1)                       # NaN counts as the minimum, like the kernels
    v = s.gather(1, i[:, None]).squeeze(1)
    idx = (i + k_offset) if want_idx else None
    k = _pack(v, i + k_offset) if (want_keys or keys is not None) else None
    if keys is not None:
        k = torch.minimum(k, keys)
    return idx, k


def vq_forward_partials(z, E, *, mode="auto", ws=None):
    idx, _ = search(z, E, mode=mode)
    z_q, sq, h = quantize(z, E, idx)
    return z_q, idx, sq, h


def pack_partials(sq_sum, hist):
    return torch.cat([sq_sum.double().reshape(1), hist.double()])


def finalize_packed(packed, n_global, D, beta):
    hist = packed[1:].to(torch.int32)
    loss, perp = finalize(packed[:1], hist, n_global, D, beta)
    return loss, perp, hist


def keys_to_idx(keys):
    return keys & 0xFFFFFFFF


def quantize(z, E, idx, *, k_offset=0, zero_skipped=False, sq_sum=None, hist=None):
    K = E.shape[0]
    local = idx - k_offset
    mine = (local >= 0) & (local < K)
    q = torch.zeros_like(z)
    q[mine] = E[local[mine]]
    d = torch.where(mine[:, None], q - z, torch.zeros_like(z))
    z_q = torch.where(mine[:, None], z + d, torch.zeros_like(z) if zero_skipped else z)
    sq = torch.tensor([float((d.double() ** 2).sum())], dtype=torch.float64)
    h = torch.bincount(local[mine], minlength=K).to(torch.int32)
    return z_q, sq, h


def finalize(sq_sum, hist, n_global, D, beta):
    m = np.float32(float(sq_sum) / (n_global * D))
    loss = torch.tensor(np.float32(m + np.float32(beta) * m))
    return loss, O.perplexity_from_counts(hist.long(), n_global)


def vq_backward(z, E, idx, hist, beta, *, g_zq=None, g_loss=None, need_dz=True, need_dE=True, k_offset=0,
                n_global=None, ws=None):
    K = E.shape[0]
    local = idx - k_offset
    mine = (local >= 0) & (local < K)
    n = z.shape[0] if n_global is None else n_global
    dz, dE = O.backward_closed_form(z[mine], E, local[mine], beta, g_zq=None if g_zq is None else g_zq[mine],
                                    g_loss=g_loss, n_global=n)
    out = None
    if need_dz:
        out = torch.zeros_like(z) if g_zq is None else g_zq.clone()
        out[mine] = dz.float()
    return out, (dE.float() if need_dE else None)


def dz_from_zq(z, z_q, g_zq, g_loss, n_global):
    c1 = float(g_loss) * 2.0 / (n_global * z.shape[1])
    return (0 if g_zq is None else g_zq) + c1 * (z - z_q)


def onehot(idx, K):
    return O.onehot(idx, K)
