"""Tensor-level wrappers over the C ABI (include/kvq.h).  PyTorch here is plumbing only: it owns device memory
and the current stream; every arithmetic step of the layer runs in libkvq's sm_100a kernels.

All functions require CUDA tensors and raise on anything else -- there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import SEARCH_MODES, check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, name: str, dtype: torch.dtype) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: kvq has no CPU fallback (got device {t.device})")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")  # the reference's z.view(-1, e_dim) has the same demand
    return t


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_WS_BYTES = {}


def workspace(N: int, D: int, K: int, device) -> torch.Tensor:
    key = (N, D, K)
    nbytes = _WS_BYTES.get(key)
    if nbytes is None:
        nbytes = _WS_BYTES[key] = _lib.load().kvq_workspace_bytes(N, D, K)
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


class _on_device:
    """`with _on_device(d)` only when d is not already current (the context manager costs ~10 us per use, which
    matters when the whole layer takes 0.1 ms at the reference's shapes)."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if device.index is None or device.index == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def device_info() -> Tuple[int, int, int]:
    import ctypes
    a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(_lib.load().kvq_device_info(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)), "kvq_device_info")
    return a.value, b.value, c.value


def code_norms(E: torch.Tensor, K_pad: Optional[int] = None) -> torch.Tensor:
    """|E_k|^2 (VectorQuantizer.py:60); entries beyond K (padding) are +inf."""
    _req(E, "E", torch.float32)
    K, D = E.shape
    K_pad = K if K_pad is None else K_pad
    out = torch.empty(K_pad, dtype=torch.float32, device=E.device)
    with _on_device(E.device):
        check(_lib.load().kvq_code_norms(E.data_ptr(), K, D, out.data_ptr(), K_pad, _stream()), "kvq_code_norms")
    return out


def search(z: torch.Tensor, E: torch.Tensor, *, mode: str = "auto", k_offset: int = 0, want_idx: bool = True,
           keys: Optional[torch.Tensor] = None, keys_accumulate: bool = False, want_keys: bool = False,
           ws: Optional[torch.Tensor] = None):
    """Fused distance + argmin (VectorQuantizer.py:59-65).  z: (N,D) fp32, E: (K,D) fp32.
    Returns (idx or None, keys or None)."""
    _req(z, "z", torch.float32)
    _req(E, "E", torch.float32)
    N, D = z.shape
    K = E.shape[0]
    if E.shape[1] != D:
        raise RuntimeError(f"z has {D} features but the codebook has {E.shape[1]}")
    idx = torch.empty(N, dtype=torch.int64, device=z.device) if want_idx else None
    if keys is None and (want_keys or keys_accumulate):
        keys = torch.full((N,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=z.device)
        keys_accumulate = True
    if keys is not None:
        _req(keys, "keys", torch.int64)
    if ws is None:
        ws = workspace(N, D, K, z.device)
    with _on_device(z.device):
        check(_lib.load().kvq_search(z.data_ptr(), E.data_ptr(), N, D, K, k_offset, SEARCH_MODES[mode], _ptr(idx),
                                     _ptr(keys), int(bool(keys_accumulate)), ws.data_ptr(), ws.numel(), _stream()),
              "kvq_search")
    return idx, keys


def keys_to_idx(keys: torch.Tensor) -> torch.Tensor:
    _req(keys, "keys", torch.int64)
    idx = torch.empty_like(keys)
    with _on_device(keys.device):
        check(_lib.load().kvq_keys_to_idx(keys.data_ptr(), keys.numel(), idx.data_ptr(), _stream()), "kvq_keys_to_idx")
    return idx


def quantize(z: torch.Tensor, E: torch.Tensor, idx: torch.Tensor, *, k_offset: int = 0, zero_skipped: bool = False,
             sq_sum: Optional[torch.Tensor] = None, hist: Optional[torch.Tensor] = None):
    """Gather + straight-through + squared-residual sum + usage histogram (VectorQuantizer.py:67-84)."""
    _req(z, "z", torch.float32); _req(E, "E", torch.float32); _req(idx, "idx", torch.int64)
    N, D = z.shape
    K = E.shape[0]
    z_q = torch.empty_like(z)
    if sq_sum is None:
        sq_sum = torch.zeros(1, dtype=torch.float64, device=z.device)
    if hist is None:
        hist = torch.zeros(K, dtype=torch.int32, device=z.device)
    with _on_device(z.device):
        check(_lib.load().kvq_quantize(z.data_ptr(), E.data_ptr(), idx.data_ptr(), N, D, K, k_offset,
                                       int(zero_skipped), z_q.data_ptr(), sq_sum.data_ptr(), hist.data_ptr(),
                                       _stream()), "kvq_quantize")
    return z_q, sq_sum, hist


def finalize(sq_sum: torch.Tensor, hist: torch.Tensor, n_global: int, D: int, beta: float):
    _req(sq_sum, "sq_sum", torch.float64); _req(hist, "hist", torch.int32)
    loss = torch.empty((), dtype=torch.float32, device=hist.device)      # fresh 0-d tensors, not views: the training loop
    perp = torch.empty((), dtype=torch.float32, device=hist.device)      # scales the loss in place (Trainer.py:104)
    with _on_device(hist.device):
        check(_lib.load().kvq_finalize(sq_sum.data_ptr(), hist.data_ptr(), n_global, D, hist.numel(), beta,
                                       loss.data_ptr(), perp.data_ptr(), _stream()), "kvq_finalize")
    return loss, perp


def vq_forward(z: torch.Tensor, E: torch.Tensor, beta: float, *, mode: str = "auto",
               ws: Optional[torch.Tensor] = None):
    """Whole forward (VectorQuantizer.py:52-93).  z: (N,D).  Returns loss, z_q, perplexity, idx (N,), hist (K,)."""
    _req(z, "z", torch.float32); _req(E, "E", torch.float32)
    N, D = z.shape
    K = E.shape[0]
    if E.shape[1] != D:
        raise RuntimeError(f"z has {D} features but the codebook has {E.shape[1]}")
    z_q = torch.empty_like(z)
    idx = torch.empty(N, dtype=torch.int64, device=z.device)
    loss = torch.empty((), dtype=torch.float32, device=z.device)         # fresh 0-d tensors, not views of one buffer: the
    perp = torch.empty((), dtype=torch.float32, device=z.device)         # training loop scales the loss in place (Trainer.py:104)
    hist = torch.empty(K, dtype=torch.int32, device=z.device)
    if ws is None:
        ws = workspace(N, D, K, z.device)
    with _on_device(z.device):
        check(_lib.load().kvq_forward(z.data_ptr(), E.data_ptr(), N, D, K, float(beta), SEARCH_MODES[mode],
                                      z_q.data_ptr(), idx.data_ptr(), loss.data_ptr(), perp.data_ptr(),
                                      hist.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "kvq_forward")
    return loss, z_q, perp, idx, hist


def vq_forward_partials(z: torch.Tensor, E: torch.Tensor, *, mode: str = "auto", ws: Optional[torch.Tensor] = None):
    """Forward without the finalisation (batch-sharded layer): returns z_q, idx (N,), sq_sum (1 float64), hist (K int32);
    the caller all-reduces the last two over the ranks and calls `finalize` with the global latent count."""
    _req(z, "z", torch.float32); _req(E, "E", torch.float32)
    N, D = z.shape
    K = E.shape[0]
    if E.shape[1] != D:
        raise RuntimeError(f"z has {D} features but the codebook has {E.shape[1]}")
    z_q = torch.empty_like(z)
    idx = torch.empty(N, dtype=torch.int64, device=z.device)
    sq_sum = torch.empty(1, dtype=torch.float64, device=z.device)
    hist = torch.empty(K, dtype=torch.int32, device=z.device)
    if ws is None:
        ws = workspace(N, D, K, z.device)
    with _on_device(z.device):
        check(_lib.load().kvq_forward_partials(z.data_ptr(), E.data_ptr(), N, D, K, SEARCH_MODES[mode], z_q.data_ptr(),
                                               idx.data_ptr(), sq_sum.data_ptr(), hist.data_ptr(), ws.data_ptr(),
                                               ws.numel(), _stream()), "kvq_forward_partials")
    return z_q, idx, sq_sum, hist


def pack_partials(sq_sum: torch.Tensor, hist: torch.Tensor) -> torch.Tensor:
    """[sq_sum | hist] as one float64 buffer of 1 + K entries: the batch-sharded forward all-reduces this once."""
    _req(sq_sum, "sq_sum", torch.float64); _req(hist, "hist", torch.int32)
    K = hist.numel()
    packed = torch.empty(K + 1, dtype=torch.float64, device=hist.device)
    with _on_device(hist.device):
        check(_lib.load().kvq_pack_partials(sq_sum.data_ptr(), hist.data_ptr(), K, packed.data_ptr(), _stream()),
              "kvq_pack_partials")
    return packed


def finalize_packed(packed: torch.Tensor, n_global: int, D: int, beta: float):
    """loss, perplexity and the global usage histogram from the all-reduced buffer of `pack_partials`."""
    _req(packed, "packed", torch.float64)
    K = packed.numel() - 1
    loss = torch.empty((), dtype=torch.float32, device=packed.device)
    perp = torch.empty((), dtype=torch.float32, device=packed.device)
    hist = torch.empty(K, dtype=torch.int32, device=packed.device)
    with _on_device(packed.device):
        check(_lib.load().kvq_finalize_packed(packed.data_ptr(), n_global, D, K, beta, loss.data_ptr(), perp.data_ptr(),
                                              hist.data_ptr(), _stream()), "kvq_finalize_packed")
    return loss, perp, hist


def vq_backward(z: torch.Tensor, E: torch.Tensor, idx: torch.Tensor, hist: Optional[torch.Tensor], beta: float, *,
                g_zq: Optional[torch.Tensor] = None, g_loss: Optional[torch.Tensor] = None, need_dz: bool = True,
                need_dE: bool = True, k_offset: int = 0, n_global: Optional[int] = None,
                ws: Optional[torch.Tensor] = None):
    """Backward (autograd of VectorQuantizer.py:72-80).  Returns (dz or None, dE or None).

    With a code shard (`k_offset` / E covering part of the codebook) ask for dz and dE in separate calls: the dE pass
    visits only the latents whose code lies in the shard, the dz-only call passes g_zq through for the others."""
    _req(z, "z", torch.float32); _req(E, "E", torch.float32); _req(idx, "idx", torch.int64)
    N, D = z.shape
    K = E.shape[0]
    if need_dz and need_dE and k_offset != 0:
        raise RuntimeError("vq_backward: with a code shard (k_offset != 0) request dz and dE in separate calls")
    if g_zq is not None:
        _req(g_zq, "g_zq", torch.float32)
    if g_loss is not None:
        _req(g_loss, "g_loss", torch.float32)
    dz = torch.empty_like(z) if need_dz else None
    dE = torch.empty_like(E) if need_dE else None
    if need_dE:
        if hist is None:
            raise RuntimeError("vq_backward: the forward histogram is required for the codebook gradient")
        _req(hist, "hist", torch.int32)
        if ws is None:
            ws = workspace(N, D, K, z.device)
    with _on_device(z.device):
        check(_lib.load().kvq_backward(z.data_ptr(), E.data_ptr(), idx.data_ptr(), _ptr(hist), _ptr(g_zq),
                                       _ptr(g_loss), N, D, K, k_offset, float(beta),
                                       N if n_global is None else n_global, _ptr(dz), _ptr(dE), _ptr(ws),
                                       0 if ws is None else ws.numel(), _stream()), "kvq_backward")
    return dz, dE


def dz_from_zq(z: torch.Tensor, z_q: torch.Tensor, g_zq: Optional[torch.Tensor], g_loss: Optional[torch.Tensor],
               n_global: int) -> torch.Tensor:
    """dz = g_zq + g_loss * 2 (z - z_q) / (n_global D) from an assembled z_q (K-sharded codebook)."""
    _req(z, "z", torch.float32); _req(z_q, "z_q", torch.float32)
    if g_zq is not None:
        _req(g_zq, "g_zq", torch.float32)
    if g_loss is not None:
        _req(g_loss, "g_loss", torch.float32)
    N, D = z.shape
    dz = torch.empty_like(z)
    with _on_device(z.device):
        check(_lib.load().kvq_dz_from_zq(z.data_ptr(), z_q.data_ptr(), _ptr(g_zq), _ptr(g_loss), N, D, n_global,
                                         dz.data_ptr(), _stream()), "kvq_dz_from_zq")
    return dz


def onehot(idx: torch.Tensor, K: int) -> torch.Tensor:
    """Dense `min_encodings` (VectorQuantizer.py:67-68)."""
    _req(idx, "idx", torch.int64)
    N = idx.numel()
    out = torch.empty(N, K, dtype=torch.float32, device=idx.device)
    with _on_device(idx.device):
        check(_lib.load().kvq_onehot(idx.data_ptr(), N, K, out.data_ptr(), _stream()), "kvq_onehot")
    return out


def forward_backward_host(z: torch.Tensor, E: torch.Tensor, g_zq: torch.Tensor, g_loss: float, beta: float, *,
                          mode: str = "auto", rows_per_chunk: int = 0, out=None):
    """End-to-end with HOST tensors (pinned recommended): copies in, runs forward + backward on the current
    device, copies every output back.  Synchronous.  Returns dict of host tensors."""
    for name, t in (("z", z), ("E", E), ("g_zq", g_zq)):
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous fp32 host tensor")
    N, D = z.shape
    K = E.shape[0]
    if out is None:
        pin = torch.cuda.is_available()
        out = dict(z_q=torch.empty(N, D, pin_memory=pin), idx=torch.empty(N, dtype=torch.int64, pin_memory=pin),
                   dz=torch.empty(N, D, pin_memory=pin), dE=torch.empty(K, D, pin_memory=pin),
                   scal=torch.empty(2, pin_memory=pin))
    check(_lib.load().kvq_forward_backward_host(
        z.data_ptr(), E.data_ptr(), g_zq.data_ptr(), float(g_loss), N, D, K, float(beta), SEARCH_MODES[mode],
        out["z_q"].data_ptr(), out["idx"].data_ptr(), out["scal"].data_ptr(), out["scal"].data_ptr() + 4,
        out["dz"].data_ptr(), out["dE"].data_ptr(), rows_per_chunk), "kvq_forward_backward_host")
    out["loss"], out["perplexity"] = out["scal"][0], out["scal"][1]
    return out


def forward_backward_host_sharded(z: torch.Tensor, E: torch.Tensor, g_zq: torch.Tensor, g_loss: float, beta: float,
                                  n_global: int, *, group=None, mode: str = "auto", rows_per_chunk: int = 0, out=None):
    """Batch-sharded end-to-end step with HOST tensors: this rank's rows through the chunked copy/compute pipeline,
    then one all-reduce(SUM) of [dE | histogram | squared-residual sum] partials over `group` and the finalisation.
    Returns dict of host tensors (z_q, idx, dz local rows; dE, loss, perplexity global)."""
    import torch.distributed as dist
    for name, t in (("z", z), ("E", E), ("g_zq", g_zq)):
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous fp32 host tensor")
    N, D = z.shape
    K = E.shape[0]
    dev = torch.device("cuda", torch.cuda.current_device())
    if out is None:
        out = dict(z_q=torch.empty(N, D, pin_memory=True), idx=torch.empty(N, dtype=torch.int64, pin_memory=True),
                   dz=torch.empty(N, D, pin_memory=True), dE=torch.empty(K, D, pin_memory=True),
                   scal=torch.empty(2, pin_memory=True),
                   dE_dev=torch.empty(K, D, device=dev), hist_dev=torch.empty(K, dtype=torch.int32, device=dev),
                   sq_dev=torch.empty(1, dtype=torch.float64, device=dev))
    check(_lib.load().kvq_forward_backward_host_sharded(
        z.data_ptr(), E.data_ptr(), g_zq.data_ptr(), float(g_loss), N, D, K, float(beta), SEARCH_MODES[mode], n_global,
        out["z_q"].data_ptr(), out["idx"].data_ptr(), out["dz"].data_ptr(), out["sq_dev"].data_ptr(),
        out["hist_dev"].data_ptr(), out["dE_dev"].data_ptr(), rows_per_chunk), "kvq_forward_backward_host_sharded")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out["dE_dev"], group=group)
        dist.all_reduce(out["hist_dev"], group=group)
        dist.all_reduce(out["sq_dev"], group=group)
    loss, perp = finalize(out["sq_dev"], out["hist_dev"], n_global, D, beta)
    out["dE"].copy_(out["dE_dev"], non_blocking=True)
    out["scal"][0:1].copy_(loss.view(1), non_blocking=True)
    out["scal"][1:2].copy_(perp.view(1), non_blocking=True)
    torch.cuda.synchronize()
    out["loss"], out["perplexity"] = out["scal"][0], out["scal"][1]
    return out


def search_peers(z: torch.Tensor, E: torch.Tensor, peer_key_ptrs, my_rank: int, *, mode: str = "auto", k_offset: int = 0,
                 ws: Optional[torch.Tensor] = None) -> None:
    """Fused search + cross-GPU argmin: MIN-combines packed keys into every rank's (peer-mapped) key buffer."""
    import ctypes
    _req(z, "z", torch.float32); _req(E, "E", torch.float32)
    N, D = z.shape
    K = E.shape[0]
    if ws is None:
        ws = workspace(N, D, K, z.device)
    arr = (ctypes.c_void_p * len(peer_key_ptrs))(*[int(p) for p in peer_key_ptrs])
    with _on_device(z.device):
        check(_lib.load().kvq_search_peers(z.data_ptr(), E.data_ptr(), N, D, K, k_offset, SEARCH_MODES[mode], arr,
                                           len(peer_key_ptrs), my_rank, ws.data_ptr(), ws.numel(), _stream()),
              "kvq_search_peers")


def quantize_shards(z: torch.Tensor, shard_ptrs, k_per: int, idx: torch.Tensor, K_total: int):
    """Gather + straight-through + loss sum + full histogram with codebook rows read from peer-mapped shards."""
    import ctypes
    _req(z, "z", torch.float32); _req(idx, "idx", torch.int64)
    N, D = z.shape
    z_q = torch.empty_like(z)
    sq_sum = torch.zeros(1, dtype=torch.float64, device=z.device)
    hist = torch.zeros(K_total, dtype=torch.int32, device=z.device)
    arr = (ctypes.c_void_p * len(shard_ptrs))(*[int(p) for p in shard_ptrs])
    with _on_device(z.device):
        check(_lib.load().kvq_quantize_shards(z.data_ptr(), arr, len(shard_ptrs), k_per, idx.data_ptr(), N, D, K_total,
                                              z_q.data_ptr(), sq_sum.data_ptr(), hist.data_ptr(), _stream()),
              "kvq_quantize_shards")
    return z_q, sq_sum, hist


def vq_backward_peers(z: torch.Tensor, E: torch.Tensor, idx: torch.Tensor, hist: torch.Tensor, beta: float, *,
                      g_zq: Optional[torch.Tensor], g_loss: torch.Tensor, need_dz: bool, n_global: int,
                      dE_peer_ptrs, dE_multicast_ptr: int, my_rank: int, ws: Optional[torch.Tensor] = None):
    """Batch-sharded backward whose codebook-gradient all-reduce is fused into the scatter-add kernel (multimem.red
    through the NVSwitch, or per-peer red).  The symmetric dE buffer must be zero on every rank.  Returns dz or None."""
    import ctypes
    _req(z, "z", torch.float32); _req(E, "E", torch.float32); _req(idx, "idx", torch.int64); _req(hist, "hist", torch.int32)
    _req(g_loss, "g_loss", torch.float32)
    if g_zq is not None:
        _req(g_zq, "g_zq", torch.float32)
    N, D = z.shape
    K = E.shape[0]
    dz = torch.empty_like(z) if need_dz else None
    if ws is None:
        ws = workspace(N, D, K, z.device)
    arr = (ctypes.c_void_p * len(dE_peer_ptrs))(*[int(p) for p in dE_peer_ptrs])
    with _on_device(z.device):
        check(_lib.load().kvq_backward_peers(z.data_ptr(), E.data_ptr(), idx.data_ptr(), hist.data_ptr(), _ptr(g_zq),
                                             g_loss.data_ptr(), N, D, K, float(beta), n_global, _ptr(dz),
                                             int(dE_multicast_ptr) if dE_multicast_ptr else None, arr, len(dE_peer_ptrs),
                                             my_rank, ws.data_ptr(), ws.numel(), _stream()), "kvq_backward_peers")
    return dz
