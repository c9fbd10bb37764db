"""ctypes binding of libkvq.so (C ABI in include/kvq.h).

The library is built in-tree by build.py.  Loading fails loudly: there is no CPU or PyTorch fallback
for any entry point -- a missing or unloadable library is an error, not a slow path.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libkvq.so")

KVQ_OK = 0
SEARCH_AUTO, SEARCH_TF32, SEARCH_FP32, SEARCH_TF32_REFINE = 0, 1, 2, 3
PROF_TAGS = ["norms", "search", "quantize", "finalize", "bwd_bucket", "bwd_segmented"]
SEARCH_MODES = {"auto": SEARCH_AUTO, "tf32": SEARCH_TF32, "fp32": SEARCH_FP32, "tf32_refine": SEARCH_TF32_REFINE}

_P = c_void_p

# name -> (restype, argtypes); mirrors include/kvq.h one to one (tests check every symbol is exported)
SIGNATURES = {
    "kvq_version": (c_int, []),
    "kvq_last_error": (c_char_p, []),
    "kvq_launch_count": (ctypes.c_longlong, []),
    "kvq_profile_enable": (c_int, [c_int]),
    "kvq_profile_collect": (c_int, [_P, _P, c_int]),
    "kvq_device_info": (c_int, [_P, _P, _P]),
    "kvq_workspace_bytes": (c_size_t, [c_int64, c_int, c_int64]),
    "kvq_code_norms": (c_int, [_P, c_int64, c_int, _P, c_int64, _P]),
    "kvq_search": (c_int, [_P, _P, c_int64, c_int, c_int64, c_int64, c_int, _P, _P, c_int, _P, c_size_t, _P]),
    "kvq_search_peers": (c_int, [_P, _P, c_int64, c_int, c_int64, c_int64, c_int, _P, c_int, c_int, _P, c_size_t, _P]),
    "kvq_quantize_shards": (c_int, [_P, _P, c_int, c_int64, _P, c_int64, c_int, c_int64, _P, _P, _P, _P]),
    "kvq_pack_key": (c_int64, [c_float, ctypes.c_uint32]),
    "kvq_keys_to_idx": (c_int, [_P, c_int64, _P, _P]),
    "kvq_quantize": (c_int, [_P, _P, _P, c_int64, c_int, c_int64, c_int64, c_int, _P, _P, _P, _P]),
    "kvq_finalize": (c_int, [_P, _P, c_int64, c_int, c_int64, c_float, _P, _P, _P]),
    "kvq_forward": (c_int, [_P, _P, c_int64, c_int, c_int64, c_float, c_int, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "kvq_forward_partials": (c_int, [_P, _P, c_int64, c_int, c_int64, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "kvq_pack_partials": (c_int, [_P, _P, c_int64, _P, _P]),
    "kvq_finalize_packed": (c_int, [_P, c_int64, c_int, c_int64, c_float, _P, _P, _P, _P]),
    "kvq_backward": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int64, c_int64, c_float, c_int64, _P, _P,
                             _P, c_size_t, _P]),
    "kvq_backward_peers": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int64, c_float, c_int64, _P, _P, _P, c_int,
                                   c_int, _P, c_size_t, _P]),
    "kvq_dz_from_zq": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int64, _P, _P]),
    "kvq_histogram": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P]),
    "kvq_cooccurrence": (c_int, [_P, _P, c_int64, c_int64, c_int64, _P, _P]),
    "kvq_kmeans_update": (c_int, [_P, _P, _P, c_int64, c_int, c_int64, _P, _P, _P, c_size_t, _P]),
    "kvq_onehot": (c_int, [_P, c_int64, c_int64, _P, _P]),
    "kvq_seq_acc": (c_int, [_P, _P, c_int64, c_int64, _P, _P, _P]),
    "kvq_recon_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "kvq_recon_loss_forward": (c_int, [_P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "kvq_recon_loss_backward": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int64, _P, _P]),
    "kvq_gemm_nt_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int64]),
    "kvq_gemm_nt": (c_int, [_P, _P, c_int64, c_int64, c_int64, _P, c_int64, _P, c_float, _P, c_size_t, _P]),
    "kvq_transpose_pad": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_int64, _P]),
    "kvq_gumbel_rows_forward": (c_int, [_P, _P, c_uint64, c_int64, c_int64, c_int64, c_float, c_float, c_int, _P, _P, _P, _P,
                                        _P]),
    "kvq_gumbel_hard_gather": (c_int, [_P, _P, _P, c_int64, c_int, c_int64, _P, _P]),
    "kvq_gumbel_rows_backward": (c_int, [_P, _P, c_uint64, _P, _P, c_int64, c_int64, c_int64, c_float, c_float, _P, _P]),
    "kvq_colsum_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "kvq_colsum": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, c_size_t, _P]),
    "kvq_replace_pct_rand_values": (c_int, [_P, c_int64, c_double, c_int64, c_int64, c_uint64, _P, _P]),
    "kvq_change_percentage_of_elements": (c_int, [_P, c_int64, c_int64, c_int, c_double, c_int64, c_int64, c_uint64,
                                                  _P, _P]),
    "kvq_forward_backward_host": (c_int, [_P, _P, _P, c_float, c_int64, c_int, c_int64, c_float, c_int, _P, _P, _P,
                                          _P, _P, _P, c_int64]),
    "kvq_forward_backward_host_sharded": (c_int, [_P, _P, _P, c_float, c_int64, c_int, c_int64, c_float, c_int, c_int64,
                                                  _P, _P, _P, _P, _P, _P, c_int64]),
    "kvq_host_release": (c_int, []),
    "kvq_search_plan": (c_int, [c_int64, c_int, c_int64, c_int, c_int, _P]),
}

_lib = None


class KvqError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libkvq.so (once).  Raises KvqError with build instructions if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KvqError(
            f"{LIB_PATH} not found. Build it with `python kindergarten-vq-vae_b200/build.py` "
            "(needs nvcc; libkvq has no CPU fallback).")
    try:
        lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    except OSError as exc:  # pragma: no cover - depends on the host
        raise KvqError(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != KVQ_OK:
        msg = load().kvq_last_error()
        raise KvqError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
