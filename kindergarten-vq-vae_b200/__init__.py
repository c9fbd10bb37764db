"""kvq -- B200-native (sm_100a) vector-quantisation bottleneck, drop-in for
dansolombrino/Kindergarten-VQ-VAE's models/shelgon3/VectorQuantizer.py.

    from kindergarten_vq_vae_b200 import VectorQuantizer

Importing the package loads libkvq.so and fails loudly if it is missing (no CPU / PyTorch fallback).
"""
from . import _lib

_lib.load()  # fail at import time, not on first use

from . import functional  # noqa: E402
from .metrics import seq_acc  # noqa: E402
from .recon import recon_loss  # noqa: E402
from .tensor_utils import change_percentage_of_elements, replace_pct_rand_values  # noqa: E402
from .vector_quantizer import VectorQuantizer  # noqa: E402
from .gumbel import GumbelQuantizer  # noqa: E402
from . import analysis  # noqa: E402
from .kmeans import codebook_init_values, kmeans2  # noqa: E402
from .sharded import BatchShardedVectorQuantizer, CodebookShardedVectorQuantizer  # noqa: E402

__all__ = ["VectorQuantizer", "GumbelQuantizer", "BatchShardedVectorQuantizer", "CodebookShardedVectorQuantizer", "functional", "analysis",
           "kmeans2", "codebook_init_values", "seq_acc", "recon_loss", "replace_pct_rand_values", "change_percentage_of_elements"]
