"""Code-usage analysis on the device.

The reference walks every sentence, word and token in Python to relate tokens to the code indices the VQ layer
assigned them (analyses/unsupervised_vq_disentanglement/unsupervised_vq_disentanglement.py:156-235).  Here one
kernel builds the (token id x code) co-occurrence table; the three result files of the reference
(`..._vq_vector_populated.txt`, `..._words_of_interest_histograms.json`, `..._vq_words_distrib.json`) are views of
that table.  Words are identified by their single token id (the reference warns when a word of interest is not a
single token, :196-200).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Mapping, Set

import torch

from . import _lib
from ._lib import check


def code_usage_by_token(input_ids: torch.Tensor, min_encoding_indices: torch.Tensor, vocab_size: int, n_e: int,
                        table: torch.Tensor = None) -> torch.Tensor:
    """(vocab_size, n_e) int32 counts; pass `table` from a previous call to accumulate over batches."""
    if not input_ids.is_cuda or not min_encoding_indices.is_cuda:
        raise RuntimeError("code_usage_by_token runs on CUDA tensors only: there is no CPU fallback")
    tok = input_ids.to(torch.int64).reshape(-1).contiguous()
    cod = min_encoding_indices.to(torch.int64).reshape(-1).contiguous()
    if tok.numel() != cod.numel():
        raise RuntimeError("input_ids and min_encoding_indices must cover the same positions")
    fresh = torch.empty(vocab_size, n_e, dtype=torch.int32, device=tok.device)
    with torch.cuda.device(tok.device):
        check(_lib.load().kvq_cooccurrence(tok.data_ptr(), cod.data_ptr(), tok.numel(), vocab_size, n_e, fresh.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream), "kvq_cooccurrence")
    if table is not None:
        table += fresh
        return table
    return fresh


def populated_codes(table: torch.Tensor) -> Set[int]:
    """`seen_v_is` of the reference (:189, :209-210)."""
    return set(torch.nonzero(table.sum(0)).flatten().tolist())


def words_of_interest_histograms(table: torch.Tensor, word_to_token_id: Mapping[str, int]) -> Dict[str, Dict[int, int]]:
    """word -> {code: count}, every code present (reference :212-226, there with range(9))."""
    host = table.cpu()
    return {w: {k: int(host[t, k]) for k in range(host.shape[1])} for w, t in word_to_token_id.items()}


def vq_words_distrib(table: torch.Tensor, token_id_to_word: Mapping[int, str]) -> Dict[int, List[str]]:
    """code -> distinct words that were mapped to it (reference :231-235)."""
    host = table.cpu()
    out: Dict[int, List[str]] = {}
    for k in range(host.shape[1]):
        toks: Iterable[int] = torch.nonzero(host[:, k]).flatten().tolist()
        words = sorted({token_id_to_word[t] for t in toks if t in token_id_to_word})
        if words:
            out[k] = words
    return out


def write_results(results_dir: str, table: torch.Tensor, word_to_token_id: Mapping[str, int],
                  token_id_to_word: Mapping[int, str]) -> Dict[str, str]:
    """Write the three result files of the reference analysis, same names and schema
    (analyses/unsupervised_vq_disentanglement/unsupervised_vq_disentanglement.py:206-235):

      dSentences_vq_vector_populated.txt            "the following VQ latent vectors were populated: {set}"      (:209-210)
      dSentences_words_of_interest_histograms.json  {word: {code: count}}, every code present                      (:212-226)
      dSentences_vq_words_distrib.json              {code: [distinct words mapped to it]}                          (:228-235)

    json.dump turns the integer code keys into strings, exactly as it does for the reference's dicts.
    Returns {name: path}."""
    import json
    import os
    os.makedirs(results_dir, exist_ok=True)
    paths = {
        "populated": os.path.join(results_dir, "dSentences_vq_vector_populated.txt"),
        "histograms": os.path.join(results_dir, "dSentences_words_of_interest_histograms.json"),
        "distrib": os.path.join(results_dir, "dSentences_vq_words_distrib.json"),
    }
    with open(paths["populated"], "w") as f:
        f.write(f"the following VQ latent vectors were populated: {str(populated_codes(table))}")
    with open(paths["histograms"], "w") as fp:
        json.dump(words_of_interest_histograms(table, word_to_token_id), fp)
    with open(paths["distrib"], "w") as fp:
        json.dump(vq_words_distrib(table, token_id_to_word), fp)
    return paths
