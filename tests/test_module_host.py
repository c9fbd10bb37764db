"""Host-side mirror of the reference interface (models/shelgon3/VectorQuantizer.py:19-29, Shelgon.py:54-58):
things that must hold before any kernel runs.  CPU only."""
import sys

import pytest
import torch

from kindergarten_vq_vae_b200 import VectorQuantizer, replace_pct_rand_values, change_percentage_of_elements
from kindergarten_vq_vae_b200.sharded import shard_bounds


def test_class_name_and_attributes():
    vq = VectorQuantizer(n_e=512, e_dim=768, beta=0.25)
    assert type(vq).__name__ == "VectorQuantizer"          # Shelgon.py:57 dispatches on this string
    assert (vq.n_e, vq.e_dim, vq.beta) == (512, 768, 0.25)
    assert isinstance(vq.embedding, torch.nn.Embedding)
    assert list(vq.state_dict().keys()) == ["embedding.weight"]   # checkpoint key, Trainer.py:243
    w = vq.embedding.weight
    assert w.shape == (512, 768) and w.requires_grad
    assert float(w.abs().max()) <= 1.0 / 512 and float(w.min()) < 0 < float(w.max())   # U(-1/K, 1/K), :29


def test_codebook_init_values_are_copied():
    init = torch.randn(10, 32)
    vq = VectorQuantizer(10, 32, 0.69, vq_codebook_init_values=init)
    assert torch.equal(vq.embedding.weight.data, init) and vq.embedding.weight.data_ptr() != init.data_ptr()


def test_state_dict_roundtrip_with_reference_layout():
    sys.path.insert(0, "/root/reference/models/shelgon3")
    try:
        from VectorQuantizer import VectorQuantizer as RefVQ   # only available in the build container
    except Exception:
        pytest.skip("reference not mounted")
    ref = RefVQ(n_e=12, e_dim=16, beta=0.1)
    ours = VectorQuantizer(n_e=12, e_dim=16, beta=0.1)
    ours.load_state_dict(ref.state_dict())
    assert torch.equal(ours.embedding.weight, ref.embedding.weight)


def test_bad_arguments():
    with pytest.raises(ValueError):
        VectorQuantizer(4, 8, 0.25, search="bf16")
    vq = VectorQuantizer(4, 8, 0.25)
    with pytest.raises(RuntimeError):
        vq.forward(torch.randn(3, 8), None)                  # not (B,S,D)


def test_perturbation_early_return_is_identity_object():
    t = torch.arange(12).reshape(3, 4)
    assert replace_pct_rand_values(t, 0.0, 0, 10) is t       # tensor_utils.py:18
    assert change_percentage_of_elements(t, 1, 0.0, 0, 10) is t   # tensor_utils.py:54


def test_shard_bounds_cover_everything_once():
    for total in (1, 7, 512, 1 << 20, 1000003):
        for parts in (1, 2, 3, 8):
            spans = [shard_bounds(total, parts, p) for p in range(parts)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            per = (total + parts - 1) // parts
            assert all(hi - lo <= per for lo, hi in spans)
