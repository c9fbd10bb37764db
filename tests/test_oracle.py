"""The oracle (oracle/vq_oracle.py) against the fixtures produced by the UNMODIFIED reference class
(tests/golden/make_golden.py).  CPU only."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, GOLDEN_CASES, load_golden
from oracle import vq_oracle as O


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_forward_matches_reference_bit_exact(name):
    torch.set_num_threads(1)
    g = load_golden(name)
    r = O.forward_fp32(g["z"], g["E"], float(g["beta"]))
    assert torch.equal(r.idx, g["idx"])                         # indices: bit exact
    assert torch.equal(r.z_q, g["z_q"])                         # fl(z + fl(q - z)): bit exact
    assert torch.allclose(r.loss, g["loss"], rtol=1e-6, atol=0)
    assert torch.allclose(r.perplexity, g["perplexity"], rtol=1e-6, atol=0)
    # the dense one-hot of the reference is one_hot(idx)
    oh = O.onehot(r.idx, g["E"].shape[0])
    assert torch.equal(oh.sum(1), g["onehot_rowsum"]) and torch.equal(oh.argmax(1), g["onehot_argmax"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_backward_closed_form_matches_reference_autograd(name):
    g = load_golden(name)
    dz, dE = O.backward_closed_form(g["z"], g["E"], g["idx"], float(g["beta"]), g_zq=g["gz"], g_loss=float(g["w"]))
    assert torch.allclose(dz.float(), g["dz"], rtol=1e-5, atol=1e-7)
    scale = g["dE"].abs().max().item() + 1e-30
    assert (dE.float() - g["dE"]).abs().max().item() <= 2e-6 * scale
    # unused codes get exact zeros (dense gradient consumed by Adam + weight decay, main.py:91)
    unused = torch.bincount(g["idx"].reshape(-1), minlength=g["E"].shape[0]) == 0
    assert torch.all(dE[unused] == 0)


def test_ties_resolve_to_lowest_index():
    g = load_golden("ties")
    K = g["E"].shape[0]
    assert int(g["idx"].max()) < K - K // 2 + (K % 2)   # duplicated upper half never wins
    i64, _, gap = O.truth_fp64(g["z"], g["E"])
    assert torch.all(gap == 0)                            # every row has an exact tie in fp64
    assert torch.equal(i64, g["idx"].reshape(-1))


def test_truth_fp64_and_parity_tolerance():
    g = load_golden("wide")
    i64, dmin, gap = O.truth_fp64(g["z"], g["E"])
    par = O.index_parity(i64, g["idx"], g["z"], g["E"], exact_fp32=True)
    assert par.unexcused == 0
    # a deliberately wrong index (second best with a large gap) must NOT be excused
    wrong = g["idx"].reshape(-1).clone()
    row = int(torch.argmax(gap))
    d = O.distances_fp32(g["z"].view(-1, g["E"].shape[1])[row:row + 1], g["E"])[0]
    wrong[row] = int(torch.topk(d, 2, largest=False).indices[1])
    par = O.index_parity(wrong, g["idx"], g["z"], g["E"])
    assert par.raw_mismatch == 1 and par.unexcused == 1


@pytest.mark.parametrize("init", ["default", "points"])
def test_config1_summary(init):
    """BASELINE config 1 (B=64,S=64,D=768,K=512): inputs regenerated from the seed, outputs from the fixture."""
    sys.path.insert(0, GOLDEN)
    from make_golden import make_inputs
    d = np.load(os.path.join(GOLDEN, f"vq_c1_{init}.npz"))
    z, E, gz = make_inputs("c1", 64, 64, 768, 512, init, 69)
    digest = hashlib.sha256(z.numpy().tobytes() + E.numpy().tobytes() + gz.numpy().tobytes()).hexdigest()
    if digest != str(d["input_sha256"]):
        pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    r = O.forward_fp32(z, E, 0.25)
    idx_ref = torch.from_numpy(d["idx"].astype(np.int64))
    par = O.index_parity(r.idx, idx_ref, z, E, exact_fp32=True)   # thread count may change fp32 summation order
    assert par.unexcused == 0 and par.raw_rate < 2e-3
    assert abs(float(r.loss) - float(d["loss"])) <= 1e-5 * float(d["loss"])
    assert abs(float(r.perplexity) - float(d["perplexity"])) <= 2e-3 * float(d["perplexity"])


def test_seq_acc_matches_reference():
    d = np.load(os.path.join(GOLDEN, "seq_acc.npz"))
    acc, per = O.seq_acc(torch.from_numpy(d["a"]), torch.from_numpy(d["b"]))
    assert float(acc) == float(d["acc"]) and np.array_equal(per.numpy(), d["per"])


def test_kshard_merge_equals_global_argmin():
    g = load_golden("wide")
    for shards in (2, 3, 8):
        idx = O.kshard_merge(g["z"], g["E"], shards)
        par = O.index_parity(idx, g["idx"], g["z"], g["E"], exact_fp32=True)
        assert par.unexcused == 0


@pytest.mark.parametrize("name", ["soft", "hard", "eval"])
def test_gumbel_oracle_matches_the_reference_class(name):
    """oracle.gumbel_forward (+ its autograd) against fixtures produced by the unmodified GumbelQuantizer
    (models/shelgon3/GumbelQuantizer.py:43-83) with torch's RNG seeded so that the Gumbel sample is the stored one."""
    d = np.load(os.path.join(GOLDEN, f"gumbel_{name}.npz"))
    t = {k: torch.from_numpy(np.array(d[k])) for k in d.files}
    z = t["z"].clone().requires_grad_(True)
    W = t["W"].clone().requires_grad_(True)
    b = t["b"].clone().requires_grad_(True)
    E = t["E"].clone().requires_grad_(True)
    z_q, diff, ind = O.gumbel_forward(z, W, b, E, t["noise"], float(t["tau"]), float(t["kld_scale"]), bool(t["hard"]))
    (diff * float(t["w"]) + (z_q * t["gz"]).sum()).backward()
    assert torch.equal(ind, t["ind"])
    assert torch.allclose(z_q, t["z_q"], rtol=1e-5, atol=1e-6)
    assert abs(float(diff) - float(t["diff"])) <= 1e-5 * abs(float(t["diff"]))
    for got, key in ((z.grad, "dz"), (W.grad, "dW"), (b.grad, "db"), (E.grad, "dE")):
        ref = t[key]
        assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max()) + 1e-9, key
