"""GPU tests of the helpers named by north_star (common/metrics.py, common/tensor_utils.py), the dense one-hot and
the host-buffer end-to-end entry point."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _kvq():
    import kindergarten_vq_vae_b200 as k
    return k


def test_seq_acc_golden_and_random():
    k = _kvq()
    d = np.load(os.path.join(GOLDEN, "seq_acc.npz"))
    acc, per = k.seq_acc(torch.from_numpy(d["a"]).to(DEV), torch.from_numpy(d["b"]).to(DEV))
    assert float(acc) == float(d["acc"]) and np.array_equal(per.cpu().numpy(), d["per"])
    a = torch.randint(0, 3, (513, 12)); b = torch.randint(0, 3, (513, 12))
    acc, per = k.seq_acc(a.to(DEV), b.to(DEV))
    ra, rp = O.seq_acc(a, b)
    assert float(acc) == float(ra) and torch.equal(per.cpu(), rp)
    with pytest.raises(AssertionError):
        k.seq_acc(a.float().to(DEV), b.to(DEV))
    with pytest.raises(AssertionError):
        k.seq_acc(a[:5].to(DEV), b.to(DEV))


def test_replace_pct_rand_values_counts_and_range():
    k = _kvq()
    t = torch.full((512, 12), -7, dtype=torch.int64, device=DEV)     # sentinel outside [low, high)
    for pct in (0.1, 0.15, 0.69, 1.0):
        out = k.replace_pct_rand_values(t, pct, 0, 30522, seed=5)
        changed = out != -7
        assert int(changed.sum()) == int(t.numel() * pct)             # exactly int(numel*pct), tensor_utils.py:29
        assert int(out[changed].min()) >= 0 and int(out[changed].max()) < 30522
        assert torch.equal(out, k.replace_pct_rand_values(t, pct, 0, 30522, seed=5))       # reproducible per seed
        assert not torch.equal(out, k.replace_pct_rand_values(t, pct, 0, 30522, seed=6))
    # positions are spread over the tensor, not clustered at the front
    out = k.replace_pct_rand_values(t, 0.5, 0, 10, seed=1)
    first, second = (out[:256] != -7).float().mean(), (out[256:] != -7).float().mean()
    assert abs(float(first) - 0.5) < 0.05 and abs(float(second) - 0.5) < 0.05


def test_change_percentage_of_elements_slices():
    k = _kvq()
    t = torch.full((20, 8), -1, dtype=torch.int64, device=DEV)
    out = k.change_percentage_of_elements(t, 1, 0.6, 5, 9, seed=3)
    cols = (out != -1).all(0)
    assert int(cols.sum()) == int(8 * 0.6) and bool(((out != -1).any(0) == cols).all())   # whole columns
    assert bool((out[:, cols] == out[0:1, cols]).all())                                  # one value per column
    assert int(out[:, cols].min()) >= 5 and int(out[:, cols].max()) < 9
    out0 = k.change_percentage_of_elements(t, 0, 0.25, 0, 100, seed=3)
    rows = (out0 != -1).all(1)
    assert int(rows.sum()) == 5 and bool((out0[rows] == out0[rows][:, :1]).all())
    with pytest.raises(ValueError):
        k.change_percentage_of_elements(t, 2, 0.5, 0, 3)


def test_host_end_to_end_matches_device_path():
    k = _kvq()
    F = k.functional
    gen = torch.Generator().manual_seed(11)
    N, D, K = 3000, 256, 700
    z = torch.randn(N, D, generator=gen).pin_memory()
    E = torch.randn(K, D, generator=gen).pin_memory()
    g = torch.randn(N, D, generator=gen).pin_memory()
    out = F.forward_backward_host(z, E, g, 1.25, 0.25, mode="tf32", rows_per_chunk=1024)
    zd, Ed, gd = z.to(DEV), E.to(DEV), g.to(DEV)
    loss, z_q, perp, idx, hist = F.vq_forward(zd, Ed, 0.25, mode="tf32")
    dz, dE = F.vq_backward(zd, Ed, idx, hist, 0.25, g_zq=gd, g_loss=torch.tensor(1.25, device=DEV))
    par = O.index_parity(out["idx"], idx.cpu(), z, E)       # chunked search may split the code range differently
    assert par.unexcused == 0
    if par.raw_mismatch == 0:
        assert torch.equal(out["z_q"], z_q.cpu())
        assert abs(float(out["loss"]) - float(loss)) <= 1e-6 * float(loss)
        assert abs(float(out["perplexity"]) - float(perp)) <= 1e-6 * float(perp)
        assert torch.allclose(out["dz"], dz.cpu(), rtol=1e-6, atol=1e-8)
        assert (out["dE"] - dE.cpu()).abs().max() <= 1e-5 * dE.abs().max()
    ref = O.forward_fp32(z, E, 0.25)
    assert O.index_parity(out["idx"], ref.idx, z, E).unexcused == 0
    # sharded entry point on one rank (no process group): same pipeline, partials finalised by the caller
    out2 = F.forward_backward_host_sharded(z, E, g, 1.25, 0.25, N, mode="tf32", rows_per_chunk=1024)
    assert torch.equal(out2["idx"], out["idx"]) and torch.equal(out2["z_q"], out["z_q"]) and torch.equal(out2["dz"], out["dz"])
    assert abs(float(out2["loss"]) - float(out["loss"])) <= 1e-6 * float(out["loss"])
    assert abs(float(out2["perplexity"]) - float(out["perplexity"])) <= 1e-6 * float(out["perplexity"])
    assert (out2["dE"] - out["dE"]).abs().max() <= 1e-6 * out["dE"].abs().max()
    F._lib.load().kvq_host_release()


@pytest.mark.parametrize("N,D,K,mode", [(700, 768, 512, "auto"), (1500, 96, 40, "fp32"), (260, 36, 9, "auto")])
def test_host_end_to_end_other_shapes(N, D, K, mode):
    """Host-buffer entry point on BERT-width latents (streaming-operand search), on a D the tensor-core kernel does
    not take, with a ragged last chunk -- against the oracle."""
    k = _kvq()
    F = k.functional
    gen = torch.Generator().manual_seed(21)
    z = torch.randn(N, D, generator=gen); E = torch.randn(K, D, generator=gen); g = torch.randn(N, D, generator=gen)
    out = F.forward_backward_host(z, E, g, 0.5, 0.25, mode=mode, rows_per_chunk=256)
    ref = O.forward_fp32(z, E, 0.25)
    par = O.index_parity(out["idx"], ref.idx, z, E, exact_fp32=(mode == "fp32"))
    assert par.unexcused == 0
    if par.raw_mismatch == 0:
        dz, dE = O.backward_closed_form(z, E, ref.idx, 0.25, g_zq=g, g_loss=0.5)
        assert torch.equal(out["z_q"], ref.z_q.view(N, D))
        assert abs(float(out["loss"]) - float(ref.loss)) <= 2e-5 * float(ref.loss)
        assert abs(float(out["perplexity"]) - float(ref.perplexity)) <= 2e-5 * float(ref.perplexity)
        assert torch.allclose(out["dz"], dz.float(), rtol=1e-5, atol=1e-7)
        assert (out["dE"] - dE.float()).abs().max() <= 1e-5 * dE.abs().max()
    F._lib.load().kvq_host_release()


def test_host_end_to_end_default_chunking_equals_device_path():
    """rows_per_chunk = 0: the library's own chunking (a short lead-in chunk, then one full wave of the search kernel per
    chunk, ragged tail).  Every output must equal the device-resident path's on the same rows."""
    k = _kvq()
    F = k.functional
    gen = torch.Generator().manual_seed(5)
    N, D, K = 100_000, 64, 300            # > 4 waves of 18944 rows: lead-in + 5 full chunks + a ragged last one
    z = torch.randn(N, D, generator=gen).pin_memory()
    E = torch.randn(K, D, generator=gen).pin_memory()
    g = torch.randn(N, D, generator=gen).pin_memory()
    out = F.forward_backward_host(z, E, g, 0.8, 0.25, mode="auto", rows_per_chunk=0)
    zd, Ed, gd = z.to(DEV), E.to(DEV), g.to(DEV)
    loss, z_q, perp, idx, hist = F.vq_forward(zd, Ed, 0.25, mode="auto")
    dz, dE = F.vq_backward(zd, Ed, idx, hist, 0.25, g_zq=gd, g_loss=torch.tensor(0.8, device=DEV))
    assert torch.equal(out["idx"], idx.cpu())              # default mode: exact pass, chunking cannot change a row's result
    assert torch.equal(out["z_q"], z_q.cpu()) and torch.allclose(out["dz"], dz.cpu(), rtol=1e-6, atol=1e-8)
    assert abs(float(out["loss"]) - float(loss)) <= 1e-6 * float(loss)
    assert abs(float(out["perplexity"]) - float(perp)) <= 1e-6 * float(perp)
    assert (out["dE"] - dE.cpu()).abs().max() <= 1e-5 * dE.abs().max()
    rows = torch.arange(0, N, 97)
    ref = O.forward_fp32(z[rows], E, 0.25)
    assert O.index_parity(out["idx"][rows], ref.idx, z[rows], E).unexcused == 0
    F._lib.load().kvq_host_release()


def test_kmeans2_matches_scipy_given_the_same_initial_centroids():
    """Device Lloyd iterations vs scipy.cluster.vq.kmeans2 (the call of vq_codebook_init_weights.py:85)."""
    from scipy.cluster.vq import kmeans2 as sp_kmeans2
    k = _kvq()
    gen = torch.Generator().manual_seed(3)
    centers = torch.randn(12, 64, generator=gen) * 4
    data = (centers[torch.randint(0, 12, (6000,), generator=gen)] + torch.randn(6000, 64, generator=gen)).contiguous()
    init = data[torch.randperm(6000, generator=gen)[:16]].clone()           # 16 > 12 true clusters
    init[15] = 100.0                                                         # a centroid that never gets members
    ref_c, ref_l = sp_kmeans2(data.numpy(), init.numpy().copy(), iter=10, minit="matrix", missing="warn")
    c, l = k.kmeans2(data.to(DEV), init.to(DEV), iter=10, minit="matrix", search="fp32")
    assert torch.equal(l.cpu(), torch.from_numpy(ref_l).long())
    assert torch.allclose(c.cpu(), torch.from_numpy(ref_c), rtol=1e-5, atol=1e-5)
    assert torch.equal(c[15].cpu(), init[15])                                # empty cluster keeps its position
    # the tensor-core search with its exact float64 re-evaluation of the two best centroids assigns the same labels
    ca, la = k.kmeans2(data.to(DEV), init.to(DEV), iter=10, minit="matrix", search="auto")
    assert torch.equal(la.cpu(), torch.from_numpy(ref_l).long())
    assert torch.allclose(ca.cpu(), torch.from_numpy(ref_c), rtol=1e-5, atol=1e-5)
    # minit='points': centroids are data points after 0 iterations, K distinct rows, reproducible per seed
    c0, _ = k.kmeans2(data.to(DEV), 9, iter=0, minit="points", seed=5)
    assert c0.shape == (9, 64) and all(bool((data == r).all(1).any()) for r in c0.cpu())
    c1, l1 = k.kmeans2(data.to(DEV), 9, iter=10, minit="points", seed=5)
    c2, l2 = k.kmeans2(data.to(DEV), 9, iter=10, minit="points", seed=5)
    assert torch.equal(l1, l2) and torch.allclose(c1, c2, rtol=1e-6, atol=1e-6)
    # within-cluster distortion does not increase over iterations (Lloyd property), large shape, tf32 search
    big = torch.randn(1 << 16, 256, device=DEV)
    cb3, lb3 = k.kmeans2(big, 512, iter=3, seed=1, search="tf32")
    cb8, lb8 = k.kmeans2(big, 512, iter=8, seed=1, search="tf32")
    d3 = float(((big - cb3[lb3]) ** 2).sum()); d8 = float(((big - cb8[lb8]) ** 2).sum())
    assert d8 <= d3 * (1 + 1e-4)
    out = k.codebook_init_values(big.view(1024, 64, 256), 9, iter=2, seed=0)
    assert out["codebook_init_values"].shape == (9, 256) and not out["codebook_init_values"].is_cuda


def test_code_usage_analysis_matches_python_loops():
    """Device co-occurrence table vs the reference's nested-loop bookkeeping (unsupervised_vq_disentanglement.py:165-235)."""
    k = _kvq()
    A = k.analysis
    gen = torch.Generator().manual_seed(9)
    B, S, V, K = 64, 12, 50, 9
    ids = torch.randint(0, V, (B, S), generator=gen)
    codes = torch.randint(0, K - 1, (B, S, 1), generator=gen)            # code K-1 is never used
    table = A.code_usage_by_token(ids.to(DEV), codes.to(DEV), V, K)
    table = A.code_usage_by_token(ids.to(DEV), codes.to(DEV), V, K, table=table)      # accumulate a second "batch"
    # reference-style bookkeeping
    words = {t: f"w{t}" for t in range(V)}
    interest = {"w3": 3, "w7": 7, "w49": 49}
    seen, per_word, per_code = set(), {w: [] for w in interest}, {}
    for _ in range(2):
        for row_ids, row_codes in zip(ids.tolist(), codes.flatten(1).tolist()):
            for t, c in zip(row_ids, row_codes):
                seen.add(c); per_code.setdefault(c, set()).add(words[t])
                if words[t] in interest:
                    per_word[words[t]].append(c)
    assert A.populated_codes(table) == seen and (K - 1) not in seen
    hist = A.words_of_interest_histograms(table, interest)
    for w in interest:
        assert hist[w] == {c: per_word[w].count(c) for c in range(K)}
    distrib = A.vq_words_distrib(table, words)
    assert {c: sorted(v) for c, v in per_code.items()} == distrib
    assert int(table.sum()) == 2 * B * S


def test_code_usage_result_files_have_the_reference_schema(tmp_path):
    """analysis.write_results: the three files of unsupervised_vq_disentanglement.py:206-235, same names and JSON layout."""
    import json
    k = _kvq()
    A = k.analysis
    gen = torch.Generator().manual_seed(10)
    B, S, V, K = 32, 12, 40, 9
    ids = torch.randint(0, V, (B, S), generator=gen)
    codes = torch.randint(0, K, (B, S, 1), generator=gen)
    table = A.code_usage_by_token(ids.to(DEV), codes.to(DEV), V, K)
    words = {t: f"w{t}" for t in range(V)}
    interest = {"w1": 1, "w2": 2}
    paths = A.write_results(str(tmp_path / "run"), table, interest, words)
    assert sorted(os.path.basename(p) for p in paths.values()) == [
        "dSentences_vq_vector_populated.txt", "dSentences_vq_words_distrib.json", "dSentences_words_of_interest_histograms.json"]
    txt = open(paths["populated"]).read()
    assert txt.startswith("the following VQ latent vectors were populated: {")
    # the reference builds the same dicts with Python loops and json.dump()s them: integer keys become strings
    per_word = {w: [] for w in interest}
    per_code = {}
    for row_ids, row_codes in zip(ids.tolist(), codes.flatten(1).tolist()):
        for t, c in zip(row_ids, row_codes):
            per_code.setdefault(c, set()).add(words[t])
            if words[t] in interest:
                per_word[words[t]].append(c)
    ref_hist = json.loads(json.dumps({w: {c: per_word[w].count(c) for c in range(9)} for w in interest}))
    assert json.load(open(paths["histograms"])) == ref_hist
    got = json.load(open(paths["distrib"]))
    assert {c: sorted(v) for c, v in got.items()} == {str(c): sorted(v) for c, v in per_code.items()}
