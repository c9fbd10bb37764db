"""Parity at the shapes BASELINE.json names, through the C ABI, with the DEFAULT search mode of the module:

  * config 1 (B=64, S=64, D=768, K=512): the CUDA path against the fixtures the unmodified reference class produced
    (tests/golden/vq_c1_*.npz; inputs are regenerated from the committed seed and checked by digest);
  * the headline shape (N=2^20 latents, D=256, K=65536): sampled rows against the oracle's reference-order fp32
    argmin (oracle.vq_oracle.forward_fp32 = models/shelgon3/VectorQuantizer.py:59-65 restated), plus size-independent
    properties, bitwise reproducibility of the codebook gradient, and torch.argmin's NaN-is-the-minimum rule.

Tolerances are those of tests/test_gpu_parity.py (stated there once).
"""
import hashlib
import math
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _kvq():
    import kindergarten_vq_vae_b200 as k
    return k


@pytest.mark.parametrize("init", ["default", "points"])
@pytest.mark.parametrize("search", ["auto", "fp32"])
def test_config1_cuda_against_reference_fixture(init, search):
    sys.path.insert(0, GOLDEN)
    from make_golden import make_inputs
    d = np.load(os.path.join(GOLDEN, f"vq_c1_{init}.npz"))
    z, E, gz = make_inputs("c1", 64, 64, 768, 512, init, 69)
    digest = hashlib.sha256(z.numpy().tobytes() + E.numpy().tobytes() + gz.numpy().tobytes()).hexdigest()
    if digest != str(d["input_sha256"]):
        pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    k = _kvq()
    vq = k.VectorQuantizer(512, 768, 0.25, vq_codebook_init_values=E, search=search).to(DEV)
    zin = z.to(DEV).requires_grad_(True)
    loss, z_q, perp, onehot, idx = vq.forward(zin, DEV)
    (loss * float(d["w"]) + (z_q * gz.to(DEV)).sum()).backward()
    torch.cuda.synchronize()
    idx_ref = torch.from_numpy(d["idx"].astype(np.int64))
    par = O.index_parity(idx.cpu(), idx_ref, z, E, exact_fp32=(search == "fp32"))
    print(f"c1[{init},{search}] index mismatches vs the reference {par.raw_mismatch}/{par.n} (unexcused {par.unexcused})")
    assert par.unexcused == 0
    # the reference's own fp32 argmin is noisy with the default init; allow its ambiguity floor, not more
    assert par.raw_rate <= (2e-2 if init == "default" else 2e-3)
    assert onehot is not None and tuple(onehot.shape) == (4096, 512)
    if par.raw_mismatch == 0:
        assert abs(float(loss) - float(d["loss"])) <= 2e-5 * float(d["loss"])
        assert abs(float(perp) - float(d["perplexity"])) <= 2e-5 * float(d["perplexity"])
        assert abs(float(z_q.double().sum()) - float(d["z_q_sum"])) <= 1e-6 * float(d["z_q_abs"])
        assert abs(float(zin.grad.double().abs().sum()) - float(d["dz_abs"])) <= 1e-6 * float(d["dz_abs"])
        rows = vq.embedding.weight.grad.double().abs().sum(1).cpu().numpy()
        assert np.abs(rows - d["dE_rows"]).max() <= 1e-5 * d["dE_rows"].max()
    else:
        # a handful of excused near-ties: the aggregates move by at most their share
        assert abs(float(loss) - float(d["loss"])) <= 1e-3 * float(d["loss"])
        assert abs(float(perp) - float(d["perplexity"])) <= 1e-2 * float(d["perplexity"])


def _headline_inputs(N, D, K, seed=69):
    """bench.py's synthetic inputs: z ~ N(0,1), data-scale codebook N(0,1) + 0.1 N(0,1)."""
    gen = torch.Generator(device=DEV).manual_seed(seed + 1)
    z = torch.randn(N, D, device=DEV, generator=gen)
    gz = torch.randn(N, D, device=DEV, generator=gen)
    gen_e = torch.Generator(device=DEV).manual_seed(seed)
    E = torch.randn(K, D, device=DEV, generator=gen_e) + 0.1 * torch.randn(K, D, device=DEV, generator=gen_e)
    return z, gz, E


def test_headline_shape_against_the_oracle():
    """N = 2^20, D = 256, K = 65536, default search: 4096 sampled rows against the reference-order fp32 argmin."""
    F = _kvq().functional
    N, D, K, beta = 1 << 20, 256, 65536, 0.25
    z, gz, E = _headline_inputs(N, D, K)
    loss, z_q, perp, idx, hist = F.vq_forward(z, E, beta, mode="auto")
    torch.cuda.synchronize()
    # ---- properties over all 2^20 rows
    assert int(hist.sum()) == N and int(idx.min()) >= 0 and int(idx.max()) < K
    assert torch.equal(hist.long(), torch.bincount(idx, minlength=K))
    q = E[idx]
    assert torch.equal(z_q, z + (q - z))
    m = float(((q - z).double() ** 2).mean())
    assert abs(float(loss) - m * (1 + beta)) <= 1e-5 * m * (1 + beta)
    p = hist.double() / N
    assert abs(float(perp) - math.exp(-float((p * torch.log(p + 1e-10)).sum()))) <= 1e-4 * float(perp)
    # ---- sampled rows against the oracle (CPU, reference evaluation order)
    rows = torch.arange(0, N, N // 4096, device=DEV)[:4096]
    zs, Ec = z[rows].cpu(), E.cpu()
    ref = O.forward_fp32(zs.view(64, 64, D), Ec, beta, row_chunk=512)
    par = O.index_parity(idx[rows].cpu(), ref.idx, zs, Ec)
    t64, _, _ = O.truth_fp64(zs, Ec, row_chunk=512)
    ours_vs_fp64 = int((idx[rows].cpu() != t64).sum())
    ref_vs_fp64 = int((ref.idx.reshape(-1) != t64).sum())
    print(f"headline sampled parity: {par.raw_mismatch}/{par.n} rows differ from the reference-order fp32 argmin, "
          f"unexcused {par.unexcused}, worst gap/tol {par.max_gap_over_tol:.3g}; vs fp64 argmin: ours {ours_vs_fp64}, "
          f"reference {ref_vs_fp64}")
    assert par.unexcused == 0
    assert par.raw_rate <= 2e-3
    assert ours_vs_fp64 <= max(2, ref_vs_fp64)          # the default mode is at least as close to the truth as the reference
    same = idx[rows].cpu() == ref.idx.reshape(-1)
    assert torch.equal(z_q[rows].cpu()[same], ref.z_q.view(-1, D)[same])
    # ---- backward: closed forms on the device + bitwise reproducibility of dE
    gl = torch.tensor(1.0, device=DEV)
    dz, dE = F.vq_backward(z, E, idx, hist, beta, g_zq=gz, g_loss=gl)
    dz2, dE2 = F.vq_backward(z, E, idx, hist, beta, g_zq=gz, g_loss=gl)
    torch.cuda.synchronize()
    assert torch.equal(dE, dE2) and torch.equal(dz, dz2)
    c1 = 2.0 / (N * D)
    assert torch.allclose(dz, gz + c1 * (z - q), rtol=1e-5, atol=1e-7)
    ref_dE = torch.zeros(K, D, device=DEV, dtype=torch.float64).index_add_(0, idx, (q - z).double()) * (c1 * beta)
    assert float((dE.double() - ref_dE).abs().max()) <= 1e-5 * float(ref_dE.abs().max())
    assert bool((dE[hist == 0] == 0).all())


def test_headline_shape_size_independent_properties():
    """N = 2^20, D = 256, K = 65536, default search -- properties that need no oracle:
      * idempotence: quantising the quantised latents (exact codebook rows) returns the same codes, residual exactly 0;
      * row independence: a row's result does not depend on where it sits (permuting the rows permutes the results bitwise:
        tile position, wave, the split tail round and the merge kernel must all be invisible);
      * linearity of the codebook gradient in the loss weight (power-of-two scaling is exact in floating point);
      * a checksum: the column sums of dE equal beta * 2 / (N D) * sum_i (E[idx_i] - z_i)."""
    F = _kvq().functional
    N, D, K, beta = 1 << 20, 256, 65536, 0.25
    z, gz, E = _headline_inputs(N, D, K)
    del gz
    loss, z_q, perp, idx, hist = F.vq_forward(z, E, beta, mode="auto")
    # idempotence (rows of a random codebook are distinct, so the nearest code of E[k] is k at distance exactly 0)
    q = E[idx]
    loss_q, z_qq, perp_q, idx_q, hist_q = F.vq_forward(q, E, beta, mode="auto")
    assert torch.equal(idx_q, idx) and torch.equal(z_qq, q) and float(loss_q) == 0.0
    assert torch.equal(hist_q, hist) and abs(float(perp_q) - float(perp)) <= 1e-6 * float(perp)
    del z_qq, q
    # row independence
    perm = torch.randperm(N, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
    zp = z[perm].contiguous()
    loss_p, z_qp, perp_p, idx_p, hist_p = F.vq_forward(zp, E, beta, mode="auto")
    assert torch.equal(idx_p, idx[perm]) and torch.equal(z_qp, z_q[perm]) and torch.equal(hist_p, hist)
    assert abs(float(loss_p) - float(loss)) <= 1e-6 * float(loss)          # summation order differs
    del zp, z_qp, idx_p
    # linearity of dE in the loss weight, and its checksum
    one, two = torch.tensor(1.0, device=DEV), torch.tensor(2.0, device=DEV)
    _, dE1 = F.vq_backward(z, E, idx, hist, beta, g_loss=one, need_dz=False)
    _, dE2 = F.vq_backward(z, E, idx, hist, beta, g_loss=two, need_dz=False)
    assert torch.equal(dE2, 2.0 * dE1)
    want = (E[idx].double().sum(0) - z.double().sum(0)) * (2.0 * beta / (N * D))
    got = dE1.double().sum(0)
    assert float((got - want).abs().max()) <= 1e-6 * float(want.abs().max()) + 1e-12


@pytest.mark.parametrize("shape", [(1 << 20, 256, 65536, "normal"), (1 << 20, 256, 4096, "collapsed"),
                                   (100000, 128, 1000, "skewed"), (4096, 768, 512, "normal"), (77, 32, 5, "normal")])
def test_codebook_gradient_is_bitwise_reproducible(shape):
    """dE must not depend on scheduling: the segment order is fixed by a stable sort, segments cut by warp ranges are
    combined in a fixed order (no floating-point atomics).  Also times the collapsed / skewed usage cases."""
    F = _kvq().functional
    N, D, K, kind = shape
    gen = torch.Generator(device=DEV).manual_seed(5)
    z = torch.randn(N, D, device=DEV, generator=gen)
    E = torch.randn(K, D, device=DEV, generator=gen)
    gz = torch.randn(N, D, device=DEV, generator=gen)
    if kind == "collapsed":
        idx = torch.full((N,), 3, dtype=torch.int64, device=DEV)            # every latent on one code
    elif kind == "skewed":
        idx = (torch.rand(N, device=DEV, generator=gen) ** 6 * K).long().clamp_(0, K - 1)   # a few very hot codes
    else:
        idx = torch.randint(0, K, (N,), device=DEV, generator=gen)
    hist = torch.bincount(idx, minlength=K).to(torch.int32)
    gl = torch.tensor(0.7, device=DEV)
    outs = []
    for _ in range(3):
        dz, dE = F.vq_backward(z, E, idx, hist, 0.25, g_zq=gz, g_loss=gl)
        outs.append((dz, dE))
    torch.cuda.synchronize()
    for dz, dE in outs[1:]:
        assert torch.equal(dE, outs[0][1]) and torch.equal(dz, outs[0][0])
    q = E[idx]
    c1 = 0.7 * 2.0 / (N * D)
    ref = torch.zeros(K, D, device=DEV, dtype=torch.float64).index_add_(0, idx, (q - z).double()) * (c1 * 0.25)
    tol = 1e-5 if kind != "collapsed" else 2e-4          # 2^20 fp32 terms in one row
    assert float((outs[0][1].double() - ref).abs().max()) <= tol * float(ref.abs().max())
    assert torch.allclose(outs[0][0], gz + c1 * (z - q), rtol=1e-5, atol=1e-7)
    assert bool((outs[0][1][hist == 0] == 0).all())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        F.vq_backward(z, E, idx, hist, 0.25, g_zq=gz, g_loss=gl)
    e1.record(); torch.cuda.synchronize()
    print(f"backward[{kind}, N={N}, D={D}, K={K}] {e0.elapsed_time(e1) / 5:.3f} ms per call (sort + segmented pass + fix-up)")


@pytest.mark.parametrize("search", ["fp32", "tf32", "auto"])
def test_nan_is_the_minimum_like_torch_argmin(search):
    """models/shelgon3/VectorQuantizer.py:65: torch.argmin treats NaN as the smallest value (first NaN wins)."""
    F = _kvq().functional
    g = torch.Generator().manual_seed(11)
    N, D, K = 600, 64, 700
    z = torch.randn(N, D, generator=g)
    E = torch.randn(K, D, generator=g)
    E[333, 5] = float("nan")          # every distance to code 333 is NaN -> every row picks 333 ...
    E[500, 0] = float("nan")          # ... not the later NaN code
    z[17, 3] = float("nan")           # a NaN latent: every distance of the row is NaN -> code 0
    z[400, :] = float("nan")
    ref = torch.argmin(O.distances_fp32(z, E), dim=1)
    assert int(ref[17]) == 0 and int(ref[0]) == 333
    idx, _ = F.search(z.to(DEV), E.to(DEV), mode=search)
    assert torch.equal(idx.cpu(), ref)
    # a large problem (tensor-core path without a code-range split) with one poisoned code and one poisoned latent
    N2, K2 = 40000, 2048
    z2 = torch.randn(N2, 64, generator=g); E2 = torch.randn(K2, 64, generator=g)
    E2[1234, 7] = float("nan"); z2[39999, 0] = float("nan")
    idx2, _ = F.search(z2.to(DEV), E2.to(DEV), mode=search)
    assert int(idx2[39999]) == 0 and bool((idx2[:39999] == 1234).all())
    # whole layer: the NaN propagates into the loss like it does in the reference
    loss, z_q, perp, idx3, hist = F.vq_forward(z.to(DEV), E.to(DEV), 0.25, mode=search)
    assert math.isnan(float(loss)) and torch.equal(idx3.cpu(), ref)


def test_tf32_search_returns_idx_together_with_accumulated_keys():
    """kvq_search(mode=tf32, idx and keys with keys_accumulate): idx must be written (it used to be left uninitialised)."""
    F = _kvq().functional
    g = torch.Generator().manual_seed(3)
    z = torch.randn(5000, 128, generator=g).to(DEV)
    E = torch.randn(900, 128, generator=g).to(DEV)
    for mode in ("tf32", "fp32", "auto"):
        plain, _ = F.search(z, E, mode=mode, want_idx=True)
        idx, keys = F.search(z, E, mode=mode, want_idx=True, want_keys=True)
        assert torch.equal(F.keys_to_idx(keys), idx)
        if mode != "auto":
            assert torch.equal(idx, plain)


def test_small_shape_step_is_a_handful_of_launches():
    """The reference's own shapes (BASELINE config 1: N = 4096, D = 768, K = 512) are launch-latency bound: the whole
    forward is 4 kernels (norms + histogram clear + key fill | search | gather + idx | finalize), the deterministic backward
    4 (two single-block sort passes, segmented pass, fix-up).  Also prints the eager and CUDA-graph step times."""
    k = _kvq()
    from kindergarten_vq_vae_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(1)
    z = torch.randn(64, 64, 768, generator=g).to(DEV).requires_grad_(True)
    E = torch.randn(512, 768, generator=g)
    gz = torch.randn(64, 64, 768, generator=g).to(DEV)
    vq = k.VectorQuantizer(512, 768, 0.25, vq_codebook_init_values=E, min_encodings=False).to(DEV)
    one = torch.ones((), device=DEV)

    def step():
        z.grad = None
        vq.embedding.weight.grad = None
        loss, z_q, perp, _, idx = vq.forward(z, DEV)
        torch.autograd.backward([loss, z_q], [one, gz])
        return idx
    for _ in range(3):
        idx = step()
    torch.cuda.synchronize()
    n0 = lib.kvq_launch_count()
    idx = step()
    torch.cuda.synchronize()
    launches = lib.kvq_launch_count() - n0
    assert launches <= 8, launches
    ref = O.forward_fp32(z.detach().cpu(), E, 0.25)
    assert O.index_parity(idx.cpu(), ref.idx, z.detach().cpu(), E, exact_fp32=True).unexcused == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        step()
    e1.record(); torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1) / 50
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    z.grad = None
    vq.embedding.weight.grad = None
    with torch.cuda.graph(graph):
        step()
    graph.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"config-1 shape fwd+bwd: {launches} library kernels per step, {eager_ms:.3f} ms eager, "
          f"{e0.elapsed_time(e1) / 50:.3f} ms as one CUDA graph")


def test_dz_only_backward_passes_the_upstream_gradient_through_for_foreign_codes():
    """kvq_backward on a code shard, dz only: latents whose code another shard owns get dz = g_zq (not garbage)."""
    F = _kvq().functional
    g = torch.Generator().manual_seed(2)
    N, D, K = 500, 64, 100
    z = torch.randn(N, D, generator=g).to(DEV)
    E = torch.randn(K, D, generator=g).to(DEV)
    gz = torch.randn(N, D, generator=g).to(DEV)
    idx = torch.randint(0, K, (N,), generator=g).to(DEV)
    gl = torch.tensor(2.0, device=DEV)
    lo, hi = 30, 70                                            # this "rank" owns codes [30, 70)
    dz, _ = F.vq_backward(z, E[lo:hi].contiguous(), idx, None, 0.25, g_zq=gz, g_loss=gl, need_dz=True, need_dE=False,
                          k_offset=lo, n_global=N)
    mine = (idx >= lo) & (idx < hi)
    c1 = 2.0 * 2.0 / (N * D)
    assert torch.equal(dz[~mine], gz[~mine])
    assert torch.allclose(dz[mine], gz[mine] + c1 * (z[mine] - E[idx[mine]]), rtol=1e-5, atol=1e-7)
    with pytest.raises(RuntimeError, match="separate calls"):
        F.vq_backward(z, E[lo:hi].contiguous(), idx, torch.zeros(hi - lo, dtype=torch.int32, device=DEV), 0.25, g_zq=gz,
                      g_loss=gl, need_dz=True, need_dE=True, k_offset=lo)
