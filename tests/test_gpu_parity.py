"""Parity of the CUDA path (through the C ABI) with the reference: committed golden vectors produced by the
unmodified reference class, the CPU oracle on seeded inputs, and size-independent properties at large sizes.

Tolerances (stated once, used throughout):
  * code indices: bit-exact, except rows whose fp64 distance gap between our code and the reference's is below
    the tolerance of oracle.vq_oracle.tf32_tolerance (tf32 search) or a few fp32 ulps of the distance (fp32 search);
    such rows are counted and reported, an unexcused row fails the test;
  * z_q: bit-exact on rows with equal indices (same fp32 expression z + (E[idx] - z));
  * loss, perplexity: 2e-5 relative (fp32 reductions in a different order; perplexity also 1 ulp of logf);
  * dz: rtol 1e-5 / atol 1e-7;  dE: 1e-5 of max|dE| (fp32 bucket sums in a different order).
"""
import math

import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _kvq():
    import kindergarten_vq_vae_b200 as k
    return k


def run_module(z, E, beta, gz=None, w=1.0, search="auto", min_encodings="auto", inplace_scale=False):
    k = _kvq()
    K, D = E.shape
    vq = k.VectorQuantizer(K, D, beta, vq_codebook_init_values=E, search=search, min_encodings=min_encodings).to(DEV)
    zin = z.to(DEV).requires_grad_(gz is not None)
    loss, z_q, perp, onehot, idx = vq.forward(zin, DEV)
    out = dict(loss=loss.detach().cpu().clone(), z_q=z_q.detach().cpu(), perplexity=perp.cpu(), idx=idx.cpu(),
               onehot=None if onehot is None else onehot.cpu())
    if gz is not None:
        if inplace_scale:
            loss *= w                                    # models/shelgon3/Trainer.py:104 multiplies in place
            total = loss + (z_q * gz.to(DEV)).sum()
        else:
            total = loss * w + (z_q * gz.to(DEV)).sum()
        total.backward()
        out["dz"] = zin.grad.cpu()
        out["dE"] = vq.embedding.weight.grad.cpu()
    torch.cuda.synchronize()
    return out


def check_against(out, ref, z, E, search, expect_exact_idx=False):
    """ref: dict with idx, z_q, loss, perplexity, optionally dz, dE (reference / oracle values)."""
    D = E.shape[1]
    par = O.index_parity(out["idx"], ref["idx"], z, E, exact_fp32=(search == "fp32"))
    assert par.unexcused == 0, f"unexcused index mismatches: {par}"
    if expect_exact_idx:
        assert par.raw_mismatch == 0, f"{par}"
    same = (out["idx"].reshape(-1) == ref["idx"].reshape(-1))
    assert torch.equal(out["z_q"].reshape(-1, D)[same], ref["z_q"].reshape(-1, D)[same])
    if par.raw_mismatch == 0:
        assert abs(float(out["loss"]) - float(ref["loss"])) <= 2e-5 * abs(float(ref["loss"]))
        assert abs(float(out["perplexity"]) - float(ref["perplexity"])) <= 2e-5 * float(ref["perplexity"])
        if "dz" in out and "dz" in ref:
            assert torch.allclose(out["dz"], ref["dz"].float(), rtol=1e-5, atol=1e-7)
            assert (out["dE"] - ref["dE"].float()).abs().max() <= 1e-5 * ref["dE"].abs().max()
    return par


@pytest.mark.parametrize("search", ["fp32", "tf32", "auto"])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_vectors(name, search):
    g = load_golden(name)
    out = run_module(g["z"], g["E"], float(g["beta"]), g["gz"], float(g["w"]), search=search)
    par = check_against(out, g, g["z"], g["E"], search, expect_exact_idx=(name == "ties" or search == "fp32"))
    assert out["onehot"] is not None
    assert torch.equal(out["onehot"], O.onehot(out["idx"], g["E"].shape[0]))       # VectorQuantizer.py:67-68
    assert out["idx"].shape == g["idx"].shape and out["idx"].dtype == torch.int64
    print(f"golden[{name},{search}] index mismatches {par.raw_mismatch}/{par.n} (all excused)")


def test_golden_with_inplace_loss_scaling():
    g = load_golden("small")
    out = run_module(g["z"], g["E"], float(g["beta"]), g["gz"], float(g["w"]) * 3.0, search="fp32", inplace_scale=True)
    dz, dE = O.backward_closed_form(g["z"], g["E"], g["idx"], float(g["beta"]), g_zq=g["gz"], g_loss=float(g["w"]) * 3.0)
    assert torch.allclose(out["dz"], dz.float(), rtol=1e-5, atol=1e-7)
    assert (out["dE"] - dE.float()).abs().max() <= 1e-5 * dE.abs().max()


def _seeded(B, S, D, K, init, seed=69):
    gen = torch.Generator().manual_seed(seed)
    z = torch.randn(B, S, D, generator=gen)
    if init == "default":
        E = (torch.rand(K, D, generator=gen) * 2 - 1) / K
    elif init == "normal":
        E = torch.randn(K, D, generator=gen)
    elif init == "collapsed":            # every latent picks code 3: worst case for histogram / bucket atomics
        E = torch.randn(K, D, generator=gen) * 50
        E[3] = 0.0
    gz = torch.randn(B, S, D, generator=gen)
    return z, E, gz


SHAPES = [
    # B, S,  D,    K,  init       -- ragged N, K not multiples of the tiles, tiny and single-code books
    (1, 1, 32, 1, "normal"),
    (1, 1, 256, 512, "normal"),
    (3, 7, 64, 10, "normal"),
    (5, 13, 128, 300, "normal"),
    (2, 129, 256, 257, "normal"),
    (4, 100, 96, 1000, "normal"),
    (8, 12, 768, 512, "default"),       # BERT latents, reference default init (streaming-operand path, D > 256)
    (8, 12, 768, 512, "normal"),
    (64, 64, 768, 512, "normal"),       # BASELINE config 1
    (3, 50, 512, 300, "normal"),        # D = 512, 1024: streaming-operand path, widest rows of the bandwidth kernels
    (2, 33, 1024, 64, "normal"),
    (1, 5000, 32, 3000, "normal"),      # narrowest D the tensor-core kernel takes
    (2, 300, 256, 2048, "collapsed"),
    (1, 200, 256, 16384, "normal"),     # few latents, many codes: code-range split + atomicMin merge
    (2, 300, 64, 140000, "normal"),     # code indices beyond 16 bits (K-sharded shards are this size), K % 256 != 0
]


@pytest.mark.parametrize("search", ["fp32", "tf32", "auto"])
@pytest.mark.parametrize("B,S,D,K,init", SHAPES)
def test_oracle_parity_shapes(B, S, D, K, init, search):
    z, E, gz = _seeded(B, S, D, K, init)
    ref_f = O.forward_fp32(z, E, 0.25)
    dz, dE = O.backward_closed_form(z, E, ref_f.idx, 0.25, g_zq=gz, g_loss=0.7)
    ref = dict(idx=ref_f.idx, z_q=ref_f.z_q, loss=ref_f.loss, perplexity=ref_f.perplexity, dz=dz, dE=dE)
    out = run_module(z, E, 0.25, gz, 0.7, search=search, min_encodings=False)
    par = check_against(out, ref, z, E, search)
    if par.raw_mismatch:   # gradients follow OUR indices: re-derive the closed form on them
        dz2, dE2 = O.backward_closed_form(z, E, out["idx"], 0.25, g_zq=gz, g_loss=0.7)
        assert torch.allclose(out["dz"], dz2.float(), rtol=1e-5, atol=1e-7)
        assert (out["dE"] - dE2.float()).abs().max() <= 1e-5 * dE2.abs().max()
    print(f"shape[{B}x{S}x{D},K={K},{init},{search}] mismatches {par.raw_mismatch}/{par.n}")


def test_fp32_only_shape_and_shape_errors():
    k = _kvq()
    z, E, gz = _seeded(2, 9, 36, 40, "normal")           # D % 32 != 0: auto picks the fp32 search
    ref = O.forward_fp32(z, E, 0.25)
    out = run_module(z, E, 0.25, gz, 1.0, search="auto")
    assert O.index_parity(out["idx"], ref.idx, z, E, exact_fp32=True).unexcused == 0
    with pytest.raises(RuntimeError, match="D %"):
        run_module(z, E, 0.25, search="tf32")            # explicit tf32 on that shape: explicit error, no fallback
    with pytest.raises(RuntimeError, match="multiple of 4"):
        run_module(torch.randn(2, 3, 6), torch.randn(5, 6), 0.25)
    vq = k.VectorQuantizer(8, 32, 0.25).to(DEV)
    with pytest.raises(RuntimeError):
        vq.forward(torch.randn(4, 6, 32, device=DEV).transpose(0, 1), DEV)   # non-contiguous: .view fails like the reference


def test_empty_batch_like_the_reference():
    """No latents at all: the reference's means over zero elements give NaN loss / perplexity, empty tensors everywhere
    else, and autograd hands the codebook an all-zero gradient (VectorQuantizer.py:55-93 run on a (0, S, D) input)."""
    k = _kvq()
    vq = k.VectorQuantizer(8, 32, 0.25).to(DEV)
    z = torch.zeros(0, 3, 32, device=DEV, requires_grad=True)
    loss, z_q, perp, onehot, idx = vq.forward(z, DEV)
    assert loss.shape == () and torch.isnan(loss) and perp.shape == () and torch.isnan(perp)
    assert z_q.shape == (0, 3, 32) and idx.shape == (0, 3, 1) and idx.dtype == torch.int64
    assert onehot is not None and onehot.shape == (0, 8)
    (loss + z_q.sum()).backward()
    assert float(vq.embedding.weight.grad.abs().max()) == 0.0 and z.grad.shape == (0, 3, 32)


def test_frozen_encoder_and_eval_modes():
    k = _kvq()
    z, E, gz = _seeded(4, 16, 64, 32, "normal")
    vq = k.VectorQuantizer(32, 64, 0.25, vq_codebook_init_values=E).to(DEV)
    # "vq-ft" (Shelgon.py:168-177): encoder frozen -> z has no grad, only the codebook trains
    loss, z_q, *_ = vq.forward(z.to(DEV), DEV)
    (loss * 2.0).backward()
    ref = O.forward_fp32(z, E, 0.25)
    _, dE = O.backward_closed_form(z, E, ref.idx, 0.25, g_loss=2.0)
    assert (vq.embedding.weight.grad.cpu() - dE.float()).abs().max() <= 1e-5 * dE.abs().max()
    # eval under no_grad (Trainer.py:365)
    with torch.no_grad():
        loss2, z_q2, perp2, _, idx2 = vq.forward(z.to(DEV), DEV)
    assert not loss2.requires_grad and torch.equal(idx2.cpu(), ref.idx)
    # decoder gradient only (loss unused): passes straight through to z, codebook gets exact zeros
    vq.zero_grad()
    zin = z.to(DEV).requires_grad_(True)
    _, z_q3, *_ = vq.forward(zin, DEV)
    (z_q3 * gz.to(DEV)).sum().backward()
    assert torch.equal(zin.grad.cpu(), gz) and float(vq.embedding.weight.grad.abs().max()) == 0.0


def test_keys_accumulate_across_codebook_shards():
    F = _kvq().functional
    z, E, _ = _seeded(2, 150, 128, 1000, "normal")
    zf, Ed = z.view(-1, 128).to(DEV), E.to(DEV)
    for mode in ("fp32", "tf32"):
        full, _ = F.search(zf, Ed, mode=mode)
        keys = None
        for lo, hi in ((0, 400), (400, 1000)):
            _, keys = F.search(zf, Ed[lo:hi].contiguous(), mode=mode, k_offset=lo, want_idx=False, keys=keys,
                               keys_accumulate=keys is not None, want_keys=True)
        merged = F.keys_to_idx(keys)
        par = O.index_parity(merged.cpu(), full.cpu(), z, E, exact_fp32=(mode == "fp32"))
        assert par.unexcused == 0 and par.raw_rate < 0.01


def test_split_tail_round_of_the_tensor_core_search():
    """More row groups than CTA pairs with a partly filled last round (20000 rows = 79 groups of 256 for 74 pairs): the
    5 trailing groups are cut into code ranges -- merged from records by the top-2 search (default mode) and by the plain
    search that writes idx directly, by the key atomics when keys are accumulated.  Results must equal the unsplit search's (KVQ_TF32_TAIL_SPLIT=0) row by row."""
    import os
    F = _kvq().functional
    gen = torch.Generator().manual_seed(7)
    N, D, K = 20000, 64, 8192
    z = torch.randn(N, D, generator=gen); E = torch.randn(K, D, generator=gen)
    zf, Ed = z.to(DEV), E.to(DEV)
    out = {}
    for split in ("1", "0"):
        os.environ["KVQ_TF32_TAIL_SPLIT"] = split
        try:
            idx_auto = F.vq_forward(zf, Ed, 0.25, mode="auto")[3]
            _, keys = F.search(zf, Ed[:5000].contiguous(), mode="tf32", want_idx=False, want_keys=True)
            _, keys = F.search(zf, Ed[5000:].contiguous(), mode="tf32", k_offset=5000, want_idx=False, keys=keys,
                               keys_accumulate=True, want_keys=True)
            idx_plain, _ = F.search(zf, Ed, mode="tf32")              # plain search writing idx directly: records too
            out[split] = (idx_auto.cpu(), F.keys_to_idx(keys).cpu(), idx_plain.cpu())
        finally:
            os.environ.pop("KVQ_TF32_TAIL_SPLIT", None)
    assert torch.equal(out["1"][0], out["0"][0])                  # default mode: exact pass on the same top-2 pairs
    assert torch.equal(out["1"][1], out["0"][1])                  # plain tf32 keys: the same scores, the same minimum
    assert torch.equal(out["1"][2], out["0"][2])
    rows = torch.arange(N - 2048, N)                              # the tail rows (and some before) against the oracle
    ref = O.forward_fp32(z[rows], E, 0.25)
    assert O.index_parity(out["1"][0][rows], ref.idx, z[rows], E).unexcused == 0
    assert O.index_parity(out["1"][1][rows], ref.idx, z[rows], E).unexcused == 0


def test_sharded_modules_world1_equal_plain_module():
    k = _kvq()
    z, E, gz = _seeded(4, 40, 64, 96, "normal")
    base = run_module(z, E, 0.25, gz, 1.0, search="fp32", min_encodings=False)
    for cls in (k.BatchShardedVectorQuantizer, k.CodebookShardedVectorQuantizer):
        vq = cls(96, 64, 0.25, vq_codebook_init_values=E, search="fp32").to(DEV)
        zin = z.to(DEV).requires_grad_(True)
        loss, z_q, perp, _, idx = vq.forward(zin, DEV)
        (loss + (z_q * gz.to(DEV)).sum()).backward()
        assert torch.equal(idx.cpu(), base["idx"]) and torch.equal(z_q.detach().cpu(), base["z_q"])
        assert abs(float(loss) - float(base["loss"])) <= 1e-6 * float(base["loss"])
        assert torch.allclose(zin.grad.cpu(), base["dz"], rtol=1e-4, atol=1e-6)
        assert (vq.embedding.weight.grad.cpu() - base["dE"]).abs().max() <= 1e-5 * base["dE"].abs().max()


# ---- large sizes: properties that need no N x K oracle ------------------------------------------------------

def _device_inputs(N, D, K, init, seed=69):
    gen = torch.Generator(device=DEV).manual_seed(seed)
    z = torch.randn(N, D, device=DEV, generator=gen)
    if init == "default":
        E = (torch.rand(K, D, device=DEV, generator=gen) * 2 - 1) / K
    elif init == "points":
        pick = torch.randperm(N, device=DEV, generator=gen)[:K]
        E = z[pick] + 0.1 * torch.randn(K, D, device=DEV, generator=gen)
    else:
        E = torch.randn(K, D, device=DEV, generator=gen)
    return z, E


@pytest.mark.parametrize("N,D,K,init", [(1 << 17, 256, 8192, "normal"), (1 << 17, 256, 8192, "points"),
                                        (1 << 16, 256, 65536, "normal"), (1 << 16, 256, 8192, "default")])
def test_large_properties(N, D, K, init):
    F = _kvq().functional
    z, E = _device_inputs(N, D, K, init)
    beta = 0.25
    loss, z_q, perp, idx, hist = F.vq_forward(z, E, beta, mode="tf32")
    torch.cuda.synchronize()
    assert int(hist.sum()) == N and int(idx.min()) >= 0 and int(idx.max()) < K
    assert torch.equal(hist.long(), torch.bincount(idx, minlength=K))
    q = E[idx]
    assert torch.equal(z_q, z + (q - z))                                     # straight-through value, bitwise
    m = ((q - z).double() ** 2).mean()
    assert abs(float(loss) - float(m * (1 + beta))) <= 1e-5 * float(m * (1 + beta))
    p = hist.double() / N
    assert abs(float(perp) - math.exp(-float((p * torch.log(p + 1e-10)).sum()))) <= 1e-4 * float(perp)
    # optimality of the chosen code on a sample of rows, judged in fp64 with the stated tf32 tolerance
    rows = torch.randperm(N, device=DEV)[:1024]
    zs = z[rows].double()
    d = (zs * zs).sum(1, keepdim=True) + (E.double() ** 2).sum(1) - 2.0 * zs @ E.double().t()
    best = d.min(1).values
    chosen = d.gather(1, idx[rows, None]).squeeze(1)
    tol = 2.0 ** -9 * zs.norm(dim=1) * E.double().norm(dim=1).max() + 4 * torch.pow(2.0, torch.floor(torch.log2(chosen)) - 23)
    assert bool(((chosen - best) <= tol).all()), float(((chosen - best) / tol).max())
    # agreement with the fp32 CUDA-core search
    idx32, _ = F.search(z, E, mode="fp32")
    rate = float((idx32 != idx).float().mean())
    print(f"large[N={N},K={K},{init}] tf32 vs fp32 search disagreement {rate:.5f}")
    assert rate < (0.05 if init == "default" else 0.005)
    # backward: closed forms on the device, against our own indices
    g = torch.randn(N, D, device=DEV)
    gl = torch.tensor(1.5, device=DEV)
    dz, dE = F.vq_backward(z, E, idx, hist, beta, g_zq=g, g_loss=gl)
    c1 = 1.5 * 2.0 / (N * D)
    assert torch.allclose(dz, g + c1 * (z - q), rtol=1e-5, atol=1e-7)
    ref = torch.zeros(K, D, device=DEV, dtype=torch.float64).index_add_(0, idx, (q - z).double()) * (c1 * beta)
    assert float((dE.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    assert bool((dE[hist == 0] == 0).all())                                   # dense gradient, exact zeros


def test_cuda_graph_capture_and_replay():
    """The library never synchronises or allocates, so a whole forward + backward of the layer captures into one
    CUDA graph; replays with new inputs in the same buffers must equal eager results."""
    k = _kvq()
    z0, E, gz0 = _seeded(8, 12, 768, 512, "normal", seed=5)
    z1, _, gz1 = _seeded(8, 12, 768, 512, "normal", seed=6)
    vq = k.VectorQuantizer(512, 768, 0.25, vq_codebook_init_values=E, min_encodings=False).to(DEV)
    zs = z0.to(DEV).requires_grad_(True)
    gs = gz0.to(DEV).clone()
    one = torch.ones((), device=DEV)

    def step():
        zs.grad = None
        vq.embedding.weight.grad = None
        loss, z_q, perp, _, idx = vq.forward(zs, DEV)
        torch.autograd.backward([loss, z_q], [one, gs])
        return loss, z_q, perp, idx

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    zs.grad = None
    vq.embedding.weight.grad = None
    with torch.cuda.graph(graph):
        loss_g, zq_g, perp_g, idx_g = step()
        dz_g, dE_g = zs.grad, vq.embedding.weight.grad
    for zin, gin in ((z0, gz0), (z1, gz1)):
        with torch.no_grad():
            zs.copy_(zin.to(DEV)); gs.copy_(gin.to(DEV))
        graph.replay()
        torch.cuda.synchronize()
        got = dict(loss=loss_g.detach().cpu().clone(), z_q=zq_g.detach().cpu().clone(), perplexity=perp_g.cpu().clone(),
                   idx=idx_g.cpu().clone(), dz=dz_g.cpu().clone(), dE=dE_g.cpu().clone())
        ref = run_module(zin, E, 0.25, gin, 1.0, search="auto", min_encodings=False)
        assert torch.equal(got["idx"], ref["idx"]) and torch.equal(got["z_q"], ref["z_q"])
        assert abs(float(got["loss"]) - float(ref["loss"])) <= 1e-6 * float(ref["loss"])
        assert torch.allclose(got["dz"], ref["dz"], rtol=1e-6, atol=1e-8)
        assert (got["dE"] - ref["dE"]).abs().max() <= 1e-5 * ref["dE"].abs().max()


@pytest.mark.parametrize("N,D,K,init", [(1 << 16, 256, 8192, "normal"), (1 << 15, 128, 4096, "default"),
                                        (20000, 256, 1000, "points"), (1 << 16, 64, 512, "normal")])
def test_tf32_refine_search_equals_exact_argmin(N, D, K, init):
    """search="tf32_refine": tensor-core search for the two best codes + exact float64 re-evaluation of the pair.
    Against a float64 argmin on the device the result may differ only where the true winner was not in the tf32 top
    two, which at these sizes should essentially never happen (and never beyond the tf32 tolerance)."""
    F = _kvq().functional
    z, E = _device_inputs(N, D, K, init)
    plain, _ = F.search(z, E, mode="tf32")
    refined, _ = F.search(z, E, mode="tf32_refine")
    Ed = E.double()
    e2 = (Ed * Ed).sum(1)
    truth = torch.empty(N, dtype=torch.int64, device=DEV)
    for s in range(0, N, 8192):
        truth[s:s + 8192] = (e2 - 2.0 * z[s:s + 8192].double() @ Ed.t()).argmin(1)
    miss_plain = int((plain != truth).sum())
    miss_ref = int((refined != truth).sum())
    print(f"refine[N={N},D={D},K={K},{init}] mismatches vs fp64 argmin: tf32 {miss_plain}, tf32_refine {miss_ref}")
    assert miss_ref <= max(1, miss_plain // 50)
    bad = (refined != truth).nonzero().flatten()
    if bad.numel():   # whatever is left must still be a near-tie
        zs = z[bad].double()
        d_ref = ((zs - Ed[refined[bad]]) ** 2).sum(1); d_tru = ((zs - Ed[truth[bad]]) ** 2).sum(1)
        tol = 2.0 ** -9 * zs.norm(dim=1) * Ed.norm(dim=1).max()
        assert bool(((d_ref - d_tru) <= tol).all())
    # the module accepts the mode and the rest of the layer follows the refined indices
    k = _kvq()
    vq = k.VectorQuantizer(K, D, 0.25, vq_codebook_init_values=E, search="tf32_refine", min_encodings=False).to(DEV)
    loss, z_q, perp, _, idx = vq.forward(z.view(N // 16, 16, D), DEV)
    assert torch.equal(idx.view(-1), refined) and torch.equal(z_q.view(N, D), z + (E[refined] - z))


def test_peer_entry_points_with_a_single_peer():
    """kvq_search_peers / kvq_quantize_shards / kvq_backward_peers (the fused NVLink exchanges) with a peer table that
    holds only this GPU's own buffers: same kernels and code paths as the multi-GPU runs, checkable on one device."""
    F = _kvq().functional
    z, E, gz = _seeded(3, 211, 128, 777, "normal", seed=9)
    zf, Ed, gd = z.view(-1, 128).to(DEV), E.to(DEV), gz.view(-1, 128).to(DEV)
    N, D, K = zf.shape[0], 128, 777
    for mode in ("tf32", "fp32"):
        ref_idx, _ = F.search(zf, Ed, mode=mode)
        keys = torch.full((N,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=DEV)
        F.search_peers(zf, Ed, [keys.data_ptr()], 0, mode=mode)
        assert torch.equal(F.keys_to_idx(keys), ref_idx)
        # two "shards" living on the same device: rows [0, 400) and [400, 777) padded to 400
        lo = Ed[:400].contiguous()
        hi = torch.zeros(400, D, device=DEV); hi[:377] = Ed[400:]
        zq_s, sq_s, hist_s = F.quantize_shards(zf, [lo.data_ptr(), hi.data_ptr()], 400, ref_idx, 800)
        zq, sq, hist = F.quantize(zf, Ed, ref_idx)
        assert torch.equal(zq_s, zq) and torch.equal(hist_s[:K], hist) and int(hist_s[K:].sum()) == 0
        assert abs(float(sq_s) - float(sq)) <= 1e-9 * float(sq)
        # backward with the reduction "fused" into the scatter-add (single replica, unicast red path)
        gl = torch.tensor(1.3, device=DEV)
        dz_ref, dE_ref = F.vq_backward(zf, Ed, ref_idx, hist, 0.25, g_zq=gd, g_loss=gl)
        dE_buf = torch.zeros(K, D, device=DEV)
        dz = F.vq_backward_peers(zf, Ed, ref_idx, hist, 0.25, g_zq=gd, g_loss=gl, need_dz=True, n_global=N,
                                 dE_peer_ptrs=[dE_buf.data_ptr()], dE_multicast_ptr=0, my_rank=0)
        assert torch.equal(dz, dz_ref)
        assert float((dE_buf - dE_ref).abs().max()) <= 1e-5 * float(dE_ref.abs().max())
        assert bool((dE_buf[hist == 0] == 0).all())
