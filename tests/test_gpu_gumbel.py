"""GumbelQuantizer (models/shelgon3/GumbelQuantizer.py:43-83) through the C ABI against (a) fixtures produced by the
unmodified reference class with a recorded Gumbel sample and (b) the oracle on larger seeded shapes.

Tolerances.  The dense contractions run in tf32 (operands rounded to 11 significant bits, fp32 accumulate), the reference
in fp32, so values carry ~1e-3 relative error of the operand norms:
  * ind: equal, except rows whose two candidates are closer in the ORACLE's perturbed logits than the tf32 bound
    2^-9 |z_n| max_k |W_k| / tau (reported; an unexcused row fails);
  * z_q, diff, dz, dW, db, dE: 4e-3 of the largest magnitude of the reference tensor (rows with equal ind for z_q)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 4e-3


def _kvq():
    import kindergarten_vq_vae_b200 as k
    return k


def _run(z, W, b, E, noise, tau, kld, st, training, gz, w):
    k = _kvq()
    K, C = W.shape
    D = E.shape[1]
    gq = k.GumbelQuantizer(C, K, D, tau, kld, st).to(DEV)
    with torch.no_grad():
        gq.proj.weight.copy_(W.view(K, C, 1)); gq.proj.bias.copy_(b); gq.embed.weight.copy_(E)
    zin = z.to(DEV).requires_grad_(True)
    z_q, diff, ind = gq.forward(zin, training, noise=noise.to(DEV))
    (diff * w + (z_q * gz.to(DEV)).sum()).backward()
    torch.cuda.synchronize()
    return dict(z_q=z_q.detach().cpu(), diff=diff.detach().cpu(), ind=ind.cpu(), dz=zin.grad.cpu(),
                dW=gq.proj.weight.grad.cpu()[:, :, 0], db=gq.proj.bias.grad.cpu(), dE=gq.embed.weight.grad.cpu())


def _compare(out, ref, z, W, b, noise, tau):
    bad = (out["ind"] != ref["ind"]).nonzero()
    if bad.numel():
        a = (torch.matmul(z, W.t()) + b + noise) / tau
        for (i, j) in bad.tolist():
            gap = abs(float(a[i, j, out["ind"][i, j]] - a[i, j, ref["ind"][i, j]]))
            tol = 2.0 ** -9 * float(z[i, j].norm()) * float(W.norm(dim=1).max()) / tau
            assert gap <= tol, f"unexcused arg-max mismatch at {(i, j)}: gap {gap} > {tol}"
    same = out["ind"] == ref["ind"]
    assert float((out["z_q"][same] - ref["z_q"][same]).abs().max()) <= RTOL * float(ref["z_q"].abs().max())
    assert abs(float(out["diff"]) - float(ref["diff"])) <= RTOL * abs(float(ref["diff"])) + 1e-9
    if not bad.numel():
        for key in ("dz", "dW", "db", "dE"):
            err = float((out[key] - ref[key]).abs().max())
            assert err <= RTOL * float(ref[key].abs().max()) + 1e-9, (key, err, float(ref[key].abs().max()))
    return int(bad.shape[0])


@pytest.mark.parametrize("name", ["soft", "hard", "eval"])
def test_gumbel_golden_vectors_from_the_reference_class(name):
    d = np.load(os.path.join(GOLDEN, f"gumbel_{name}.npz"))
    t = {k: torch.from_numpy(np.array(d[k])) for k in d.files}
    training = name != "eval"
    st = bool(t["hard"]) if training else False
    out = _run(t["z"], t["W"], t["b"], t["E"], t["noise"], float(t["tau"]), float(t["kld_scale"]), st, training, t["gz"],
               float(t["w"]))
    n_bad = _compare(out, t, t["z"], t["W"], t["b"], t["noise"], float(t["tau"]))
    assert out["ind"].shape == t["ind"].shape and out["ind"].dtype == torch.int64
    print(f"gumbel golden[{name}]: arg-max rows differing from the reference {n_bad}/{t['ind'].numel()} (all excused)")


@pytest.mark.parametrize("B,S,C,K,D,hard", [(64, 12, 768, 512, 768, True), (64, 12, 768, 512, 768, False),
                                             (7, 5, 96, 9, 64, False), (33, 12, 256, 1000, 128, True)])
def test_gumbel_against_the_oracle(B, S, C, K, D, hard):
    g = torch.Generator().manual_seed(B * 7 + K)
    z = torch.randn(B, S, C, generator=g)
    W = (torch.rand(K, C, generator=g) * 2 - 1) / C ** 0.5          # Conv1d default init scale
    b = (torch.rand(K, generator=g) * 2 - 1) / C ** 0.5
    E = torch.randn(K, D, generator=g)
    gz = torch.randn(B, S, D, generator=g)
    noise = -torch.empty(B, S, K).exponential_(generator=g).log()
    tau, kld, w = 0.9, 5e-4, 2.0
    zr, Wr, br, Er = (x.clone().requires_grad_(True) for x in (z, W, b, E))
    z_q, diff, ind = O.gumbel_forward(zr, Wr, br, Er, noise, tau, kld, hard)
    (diff * w + (z_q * gz).sum()).backward()
    ref = dict(z_q=z_q.detach(), diff=diff.detach(), ind=ind, dz=zr.grad, dW=Wr.grad, db=br.grad, dE=Er.grad)
    out = _run(z, W, b, E, noise, tau, kld, hard, True, gz, w)
    n_bad = _compare(out, ref, z, W, b, noise, tau)
    print(f"gumbel[{B}x{S}, C={C}, K={K}, D={D}, hard={hard}]: arg-max rows differing from the oracle {n_bad}/{B * S}")


def test_gumbel_device_sample_is_gumbel_and_repeatable():
    """Without an explicit sample the layer draws Gumbel(0,1) noise on the device: same seed -> same result, and with
    constant logits the code frequencies are uniform (the Gumbel-max trick samples from softmax(logits))."""
    k = _kvq()
    K, C, D = 16, 64, 32
    gq = k.GumbelQuantizer(C, K, D, 1.0, 5e-4, True).to(DEV)
    with torch.no_grad():
        gq.proj.weight.zero_(); gq.proj.bias.zero_()
    z = torch.randn(512, 12, C, device=DEV)
    _, _, i1 = gq.forward(z, True, seed=123)
    _, _, i2 = gq.forward(z, True, seed=123)
    _, _, i3 = gq.forward(z, True, seed=124)
    assert torch.equal(i1, i2) and not torch.equal(i1, i3)
    freq = torch.bincount(i1.reshape(-1), minlength=K).float() / i1.numel()
    assert float((freq - 1.0 / K).abs().max()) < 0.02
    # state-dict keys of the reference class
    assert sorted(gq.state_dict().keys()) == ["embed.weight", "proj.bias", "proj.weight"]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        k.GumbelQuantizer(C, K, D, 1.0, 5e-4, True).forward(torch.randn(2, 3, C), True)


@pytest.mark.parametrize("M,n,Kc", [(512, 768, 24576), (300, 100, 4096), (512, 512, 1024), (24576, 512, 768)])
def test_gemm_nt_with_and_without_the_contraction_split(M, n, Kc):
    """kvq_gemm_nt on skinny outputs with a long contraction (the weight-gradient shapes): the contraction is cut into
    ranges added in a fixed order.  Same result as the unsplit run up to fp32 summation order, both within the tf32 bound
    of the fp64 product; bias and the zero padding of columns [n, ldc) survive the split; bitwise reproducible."""
    from kindergarten_vq_vae_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(M + n)
    A = torch.randn(M, Kc, device=DEV, generator=g)
    B = torch.randn(n, Kc, device=DEV, generator=g)
    bias = torch.randn(n, device=DEV, generator=g)
    ldc = (n + 31) // 32 * 32
    stream = torch.cuda.current_stream().cuda_stream

    def run(with_ws):
        C = torch.full((M, ldc), 7.0, device=DEV)
        nbytes = lib.kvq_gemm_nt_workspace_bytes(M, n, Kc, ldc) if with_ws else 0
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=DEV)
        _lib.check(lib.kvq_gemm_nt(A.data_ptr(), B.data_ptr(), M, n, Kc, C.data_ptr(), ldc, bias.data_ptr(), 0.5,
                                   ws.data_ptr() if nbytes else None, nbytes, stream), "kvq_gemm_nt")
        return C, nbytes
    C_split, nbytes = run(True)
    C_split2, _ = run(True)
    C_plain, _ = run(False)
    torch.cuda.synchronize()
    assert (nbytes > 0) == (M * n <= 512 * 768 and Kc >= 1024)          # the skinny shapes split, the tall one does not
    assert torch.equal(C_split, C_split2)
    ref = 0.5 * (A.double() @ B.double().t()) + bias.double()
    bound = 2.0 ** -9 * 0.5 * A.norm(dim=1).double()[:, None] * B.norm(dim=1).double()[None, :] + 1e-3
    for C in (C_split, C_plain):
        assert bool(((C[:, :n].double() - ref).abs() <= bound).all())
        assert bool((C[:, n:] == 0).all())
    assert float((C_split - C_plain).abs().max()) <= 1e-4 * float(ref.abs().max())
