"""world_size-2 gloo runs of the two sharded modules on CPU.  The local compute steps are the oracle-backed
stand-in (tests/cpu_backend.py); what is under test is the orchestration in kindergarten-vq-vae_b200/sharded.py:
which collectives run, global normalisation, shard offsets, gradient assembly."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    import cpu_backend
    import kindergarten_vq_vae_b200 as kvq
    from kindergarten_vq_vae_b200 import BatchShardedVectorQuantizer, CodebookShardedVectorQuantizer
    kvq.sharded.F = cpu_backend          # the test seam: the sharded layers call everything through `sharded.F`
    g = load_golden("wide")
    z, E, gz, beta, w = g["z"], g["E"], g["gz"], float(g["beta"]), float(g["w"])
    B = z.shape[0]
    try:
        if mode == "batch":
            vq = BatchShardedVectorQuantizer(E.shape[0], E.shape[1], beta, E)
            lo, hi = rank * B // world, (rank + 1) * B // world
            zl = z[lo:hi].clone().requires_grad_(True)
            loss, z_q, perp, _, idx = vq.forward(zl, "cpu")
            (loss * w + (z_q * gz[lo:hi]).sum()).backward()
            out = dict(loss=loss.detach(), perp=perp, z_q=z_q.detach(), idx=idx, dz=zl.grad, dE=vq.embedding.weight.grad,
                       lo=lo, hi=hi)
        else:
            if mode == "codebook_uneven":
                E = E[:511].contiguous()            # 511 codes over 2 ranks: shards of 256 and 255 (+1 padding row)
            # "auto" must settle on the NCCL-style collectives when peer memory is not available (CPU / gloo here)
            vq = CodebookShardedVectorQuantizer(E.shape[0], E.shape[1], beta, E,
                                                exchange="nccl" if mode == "codebook" else "auto")
            zl = z.clone().requires_grad_(True)
            loss, z_q, perp, _, idx = vq.forward(zl, "cpu")
            (loss * w + (z_q * gz).sum()).backward()
            out = dict(loss=loss.detach(), perp=perp, z_q=z_q.detach(), idx=idx, dz=zl.grad, dE=vq.embedding.weight.grad,
                       k_offset=vq.k_offset, k_valid=vq.k_valid)
        q.put((rank, {k: (v.detach().numpy() if torch.is_tensor(v) else v) for k, v in out.items()}))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def _run(mode, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    [p.start() for p in procs]
    raw = dict(q.get(timeout=120) for _ in range(world))
    res = {r: {k: (torch.from_numpy(v) if hasattr(v, 'dtype') and hasattr(v, 'shape') else v) for k, v in d.items()}
           for r, d in raw.items()}
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    return res


def test_batch_sharded_equals_single_device_reference():
    g = load_golden("wide")
    res = _run("batch")
    z_q = torch.cat([res[r]["z_q"] for r in sorted(res)])
    idx = torch.cat([res[r]["idx"] for r in sorted(res)])
    dz = torch.cat([res[r]["dz"] for r in sorted(res)])
    assert torch.equal(idx, g["idx"]) and torch.equal(z_q, g["z_q"])
    for r in res:   # loss / perplexity / dE are global quantities, identical on every rank
        assert torch.allclose(res[r]["loss"], g["loss"], rtol=1e-6)
        assert torch.allclose(res[r]["perp"], g["perplexity"], rtol=1e-5)
        assert (res[r]["dE"] - g["dE"]).abs().max() <= 2e-6 * g["dE"].abs().max()
    assert torch.allclose(dz, g["dz"], rtol=1e-5, atol=1e-7)


def test_codebook_sharded_equals_single_device_reference():
    from oracle import vq_oracle as O
    g = load_golden("wide")
    res = _run("codebook")
    for r in res:
        par = O.index_parity(res[r]["idx"], g["idx"], g["z"], g["E"], exact_fp32=True)
        assert par.unexcused == 0 and par.raw_mismatch == 0
        assert torch.equal(res[r]["z_q"], g["z_q"])
        assert torch.allclose(res[r]["loss"], g["loss"], rtol=1e-6)
        assert torch.allclose(res[r]["perp"], g["perplexity"], rtol=1e-5)
        assert torch.allclose(res[r]["dz"], g["dz"], rtol=1e-4, atol=1e-6)
    dE = torch.cat([res[r]["dE"][: res[r]["k_valid"]] for r in sorted(res)])
    assert (dE - g["dE"]).abs().max() <= 2e-6 * g["dE"].abs().max()


def test_codebook_sharded_uneven_shards():
    from oracle import vq_oracle as O
    g = load_golden("wide")
    E = g["E"][:511].contiguous()
    ref = O.forward_fp32(g["z"], E, float(g["beta"]))
    dz, dE_ref = O.backward_closed_form(g["z"], E, ref.idx, float(g["beta"]), g_zq=g["gz"], g_loss=float(g["w"]))
    res = _run("codebook_uneven")
    assert res[0]["k_valid"] == 256 and res[1]["k_valid"] == 255
    for r in res:
        assert torch.equal(res[r]["idx"], ref.idx) and torch.equal(res[r]["z_q"], ref.z_q)
        assert torch.allclose(res[r]["loss"], ref.loss, rtol=1e-6)
        assert torch.allclose(res[r]["perp"], ref.perplexity, rtol=1e-5)
        assert res[r]["dE"].shape[0] == 256                     # parameter keeps the padded shape
    assert torch.all(res[1]["dE"][255] == 0)                    # padding row: exact zero gradient
    dE = torch.cat([res[r]["dE"][: res[r]["k_valid"]] for r in sorted(res)])
    assert (dE - dE_ref.float()).abs().max() <= 2e-6 * dE_ref.abs().max()
