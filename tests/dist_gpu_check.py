"""Multi-GPU parity check, launched under torchrun by tests/test_gpu_multi.py (one rank per GPU, NCCL).

Every rank builds the same seeded problem and evaluates the ORACLE on it (oracle.vq_oracle on the CPU: the reference's
fp32 evaluation order for the forward, closed-form float64 gradients for the backward).  The batch-sharded and the
codebook-sharded layers, each with both exchanges, are then run across the ranks and compared with the oracle using the
tolerances of tests/test_gpu_parity.py."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    import kindergarten_vq_vae_b200 as kvq
    from oracle import vq_oracle as O
    torch.set_num_threads(max(1, (os.cpu_count() or 8) // world))

    for (B, S, D, K, search) in [(8 * world, 64, 128, 1000, "fp32"), (16 * world, 64, 256, 4096, "auto")]:
        g = torch.Generator().manual_seed(1234)
        z = torch.randn(B, S, D, generator=g)
        E = torch.randn(K, D, generator=g)
        gz = torch.randn(B, S, D, generator=g)
        beta, w = 0.25, 1.5
        exact = search == "fp32"
        # ---- the referee: the oracle on the whole problem (identical on every rank by construction)
        ref = O.forward_fp32(z, E, beta)
        dz_ref, dE_ref = O.backward_closed_form(z, E, ref.idx, beta, g_zq=gz, g_loss=w)
        idx0 = ref.idx.reshape(-1)

        def check_rows(tag, idx1, zq1, rows_z, rows_ref_idx, rows_ref_zq):
            par = O.index_parity(idx1.cpu(), rows_ref_idx, rows_z, E, exact_fp32=exact)
            assert par.unexcused == 0, f"{tag}: {par}"
            same = (idx1.cpu().reshape(-1) == rows_ref_idx.reshape(-1))
            assert torch.equal(zq1.detach().cpu().reshape(-1, D)[same], rows_ref_zq.reshape(-1, D)[same]), tag
            return par

        # ---- batch-sharded: NCCL all-reduce of dE, then the all-reduce fused into the scatter-add kernel ----
        lo, hi = rank * B // world, (rank + 1) * B // world
        for exchange in ("nccl", "nvlink"):
            vq = kvq.BatchShardedVectorQuantizer(K, D, beta, vq_codebook_init_values=E, search=search,
                                                 exchange=exchange).to(dev)
            for rep in range(2):            # twice: the symmetric gradient buffer is reused across steps
                vq.zero_grad()
                zl = z[lo:hi].to(dev).requires_grad_(True)
                l1, q1, p1, _, i1 = vq.forward(zl, dev)
                (l1 * w + (q1 * gz[lo:hi].to(dev)).sum()).backward()
            par = check_rows(f"batch-sharded[{exchange}]", i1, q1, z[lo:hi], ref.idx[lo:hi], ref.z_q[lo:hi])
            bad = torch.tensor([par.raw_mismatch], device=dev); dist.all_reduce(bad)
            if int(bad) == 0:               # every rank agrees with the oracle row for row: global quantities must match
                assert abs(float(l1) - float(ref.loss)) <= 2e-5 * float(ref.loss), (float(l1), float(ref.loss))
                assert abs(float(p1) - float(ref.perplexity)) <= 2e-5 * float(ref.perplexity)
                assert torch.allclose(zl.grad.cpu(), dz_ref[lo:hi].float(), rtol=1e-5, atol=1e-7)
                dE = vq.embedding.weight.grad.cpu()
                err = float((dE - dE_ref.float()).abs().max()); scale = float(dE_ref.abs().max())
                assert err <= 2e-5 * scale, (exchange, err, scale)
                assert bool((dE[torch.bincount(idx0, minlength=K) == 0] == 0).all())
            if rank == 0:
                mc = vq._grad_peer.multicast_ptr if vq._grad_peer is not None else None
                print(f"  batch-sharded exchange={exchange}: rows differing from the oracle (all ranks) {int(bad)} "
                      f"multicast_ptr={mc}", flush=True)

        # ---- codebook-sharded: fused NVLink peer-memory exchange (the default when it is available), then NCCL ----
        for exchange in ("auto", "nvlink", "nccl"):
            cq = kvq.CodebookShardedVectorQuantizer(K, D, beta, vq_codebook_init_values=E, search=search,
                                                    exchange=exchange).to(dev)
            for rep in range(2):            # twice: the peer buffers are reused across forwards
                cq.zero_grad()
                zc = z.to(dev).requires_grad_(True)
                l2, q2, p2, _, i2 = cq.forward(zc, dev)
                (l2 * w + (q2 * gz.to(dev)).sum()).backward()
            par = check_rows(f"codebook-sharded[{exchange}]", i2, q2, z, ref.idx, ref.z_q)
            if par.raw_mismatch == 0:
                assert abs(float(l2) - float(ref.loss)) <= 2e-5 * float(ref.loss), exchange
                assert abs(float(p2) - float(ref.perplexity)) <= 2e-5 * float(ref.perplexity), exchange
                assert torch.allclose(zc.grad.cpu(), dz_ref.float(), rtol=1e-4, atol=1e-6), exchange
                dE_shard = dE_ref[cq.k_offset:cq.k_offset + cq.k_valid].float()
                dE = cq.embedding.weight.grad[: cq.k_valid].cpu()
                assert float((dE - dE_shard).abs().max()) <= 2e-5 * float(dE_ref.abs().max()), exchange
            if rank == 0:
                print(f"  codebook-sharded exchange={exchange} (resolved: {cq.exchange}): rows differing from the oracle "
                      f"{par.raw_mismatch}/{par.n}, unexcused {par.unexcused}", flush=True)
        if rank == 0:
            print(f"dist check OK: world={world} B={B} S={S} D={D} K={K} {search}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
