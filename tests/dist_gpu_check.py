"""Multi-GPU parity check, launched under torchrun by tests/test_gpu_multi.py (one rank per GPU, NCCL).

Every rank builds the same seeded problem, runs the single-device VectorQuantizer on the full problem as the
referee, then the batch-sharded and codebook-sharded layers across the ranks, and compares."""
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

    for (B, S, D, K, search) in [(8 * world, 64, 128, 1000, "fp32"), (16 * world, 64, 256, 4096, "tf32")]:
        g = torch.Generator().manual_seed(1234)
        z = torch.randn(B, S, D, generator=g)
        E = torch.randn(K, D, generator=g)
        gz = torch.randn(B, S, D, generator=g)
        beta, w = 0.25, 1.5
        # referee: plain module on the whole problem (every rank computes it; identical by construction)
        ref = kvq.VectorQuantizer(K, D, beta, vq_codebook_init_values=E, search=search, min_encodings=False).to(dev)
        zr = z.to(dev).requires_grad_(True)
        l0, q0, p0, _, i0 = ref.forward(zr, dev)
        (l0 * w + (q0 * gz.to(dev)).sum()).backward()

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
            same = (i1 == i0[lo:hi]).reshape(-1)
            frac = float(same.float().mean())
            assert frac > 0.999, f"batch-sharded[{exchange}] idx agreement {frac}"     # split searches may flip exact near-ties
            assert torch.equal(q1.detach().reshape(-1, D)[same], q0.detach()[lo:hi].reshape(-1, D)[same])
            fr = torch.tensor([frac], device=dev); dist.all_reduce(fr, op=dist.ReduceOp.MIN)
            if float(fr) == 1.0:
                assert abs(float(l1) - float(l0)) <= 1e-5 * float(l0), (float(l1), float(l0))
                assert abs(float(p1) - float(p0)) <= 1e-5 * float(p0)
                assert torch.allclose(zl.grad, zr.grad[lo:hi], rtol=1e-5, atol=1e-7)
                dE = vq.embedding.weight.grad
                err = float((dE - ref.embedding.weight.grad).abs().max()); scale = float(ref.embedding.weight.grad.abs().max())
                assert err <= 2e-5 * scale, (exchange, err, scale)
            if rank == 0:
                mc = vq._grad_peer.multicast_ptr if vq._grad_peer is not None else None
                print(f"  batch-sharded exchange={exchange}: idx agreement {frac:.6f} multicast_ptr={mc}", flush=True)

        # ---- codebook-sharded: NCCL exchange, then the fused NVLink peer-memory exchange ----
        for exchange in ("nccl", "nvlink"):
            cq = kvq.CodebookShardedVectorQuantizer(K, D, beta, vq_codebook_init_values=E, search=search,
                                                    exchange=exchange).to(dev)
            for rep in range(2):            # twice: the peer buffers are reused across forwards
                cq.zero_grad()
                zc = z.to(dev).requires_grad_(True)
                l2, q2, p2, _, i2 = cq.forward(zc, dev)
                (l2 * w + (q2 * gz.to(dev)).sum()).backward()
            same = (i2 == i0).reshape(-1)
            frac = float(same.float().mean())
            assert frac > 0.999, f"codebook-sharded[{exchange}] idx agreement {frac}"
            assert torch.equal(q2.detach().reshape(-1, D)[same], q0.detach().reshape(-1, D)[same]), exchange
            if frac == 1.0:
                assert abs(float(l2) - float(l0)) <= 1e-5 * float(l0), exchange
                assert abs(float(p2) - float(p0)) <= 1e-5 * float(p0), exchange
                assert torch.allclose(zc.grad, zr.grad, rtol=1e-4, atol=1e-6), exchange
                dE_ref = ref.embedding.weight.grad[cq.k_offset:cq.k_offset + cq.k_valid]
                dE = cq.embedding.weight.grad[: cq.k_valid]
                assert float((dE - dE_ref).abs().max()) <= 2e-5 * float(ref.embedding.weight.grad.abs().max()), exchange
            if rank == 0:
                print(f"  codebook-sharded exchange={exchange}: idx agreement {frac:.6f}", flush=True)
        if rank == 0:
            print(f"dist check OK: world={world} B={B} S={S} D={D} K={K} {search}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
