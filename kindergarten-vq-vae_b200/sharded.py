"""Multi-GPU forms of the VQ layer (new capability: the reference is single-device).

One process per GPU, `torch.distributed` (NCCL over NVLink / NVSwitch) for the exchange steps.  The path shards in
exactly two natural ways, and each needs only reductions:

* BatchShardedVectorQuantizer -- latents are sharded by rows, the codebook is replicated.  Local search, gather,
  straight-through, dz and bucketed dE; then ONE exchange per direction:
    forward : all-reduce(SUM) of [sum of squared residuals (f64), usage histogram (i32)]
    backward: all-reduce(SUM) of the dense codebook gradient dE (K x D fp32)
  Everything is normalised by the GLOBAL N*D so the result equals the reference on the concatenated batch.

* CodebookShardedVectorQuantizer -- the codebook is sharded by rows (huge K), latents are replicated.  Each rank
  searches its shard and emits packed (score, index) int64 keys; the cross-GPU argmin is
    all-reduce(MIN) over the N keys   (signed order of the key == (score, index) order: lowest index wins ties)
  then each rank gathers the rows it owns (zeros elsewhere) and z_q is assembled with an all-reduce(SUM); the
  histogram shards are all-gathered for the perplexity.  The codebook gradient needs no communication at all:
  every rank bucket-sums only the latents that chose one of its codes.

Every local compute step goes through the module-level name `F` (the CUDA functional module, i.e. the C ABI).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn
from torch import Tensor

from . import functional as F
from .vector_quantizer import ONEHOT_AUTO_BYTES


def _world(group) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _rank(group) -> int:
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def z_is_cuda(device) -> bool:
    return torch.device(device).type == "cuda"


def shard_bounds(total: int, parts: int, part: int):
    """Contiguous equal shards (the last may be short): [lo, hi) of shard `part`."""
    per = (total + parts - 1) // parts
    lo = min(part * per, total)
    return lo, min(lo + per, total)


# ------------------------------------------------------------------------------------------------
# batch-sharded (data parallel)
# ------------------------------------------------------------------------------------------------
class _GradPeerMemory:
    """Symmetric (peer-mapped, multicast-capable where the fabric allows) buffer for the codebook gradient."""

    def __init__(self, group, K: int, D: int, device):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.buf = symm.empty(K, D, dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.buf, self.group)
        self.multicast_ptr = int(getattr(self.handle, "multicast_ptr", 0) or 0)


class _BatchShardedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, E, beta, mode, group, n_global, grad_peer=None):
        N, D = z.shape
        # local forward of the layer with the default search machinery (incl. the fused exact top-2 re-evaluation);
        # its loss / perplexity are per-rank values and are recomputed below from the all-reduced partials
        z_q, idx, sq_sum, hist_local = F.vq_forward_partials(z, E, mode=mode)
        # the two partials the loss / perplexity need (8 B and 4K B) travel as ONE float64 buffer: one all-reduce
        packed = F.pack_partials(sq_sum, hist_local)
        if _world(group) > 1:
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        loss, perplexity, hist = F.finalize_packed(packed, n_global, D, beta)
        ctx.save_for_backward(z, E, idx, hist_local)
        ctx.beta, ctx.group, ctx.n_global, ctx.grad_peer = beta, group, n_global, grad_peer
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(perplexity, idx, hist)
        return loss, z_q, perplexity, idx, hist

    @staticmethod
    def backward(ctx, g_loss, g_zq, *_):
        z, E, idx, hist_local = ctx.saved_tensors
        need_dz, need_dE = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if g_zq is not None:
            g_zq = g_zq.contiguous()
        gp = ctx.grad_peer
        if gp is not None and need_dE and _world(ctx.group) > 1:
            # fused exchange: the scatter-add kernel reduces its bucket sums into every rank's dE over NVLink / NVSwitch
            gl = (torch.zeros((), device=z.device) if g_loss is None
                  else g_loss.detach().to(torch.float32).contiguous())
            gp.buf.zero_()
            gp.handle.barrier(channel=0)                  # every replica is zero before any remote reduction lands
            dz = F.vq_backward_peers(z, E, idx, hist_local, ctx.beta, g_zq=g_zq, g_loss=gl, need_dz=need_dz,
                                                 n_global=ctx.n_global, dE_peer_ptrs=gp.handle.buffer_ptrs,
                                                 dE_multicast_ptr=gp.multicast_ptr, my_rank=_rank(ctx.group))
            gp.handle.barrier(channel=1)                  # all ranks' reductions are complete
            return dz, gp.buf.clone(), None, None, None, None, None
        if g_loss is None:
            dz = g_zq if need_dz else None
            dE = torch.zeros_like(E) if need_dE else None
        else:
            g_loss = g_loss.detach().to(torch.float32).contiguous()
            dz, dE = F.vq_backward(z, E, idx, hist_local, ctx.beta, g_zq=g_zq, g_loss=g_loss,
                                   need_dz=need_dz, need_dE=need_dE, n_global=ctx.n_global)
        if need_dE and _world(ctx.group) > 1:
            dist.all_reduce(dE, op=dist.ReduceOp.SUM, group=ctx.group)   # codebook gradient of the global batch
        return dz, dE, None, None, None, None, None


class BatchShardedVectorQuantizer(nn.Module):
    """VectorQuantizer over a row-sharded batch; outputs equal the single-device layer on the concatenated batch.
    Every rank must pass the same number of latents unless `n_global` is given to forward().  The codebook
    gradient is all-reduced here -- do not also wrap `embedding.weight` in DistributedDataParallel."""

    def __init__(self, n_e, e_dim, beta, vq_codebook_init_values: Tensor = None, *, process_group=None,
                 search: str = "auto", min_encodings=False, exchange: str = "nccl"):
        """exchange="nccl" (default; the all-reduce keeps dE bitwise reproducible): all-reduce(SUM) of dE after the backward
        kernel.  exchange="nvlink": the all-reduce is fused
        into the scatter-add kernel (multimem.red through the NVSwitch when the symmetric buffer has a multicast
        address, else one system-scope red per peer); needs torch symmetric memory, world <= 8."""
        super().__init__()
        self.n_e, self.e_dim, self.beta = n_e, e_dim, beta
        self.search, self.group = search, process_group
        if exchange not in ("nccl", "nvlink"):
            raise ValueError("exchange must be 'nccl' or 'nvlink'")
        self.exchange = exchange
        self._grad_peer = None
        self.return_min_encodings = min_encodings
        self.embedding = nn.Embedding(n_e, e_dim)
        if vq_codebook_init_values is not None:
            self.embedding.weight.data.copy_(vq_codebook_init_values)
        else:
            self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)

    @torch.compiler.disable
    def forward(self, z: Tensor, device=None, n_global: Optional[int] = None):
        batch_size, seq_len, _ = z.shape
        zf = z.view((-1, self.e_dim))
        if n_global is None:
            n_global = zf.shape[0] * _world(self.group)
        if self.exchange == "nvlink" and _world(self.group) > 1 and self._grad_peer is None:
            self._grad_peer = _GradPeerMemory(self.group, self.n_e, self.e_dim, z.device)
        loss, z_q, perplexity, idx, _ = _BatchShardedFn.apply(zf, self.embedding.weight, float(self.beta),
                                                               self.search, self.group, int(n_global), self._grad_peer)
        want = self.return_min_encodings
        if want == "auto":
            want = zf.shape[0] * self.n_e * 4 <= ONEHOT_AUTO_BYTES
        onehot = F.onehot(idx, self.n_e) if want else None
        return loss, z_q.view(z.shape), perplexity, onehot, idx.reshape((batch_size, seq_len, 1))


# ------------------------------------------------------------------------------------------------
# codebook-sharded (very large K)
# ------------------------------------------------------------------------------------------------
class _CodebookShardedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, E_param, beta, mode, group, k_offset, k_total):
        N, D = z.shape
        world = _world(group)
        k_valid = max(0, min(E_param.shape[0], k_total - k_offset))     # rows past the end of the codebook are padding
        E_local = E_param[:k_valid]
        _, keys = F.search(z, E_local, mode=mode, k_offset=k_offset, want_idx=False, want_keys=True)
        if world > 1:
            dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)      # cross-GPU (distance, index) argmin
        idx = F.keys_to_idx(keys)
        z_q, sq_sum, hist_local = F.quantize(z, E_local, idx, k_offset=k_offset, zero_skipped=True)
        if world > 1:
            dist.all_reduce(z_q, op=dist.ReduceOp.SUM, group=group)       # each row is non-zero on exactly one rank
            dist.all_reduce(sq_sum, op=dist.ReduceOp.SUM, group=group)
            per = E_param.shape[0]
            padded = torch.zeros(per, dtype=hist_local.dtype, device=hist_local.device)
            padded[:k_valid] = hist_local
            parts = [torch.empty(per, dtype=hist_local.dtype, device=hist_local.device) for _ in range(world)]
            dist.all_gather(parts, padded, group=group)
            hist = torch.cat(parts)[:k_total].contiguous()
        else:
            hist = hist_local
        loss, perplexity = F.finalize(sq_sum, hist, N, D, beta)
        ctx.save_for_backward(z, E_param, idx, hist_local, z_q)
        ctx.beta, ctx.k_offset, ctx.k_valid = beta, k_offset, k_valid
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(perplexity, idx, hist)
        return loss, z_q, perplexity, idx, hist

    @staticmethod
    def backward(ctx, g_loss, g_zq, *_):
        z, E_param, idx, hist_local, z_q = ctx.saved_tensors
        need_dz, need_dE = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if g_zq is not None:
            g_zq = g_zq.contiguous()
        if g_loss is None:
            return (g_zq if need_dz else None), (torch.zeros_like(E_param) if need_dE else None), None, None, None, \
                None, None
        g_loss = g_loss.detach().to(torch.float32).contiguous()
        dz = F.dz_from_zq(z, z_q, g_zq, g_loss, z.shape[0]) if need_dz else None
        dE = None
        if need_dE:   # local: only latents that chose one of this rank's codes contribute
            dE = _local_codebook_grad(z, E_param, ctx.k_valid, idx, hist_local, ctx.beta, g_loss, ctx.k_offset)
        return dz, dE, None, None, None, None, None


def _local_codebook_grad(z, E_param, k_valid, idx, hist_local, beta, g_loss, k_offset):
    """dE of this rank's shard (padding rows, if any, get exact zeros)."""
    _, dE = F.vq_backward(z, E_param[:k_valid], idx, hist_local, beta, g_zq=None, g_loss=g_loss, need_dz=False,
                                need_dE=True, k_offset=k_offset)
    if k_valid == E_param.shape[0]:
        return dE
    full = torch.zeros_like(E_param)
    full[:k_valid] = dE
    return full


class _PeerMemory:
    """NVLink-peer-mapped buffers of the codebook-sharded layer (torch symmetric memory = CUDA VMM + IPC handles):
    one packed-key buffer per rank (N int64) and one codebook-shard mirror per rank (k_per x D fp32)."""

    def __init__(self, group, k_per: int, D: int, device):
        import torch.distributed._symmetric_memory as symm
        self.symm = symm
        self.group = group if group is not None else dist.group.WORLD
        self.device = device
        self.mirror = symm.empty(k_per, D, dtype=torch.float32, device=device)
        self.mirror_h = symm.rendezvous(self.mirror, self.group)
        self.keys = None
        self.keys_h = None

    def keys_for(self, N: int):
        if self.keys is None or self.keys.numel() != N:
            self.keys = self.symm.empty(N, dtype=torch.int64, device=self.device)
            self.keys_h = self.symm.rendezvous(self.keys, self.group)
        return self.keys, self.keys_h


class _CodebookShardedFusedFn(torch.autograd.Function):
    """Codebook-sharded forward with NO collective library call: the search kernel MIN-combines packed keys into every
    rank's key buffer over NVLink (system-scope atomics on peer memory), the gather kernel reads winning rows from the
    owner's shard through the peer mapping; three device-side barriers order the phases."""

    @staticmethod
    def forward(ctx, z, E_param, beta, mode, peer: _PeerMemory, rank, world, k_offset, k_total):
        N, D = z.shape
        k_per = E_param.shape[0]
        k_valid = max(0, min(k_per, k_total - k_offset))
        E_local = E_param[:k_valid]
        keys, kh = peer.keys_for(N)
        keys.fill_(torch.iinfo(torch.int64).max)
        peer.mirror.copy_(E_param.detach())
        kh.barrier(channel=0)                       # every rank's key buffer is initialised, every mirror refreshed
        F.search_peers(z, E_local, kh.buffer_ptrs, rank, mode=mode, k_offset=k_offset)
        kh.barrier(channel=1)                       # all remote atomics have landed: keys hold the global argmin
        idx = F.keys_to_idx(keys)
        z_q, sq_sum, hist_all = F.quantize_shards(z, peer.mirror_h.buffer_ptrs, k_per, idx, k_per * world)
        kh.barrier(channel=0)                       # peers are done reading this rank's mirror
        hist = hist_all[:k_total].contiguous()
        loss, perplexity = F.finalize(sq_sum, hist, N, D, beta)
        hist_local = hist_all[k_offset:k_offset + k_valid].contiguous()
        ctx.save_for_backward(z, E_param, idx, hist_local, z_q)
        ctx.beta, ctx.k_offset, ctx.k_valid = beta, k_offset, k_valid
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(perplexity, idx, hist)
        return loss, z_q, perplexity, idx, hist

    @staticmethod
    def backward(ctx, g_loss, g_zq, *_):
        z, E_param, idx, hist_local, z_q = ctx.saved_tensors
        need_dz, need_dE = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if g_zq is not None:
            g_zq = g_zq.contiguous()
        none = (None,) * 7
        if g_loss is None:
            return ((g_zq if need_dz else None), (torch.zeros_like(E_param) if need_dE else None)) + none
        g_loss = g_loss.detach().to(torch.float32).contiguous()
        dz = F.dz_from_zq(z, z_q, g_zq, g_loss, z.shape[0]) if need_dz else None
        dE = None
        if need_dE:
            dE = _local_codebook_grad(z, E_param, ctx.k_valid, idx, hist_local, ctx.beta, g_loss, ctx.k_offset)
        return (dz, dE) + none


class CodebookShardedVectorQuantizer(nn.Module):
    """VectorQuantizer whose codebook rows are sharded over the ranks of `process_group` (latents replicated).
    `embedding` holds only this rank's shard: rows [k_offset, k_offset + k_local) of the global (n_e, e_dim)
    codebook (all shards have ceil(n_e / world) rows; the tail of the last one is padding the kernels never see)."""

    def __init__(self, n_e, e_dim, beta, vq_codebook_init_values: Tensor = None, *, process_group=None,
                 search: str = "auto", exchange: str = "auto"):
        """exchange="nvlink": fused path -- the kernels do the exchange themselves over NVLink peer memory (system-scope
        atomics for the argmin, peer loads for the winning rows; needs torch symmetric memory on one NVLink domain,
        world <= 8).  exchange="nccl": all-reduce(MIN) of keys + all-reduce(SUM) of the N x D z_q partials (works
        everywhere, moves 2 x N x D x 4 bytes more).  exchange="auto" (default): "nvlink" when the symmetric-memory
        rendezvous succeeds on every rank, else "nccl"."""
        super().__init__()
        self.n_e, self.e_dim, self.beta = n_e, e_dim, beta
        self.search, self.group = search, process_group
        if exchange not in ("auto", "nccl", "nvlink"):
            raise ValueError("exchange must be 'auto', 'nccl' or 'nvlink'")
        self.exchange = exchange
        self._peer = None
        world, rank = _world(process_group), _rank(process_group)
        self.world, self.rank = world, rank
        self.k_per = (n_e + world - 1) // world
        self.k_offset = rank * self.k_per
        lo, hi = shard_bounds(n_e, world, rank)
        self.k_valid = hi - lo
        if self.k_valid < 1:
            raise ValueError(f"n_e={n_e} codes cannot be sharded over {world} ranks (rank {rank} would own none)")
        self.embedding = nn.Embedding(self.k_per, e_dim)
        with torch.no_grad():
            if vq_codebook_init_values is not None:
                self.embedding.weight[: self.k_valid].copy_(vq_codebook_init_values[lo:hi])
            else:
                g = torch.Generator().manual_seed(0x5EED + rank)
                self.embedding.weight.copy_((torch.rand(self.k_per, e_dim, generator=g) * 2 - 1) / n_e)
            if self.k_valid < self.k_per:
                self.embedding.weight[self.k_valid:] = 0.0            # padding rows: never searched, zero gradient

    @torch.compiler.disable
    def forward(self, z: Tensor, device=None):
        batch_size, seq_len, _ = z.shape
        zf = z.view((-1, self.e_dim))
        if self.exchange == "auto" and self.world > 1:
            self.exchange = self._resolve_exchange(z.device)
        if self.exchange == "nvlink" and self.world > 1:
            if self._peer is None:
                self._peer = _PeerMemory(self.group, self.k_per, self.e_dim, z.device)
            loss, z_q, perplexity, idx, _ = _CodebookShardedFusedFn.apply(
                zf, self.embedding.weight, float(self.beta), self.search, self._peer, self.rank, self.world,
                self.k_offset, self.n_e)
            return loss, z_q.view(z.shape), perplexity, None, idx.reshape((batch_size, seq_len, 1))
        loss, z_q, perplexity, idx, _ = _CodebookShardedFn.apply(zf, self.embedding.weight, float(self.beta),
                                                                  self.search, self.group, self.k_offset, self.n_e)
        return loss, z_q.view(z.shape), perplexity, None, idx.reshape((batch_size, seq_len, 1))

    def _resolve_exchange(self, device) -> str:
        """"nvlink" iff every rank can set up the peer-mapped buffers (torch symmetric memory: CUDA VMM + fabric / IPC
        handles); decided once, collectively, so that all ranks take the same path."""
        ok = 1
        try:
            if not z_is_cuda(device) or self.world > 8:
                raise RuntimeError("fused exchange needs CUDA and world <= 8")
            self._peer = _PeerMemory(self.group, self.k_per, self.e_dim, device)
        except Exception:
            self._peer = None
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            return "nvlink"
        self._peer = None
        return "nccl"
