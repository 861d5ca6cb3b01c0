"""Data parallelism over the batch.  One process per GPU.

Inference shards the batch and needs no collective.  Training keeps a replica per rank and
all-reduces (mean) the gradients of the parameters the step touched -- the generator's in a
G-step, the discriminator's in a D-step -- over NCCL/NVLink, then clips and steps identically
on every rank.  Gradients live in ONE flat buffer per sub-network (parameters' ``.grad`` are
views into it), so the exchange is a single bucketed all-reduce with no pack/unpack copies,
and the host-side coin flips (gan.py:105, jlcss.py:127) are drawn from a generator that every
rank seeds identically.  BatchNorm uses per-rank batch statistics (DDP semantics): parity is
defined per rank against the reference on that rank's shard, gradients against the mean of
the per-shard reference gradients."""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGrads:
    """Re-homes the ``.grad`` of every parameter of ``module`` into one contiguous buffer."""

    def __init__(self, module, bucket_bytes=32 << 20):
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        p0 = self.params[0]
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=p0.dtype, device=p0.device)
        off = 0
        self.offsets = []
        for p in self.params:
            if p.dtype != p0.dtype or p.device != p0.device:
                raise ValueError("all parameters must share dtype/device")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            self.offsets.append(off)
            off += p.numel()
        # bucket boundaries (in elements), filled back-to-front: later parameters finish backward first
        per = max(1, bucket_bytes // self.flat.element_size())
        self.buckets = []
        end = self.numel
        while end > 0:
            beg = max(0, end - per)
            self.buckets.append((beg, end))
            end = beg

    def zero(self):
        self.flat.zero_()
        for p, off in zip(self.params, self.offsets):      # re-attach if someone set grads to None
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + off * self.flat.element_size():
                p.grad = self.flat[off:off + p.numel()].view_as(p)

    def allreduce_mean(self, group=None, async_op=False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return []
        ws = dist.get_world_size(group)
        works = []
        for beg, end in self.buckets:
            chunk = self.flat[beg:end]
            chunk.div_(ws)
            works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=group, async_op=async_op))
        return works


def shard_batch(tensors, rank, world):
    """Contiguous batch shard for inference / per-rank training data (no collective)."""
    out = []
    for t in tensors:
        B = t.shape[0]
        if B % world:
            raise ValueError("batch %d is not divisible by world size %d" % (B, world))
        per = B // world
        out.append(t[rank * per:(rank + 1) * per])
    return out


def sync_host_rng(seed: int):
    """All ranks must take the same D/G and curriculum branches: seed the CPU generator
    that `torch.rand(1)` draws from identically everywhere."""
    torch.manual_seed(seed)
