"""Data parallelism over the batch.  One process per GPU.

Inference shards the batch and needs no collective.  Training keeps a replica per rank and
all-reduces (mean) the gradients of the parameters the step touched -- the generator's in a
G-step, the discriminator's in a D-step -- over NCCL/NVLink, then clips and steps identically
on every rank.  Gradients live in ONE flat buffer per sub-network
(``train_step.FlatState``; parameters' ``.grad`` are views into it), so the exchange needs no
pack/unpack copies: the generator's buffer is all-reduced in four contiguous buckets launched
from backward hooks on a communication stream while backward still runs, as fp32
(``train_step.TrainStep(overlap_allreduce=True, exchange_dtype="fp32")``, the defaults).  The
host-side coin flips (gan.py:105, jlcss.py:127) come from a generator OWNED by TrainStep that
every rank seeds identically, so nothing else that consumes the process-global generator can
make ranks pick different step kinds.  BatchNorm uses per-rank batch statistics (DDP semantics): parity is
defined per rank against the reference on that rank's shard, gradients against the mean of
the per-shard reference gradients."""
from __future__ import annotations

import torch


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
    """Seeds the process-global CPU generator identically on every rank (data order, anything user code draws).
    The D/G coin and the curriculum draw do NOT depend on it in a data-parallel run: TrainStep draws them from its own
    generator (``rng_seed``)."""
    torch.manual_seed(seed)
