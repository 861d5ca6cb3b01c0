"""Data-parallel train step on 2 ranks (gloo, CPU; kernels routed to the CPU specification): every rank runs
TrainStep on its own shard, the flat gradient buffer is all-reduced (mean) and the fused clip+Adam leaves identical
parameters on both ranks, equal to a single process that averages the two shards' gradients."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out, overlap=None, precision="fp32"):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_emu
    import mixstage_b200 as M
    import mixstage_oracle as O
    from mixstage_b200 import parallel, train_step
    from model_cases import build
    mpatch = pytest.MonkeyPatch()
    cpu_emu.install(mpatch)
    mpatch.setattr(train_step, "call", M._lib.call)
    mpatch.setattr(train_step, "stream", lambda: None)
    M.set_precision(precision)
    spec = O.Spec(num_speakers=4)
    B, T = 4, 64
    G, D, gan = build(spec, T, "cpu", torch.float64)
    G.thresh.value, G.thresh.iters = 1.0, 1000
    parallel.sync_host_rng(11212)
    if rank == 1:
        torch.rand(3)              # ranks may consume the global generator differently: TrainStep draws from its own
    ts = M.TrainStep(gan, use_graphs=False, overlap_allreduce=bool(overlap))
    full = O.synth_inputs(world * B, T, spec)
    audio, pose, labels, style = parallel.shard_batch(full, rank, world)
    kinds = []
    # the tensor-core mode learns its exchange buckets in the first generator step and sends them from hooks from the second
    nsteps = 3 if (overlap is not None and precision != "fp32") else 2
    for i in range(nsteps):
        # coin flips from the shared host RNG; the overlap comparison forces the step kinds
        ts.step(audio, labels, pose, style, kind=("G", "D", "G")[i] if overlap is not None else None)
        kinds.append(ts.last_kind)
    res = {"kinds": kinds, "pG": ts.fG.p.clone(), "pD": ts.fD.p.clone(), "gG": ts.fG.g.clone(), "gD": ts.fD.g.clone(),
           "steps": (int(ts.fG.step_count), int(ts.fD.step_count))}
    torch.save(res, os.path.join(out, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_two_rank_train_step(tmp_path):
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    assert r0["kinds"] == r1["kinds"]                       # both ranks took the same D/G branches
    assert r0["steps"] == r1["steps"] and sum(r0["steps"]) == 2
    for k in ("pG", "pD", "gG", "gD"):
        assert torch.equal(r0[k], r1[k]), k                 # replicas stay bit-identical
    assert float(r0["gG"].abs().max()) > 0 or float(r0["gD"].abs().max()) > 0


@pytest.mark.timeout(1500)
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_overlapped_allreduce_equals_single_allreduce(tmp_path, precision):
    """TrainStep(overlap_allreduce=True): the generator's gradients are exchanged segment by segment from backward hooks
    (decoder / logits / classifier first, then style embedding, UNet, and the rest at the end).  Same element-wise
    operations as the single all-reduce, so parameters and gradients must be bit-identical.  In the tensor-core mode the conv
    weight gradients live in fp32 accumulators until a conversion launch moves them into the flat buffer (chains, one
    weight-gradient launch per chain): every bucket must convert ITS accumulators before it is reduced."""
    world = 2
    res = {}
    for overlap in (False, True):
        d = tmp_path / ("overlap%d" % overlap)
        d.mkdir()
        mp.spawn(_worker, args=(world, 31500 + os.getpid() % 2000 + int(overlap) + (7 if precision != "fp32" else 0), str(d), overlap,
                                precision), nprocs=world, join=True)
        res[overlap] = [torch.load(os.path.join(d, "rank%d.pt" % r)) for r in range(world)]
    for r in range(world):
        assert res[False][r]["kinds"] == res[True][r]["kinds"] == (["G", "D"] if precision == "fp32" else ["G", "D", "G"])
        for k in ("pG", "pD", "gG", "gD"):
            assert torch.equal(res[False][r][k], res[True][r][k]), (r, k)
    assert torch.equal(res[True][0]["pG"], res[True][1]["pG"])


def test_shard_batch():
    from mixstage_b200 import parallel
    t = torch.arange(24).view(8, 3)
    a, = parallel.shard_batch([t], 1, 4)
    assert a.tolist() == t[2:4].tolist()
    with pytest.raises(ValueError):
        parallel.shard_batch([t], 0, 3)
