"""Data-parallel consistency check on real GPUs (torchrun): after a few TrainStep steps every rank must hold bit-identical
parameters; reports the generator segments that differ.  Usage: torchrun --nproc-per-node 2 tools/dp_check.py [--no-graphs] [--no-overlap]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--steps", type=int, default=6)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_MAX_CTAS", "16")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import mixstage_b200 as M
    import mixstage_oracle as O
    from mixstage_b200 import parallel
    from model_cases import build
    M.set_precision("bf16x3")
    spec = O.Spec(num_speakers=4)
    G, D, gan = build(spec, 64, torch.device("cuda", local), torch.float64)
    gan.train()
    G.thresh.value, G.thresh.iters = 1.0, 1000
    parallel.sync_host_rng(11212)
    ts = M.TrainStep(gan, use_graphs=not a.no_graphs, overlap_allreduce=not a.no_overlap)
    batch = [t.cuda() for t in O.synth_inputs(16, 64, spec, seed=11212 + rank)]
    audio, pose, labels, style = batch
    for i in range(a.steps):
        ts.step(audio, labels, pose, style, kind="G" if i % 2 == 0 else "D")
        torch.cuda.synchronize()
        for name, f in (("G", ts.fG), ("D", ts.fD)):
            for what, buf in (("p", f.p), ("g", f.g)):
                mine = buf.clone()
                other = buf.clone()
                dist.broadcast(other, src=0)
                if rank == 1 and not torch.equal(mine, other):
                    bad = []
                    for top, (b0, e0) in f.segments.items():
                        if not torch.equal(mine[b0:e0], other[b0:e0]):
                            d = (mine[b0:e0] - other[b0:e0]).abs().max().item()
                            bad.append("%s(%.2e)" % (top, d))
                    print("step %d: %s.%s differs between ranks in: %s" % (i, name, what, ", ".join(bad)), flush=True)
    if rank == 0:
        print("dp_check done (graphs=%s overlap=%s)" % (not a.no_graphs, not a.no_overlap), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
