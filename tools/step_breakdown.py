"""Where a train step's time goes: (1) CUDA-graph replay time of the G-step and the D-step separately, (2) every C-ABI
launch of one eager G-step / D-step timed with CUDA events behind a spin kernel (host latency hidden), in launch order
and aggregated per entry point.  torch-side kernels (memsets, copies, stack, casts done by torch) are the remainder.

  python tools/step_breakdown.py [--batch 16] [--speakers 4] [--precision bf16x3] [--order]"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--speakers", type=int, default=4)
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--order", action="store_true", help="print every launch in order")
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    import mixstage_b200 as M
    import mixstage_oracle as O
    from mixstage_b200 import ops
    from model_cases import build
    spec = O.Spec(num_speakers=a.speakers)
    M.set_precision(a.precision)
    G, D, gan = build(spec, 64, "cuda", torch.float64)
    gan.train()
    G.thresh.value, G.thresh.iters = 1.0, 1000
    ts = M.TrainStep(gan)
    batch = [t.cuda() for t in O.synth_inputs(a.batch, 64, spec)]
    audio, pose, labels, style = batch
    args = (audio, labels, pose, style)
    for i in range(6):
        ts.step(*args, kind="G" if i % 2 == 0 else "D")
    torch.cuda.synchronize()
    for kind in ("G", "D"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            ts.step(*args, kind=kind)
        e1.record()
        torch.cuda.synchronize()
        print("graph replay %s-step: %.3f ms/step (%d launches/graph)" % (
            kind, e0.elapsed_time(e1) / a.reps, ts.kernels_per_graph.get((kind, False), -1)))
    # eager, instrumented
    ts.use_graphs = False
    recs = []
    orig = ops.call

    def timed(name, *x):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *x)
        e1.record()
        recs.append((name, e0, e1, torch.cuda.current_stream().cuda_stream))

    import mixstage_b200.train_step as TS
    for kind in ("G", "D"):
        ts.step(*args, kind=kind)      # eager warm-up
        torch.cuda.synchronize()
        recs.clear()
        ops.call = TS.call = timed
        try:
            torch.cuda._sleep(int(300 * 1.9e6))
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            ts.step(*args, kind=kind)
            s1.record()
        finally:
            ops.call = TS.call = orig
        torch.cuda.synchronize()
        tot = s0.elapsed_time(s1) * 1e3
        agg = collections.defaultdict(lambda: [0, 0.0])
        main_stream = recs[0][3]
        t_main = 0.0
        for n, e0, e1, st in recs:
            us = e0.elapsed_time(e1) * 1e3
            agg[n][0] += 1
            agg[n][1] += us
            if st == main_stream:
                t_main += us
            if a.order:
                print("  %-34s %8.1f us %s" % (n, us, "" if st == main_stream else "(side)"))
        ksum = sum(v[1] for v in agg.values())
        print("eager %s-step behind a spin: %.1f us wall on the stream; %d C-ABI launches, sum %.1f us (main stream %.1f us)" % (
            kind, tot, len(recs), ksum, t_main))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print("  %-34s %4d  %9.1f us  %5.1f%%  avg %6.1f" % (k, v[0], v[1], 100 * v[1] / ksum, v[1] / v[0]))


if __name__ == "__main__":
    main()
