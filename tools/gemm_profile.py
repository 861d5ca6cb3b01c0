"""Per-launch timing (CUDA events, warm, eager) of every GEMM-class C-ABI call in one G-step and one D-step,
with the launch geometry.  python tools/gemm_profile.py [precision] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
import mixstage_b200 as M
import mixstage_oracle as O
from mixstage_b200 import _lib, ops, igemm
from model_cases import build

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
spec = O.Spec(num_speakers=4)
M.set_precision(prec)
G, D, gan = build(spec, 64, "cuda", torch.float64)
G.thresh.value, G.thresh.iters = 1.0, 1000
ts = M.TrainStep(gan, use_graphs=False)
audio, pose, labels, style = [t.cuda() for t in O.synth_inputs(B, 64, spec)]
batch = (audio, labels, pose, style)
for k in "GDGD":
    ts.step(*batch, kind=k)
torch.cuda.synchronize()
records = []
orig = ops.call


def info(name, a):
    for x in a:
        if isinstance(x, _lib.IgemmDesc):
            d = x
            tm = igemm._tiles_m(d)
            npass = 3 if d.planes == 2 else 1
            if name == "ms_wgrad_bf16":
                kpad = d.cchunks * 64
                tiles = ((d.class_n + 127) // 128) * ((kpad + 255) // 256) * d.num_classes * d.ntaps
                return "wgrad ct=%d tiles=%d split=%d out=%s cls=%d n=%d taps=%d cch=%d" % (d.wgrad_c_tile, tiles, max(1, d.split_k), list(d.out_dims), d.num_classes, d.class_n, d.ntaps, d.cchunks)
            nt = (d.class_n + d.block_n - 1) // d.block_n
            return "igemm grid=(%d,%d,%d) bn=%d numk=%d out=%s cls=%d n=%d taps=%d cch=%d" % (
                tm, nt * d.num_classes, max(1, d.split_k), d.block_n, d.ntaps * d.cchunks * npass, list(d.out_dims), d.num_classes, d.class_n, d.ntaps, d.cchunks)
        if isinstance(x, _lib.ConvDesc):
            return "simt B%d %dx%d Cin%d Cout%d k%dx%d s%d g%d" % (x.B, x.H, x.W, x.Cin, x.Cout, x.kh, x.kw, x.sw, x.groups)
    return ""


def timed(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig(name, *a)
    e1.record()
    records.append((name, info(name, a), e0, e1))


ops.call = timed
from mixstage_b200 import train_step
train_step.call = timed
for kind in "GD":
    records.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    ts.step(*batch, kind=kind)
    t1.record()
    torch.cuda.synchronize()
    tot = {}
    print("==== %s-step: %d C-ABI launches, eager wall %.2f ms" % (kind, len(records), t0.elapsed_time(t1)))
    for name, inf, e0, e1 in records:
        ms = e0.elapsed_time(e1) * 1e3
        tot.setdefault(name, [0, 0.0])
        tot[name][0] += 1
        tot[name][1] += ms
        if inf:
            print("%8.1f us  %-22s %s" % (ms, name, inf))
    print("---- totals")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-28s %4d  %9.1f us" % (k, v[0], v[1]))
    print("sum %.1f us" % sum(v[1] for v in tot.values()))
