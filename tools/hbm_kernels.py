"""Achieved HBM bandwidth of the memory-bound kernels of the path at the sizes of a B = 1024 batch (SURVEY.md section 8d:
"quote GB/s only for B >= 1024"): style gather + concat (+ operand planes), L1 loss forward / backward, velocity, gradient
norm, clip + Adam.  CUDA events around ONE launch each, median of 20, a 512 MB buffer rewritten between launches (L2 flush);
algorithmic bytes / time against MEASURED_PEAKS.json's hbm_gbs (fallback 6650).
Usage: python tools/hbm_kernels.py > profiles/r02_hbm_kernels_b1024.txt"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mixstage_b200 import _lib  # noqa: E402
from mixstage_b200._lib import call, ptr  # noqa: E402


def main():
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        src = "MEASURED_PEAKS.json"
    except Exception:
        peak, src = 6650.0, "fallback"
    dev = torch.device("cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    st = lambda: torch.cuda.current_stream().cuda_stream      # noqa: E731

    def timed(fn, reps=20):
        ts = []
        for _ in range(3):
            fn()
        for _ in range(reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return statistics.median(ts)

    B, T, C, sd, S, P = 1024, 64, 256, 10, 4, 96
    rows = B * T
    print("# HBM-bound kernels at B = %d (T = %d): algorithmic bytes / median launch time, peak %.0f GB/s (%s)" % (B, T, peak, src))
    print("# %-44s %10s %12s %10s %8s" % ("kernel", "us", "MB (algo)", "GB/s", "frac"))

    def report(name, us, nbytes):
        gbs = nbytes / (us * 1e-6) / 1e9
        print("%-46s %10.1f %12.1f %10.0f %8.2f" % (name, us, nbytes / 1e6, gbs, gbs / peak))

    x = torch.randn(rows, C, device=dev)
    emb = torch.randn(S, sd, dtype=torch.float64, device=dev)
    idx = torch.randint(0, S, (B,), device=dev)
    out = torch.empty(rows, C + sd, device=dev)
    rs = 272
    ps = rows * rs
    planes = torch.empty(2 * ps, dtype=torch.bfloat16, device=dev)
    us = timed(lambda: call("ms_style_concat_fwd_f32", ptr(x), rows, C, ptr(idx), None, T, ptr(emb), 1, S, sd, ptr(out), st()))
    report("style_concat_fwd (round 1, scalar)", us, rows * (C * 4 + (C + sd) * 4))
    us = timed(lambda: call("ms_style_concat_planes_fwd_f32", ptr(x), rows, C, ptr(idx), None, T, ptr(emb), 1, S, sd, ptr(out),
                            None, 0, 0, rs, st()))
    report("style_concat rows kernel, fp32 only", us, rows * (C * 4 + (C + sd) * 4))
    us = timed(lambda: call("ms_style_concat_planes_fwd_f32", ptr(x), rows, C, ptr(idx), None, T, ptr(emb), 1, S, sd, ptr(out),
                            ptr(planes), 3, ps, rs, st()))
    report("style_concat rows kernel + hi/lo planes", us, rows * (C * 4 + (C + sd) * 4 + 2 * rs * 2))
    us = timed(lambda: call("ms_to_planes", ptr(out), rows, C + sd, rs, ptr(planes), 3, ps, st()))
    report("to_planes (the pass the fused form removes)", us, rows * ((C + sd) * 4 + 2 * rs * 2))
    dout = torch.randn(rows, C + sd, device=dev)
    dx = torch.empty(rows, C, device=dev)
    demb = torch.zeros(S, sd, device=dev)
    us = timed(lambda: call("ms_style_concat_bwd_f32", ptr(dout), rows, C, ptr(idx), None, T, ptr(emb), 1, S, sd, ptr(dx), ptr(demb),
                            None, st()))
    report("style_concat_bwd (slice + scatter-add)", us, rows * ((C + sd) * 4 + C * 4))

    n = B * T * P
    a, b = torch.randn(n, device=dev), torch.randn(n, device=dev)
    acc = torch.zeros(1, dtype=torch.float64, device=dev)
    sgn = torch.empty(n, device=dev)
    da = torch.empty(n, device=dev)
    g = torch.ones(1, device=dev)
    us = timed(lambda: call("ms_l1_fwd_f32", ptr(a), ptr(b), 0.0, n, ptr(acc), ptr(sgn), st()))
    report("l1_fwd with sign tensor (round 1)", us, n * 12)
    us = timed(lambda: call("ms_l1_fwd_f32", ptr(a), ptr(b), 0.0, n, ptr(acc), None, st()))
    report("l1_fwd, no sign tensor, 16-byte loads", us, n * 8)
    us = timed(lambda: call("ms_l1_bwd_f32", ptr(sgn), ptr(g), n, ptr(da), st()))
    report("l1_bwd from sign tensor (round 1)", us, n * 8)
    us = timed(lambda: call("ms_l1_bwd_ab_f32", ptr(a), ptr(b), 0.0, ptr(g), n, ptr(da), st()))
    report("l1_bwd from the operands", us, n * 12)
    pose = torch.randn(B, T, P, device=dev)
    vel = torch.empty_like(pose)
    us = timed(lambda: call("ms_velocity_fwd_f32", ptr(pose), B, T, P, ptr(vel), st()))
    report("velocity_fwd", us, n * 8)

    npar = 14_900_000
    p, gr, m, v = (torch.randn(npar, dtype=torch.float64, device=dev) for _ in range(4))
    v.abs_()
    sq = torch.zeros(1, dtype=torch.float64, device=dev)
    stepc = torch.ones(1, dtype=torch.int64, device=dev)
    lr = torch.full((1,), 1e-4, dtype=torch.float64, device=dev)
    us = timed(lambda: call("ms_grad_sqnorm", ptr(gr), 1, npar, ptr(sq), ptr(stepc), st()))
    report("grad_sqnorm (fp64, fixed-order)", us, npar * 8)
    us = timed(lambda: call("ms_clip_adam", ptr(p), ptr(gr), ptr(m), ptr(v), 1, npar, ptr(sq), ptr(stepc), 1e-4, 0.9, 0.999, 1e-8, 1.0,
                            ptr(lr), st()))
    report("clip_adam (fp64 p, g, m, v)", us, npar * 56)
    us = timed(lambda: call("ms_clip_adam_mixed", ptr(p), ptr(gr), ptr(m), ptr(v), 1, 1, npar, ptr(sq), ptr(stepc), 1e-4, 0.9, 0.999,
                            1e-8, 1.0, ptr(lr), st()))
    report("clip_adam_mixed (fp64 p, g, m, v; 4 per thread)", us, npar * 56)
    m32, v32 = m.float(), v.float()
    us = timed(lambda: call("ms_clip_adam_mixed", ptr(p), ptr(gr), ptr(m32), ptr(v32), 1, 0, npar, ptr(sq), ptr(stepc), 1e-4, 0.9,
                            0.999, 1e-8, 1.0, ptr(lr), st()))
    report("clip_adam_mixed (fp64 p, g; fp32 m, v)", us, npar * 40)


if __name__ == "__main__":
    main()
