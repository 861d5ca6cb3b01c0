"""Phase timing inside the fused training blocks (csrc/conv_train.cu) for single layers of the generator's geometries:
CUDA-event time of the launch (behind a spin, so host latency is hidden) and CTA 0's %globaltimer stamps at the phase
boundaries (MS_PHASE_TS=1).  Usage: MS_PHASE_TS=1 python tools/block_phases.py [--precision bf16x3] [names...]"""
import argparse
import ctypes
import os
import sys

os.environ.setdefault("MS_PHASE_TS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

FWD = ["start", "setup", "gemm", "bar1", "stats", "bar2", "norm"]
BWD = ["start", "setup", "reduce", "bar1", "apply", "bar2", "gemm"]


def stamps():
    from mixstage_b200 import _lib
    buf = (ctypes.c_ulonglong * 16)()
    _lib.call("ms_debug_phase_ts", ctypes.cast(buf, ctypes.c_void_p))
    return list(buf)


def main():
    from test_fused_blocks_gpu import GEOMS, _make
    from mixstage_b200 import ops
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("names", nargs="*")
    a = ap.parse_args()
    ops.set_precision(a.precision)
    names = a.names or list(GEOMS)
    for name in names:
        g = GEOMS[name]
        m = _make(g)
        m.train()
        Gr = g.get("groups", 1)
        torch.manual_seed(11)
        shape = (g["B"], g["H"], g["W"], g["cin"] * Gr) if g.get("two_d") else (g["B"], 1, g["L"], g["cin"] * Gr)
        x = torch.randn(*shape, device="cuda", requires_grad=True)
        res = torch.randn(g["B"], 1, 2 * g["L"], g["cout"] * Gr, device="cuda", requires_grad=True) if g.get("up2") else None
        rec = {}
        orig = ops.call

        def timed(n, *args):
            if n in ("ms_conv_block_train_fwd", "ms_conv_block_train_bwd", "ms_wgrad_bf16_acc", "ms_wgrad_bf16"):
                torch.cuda.synchronize()
                torch.cuda._sleep(int(2e6))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                orig(n, *args)
                e1.record()
                torch.cuda.synchronize()
                rec[n] = (e0.elapsed_time(e1) * 1e3, stamps() if "block" in n else None)
            else:
                orig(n, *args)

        for it in range(3):
            ops.call = timed if it == 2 else orig
            try:
                y = m(x, residual=res, up2=True) if res is not None else m(x)
                y.backward(torch.randn_like(y))
            finally:
                ops.call = orig
            torch.cuda.synchronize()
        print("== %s  %s" % (name, g))
        for n, (us, ts) in rec.items():
            line = "  %-26s event %7.1f us" % (n, us)
            if ts:
                labels = FWD if n.endswith("fwd") else BWD
                t0 = ts[0]
                parts = []
                for i in range(1, 7):
                    if ts[i] >= ts[i - 1] and ts[i] - t0 < 10_000_000:
                        parts.append("%s %.1f" % (labels[i], (ts[i] - ts[i - 1]) / 1e3))
                line += "   in-kernel %.1f us: %s" % ((max(ts[:7]) - t0) / 1e3, ", ".join(parts))
            print(line)


if __name__ == "__main__":
    main()
