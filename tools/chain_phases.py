"""Per-block phase timing inside the chain launches (csrc/conv_train.cu) of the generator's conv stacks at a given batch:
CTA 0's %globaltimer stamps, 8 per block.  Usage: MS_PHASE_TS=1 python tools/chain_phases.py [--batch 16] [--precision bf16x3]"""
import argparse
import ctypes
import os
import sys

os.environ.setdefault("MS_PHASE_TS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

FWD = ["top", "gemm@", "gemm", "bar1", "stats", "bar2", "norm"]
BWD = ["top", "red@", "reduce", "bar1", "apply", "bar2", "gemm"]


def stamps():
    from mixstage_b200 import _lib
    buf = (ctypes.c_ulonglong * (8 * _lib.CHAIN_MAX))()
    _lib.call("ms_debug_phase_ts", ctypes.cast(buf, ctypes.c_void_p))
    return list(buf)


def show(tag, n, ts, us, labels):
    t0 = ts[0]
    print("  %s: %d blocks, event %.1f us, in-kernel %.1f us" % (tag, n, us, (max(ts[:8 * n]) - t0) / 1e3))
    for li in range(n):
        s = ts[8 * li:8 * li + 8]
        parts = []
        prev = s[0]
        for k in range(1, 7):
            if s[k] >= prev and s[k] - t0 < 50_000_000 and s[k] >= s[0]:
                parts.append("%s %.1f" % (labels[k], (s[k] - prev) / 1e3))
                prev = s[k]
        print("    block %2d  +%.1f us  total %.1f: %s" % (li, (s[0] - t0) / 1e3, (prev - s[0]) / 1e3, ", ".join(parts)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--eval", action="store_true", help="inference chains (eval mode, no_grad)")
    a = ap.parse_args()
    from mixstage_b200 import layers, ops
    import torch.nn as nn
    ops.set_precision(a.precision)
    B = a.batch
    torch.manual_seed(3)
    mods = {
        "audio_encoder.conv.1-7": (nn.Sequential(*list(layers.AudioEncoder().conv)[1:]), (B, 64, 64, 64)),
        "unet": (layers.UNet1D(256, 256), (B, 1, 64, 256)),
        "classify_cluster.conv": (nn.Sequential(*list(layers.ClusterClassify(input_channels=266).conv)), (B, 1, 64, 266)),
        "decoder": (nn.Sequential(layers.ConvNormRelu(266, 2048, leaky=True), *[layers.ConvNormRelu(256, 256, leaky=True, groups=8) for _ in range(3)]), (B, 1, 64, 266)),
        "pose_style.conv.0-5": (nn.Sequential(*list(layers.PoseStyleEncoder().conv)[:6]), (B, 1, 64, 96)),
    }
    for name, (m, shape) in mods.items():
        m = m.to("cuda", torch.float64).train(not a.eval)
        x = torch.randn(*shape, device="cuda", requires_grad=not a.eval)
        rec = {}
        orig = ops.call

        def timed(n, *args):
            if n in ("ms_conv_chain_fwd", "ms_conv_chain_bwd", "ms_wgrad_bf16_acc_multi"):
                torch.cuda.synchronize()
                torch.cuda._sleep(int(2e6))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                orig(n, *args)
                e1.record()
                torch.cuda.synchronize()
                rec[n] = (e0.elapsed_time(e1) * 1e3, stamps() if "chain" in n else None, args[1])
            else:
                orig(n, *args)

        for it in range(3):
            ops.call = timed if it == 2 else orig
            try:
                with torch.set_grad_enabled(not a.eval):
                    if isinstance(m, layers.UNet1D):
                        y = m(x)
                    else:
                        y = layers._run(list(m), x)
                if not a.eval:
                    y.backward(torch.randn_like(y))
            finally:
                ops.call = orig
            torch.cuda.synchronize()
        print("== %s  input %s" % (name, shape))
        for n, (us, ts, cnt) in rec.items():
            if ts:
                show(n, cnt, ts, us, FWD if n.endswith("fwd") else BWD)
            else:
                print("  %s: %d blocks, event %.1f us" % (n, cnt, us))


if __name__ == "__main__":
    main()
