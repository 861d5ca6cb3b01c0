"""Diagnostic: per-parameter gradient error of the CUDA path vs the fp64 oracle (and the
oracle's own fp32-vs-fp64 spread for calibration).  python tools/grad_profile.py [case]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from model_cases import run_case  # noqa: E402
from oracle_cases import run_oracle  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_gstep"
got = run_case(name, "cuda", torch.float64)
ref = run_oracle(name)
r32 = run_oracle(name, torch.float32)
print("pose rel: cuda %.2e oracle32 %.2e" % (
    float((got["pose"].cpu() - ref["pose"]).norm() / ref["pose"].norm()),
    float((r32["pose"].double() - ref["pose"]).norm() / ref["pose"].norm())))
for n, p in got["G"].named_parameters():
    r = ref["sd"][n].grad
    if r is None or p.grad is None or float(r.abs().max()) == 0:
        continue
    if n.endswith("conv.weight") or n.endswith("emb.weight") or n.endswith("logits.weight") or n.endswith("norm.weight"):
        e = float((p.grad.cpu().double() - r).norm() / r.norm())
        e32 = float((r32["sd"][n].grad.double() - r).norm() / r.norm())
        print("%-48s cuda %.2e   oracle-fp32 %.2e" % (n, e, e32))
