"""Diagnostic: discriminator fwd/bwd, GPU kernels vs CPU spec, layer by layer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import mixstage_oracle as O
import mixstage_b200 as M
from oracle_cases import D_SEED
import cpu_emu
from mixstage_b200 import ops

torch.manual_seed(0)
B = 16
x0 = torch.randn(B, 64, 96)

def run(device):
    D = M.Speech2Gesture_D(in_channels=96)
    D.load_state_dict(O.synth_state(O.d_state_shapes(96), D_SEED, torch.float32))
    D = D.to(device).train()
    x = x0.to(device).requires_grad_(True)
    acts, grads = {}, {}
    h = ops.velocity(x)
    h = h.view(B, 1, 64, 96)
    names = ["conv1", "conv2", "conv3", "logits"]
    fns = [lambda t: D._conv1(D.conv1[0], t), lambda t: D.conv2[0](t), lambda t: D.conv3(t), lambda t: D._logits(D.logits, t)]
    for n, f in zip(names, fns):
        h = f(h)
        acts[n] = h.detach().cpu().clone()
        h.register_hook(lambda g, n=n: grads.__setitem__(n, g.detach().cpu().clone()))
    loss = ops.l1_mean(h.reshape(-1).contiguous(), None, 1.0)
    loss.backward()
    grads["x"] = x.grad.detach().cpu().clone()
    pg = {n: p.grad.detach().cpu().clone() for n, p in D.named_parameters()}
    return acts, grads, pg, float(loss)

ga, gg, gp, gl = run("cuda")
class MP:
    def setattr(s, obj, name, val, raising=True): setattr(obj, name, val)
cpu_emu.install(MP())
ca, cg, cp, cl = run("cpu")
print("loss", gl, cl)
rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
for n in ga: print("act  %-8s rel %.3e  min|z| %.3e" % (n, rel(ga[n], ca[n]), float(ca[n].abs().min())))
for n in gg: print("grad %-8s rel %.3e" % (n, rel(gg[n], cg[n])))
for n in gp: print("pgrad %-28s rel %.3e" % (n, rel(gp[n], cp[n])))
s_g, s_c = ga["logits"].reshape(-1), ca["logits"].reshape(-1)
print("sign flips in |score-1|:", int(((s_g - 1).sign() != (s_c - 1).sign()).sum()), "min |score-1|", float((s_c - 1).abs().min()))
