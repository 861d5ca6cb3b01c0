"""Diagnostic: GPU vs CPU-spec (tests/cpu_emu.py) run of the same module graph, G-step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import mixstage_oracle as O
from model_cases import build, MOD
from oracle_cases import CFG2
import cpu_emu
from mixstage_b200 import _lib, ops

def run(device):
    G, D, gan = build(CFG2, 64, device, torch.float64)
    audio, pose, labels, style = (t.to(device) for t in O.synth_inputs(16, 64, CFG2))
    gan.train(); gan.force_step = "G"; G.thresh.value, G.thresh.iters = 1.0, 1000
    grads = {}
    fake, losses, _ = gan([audio, labels], pose, input_modalities=MOD, style=style, sample_flag=0, description="train", desc="train")
    fake.register_hook(lambda g: grads.__setitem__("dfake", g.detach().cpu().clone()))
    terms = {}
    for i, l in enumerate(losses):
        for p in list(G.parameters()) + list(D.parameters()):
            p.grad = None
        l.backward(retain_graph=True)
        terms[i] = grads.get("dfake")
        grads.pop("dfake", None)
        terms["logits_w_%d" % i] = None if G.logits.weight.grad is None else G.logits.weight.grad.detach().cpu().clone()
    return terms, [float(l) for l in losses]

gpu, lg = run("cuda")
class MP:
    def __init__(s): s.saved = []
    def setattr(s, obj, name, val, raising=True):
        s.saved.append((obj, name, getattr(obj, name, None))); setattr(obj, name, val)
mp = MP(); cpu_emu.install(mp)
cpu, lc = run("cpu")
print("losses gpu", lg); print("losses cpu", lc)
for k in gpu:
    a, b = gpu[k], cpu[k]
    if a is None or b is None:
        print(k, "None", a is None, b is None); continue
    print("%-14s rel %.3e  |ref| %.3e" % (str(k), float((a - b).norm() / (b.norm() + 1e-30)), float(b.norm())))
