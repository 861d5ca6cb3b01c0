"""Live pin of the oracle (and of the drop-in install) against the UNMODIFIED reference modules, whenever
/root/reference is present (the build container; skipped on the GPU box, which has no reference tree)."""
import os
import sys

import pytest
import torch

import mixstage_oracle as O
import ref_loader
from oracle_cases import CASES, run_oracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present")


@pytest.mark.parametrize("name,step", [("cfg2_gstep", "G"), ("cfg2_dstep", "D")])
def test_oracle_equals_executed_reference(name, step):
    import make_golden
    ns = ref_loader.load()
    spec, B, T, kind, kw = CASES[name]
    got = make_golden.run_gan(ns, spec, B, T, step)
    res = run_oracle(name)
    assert float((res["pose"] - torch.from_numpy(got["pose"]).double()).abs().max()) < 2e-6
    for a, b in zip(res["losses"], got["losses"]):
        assert abs(a - b) < 1e-10 * max(1.0, abs(b))
    for n, v in zip(got["g_grad_names"], got["g_grad_norms"]):
        g = res["sd"][str(n)].grad
        assert abs((0.0 if g is None else float(g.norm())) - v) <= 1e-9 * max(1.0, v) + 1e-12, n


def test_install_patches_reference_names():
    """mixstage_b200.install(namespace) puts the B200 classes under the names the reference's trainer evals
    (trainer.py:1049,1076) without touching reference files."""
    import mixstage_b200 as M
    ns = {}
    M.install(ns)
    assert ns["JointLateClusterSoftStyle4_G"] is M.JointLateClusterSoftStyle4_G
    assert ns["JointLateClusterSoftStyle4_D"] is M.Speech2Gesture_D
    ref = ref_loader.load()
    spec = O.Spec(num_speakers=4)
    kw = dict(time_steps=64, out_feats=96, num_clusters=8, style_dict={i: i for i in range(4)}, style_dim=10,
              shape={"audio/log_mel_400": [64, 64]})
    ours, theirs = eval("JointLateClusterSoftStyle4_G", ns)(**kw), ref.G(**kw)
    a, b = ours.state_dict(), theirs.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert tuple(a[k].shape) == tuple(b[k].shape), k
    theirs.load_state_dict(a)          # checkpoints are interchangeable
