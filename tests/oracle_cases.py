"""Shared definitions of the golden cases (tests/golden/*.npz, made by oracle/make_golden.py)."""
import numpy as np
import torch

import mixstage_oracle as O

CFG1 = O.Spec(num_speakers=2)
CFG2 = O.Spec(num_speakers=4)
CFG5 = O.Spec(num_speakers=25, num_clusters=16, argmax=0, time_steps=256)
STAGE = O.Spec(num_speakers=4, num_clusters=1)      # StAGE variant: one sub-decoder (src/jobs/stage.py)

# name -> (spec, B, T, kind, kwargs)
CASES = {
    "cfg1_eval_sample": (CFG1, 16, 64, "g", dict(training=False, sample_flag=1, description="test")),
    "cfg1_train_fwd": (CFG1, 16, 64, "g", dict(training=True, sample_flag=0, description="train")),
    "cfg2_gstep": (CFG2, 16, 64, "gan", dict(step="G")),
    "cfg2_dstep": (CFG2, 16, 64, "gan", dict(step="D")),
    "cfg2_eval": (CFG2, 16, 64, "gan", dict(step="eval")),
    "cfg2_pose_branch": (CFG2, 16, 64, "gan", dict(step="G", use_pose_encoder=True)),
    "cfg5_stress_small": (CFG5, 2, 256, "gan", dict(step="G")),
    "sample_long": (CFG2, 2, 64, "g_long", dict(training=False, sample_flag=1, description="test")),
    "stage_k1_gstep": (STAGE, 8, 64, "gan", dict(step="G")),
}
G_SEED, D_SEED = 7, 8


def leafify(sd):
    """state dict -> same dict with float non-buffer tensors as autograd leaves
    (style_dec_gr aliases style_dec, as in the reference)."""
    out = {}
    for k, v in sd.items():
        if k.startswith("style_dec_gr.models.0."):
            continue
        if v.is_floating_point() and "running_" not in k and k != "eye":
            out[k] = v.clone().requires_grad_(True)
        else:
            out[k] = v.clone()
    for k in sd:
        if k.startswith("style_dec_gr.models.0."):
            out[k] = out["style_dec." + k[len("style_dec_gr.models.0."):]]
    return out


def run_oracle(name, dtype=torch.float64):
    spec, B, T, kind, kw = CASES[name]
    sd = leafify(O.synth_state(O.g_state_shapes(spec), G_SEED, dtype))
    sdd = leafify(O.synth_state(O.d_state_shapes(spec.out_feats), D_SEED, dtype))
    audio, pose, labels, style = O.synth_inputs(B, T, spec, dtype=dtype)
    res = {"sd": sd, "sdd": sdd}
    if kind == "gan":
        lg, ld = O.BNLog(), O.BNLog()
        fake, losses, aux = O.gan_forward(sd, sdd, spec, audio, labels, pose, style, log_g=lg, log_d=ld, **kw)
        if kw["step"] != "eval":
            sum(losses).backward()
        res.update(log_g=lg, log_d=ld)
    else:
        if kind == "g_long":
            audio, pose, labels = audio.reshape(1, B * T, -1), pose.reshape(1, B * T, -1), labels.reshape(1, B * T)
        lg = O.BNLog()
        with torch.no_grad():
            fake, losses, aux = O.g_forward(sd, spec, audio, labels, pose, style, log=lg, **kw)
        res.update(log_g=lg)
    res.update(pose=fake.detach(), losses=[float(l.detach()) for l in losses], aux=aux)
    return res


def load_golden(golden_dir, name):
    import os
    return dict(np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False))
