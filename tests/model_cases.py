"""Builds mixstage_b200 models for the golden cases and runs them (shared by the CPU
host-logic tests, which route kernels to tests/cpu_emu.py, and the GPU parity tests)."""
import torch

import mixstage_b200 as M
import mixstage_oracle as O
from oracle_cases import CASES, D_SEED, G_SEED

MOD = ["audio/log_mel_400"]


def build(spec, T, device, dtype=torch.float64):
    G = M.JointLateClusterSoftStyle4_G(
        time_steps=T, out_feats=spec.out_feats, num_clusters=spec.num_clusters,
        style_dict={i: i for i in range(spec.num_speakers)}, style_dim=spec.style_dim, lambda_id=spec.lambda_id,
        train_only=spec.train_only, softmax=spec.softmax, argmax=spec.argmax, some_grad_flag=spec.some_grad_flag,
        shape={MOD[0]: [T, spec.mel_bins]})
    D = M.JointLateClusterSoftStyle4_D(in_channels=spec.out_feats)
    G.load_state_dict(O.synth_state(O.g_state_shapes(spec), G_SEED, torch.float32))
    D.load_state_dict(O.synth_state(O.d_state_shapes(spec.out_feats), D_SEED, torch.float32))
    gan = M.GAN(G, D, criterion="L1Loss", no_grad=0, input_modalities=MOD)
    gan = gan.to(device=device, dtype=dtype)
    return G, D, gan


def run_case(name, device, dtype=torch.float64, precision="fp32"):
    """Runs case `name` through the mixstage_b200 classes exactly as the reference's trainer
    would (GAN.forward / G.forward); returns a dict comparable with oracle_cases.run_oracle."""
    from mixstage_b200 import ops
    old = ops.get_precision()
    ops.set_precision(precision)
    try:
        return _run_case(name, device, dtype)
    finally:
        ops.set_precision(old)


def _run_case(name, device, dtype):
    spec, B, T, kind, kw = CASES[name]
    G, D, gan = build(spec, T, device, dtype)
    audio, pose, labels, style = O.synth_inputs(B, T, spec, dtype=dtype)
    audio, pose, labels, style = (t.to(device) for t in (audio, pose, labels, style))
    res = {"G": G, "D": D, "gan": gan}
    if kind == "gan":
        step = kw["step"]
        if kw.get("use_pose_encoder"):
            G.thresh.value, G.thresh.iters = 0.0, 0
        else:
            G.thresh.value, G.thresh.iters = 1.0, 1000
        if step == "eval":
            gan.eval()
            with torch.no_grad():
                fake, losses, args = gan([audio, labels], pose, input_modalities=MOD, style=style, sample_flag=0,
                                         description="dev", desc="dev")
        else:
            gan.train()
            gan.force_step = step
            fake, losses, args = gan([audio, labels], pose, input_modalities=MOD, style=style, sample_flag=0,
                                     description="train", desc="train")
            sum(losses).backward()
    else:
        if kind == "g_long":
            audio, pose, labels = audio.reshape(1, B * T, -1), pose.reshape(1, B * T, -1), labels.reshape(1, B * T)
        G.train(kw["training"])
        G.thresh.value, G.thresh.iters = 1.0, 1000
        with torch.no_grad():
            fake, losses = G([audio, labels], pose, input_modalities=MOD, style=style, sample_flag=kw["sample_flag"],
                             description=kw["description"])
    res.update(pose=fake.detach(), losses=[float(l.detach()) for l in losses], labels_cap_soft=G.labels_cap_soft.detach())
    return res
