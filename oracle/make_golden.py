"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (build container only).

    python oracle/make_golden.py            # needs /root/reference

The reference ships no golden vectors, so parity is pinned on outputs of
the reference's own modules run here on seeded synthetic inputs/weights
(``mixstage_oracle.synth_state`` / ``synth_inputs``; G weights seed 7, D weights seed 8,
inputs seed 11212).  Tensors are stored as fp32 (they come from an fp64 run), losses and
norms as fp64.  Cases (BASELINE.json configs):

  cfg1_eval_sample   config 1: B16 T64 S2 K8 argmax=1, eval, sample_flag=1 ('test')
  cfg1_train_fwd     config 1 model in train mode, description='train' (pose-style path)
  cfg2_gstep         config 2: B16 S4 GAN generator step (losses, grads, running stats)
  cfg2_dstep         config 2: discriminator step
  cfg2_eval          config 2 model, GAN eval branch (dev loop)
  cfg2_pose_branch   curriculum branch (pose encoder instead of audio encoder), G-step
  cfg5_stress_small  config 5 shape at B=2: T256 S25 K16 argmax=0 (soft style), G-step
  sample_long        sampling layout: batch 1 x (2*64) frames, style (2,64) (trainer.py:778-786)
  stage_k1_gstep     StAGE variant (num_clusters=1, src/jobs/stage.py): B8 S4, G-step          
  prep_pvs_k8 / prep_pva_k16   KMeans.predict + ZNorm on a raw pose batch (src/data/transform.py)     
  metrics_l1_vel_pck     L1 / VelL1 / PCK running averages over two batches (src/evaluation/metrics.py)  (§8f row 4)
  s2g_eval / s2g_train   Speech2Gesture_G baseline (speech2gesture.py:13-40): eval forward; train forward + L1 backward

    python oracle/make_golden.py [case ...]     # default: all cases
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import mixstage_oracle as O          # noqa: E402
import ref_loader                    # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
MOD = ["audio/log_mel_400"]
FULL_GRADS_G = ["style_emb.emb.weight", "classify_cluster.logits.weight", "logits.bias",
                "audio_encoder.conv.0.conv.weight", "unet.conv1.4.norm.weight",
                "pose_style_encoder.conv.6.conv.weight", "decoder.3.norm.bias"]
FULL_GRADS_D = ["logits.weight", "conv2.0.norm.weight", "conv1.0.bias"]
FULL_STATS = ["unet.conv1.4.norm.running_mean", "unet.conv1.4.norm.running_var",
              "pose_style_encoder.conv.6.norm.running_mean", "pose_style_encoder.conv.6.norm.running_var",
              "audio_encoder.conv.7.norm.running_var", "decoder.0.norm.running_mean"]


class FixedRand:
    """Replaces torch.rand for the two host coin flips (gan.py:105, jlcss.py:127)."""

    def __init__(self, vals):
        self.vals = list(vals)

    def __call__(self, *a, **k):
        return torch.tensor([self.vals.pop(0)])


def build(ns, spec, T):
    G = ns.G(time_steps=T, out_feats=spec.out_feats, num_clusters=spec.num_clusters,
             style_dict={i: i for i in range(spec.num_speakers)}, style_dim=spec.style_dim,
             lambda_id=spec.lambda_id, train_only=spec.train_only, softmax=spec.softmax,
             argmax=spec.argmax, some_grad_flag=spec.some_grad_flag,
             shape={MOD[0]: [T, spec.mel_bins]}).double()
    G.load_state_dict(O.synth_state(O.g_state_shapes(spec), 7))
    D = ns.D(in_channels=spec.out_feats).double()
    D.load_state_dict(O.synth_state(O.d_state_shapes(spec.out_feats), 8))
    gan = ns.GAN(G, D, criterion="L1Loss", no_grad=0, input_modalities=MOD).double()
    return G, D, gan


def f32(t):
    return t.detach().to(torch.float32).numpy()


def grad_norms(module):
    names, vals = [], []
    for n, p in module.named_parameters():
        if p.grad is not None:
            names.append(n)
            vals.append(float(p.grad.norm()))
    return np.array(names), np.array(vals, dtype=np.float64)


def run_gan(ns, spec, B, T, step, pose_branch=False):
    G, D, gan = build(ns, spec, T)
    audio, pose, labels, style = O.synth_inputs(B, T, spec)
    sd0 = {k: v.clone() for k, v in G.state_dict().items()}
    sdd0 = {k: v.clone() for k, v in D.state_dict().items()}
    if step == "eval":
        gan.eval()
        kw = dict(sample_flag=0, description="dev")
    else:
        gan.train()
        kw = dict(sample_flag=0, description="train")
    if pose_branch:
        G.thresh.value, G.thresh.iters = 0.0, 0          # rand(0.5) > 0 -> pose encoder
    else:
        G.thresh.value, G.thresh.iters = 1.0, 1000
    orig = torch.rand
    coin = {"G": 0.9, "D": 0.1, "eval": 0.9}[step]
    torch.rand = FixedRand([coin, 0.5] if step != "eval" else [0.5])
    try:
        ctx = torch.no_grad() if step == "eval" else torch.enable_grad()
        with ctx:
            fake, il, args = gan([audio.clone(), labels], pose, input_modalities=MOD, style=style, desc="x", **kw)
            if step != "eval":
                sum(il).backward()
    finally:
        torch.rand = orig
    out = {"pose": f32(fake), "losses": np.array([float(l) for l in il], dtype=np.float64),
           "labels_cap_soft": f32(G.labels_cap_soft),
           "cluster_argmax": G.labels_cap_soft.argmax(-1).numpy().astype(np.int16)}
    if step != "eval":
        gn, gv = grad_norms(G)
        dn, dv = grad_norms(D)
        out.update(g_grad_names=gn, g_grad_norms=gv, d_grad_names=dn, d_grad_norms=dv)
        gp = dict(G.named_parameters())
        dp = dict(D.named_parameters())
        for k in FULL_GRADS_G:
            if gp[k].grad is not None:
                out["ggrad/" + k] = f32(gp[k].grad)
        for k in FULL_GRADS_D:
            if dp[k].grad is not None:
                out["dgrad/" + k] = f32(dp[k].grad)
        sd1 = G.state_dict()
        sdd1 = D.state_dict()
        nbt = {k: int(sd1[k] - sd0[k]) for k in sd1 if k.endswith("num_batches_tracked")}
        out["nbt_names"] = np.array(list(nbt.keys()))
        out["nbt_incr"] = np.array(list(nbt.values()), dtype=np.int64)
        nbt_d = {k: int(sdd1[k] - sdd0[k]) for k in sdd1 if k.endswith("num_batches_tracked")}
        out["d_nbt_names"] = np.array(list(nbt_d.keys()))
        out["d_nbt_incr"] = np.array(list(nbt_d.values()), dtype=np.int64)
        for k in FULL_STATS:
            out["gstat/" + k] = f32(sd1[k])
        for k in ("conv2.0.norm.running_mean", "conv3.norm.running_var"):
            out["dstat/" + k] = f32(sdd1[k])
    return out


def run_g_only(ns, spec, B, T, training, sample_flag, description, long_layout=False):
    G, D, gan = build(ns, spec, T)
    audio, pose, labels, style = O.synth_inputs(B, T, spec)
    if long_layout:                                  # trainer.py:778-786
        audio = audio.reshape(1, B * T, -1)
        pose = pose.reshape(1, B * T, -1)
        labels = labels.reshape(1, B * T)
    G.train(training)
    G.thresh.value, G.thresh.iters = 1.0, 1000
    orig = torch.rand
    torch.rand = FixedRand([0.5])
    try:
        with torch.no_grad():
            out, il = G([audio.clone(), labels], pose, input_modalities=MOD, style=style,
                        sample_flag=sample_flag, description=description)
    finally:
        torch.rand = orig
    return {"pose": f32(out), "losses": np.array([float(l) for l in il], dtype=np.float64),
            "labels_cap_soft": f32(G.labels_cap_soft),
            "cluster_argmax": G.labels_cap_soft.argmax(-1).numpy().astype(np.int16)}


S2G_SEED = 9
S2G_GRADS = ["logits.weight", "decoder.3.norm.weight", "unet.conv1.4.norm.weight", "audio_encoder.conv.0.conv.weight"]


def run_s2g(ns, P, B, T, training):
    """Speech2Gesture_G on the synthetic audio batch; train mode also back-propagates mean|pose - y| (the trainer's
    L1 criterion, trainer.py:1268-1285 with args.loss = L1Loss) and records gradients / BatchNorm side effects."""
    spec = O.Spec(num_speakers=4, out_feats=P)
    G = ns.S2G(time_steps=T, out_feats=P).double()
    G.load_state_dict(O.synth_state(O.s2g_state_shapes(P), S2G_SEED))
    audio, pose, _, _ = O.synth_inputs(B, T, spec)
    G.train(training)
    sd0 = {k: v.clone() for k, v in G.state_dict().items()}
    if training:
        out, il = G(audio.clone(), pose)
        loss = (out - pose).abs().mean()
        loss.backward()
    else:
        with torch.no_grad():
            out, il = G(audio.clone(), pose)
        loss = (out - pose).abs().mean()
    res = {"pose": f32(out), "losses": np.array([float(loss)], dtype=np.float64)}
    if training:
        gp = dict(G.named_parameters())
        names = [n for n, p in gp.items() if p.grad is not None]
        res["grad_names"] = np.array(names)
        res["grad_norms"] = np.array([float(gp[n].grad.norm()) for n in names], dtype=np.float64)
        for n in S2G_GRADS:
            res["grad/" + n] = f32(gp[n].grad)
        sd1 = G.state_dict()
        nbt = {k: int(sd1[k] - sd0[k]) for k in sd1 if k.endswith("num_batches_tracked")}
        res["nbt_names"] = np.array(list(nbt.keys()))
        res["nbt_incr"] = np.array(list(nbt.values()), dtype=np.int64)
        for k in ("unet.conv1.4.norm.running_mean", "decoder.0.norm.running_var"):
            res["gstat/" + k] = f32(sd1[k])
    return res


PREP_FEATS = ["pose", "velocity", "speed"]
PREP_MASK = [0, 7, 8, 9]


def run_prep(feats, K):
    """KMeans.predict / ZNorm.znorm of the reference (its own function bodies, ref_loader.load_transform_functions) on a
    synthetic raw pose batch.  RemoveJoints' slicing lives in un-vendored pycasper: the oracle's remove_joints stands in."""
    import types
    fns = ref_loader.load_transform_functions()
    x, mean, var, centers = O.synth_prep(4, 64, K, feats)
    xr = O.remove_joints(x, PREP_MASK)
    km = types.SimpleNamespace(feats=feats, centers=centers)
    km.get_feats = lambda t: fns["KMeans.get_feats"](km, t)
    labels = fns["KMeans.predict"](km, xr.clone())
    soft = fns["KMeans.predict"](km, xr.clone(), soft_labels=True)
    zn = fns["ZNorm.znorm"](None, x, [mean.view(1, 1, -1), var.view(1, 1, -1)])
    y = O.remove_joints(zn, PREP_MASK)
    inv = fns["ZNorm.inv_znorm"](None, zn[..., 4:5].expand(-1, -1, x.shape[-1]).contiguous(), [mean.view(1, 1, -1), var.abs().view(1, 1, -1)])
    return {"labels": labels.numpy().astype(np.int16), "soft": soft.numpy(), "y": y.numpy(), "inv": inv[:1].numpy(),
            "pose": np.zeros((1,), dtype=np.float32), "losses": np.zeros((0,))}


def run_metrics():
    """The reference's own L1 / VelL1 / PCK classes (ref_loader.load_metric_classes) driven as calculate_metrics drives them
    (trainer.py:886-907) on two synthetic batches; stores get_averages('test')."""
    env = ref_loader.load_metric_classes()
    mean, var, batches = O.synth_metric_batches()
    l1, vel, pck = env["L1"](), env["VelL1"](), env["PCK"](num_joints=52)
    for y_cap, y_ in batches:
        l1(y_cap, y_, PREP_MASK)
        vel(y_cap, y_, PREP_MASK)
        yu = (y_cap * (var.view(1, 1, -1) ** 0.5) + mean.view(1, 1, -1))          # ZNorm.inv_znorm, transform.py:228
        gu = (y_ * (var.view(1, 1, -1) ** 0.5) + mean.view(1, 1, -1))
        yu = yu.view(yu.shape[0], yu.shape[1], 2, -1).view(-1, 2, 52).clone()
        gu = gu.view(gu.shape[0], gu.shape[1], 2, -1).view(-1, 2, 52).clone()
        yu[..., 0] = 0
        gu[..., 0] = 0
        pck(yu, gu, PREP_MASK)
    av = {}
    for m in (l1, vel, pck):
        av.update(m.get_averages("test"))
    keys = sorted(av)
    return {"keys": np.array(keys), "values": np.array([av[k] for k in keys], dtype=np.float64),
            "pose": np.zeros((1,), dtype=np.float32), "losses": np.zeros((0,))}


def main():
    torch.set_num_threads(os.cpu_count())
    ns = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    cfg1 = O.Spec(num_speakers=2)
    cfg2 = O.Spec(num_speakers=4)
    cfg5 = O.Spec(num_speakers=25, num_clusters=16, argmax=0, time_steps=256)
    cases = {
        "cfg1_eval_sample": lambda: run_g_only(ns, cfg1, 16, 64, False, 1, "test"),
        "cfg1_train_fwd": lambda: run_g_only(ns, cfg1, 16, 64, True, 0, "train"),
        "cfg2_gstep": lambda: run_gan(ns, cfg2, 16, 64, "G"),
        "cfg2_dstep": lambda: run_gan(ns, cfg2, 16, 64, "D"),
        "cfg2_eval": lambda: run_gan(ns, cfg2, 16, 64, "eval"),
        "cfg2_pose_branch": lambda: run_gan(ns, cfg2, 16, 64, "G", pose_branch=True),
        "cfg5_stress_small": lambda: run_gan(ns, cfg5, 2, 256, "G"),
        "sample_long": lambda: run_g_only(ns, cfg2, 2, 64, False, 1, "test", long_layout=True),
        "stage_k1_gstep": lambda: run_gan(ns, O.Spec(num_speakers=4, num_clusters=1), 8, 64, "G"),
        "s2g_eval": lambda: run_s2g(ns, 96, 8, 64, False),
        "s2g_train": lambda: run_s2g(ns, 96, 8, 64, True),
        "prep_pvs_k8": lambda: run_prep(PREP_FEATS, 8),
        "prep_pva_k16": lambda: run_prep(["pose", "velocity", "acceleration"], 16),
        "metrics_l1_vel_pck": run_metrics,
    }
    only = sys.argv[1:]
    for name, fn in cases.items():
        if only and name not in only:
            continue
        res = fn()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
        print(name, "losses", res["losses"], "pose", res["pose"].shape)
    if only:
        return
    # state_dict key/shape contract of the reference classes
    G, D, _ = build(ns, cfg2, 64)
    with open(os.path.join(OUT, "state_dict_keys.txt"), "w") as f:
        for k, v in G.state_dict().items():
            f.write("G %s %s %s\n" % (k, "x".join(map(str, v.shape)) or "-", str(v.dtype).replace("torch.", "")))
        for k, v in D.state_dict().items():
            f.write("D %s %s %s\n" % (k, "x".join(map(str, v.shape)) or "-", str(v.dtype).replace("torch.", "")))


if __name__ == "__main__":
    main()
