"""CPU oracle for the Mix-StAGE generator hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (functional, table-driven, plain torch CPU ops in
fp64 by default) of the reference algorithm for the path named in
BASELINE.json's north_star: forward/backward of ``JointLateClusterSoftStyle4_G``
plus the pose discriminator and the L1/GAN losses.  It is the checker for the
CUDA path; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product package
(``mixstage_b200``) never does.

Pinning status: the reference ships no tests or golden vectors, so
the oracle is pinned against *outputs of the reference itself executed in the
build container* -- ``oracle/make_golden.py`` imports the unmodified reference
modules (``oracle/ref_loader.py``), runs them on the seeded synthetic inputs
defined below and commits the results under ``tests/golden``;
``tests/test_oracle_golden.py`` checks this file against those vectors on every
run, and ``tests/test_oracle_vs_reference.py`` re-checks it live whenever
``/root/reference`` is present.  Two pycasper symbols on the path (``some_grad``,
``LambdaScheduler``) are un-vendored: their semantics are defined in
``oracle/ref_loader.py`` and are "parity unpinned" (SURVEY.md §8c).

Rows either side of the path, same rules: ``s2g_forward`` (Speech2Gesture_G) and the K=1 StAGE case are
pinned to golden vectors of the executed reference classes; ``kmeans_feats`` / ``kmeans_predict`` / ``znorm`` /
``inv_znorm`` and ``pose_metrics`` are pinned to the reference's OWN function / class bodies (src/data/transform.py,
src/evaluation/metrics.py), cut out of the source with ``ast`` and executed unchanged because the modules around them
import h5py/librosa (``ref_loader.load_transform_functions`` / ``load_metric_classes``).  ``remove_joints`` stands in for
pycasper's un-vendored ``remove_slices``: parity unpinned.

Reference citations are relative to /root/reference/src/model/.
"""
from __future__ import annotations

import dataclasses
import zlib
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.2      # layers.py:72-73
BN_EPS = 1e-5          # nn.BatchNorm default used at layers.py:64,70
BN_MOMENTUM = 0.1


@dataclasses.dataclass
class Spec:
    """Constructor arguments of JointLateClusterSoftStyle4_G that shape the path
    (joint_late_cluster_soft_style.py:26-30)."""
    time_steps: int = 64
    out_feats: int = 96
    num_clusters: int = 8
    num_speakers: int = 4          # len(style_dict)
    style_dim: int = 10
    lambda_id: float = 0.1
    train_only: int = 1
    softmax: int = 1
    argmax: int = 1
    some_grad_flag: int = 1
    in_channels: int = 256
    mel_bins: int = 64
    text_channels: int = 300


# --------------------------------------------------------------------------
# architecture tables: (name, c_in, c_out, kernel, stride, padding, groups)
# --------------------------------------------------------------------------
def _pad(k, s):
    # layers.py:46-55 -- padding = int((k - s) / 2) per dim
    return int((k - s) / 2)


def audio_encoder_table():
    """layers.py:166-185.  All Conv2d.  The last block has kernel (3, 8), stride 1
    and therefore padding (1, 3) via the tuple/int rule at layers.py:49-50."""
    chans = [(1, 64, False), (64, 64, True), (64, 128, False), (128, 128, True),
             (128, 256, False), (256, 256, True), (256, 256, False)]
    rows = []
    for i, (ci, co, down) in enumerate(chans):
        k, s = (4, 2) if down else (3, 1)
        rows.append(("audio_encoder.conv.%d" % i, ci, co, (k, k), (s, s), (_pad(k, s),) * 2, 1))
    rows.append(("audio_encoder.conv.7", 256, 256, (3, 8), (1, 1), (_pad(3, 1), _pad(8, 1)), 1))
    return rows


def seq_encoder_table(prefix, c_in):
    """PoseEncoder layers.py:208-224 / TextEncoder1D layers.py:348-363: six k3 s1 blocks."""
    chans = [c_in, 64, 64, 128, 128, 256, 256]
    return [("%s.conv.%d" % (prefix, i), chans[i], chans[i + 1], 3, 1, 1, 1) for i in range(6)]


def pose_style_encoder_table(c_in, num_speakers):
    """layers.py:253-271: one k3 block then six stride-2 k4 blocks, last -> num_speakers."""
    chans = [c_in, 64, 64, 128, 128, 256, 256, num_speakers]
    rows = [("pose_style_encoder.conv.0", chans[0], chans[1], 3, 1, 1, 1)]
    for i in range(1, 7):
        rows.append(("pose_style_encoder.conv.%d" % i, chans[i], chans[i + 1], 4, 2, 1, 1))
    return rows


def unet_table(c=256, depth=5):
    """layers.py:118-132."""
    rows = [("unet.pre_downsampling_conv.%d" % i, c, c, 3, 1, 1, 1) for i in range(2)]
    rows += [("unet.conv1.%d" % i, c, c, 4, 2, 1, 1) for i in range(depth)]
    rows += [("unet.conv2.%d" % i, c, c, 3, 1, 1, 1) for i in range(depth)]
    return rows


def classify_table(c_in):
    """layers.py:454-457 (the 1x1 logits conv at :459 is listed separately)."""
    return [("classify_cluster.conv.%d" % i, c_in if i == 0 else 256, 256, 3, 1, 1, 1) for i in range(6)]


def decoder_table(spec: Spec):
    """joint_late_cluster_soft_style.py:69-77: channels are multiplied by groups
    inside ConvNormRelu (layers.py:58-59)."""
    K = spec.num_clusters
    c = spec.in_channels
    rows = [("decoder.0", (spec.style_dim + c), c, 3, 1, 1, K)]
    rows += [("decoder.%d" % i, c, c, 3, 1, 1, K) for i in range(1, 4)]
    return rows


def disc_table(c_in):
    """speech2gesture.py:74-90 with n_downsampling=2, out_channels=64.
    conv1.0 has no norm; conv3 is k4 s1 p=int((4-1)/2)=1; logits k4 s1 p0."""
    return [("conv2.0", 64, 128, 4, 2, 1, 1), ("conv3", 128, 256, 4, 1, 1, 1)]


def g_state_shapes(spec: Spec) -> Dict[str, Tuple[Tuple[int, ...], str]]:
    """Every state_dict entry of the reference G (SURVEY.md §8a contract): key ->
    (shape, kind).  kind selects the synthetic initialiser in ``synth_state``."""
    out: Dict[str, Tuple[Tuple[int, ...], str]] = {}
    K, S, P, c = spec.num_clusters, spec.num_speakers, spec.out_feats, spec.in_channels

    def block(name, ci, co, k, groups=1):
        ks = tuple(k) if isinstance(k, tuple) else (k,)
        out[name + ".conv.weight"] = ((co * groups, ci) + ks, "conv_w")
        out[name + ".conv.bias"] = ((co * groups,), "conv_b:" + str(ci * _prod(ks)))
        out[name + ".norm.weight"] = ((co * groups,), "bn_w")
        out[name + ".norm.bias"] = ((co * groups,), "bn_b")
        out[name + ".norm.running_mean"] = ((co * groups,), "bn_rm")
        out[name + ".norm.running_var"] = ((co * groups,), "bn_rv")
        out[name + ".norm.num_batches_tracked"] = ((), "nbt")

    out["eye"] = ((K, K), "eye")
    for (n, ci, co, k, s, p, g) in audio_encoder_table():
        block(n, ci, co, k)
    for (n, ci, co, k, s, p, g) in seq_encoder_table("text_encoder", spec.text_channels):
        block(n, ci, co, k)
    for (n, ci, co, k, s, p, g) in seq_encoder_table("pose_encoder", P):
        block(n, ci, co, k)
    for (n, ci, co, k, s, p, g) in unet_table(c):
        block(n, ci, co, k)
    for (n, ci, co, k, s, p, g) in pose_style_encoder_table(P, S):
        block(n, ci, co, k)
    out["style_emb.emb.weight"] = ((S, spec.style_dim), "emb")
    for pre in ("style_dec", "style_dec_gr.models.0"):      # aliased modules, both key sets exist
        for i in range(2):
            block("%s.%d" % (pre, i), c, c, 3, groups=spec.style_dim)
    for (n, ci, co, k, s, p, g) in decoder_table(spec):
        block(n, ci, co, k, groups=g)
    block("concat_encoder.0", 512, 256, 3)
    out["logits.weight"] = ((P * K, c, 1), "conv_w")
    out["logits.bias"] = ((P * K,), "conv_b:%d" % c)
    for (n, ci, co, k, s, p, g) in classify_table(c + spec.style_dim):
        block(n, ci, co, k)
    out["classify_cluster.logits.weight"] = ((K, 256, 1), "conv_w")
    out["classify_cluster.logits.bias"] = ((K,), "conv_b:256")
    block("smoothen", P, P, 3)
    return out


def d_state_shapes(c_in: int) -> Dict[str, Tuple[Tuple[int, ...], str]]:
    out: Dict[str, Tuple[Tuple[int, ...], str]] = {}
    out["conv1.0.weight"] = ((64, c_in, 4), "conv_w")
    out["conv1.0.bias"] = ((64,), "conv_b:%d" % (c_in * 4))
    for (n, ci, co, k, s, p, g) in disc_table(c_in):
        out[n + ".conv.weight"] = ((co, ci, k), "conv_w")
        out[n + ".conv.bias"] = ((co,), "conv_b:%d" % (ci * k))
        out[n + ".norm.weight"] = ((co,), "bn_w")
        out[n + ".norm.bias"] = ((co,), "bn_b")
        out[n + ".norm.running_mean"] = ((co,), "bn_rm")
        out[n + ".norm.running_var"] = ((co,), "bn_rv")
        out[n + ".norm.num_batches_tracked"] = ((), "nbt")
    out["logits.weight"] = ((1, 256, 4), "conv_w")
    out["logits.bias"] = ((1,), "conv_b:1024")
    return out


def _prod(t):
    r = 1
    for v in t:
        r *= v
    return r


def synth_state(shapes, seed: int, dtype=torch.float64) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights, independent of module construction order.

    Every tensor is drawn from its own generator seeded by (seed, crc32(key)); the
    values are generated in fp32 and then widened so fp32 and fp64 runs see the very
    same numbers.  Unlike PyTorch's defaults the BatchNorm affine parameters and
    running statistics are non-trivial so that a path which ignores them fails."""
    sd = {}
    alias = {}
    for key, (shape, kind) in shapes.items():
        if key.startswith("style_dec_gr.models.0."):
            alias[key] = "style_dec." + key[len("style_dec_gr.models.0."):]
            continue
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63))
        if kind == "conv_w":
            fan_in = _prod(shape[1:])
            b = 1.0 / fan_in ** 0.5
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * b
        elif kind.startswith("conv_b:"):
            b = 1.0 / int(kind.split(":")[1]) ** 0.5
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * b
        elif kind == "bn_w":
            t = 0.5 + torch.rand(shape, generator=g, dtype=torch.float32)
        elif kind == "bn_b":
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * 0.2
        elif kind == "bn_rm":
            t = torch.randn(shape, generator=g, dtype=torch.float32) * 0.1
        elif kind == "bn_rv":
            t = 0.5 + torch.rand(shape, generator=g, dtype=torch.float32)
        elif kind == "emb":
            t = torch.randn(shape, generator=g, dtype=torch.float32)
        elif kind == "eye":
            t = torch.eye(shape[0], dtype=torch.float32)
        elif kind == "nbt":
            sd[key] = torch.zeros((), dtype=torch.int64)
            continue
        else:
            raise KeyError(kind)
        sd[key] = t.to(dtype)
    for k, src in alias.items():
        sd[k] = sd[src]
    return sd


def synth_inputs(B: int, T: int, spec: Spec, seed: int = 11212, dtype=torch.float64):
    """SURVEY.md §8d synthetic batch: audio ~ N(0,1) (B,T,64), pose ~ N(0,1) (B,T,P),
    labels = randint(0,K) (B,T) int64, style = arange(B) % S broadcast over T."""
    g = torch.Generator().manual_seed(seed)
    audio = torch.randn(B, T, spec.mel_bins, generator=g, dtype=torch.float32).to(dtype)
    pose = torch.randn(B, T, spec.out_feats, generator=g, dtype=torch.float32).to(dtype)
    labels = torch.randint(0, spec.num_clusters, (B, T), generator=g, dtype=torch.int64)
    style = (torch.arange(B) % spec.num_speakers)[:, None].expand(B, T).contiguous()
    return audio, pose, labels, style


# --------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------
class BNLog:
    """Collects running-stat updates made by train-mode BatchNorm: key -> new tensor,
    plus how many times each block ran (num_batches_tracked increments)."""

    def __init__(self):
        self.updates: Dict[str, torch.Tensor] = {}
        self.counts: Dict[str, int] = {}


def conv_norm_relu(x, sd, name, stride, padding, groups, training, log: Optional[BNLog]):
    """ConvNormRelu.forward, layers.py:78: relu(norm(dropout(conv(x)))) with p=0
    dropout (identity), BatchNorm (batch statistics when training, running
    statistics otherwise) and LeakyReLU(0.2)."""
    w, b = sd[name + ".conv.weight"], sd[name + ".conv.bias"]
    if w.dim() == 4:
        z = F.conv2d(x, w, b, stride=stride, padding=padding, groups=groups)
        red = (0, 2, 3)
        shp = (1, -1, 1, 1)
    else:
        z = F.conv1d(x, w, b, stride=stride, padding=padding, groups=groups)
        red = (0, 2)
        shp = (1, -1, 1)
    gamma, beta = sd[name + ".norm.weight"], sd[name + ".norm.bias"]
    rm_key, rv_key = name + ".norm.running_mean", name + ".norm.running_var"
    # a block may run twice per step (pose_style_encoder): chain the running stats
    rm = log.updates.get(rm_key, sd[rm_key]) if log is not None else sd[rm_key]
    rv = log.updates.get(rv_key, sd[rv_key]) if log is not None else sd[rv_key]
    if training:
        n = z.numel() // z.shape[1]
        mean = z.mean(dim=red)
        var = ((z - mean.view(shp)) ** 2).mean(dim=red)          # biased, used to normalise
        if log is not None:
            with torch.no_grad():
                log.updates[rm_key] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean.detach()
                log.updates[rv_key] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var.detach() * (n / max(n - 1, 1))
                log.counts[name] = log.counts.get(name, 0) + 1
    else:
        mean, var = rm, rv
    zh = (z - mean.view(shp)) / torch.sqrt(var.view(shp) + BN_EPS)
    a = zh * gamma.view(shp) + beta.view(shp)
    return F.leaky_relu(a, LEAKY_SLOPE)


def run_table(x, sd, table, training, log):
    for (n, ci, co, k, s, p, g) in table:
        x = conv_norm_relu(x, sd, n, s, p, g, training, log)
    return x


def audio_encoder(x, sd, time_steps, training, log):
    """AudioEncoder.forward, layers.py:191-199.  x: (B, 1, T, F) -> (B, 256, T)."""
    x = run_table(x, sd, audio_encoder_table(), training, log)
    x = F.interpolate(x, size=(time_steps, 1), mode="bilinear")
    return x.squeeze(-1)


def unet1d(x, sd, training, log, depth=5):
    """UNet1D.forward, layers.py:134-157."""
    T = x.shape[-1]
    assert T >= 2 ** (depth - 1) and T % (2 ** depth) == 0, "layers.py:136-138"
    tab = unet_table()
    x = run_table(x, sd, tab[:2], training, log)
    res = [x]
    for i in range(depth):
        x = run_table(x, sd, [tab[2 + i]], training, log)
        if i < depth - 1:
            res.append(x)
    for i in range(depth):
        x = F.interpolate(x, scale_factor=2, mode="nearest") + res[depth - i - 1]
        x = run_table(x, sd, [tab[2 + depth + i]], training, log)
    return x


def pose_style_encoder(y, sd, spec: Spec, training, log):
    """PoseStyleEncoder.forward, layers.py:279-289.  y: (B, T, P) -> (B, S)."""
    x = run_table(y.transpose(1, 2), sd, pose_style_encoder_table(spec.out_feats, spec.num_speakers), training, log)
    return x.mean(-1)


def style_emb(sd, pose_style, mode):
    """EmbLin.forward, layers.py:659-663."""
    w = sd["style_emb.emb.weight"]
    if mode == "lin":
        return pose_style.matmul(w)
    return w[pose_style]


def mixture(x, w, K):
    """index_select_outputs, joint_late_cluster_soft_style.py:106-115.
    x: (B, K*P, T); w: (B, T, K) -> (B, T, P)."""
    x = x.transpose(2, 1)
    x = x.reshape(x.shape[0], x.shape[1], K, -1)
    w = w.reshape(x.shape[0], x.shape[1], x.shape[2])
    return (x * w.unsqueeze(-1)).sum(dim=-2)


def g_forward(sd, spec: Spec, audio, labels, y, style, *, training: bool, sample_flag: int,
              description: str, use_pose_encoder: bool = False, log: Optional[BNLog] = None,
              freeze_pose_style_out: Optional[bool] = None):
    """JointLateClusterSoftStyle4_G.forward, joint_late_cluster_soft_style.py:117-209.

    ``use_pose_encoder`` replaces the curriculum coin flip at :127 (deterministic
    here).  Returns (pose (B,T,P), [cluster_CE, id_in*lambda, id_out*lambda], aux)."""
    T = spec.time_steps if audio is None else audio.shape[-2]
    if use_pose_encoder and training:
        x = run_table(y.transpose(1, 2), sd, seq_encoder_table("pose_encoder", spec.out_feats), training, log)
    else:
        a = audio.unsqueeze(1) if audio.dim() == 3 else audio                # :138-139
        x = audio_encoder(a, sd, T, training, log)                          # :140
    x = unet1d(x, sd, training, log)                                         # :149
    x = x.transpose(2, 1)                                                    # :151 (B,T,256)
    aux = {}
    flag = (not sample_flag) and (description == "train" or not spec.train_only)   # :154
    if flag:
        mode = "lin"
        score = pose_style_encoder(y, sd, spec, training, log)               # :158
        id_in = F.cross_entropy(score, style[:, 0])                          # :159
        aux["pose_style_score"] = score
        score_t = score.unsqueeze(1).expand(score.shape[0], x.shape[1], score.shape[-1])
        if spec.softmax:
            pose_style = torch.softmax(score_t, dim=-1)
            if spec.argmax:
                pose_style = torch.argmax(pose_style, dim=-1)
                mode = "emb"
        else:
            pose_style = score_t
    else:
        pose_style = style
        mode = "emb" if style.dim() == 2 else "lin"                          # :169-173
        id_in = torch.zeros((), dtype=x.dtype)
    aux["style_index"] = pose_style if mode == "emb" else None
    ls = style_emb(sd, pose_style, mode)                                     # :175
    if x.shape[1] != ls.shape[1]:
        ls = ls.reshape(x.shape[0], -1, ls.shape[-1])                        # :177-178
    x = torch.cat([x, ls], dim=-1).transpose(2, 1)                           # :180 (B,266,T)
    h = run_table(x, sd, classify_table(spec.in_channels + spec.style_dim), training, log)
    score_c = F.conv1d(h, sd["classify_cluster.logits.weight"], sd["classify_cluster.logits.bias"])
    score_c = score_c.transpose(2, 1)                                        # :183 (B,T,K)
    ce = F.cross_entropy(score_c.reshape(-1, score_c.shape[-1]), labels.reshape(-1))   # :184
    soft = torch.softmax(score_c, dim=-1)                                    # :186
    aux["labels_cap_soft"] = soft
    aux["labels_score"] = score_c
    K = spec.num_clusters
    x = torch.cat([x] * K, dim=1)                                            # :190
    x = run_table(x, sd, decoder_table(spec), training, log)                 # :192
    x = F.conv1d(x, sd["logits.weight"], sd["logits.bias"], groups=K)        # :193
    x = mixture(x, soft, K)                                                  # :194
    if flag:
        frozen = bool(spec.some_grad_flag) if freeze_pose_style_out is None else freeze_pose_style_out
        sd_out = sd
        if frozen:                                                           # some_grad, :198-200
            sd_out = {k: (v.detach() if k.startswith("pose_style_encoder.") else v) for k, v in sd.items()}
        score_out = pose_style_encoder(x, sd_out, spec, training, log)
        id_out = F.cross_entropy(score_out, style[:, 0])                     # :203
    else:
        id_out = torch.zeros((), dtype=x.dtype)
    return x, [ce, id_in * spec.lambda_id, id_out * spec.lambda_id], aux


def d_forward(sd, x, training, log: Optional[BNLog] = None):
    """Speech2Gesture_D.forward, speech2gesture.py:92-100.  x: (B,T,P) -> (B,L')."""
    h = x.transpose(-1, -2)
    h = F.leaky_relu(F.conv1d(h, sd["conv1.0.weight"], sd["conv1.0.bias"], stride=2, padding=1), LEAKY_SLOPE)
    h = run_table(h, sd, disc_table(x.shape[-1]), training, log)
    h = F.conv1d(h, sd["logits.weight"], sd["logits.bias"])
    return h.transpose(-1, -2).squeeze(-1)


def velocity(x):
    """GAN.get_velocity with joint=False, gan.py:47-52."""
    return torch.cat([torch.zeros_like(x[..., 0:1, :]), x[..., 1:, :] - x[..., :-1, :]], dim=-2)


def l1_mean(a, b):
    """GAN.get_loss / get_gan_loss with criterion L1Loss(reduction='none') and unit
    sample weights, gan.py:64-75."""
    return (a - b).abs().mean()


def gan_forward(sd_g, sd_d, spec: Spec, audio, labels, y, style, *, step: str,
                lambda_D: float = 1.0, lambda_gan: float = 1.0, use_pose_encoder=False,
                log_g: Optional[BNLog] = None, log_d: Optional[BNLog] = None):
    """GAN.forward, gan.py:86-164, with the coin flip at :105 replaced by ``step``:
    'G' (generator step, :135-152), 'D' (discriminator step, :106-132) or 'eval'
    (:153-161).  Returns (fake_pose, losses list in the reference's order, aux)."""
    kw = dict(sample_flag=0, description="train")
    if step == "G":
        fake, part, aux = g_forward(sd_g, spec, audio, labels, y, style, training=True,
                                    use_pose_encoder=use_pose_encoder, log=log_g, **kw)
        score = d_forward(sd_d, velocity(fake), True, log_d)                 # no_grad=0: grads flow
        g_gan = lambda_gan * l1_mean(score, torch.ones_like(score))
        pose = l1_mean(fake, y)
        aux["fake_score"] = score
        return fake, [pose, g_gan] + part, aux
    if step == "D":
        with torch.no_grad():                                                # :106-110 G.eval()
            fake, part, aux = g_forward(sd_g, spec, audio, labels, y, style, training=False, **kw)
        fs = d_forward(sd_d, velocity(fake).detach(), True, log_d)
        fake_d = lambda_D * l1_mean(fs, torch.zeros_like(fs))
        rs = d_forward(sd_d, velocity(y), True, log_d)
        real_d = l1_mean(rs, torch.ones_like(rs))
        aux["fake_score"], aux["real_score"] = fs, rs
        return fake, [real_d, fake_d] + part, aux
    if step == "eval":
        fake, part, aux = g_forward(sd_g, spec, audio, labels, y, style, training=False,
                                    sample_flag=0, description="dev")
        return fake, [l1_mean(fake, y), torch.zeros(())] + part, aux
    raise ValueError(step)


# --------------------------------------------------------------------------
# next-row helpers: clip_grad_norm_ + Adam
# --------------------------------------------------------------------------
def clip_and_adam(params: List[torch.Tensor], grads: List[torch.Tensor], m, v, step: int,
                  lr=1e-4, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0):
    """trainer.py:1138-1146: clip_grad_norm_(params, 1) then Adam(lr).step().
    Pure-tensor restatement; updates params/m/v in place and returns the pre-clip norm."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads))
    coef = min(1.0, max_norm / (float(total) + 1e-6))
    b1, b2 = betas
    for p, g, mi, vi in zip(params, grads, m, v):
        g = g * coef
        mi.mul_(b1).add_(g, alpha=1 - b1)
        vi.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (vi / (1 - b2 ** step)).sqrt() + eps
        p.sub_(lr * (mi / (1 - b1 ** step)) / denom)
    return total



# --------------------------------------------------------------------------
# Speech2Gesture_G baseline
# --------------------------------------------------------------------------
def s2g_decoder_table(c=256):
    """speech2gesture.py:23-27: four k3 s1 ConvNormRelu blocks, no groups."""
    return [("decoder.%d" % i, c, c, 3, 1, 1, 1) for i in range(4)]


def s2g_state_shapes(out_feats: int, c: int = 256) -> Dict[str, Tuple[Tuple[int, ...], str]]:
    """state_dict of the reference Speech2Gesture_G (speech2gesture.py:20-28): audio_encoder, unet, decoder.{0-3},
    logits (out_feats, 256, 1)."""
    out: Dict[str, Tuple[Tuple[int, ...], str]] = {}

    def block(name, ci, co, k):
        ks = tuple(k) if isinstance(k, tuple) else (k,)
        out[name + ".conv.weight"] = ((co, ci) + ks, "conv_w")
        out[name + ".conv.bias"] = ((co,), "conv_b:" + str(ci * _prod(ks)))
        out[name + ".norm.weight"] = ((co,), "bn_w")
        out[name + ".norm.bias"] = ((co,), "bn_b")
        out[name + ".norm.running_mean"] = ((co,), "bn_rm")
        out[name + ".norm.running_var"] = ((co,), "bn_rv")
        out[name + ".norm.num_batches_tracked"] = ((), "nbt")

    for (n, ci, co, k, s, p, g) in audio_encoder_table() + unet_table(c) + s2g_decoder_table(c):
        block(n, ci, co, k)
    out["logits.weight"] = ((out_feats, c, 1), "conv_w")
    out["logits.bias"] = ((out_feats,), "conv_b:%d" % c)
    return out


def s2g_forward(sd, audio, time_steps, training, log: Optional[BNLog] = None):
    """Speech2Gesture_G.forward, speech2gesture.py:30-40: audio (B,T,F) -> unsqueeze(1) -> audio_encoder -> unet ->
    decoder -> 1x1 logits -> transpose.  Returns (pose (B,T,P), [])."""
    x = audio.unsqueeze(1) if audio.dim() == 3 else audio
    x = audio_encoder(x, sd, time_steps, training, log)
    x = unet1d(x, sd, training, log)
    x = run_table(x, sd, s2g_decoder_table(), training, log)
    x = F.conv1d(x, sd["logits.weight"], sd["logits.bias"])
    return x.transpose(-1, -2), []



# --------------------------------------------------------------------------
# Pose preprocessing in front of the hot path, src/data/transform.py
# --------------------------------------------------------------------------
def remove_joints(x, mask, num_joints=52):
    """RemoveJoints.__call__ (transform.py:499-508): view (B,T,2,J), drop the joints in `mask` along the last dimension,
    flatten back.  The slicing itself is pycasper.torchUtils.remove_slices (un-vendored, parity unpinned): taken to remove
    exactly the listed indices and keep the order of the rest."""
    B, T, _ = x.shape
    keep = [j for j in range(num_joints) if j not in set(mask)]
    return x.reshape(B, T, 2, num_joints)[..., keep].reshape(B, T, -1)


def znorm(x, mean, var, eps=1e-8):
    """ZNorm.znorm (transform.py:221-226)."""
    mask_std = (var >= 0).to(torch.double)
    std = (var * mask_std) ** 0.5
    mask = (std == 0).to(torch.double)
    std = (mask * eps) + (1 - mask) * std
    return (x - mean) / std


def inv_znorm(x, mean, var):
    """ZNorm.inv_znorm (transform.py:228-229)."""
    return x * (var ** 0.5) + mean


def kmeans_feats(x, feats):
    """KMeans.get_feats (transform.py:352-380), 'spatial' excluded."""
    out = []
    for feat in feats:
        if feat == "pose":
            out.append(x)
        elif feat == "velocity":
            v = torch.zeros_like(x)
            v[:, 1:, :] = x[:, 1:] - x[:, :-1]
            out.append(v)
        elif feat == "speed":
            v = torch.zeros_like(x)
            v[:, 1:, :] = x[:, 1:] - x[:, :-1]
            v = v.reshape(v.shape[0], v.shape[1], 2, -1)
            out.append((v ** 2).sum(dim=-2) ** 0.5)
        elif feat == "acceleration":
            v = torch.zeros_like(x)
            v[:, 1:, :] = x[:, 1:] - x[:, :-1]
            a = torch.zeros_like(x)
            a[:, 1:, :] = v[:, 1:] - v[:, :-1]
            out.append(a)
        else:
            raise KeyError(feat)
    return torch.cat(out, dim=-1)


def kmeans_predict(x, centers, feats, soft_labels=False):
    """KMeans.predict (transform.py:392-407): squared distances of the frame features to the centres; hard labels = index
    of the minimum, soft labels = softmax(-mse / mean(mse))."""
    f = kmeans_feats(x.double(), feats)
    shp = list(f.shape)
    f = f.view(-1, 1, shp[-1])
    mse = ((centers.view(1, *centers.shape) - f) ** 2).sum(dim=-1)
    if soft_labels:
        return F.softmax(-mse / mse.mean(-1).unsqueeze(-1), dim=-1).view(shp[:-1] + [centers.shape[0]])
    return mse.min(dim=-1)[1].view(shp[:-1])


def synth_prep(B, T, K, feats, seed=31, num_joints=52, mask=(0, 7, 8, 9)):
    """Synthetic raw pose batch, ZNorm statistics (one zero and one negative variance exercise the eps rules) and k-means
    centres (features of random frames of a second batch, so that nearest-centre margins are realistic)."""
    g = torch.Generator().manual_seed(seed)
    Pr = 2 * num_joints
    scale = 20.0 + 100.0 * torch.rand(Pr, generator=g, dtype=torch.float64)
    off = 50.0 * torch.randn(Pr, generator=g, dtype=torch.float64)
    walk = torch.cumsum(torch.randn(B, T, Pr, generator=g, dtype=torch.float64) * 0.05, dim=1)
    x = (torch.randn(B, 1, Pr, generator=g, dtype=torch.float64) + walk) * scale + off
    mean = off + torch.randn(Pr, generator=g, dtype=torch.float64)
    var = scale ** 2 * (0.5 + torch.rand(Pr, generator=g, dtype=torch.float64))
    var[3] = 0.0
    var[60] = -1.0
    walk2 = torch.cumsum(torch.randn(K, 4, Pr, generator=g, dtype=torch.float64) * 0.05, dim=1)
    x2 = (torch.randn(K, 1, Pr, generator=g, dtype=torch.float64) + walk2) * scale + off
    centers = kmeans_feats(remove_joints(x2, mask, num_joints), feats)[:, -1, :].contiguous()
    return x, mean, var, centers



def synth_metric_batches(seed=41, num_joints=52):
    """Two batches (B=3 and B=2, T=16) of normalised full-width poses: ground truth and a noisy prediction."""
    g = torch.Generator().manual_seed(seed)
    Pr = 2 * num_joints
    mean = 30.0 * torch.randn(Pr, generator=g, dtype=torch.float64)
    var = (20.0 + 60.0 * torch.rand(Pr, generator=g, dtype=torch.float64)) ** 2
    out = []
    for B in (3, 2):
        gt = torch.randn(B, 1, Pr, generator=g, dtype=torch.float64) + torch.cumsum(
            torch.randn(B, 16, Pr, generator=g, dtype=torch.float64) * 0.1, dim=1)
        noise = torch.randn(B, 16, Pr, generator=g, dtype=torch.float64) * torch.rand(1, 1, Pr, generator=g, dtype=torch.float64) * 0.6
        out.append((gt + noise, gt))
    return mean, var, out


def pose_metrics(batches, mean, var, mask, alphas=(0.1, 0.2), num_joints=52, desc="test"):
    """L1 / VelL1 / PCK as TrainerBase.calculate_metrics drives them (trainer.py:865-907; metrics.py:94-131, 247-303),
    with AverageMeter's weighting (metrics.py:36-62)."""
    J = num_joints
    keep = [j for j in range(J) if j not in set(mask)]
    l1_s = vel_s = 0.0
    n = 0
    pj = torch.zeros(len(alphas), J, dtype=torch.float64)
    pa = torch.zeros(len(alphas), dtype=torch.float64)
    pn = pan = 0
    ps, pc = 0.0, 0
    for y, gt in batches:
        B, T, _ = y.shape
        y4, g4 = y.view(B, T, 2, J), gt.view(B, T, 2, J)
        l1_s += float(F.l1_loss(y4[..., keep], g4[..., keep])) * B
        vel_s += float(F.l1_loss((y4[:, 1:] - y4[:, :-1])[..., keep], (g4[:, 1:] - g4[:, :-1])[..., keep])) * B
        n += B
        yu = inv_znorm(y, mean.view(1, 1, -1), var.view(1, 1, -1)).reshape(-1, 2, J).clone()
        gu = inv_znorm(gt, mean.view(1, 1, -1), var.view(1, 1, -1)).reshape(-1, 2, J).clone()
        yu[..., 0] = 0
        gu[..., 0] = 0
        dist = ((yu - gu) ** 2).sum(dim=1) ** 0.5
        h = gu[:, 0, :].max(dim=-1).values - gu[:, 0, :].min(dim=-1).values
        w = gu[:, 1, :].max(dim=-1).values - gu[:, 1, :].min(dim=-1).values
        Fr = B * T
        for a, al in enumerate(alphas):
            thresh = al * torch.max(torch.stack([h, w], dim=-1), dim=-1, keepdim=True).values
            pck = (dist < thresh).to(torch.float)
            pj[a] += pck.mean(dim=0).double() * Fr
            pa[a] += float(pck[:, keep].mean()) * Fr * len(keep)
        pn += Fr
        pan += Fr * len(keep)
        for a in range(len(alphas)):
            ps += float(pa[a] / pan) * Fr * len(keep)
            pc += Fr * len(keep)
    out = {"%s_L1" % desc: l1_s / n, "%s_VelL1" % desc: vel_s / n}
    for a, al in enumerate(alphas):
        for j in range(J):
            out["%s_pck_%s_%d" % (desc, al, j)] = float(pj[a, j] / pn)
        out["%s_pck_%s" % (desc, al)] = float(pa[a] / pan)
    out["%s_pck" % desc] = ps / pc
    return out


def flops_per_sequence(spec: Spec, T: int, train_description: bool) -> float:
    """Algorithmic conv FLOPs (2*MAC) of one forward per sequence, SURVEY.md §8d /
    Appendix A closed forms: MAC = L_out * C_out_total * (C_in/groups) * k."""
    mac = 0.0
    H, W = T, spec.mel_bins
    for (n, ci, co, k, s, p, g) in audio_encoder_table():
        H = (H + 2 * p[0] - k[0]) // s[0] + 1
        W = (W + 2 * p[1] - k[1]) // s[1] + 1
        mac += H * W * co * ci * k[0] * k[1]
    def seq(table, L):
        m = 0.0
        for (n, ci, co, k, s, p, g) in table:
            L = (L + 2 * p - k) // s + 1
            m += L * co * g * ci * k
        return m
    # unet: explicit because of the down/up lengths
    c = spec.in_channels
    mac += 2 * T * c * c * 3
    L = T
    for i in range(5):
        L //= 2
        mac += L * c * c * 4
    for i in range(5):
        L *= 2
        mac += L * c * c * 3
    mac += seq(classify_table(c + spec.style_dim), T) + T * spec.num_clusters * 256
    mac += seq(decoder_table(spec), T) + T * spec.num_clusters * spec.out_feats * c
    if train_description:
        mac += 2 * seq(pose_style_encoder_table(spec.out_feats, spec.num_speakers), T)
        P = spec.out_feats
        L1 = T // 2
        L2 = L1 // 2
        L3 = L2 - 1
        L4 = L3 - 3
        mac += L1 * 64 * P * 4 + L2 * 128 * 64 * 4 + L3 * 256 * 128 * 4 + L4 * 256 * 4
    return 2.0 * mac
