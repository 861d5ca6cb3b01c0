"""Host-side mirror of the reference layer classes on the generator hot path
(reference: src/model/layers.py).  Same class names, constructor arguments, parameter
names and state_dict layout; the arithmetic runs in hand-written CUDA (ops.py).

Internal activation layout is channels-last fp32 (B, H, W, C) with H == 1 for the 1-D
stacks, so none of the reference's transposes exist here.  ``nn.Conv*``/``nn.BatchNorm*``
objects are used purely as parameter containers (identical names, shapes, default
initialisation and ``.double()`` behaviour); their forward is never called."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import MixStageError


def num_powers_of_two(x):
    n = 0
    while x > 1 and x % 2 == 0:
        x //= 2
        n += 1
    return n


class ConvNormRelu(nn.Module):
    """conv -> BatchNorm -> (Leaky)ReLU, reference layers.py:32-78.  Same ctor contract:
    default k3/s1, ``downsample`` k4/s2, padding (k-s)//2, channels multiplied by groups."""

    def __init__(self, in_channels, out_channels, type='1d', leaky=False, downsample=False,
                 kernel_size=None, stride=None, padding=None, p=0, groups=1):
        super().__init__()
        if p != 0:
            raise NotImplementedError("mixstage_b200: dropout p>0 is outside the accelerated path (reference jobs use p=0)")
        if kernel_size is None and stride is None:
            kernel_size, stride = (4, 2) if downsample else (3, 1)
        if padding is None:           # layers.py:46-55
            if isinstance(kernel_size, int) and isinstance(stride, tuple):
                padding = tuple(int((kernel_size - st) / 2) for st in stride)
            elif isinstance(kernel_size, tuple) and isinstance(stride, int):
                padding = tuple(int((ks - stride) / 2) for ks in kernel_size)
            elif isinstance(kernel_size, tuple) and isinstance(stride, tuple):
                padding = tuple(int((ks - st) / 2) for ks, st in zip(kernel_size, kernel_size))
            else:
                padding = int((kernel_size - stride) / 2)
        in_channels, out_channels = in_channels * groups, out_channels * groups
        if type == '1d':
            self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, stride, padding, groups=groups)
            self.norm = nn.BatchNorm1d(out_channels)
            kh, kw = 1, self.conv.kernel_size[0]
            sh, sw = 1, self.conv.stride[0]
            ph, pw = 0, self.conv.padding[0]
        elif type == '2d':
            self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, groups=groups)
            self.norm = nn.BatchNorm2d(out_channels)
            (kh, kw), (sh, sw), (ph, pw) = self.conv.kernel_size, self.conv.stride, self.conv.padding
        else:
            raise ValueError(type)
        self.cfg = ops.ConvCfg(kh, kw, sh, sw, ph, pw, groups, 0.2 if leaky else 0.0, has_bn=True, act=True)
        self._packed = ops.PackedWeight()

    def forward(self, x, residual=None, up2=False, want="both", row_w=None):
        """x: channels-last (B, H, W, C) fp32.  With ``up2`` the output is
        ``upsample2(act(bn(conv(x)))) + residual`` (UNet1D decoder step, layers.py:151).
        ``want`` ("planes" | "f32" | "both"): which forms of the output the consumers read; only the inference fast
        path (ops.conv_block) acts on it."""
        n = self.norm
        return ops.conv_block(x, self.conv.weight, self.conv.bias, n.weight, n.bias, self.cfg, self._packed,
                              (n.running_mean, n.running_var, n.num_batches_tracked), self.training,
                              residual=residual, up2=up2, want=want, row_w=row_w)


class PlainConv(object):
    """Helper (not a Module): runs an nn.Conv1d's parameters as conv (+ optional LeakyReLU)."""

    def __init__(self, conv, slope=None):
        k, s, p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        self.cfg = ops.ConvCfg(1, k, 1, s, 0, p, conv.groups, 0.0 if slope is None else slope, has_bn=False,
                               act=slope is not None)
        self.packed = ops.PackedWeight()

    def __call__(self, conv, x):
        return ops.conv_block(x, conv.weight, conv.bias, None, None, self.cfg, self.packed, None, False, want="f32")


def _chain_block(m):
    n = m.norm
    return ops.ChainBlock(m.conv.weight, m.conv.bias, n.weight, n.bias, m.cfg, m._packed,
                          (n.running_mean, n.running_var, n.num_batches_tracked))


def _run(blocks, x, last="f32", res_from=None):
    """A chain of ConvNormRelu blocks: every intermediate activation is only read by the next convolution (operand planes
    suffice); the last one is wanted as `last`.  Runs of blocks that qualify leave as ONE launch (ops.conv_chain)."""
    blocks = list(blocks)
    return ops.conv_chain([_chain_block(b) for b in blocks], x, blocks[0].training, res_from=res_from, last=last)


class UNet1D(nn.Module):
    """reference layers.py:80-157."""

    def __init__(self, input_channels, output_channels, max_depth=5, kernel_size=None, stride=None, p=0, groups=1):
        super().__init__()
        self.pre_downsampling_conv = nn.ModuleList([])
        self.conv1 = nn.ModuleList([])
        self.conv2 = nn.ModuleList([])
        self.max_depth = max_depth
        self.groups = groups
        kw = dict(type='1d', leaky=True, kernel_size=kernel_size, stride=stride, p=p, groups=groups)
        self.pre_downsampling_conv.append(ConvNormRelu(input_channels, output_channels, downsample=False, **kw))
        self.pre_downsampling_conv.append(ConvNormRelu(output_channels, output_channels, downsample=False, **kw))
        for _ in range(max_depth):
            self.conv1.append(ConvNormRelu(output_channels, output_channels, downsample=True, **kw))
        for _ in range(max_depth):
            self.conv2.append(ConvNormRelu(output_channels, output_channels, downsample=False, **kw))

    def forward(self, x):
        T = x.shape[2]
        assert T / (2 ** (self.max_depth - 1)) >= 1, \
            'Input size is {}. It must be >= {}'.format(T, 2 ** (self.max_depth - 1))
        assert num_powers_of_two(T) >= self.max_depth, \
            'Input size is {}. It must be a multiple of 2^(max_depth) = 2^{} = {}'.format(T, self.max_depth, 2 ** self.max_depth)
        # the whole UNet as one chain: pre (2), down (5), up (5).  The last down block and the first four up blocks produce
        # `upconv(x) + residual` directly (layers.py:150-151): block i with res_from[i] = j adds the output of block j
        d = self.max_depth
        blocks = list(self.pre_downsampling_conv) + list(self.conv1) + list(self.conv2)
        res_from = [None] * len(blocks)
        res_idx = [1] + [2 + i for i in range(d - 1)]          # residuals[0] = pre output, residuals[i + 1] = conv1[i] output
        res_from[2 + d - 1] = res_idx[d - 1]
        for i in range(d - 1):
            res_from[2 + d + i] = res_idx[d - i - 2]
        return _run(blocks, x, last="f32", res_from=res_from)


class AudioEncoder(nn.Module):
    """reference layers.py:159-199.  forward takes (B, T, F, 1) channels-last, returns (B, 1, T, 256)."""

    def __init__(self, output_feats=64, input_channels=1, kernel_size=None, stride=None, p=0, groups=1):
        super().__init__()
        kw = dict(type='2d', leaky=True, kernel_size=kernel_size, stride=stride, p=p, groups=groups)
        self.conv = nn.ModuleList([])
        for ci, co, down in [(input_channels, 64, False), (64, 64, True), (64, 128, False), (128, 128, True),
                             (128, 256, False), (256, 256, True), (256, 256, False)]:
            self.conv.append(ConvNormRelu(ci, co, downsample=down, **kw))
        self.conv.append(ConvNormRelu(256, 256, type='2d', leaky=True, downsample=False,
                                      kernel_size=(3, 8), stride=1, p=p, groups=groups))
        self.prune_eval_columns = True      # inference: only the output column the bilinear resize reads (see _pruned_last)
        self._alt = {}                      # input width -> (pruned ConvCfg, its PackedWeight)
        self.mid_mark = None                # set by the generator while a gradient-exchange hook is active (G._mark)

    def _pruned_last(self, W):
        """Eval only.  The bilinear resize to width 1 (layers.py:197) reads exactly the CENTRE output column of the last
        block (kernel (3, 8), padding (1, 3) on 8 mel columns -> 7 output columns, column 3 selected with weight 1: probe in
        SURVEY.md Appendix A note 4), so inference computes that column only: the same convolution with the width padding
        reduced until one output column is left -- 1/7 of the block's MACs, every tap still inside the input.  (Training
        needs all columns: they enter the BatchNorm statistics.)  Returns a ChainBlock or None when the geometry does not
        reduce to a single exact column."""
        m = self.conv[-1]
        c = m.cfg
        if c.sw != 1 or W > c.kw or (c.kw - W) % 2:
            return None
        pw = (c.kw - W) // 2
        wo_full = W + 2 * c.pw - c.kw + 1
        if pw > c.pw or wo_full % 2 == 0 or c.pw - pw != (wo_full - 1) // 2:
            return None
        alt = self._alt.get(W)
        if alt is None:
            cfg = ops.ConvCfg(c.kh, c.kw, c.sh, c.sw, c.ph, pw, c.groups, c.slope, has_bn=c.has_bn, act=c.act)
            alt = self._alt[W] = (cfg, ops.PackedWeight())
        n = m.norm
        return ops.ChainBlock(m.conv.weight, m.conv.bias, n.weight, n.bias, alt[0], alt[1],
                              (n.running_mean, n.running_var, n.num_batches_tracked))

    def forward(self, x, time_steps=None):
        if time_steps is None:
            time_steps = x.shape[1]
        blocks = [_chain_block(b) for b in self.conv]
        if not self.training and not torch.is_grad_enabled() and self.prune_eval_columns:
            # width of the last block's input: the mel axis halves at every stride-2 block
            W = x.shape[2]
            for b in list(self.conv)[:-1]:
                W = ops.conv_out(W, b.cfg.kw, b.cfg.sw, b.cfg.pw)
            last = self._pruned_last(W)
            if last is not None:
                blocks[-1] = last
        mark = self.mid_mark
        if mark is not None and self.training and torch.is_grad_enabled():
            # data-parallel training: the four late blocks hold 93 % of the encoder's weights; a hook on the activation
            # between the halves lets their gradients travel while the early (large-map) blocks still run backward
            x = mark(ops.conv_chain(blocks[:4], x, True, last="f32"))
            x = ops.conv_chain(blocks[4:], x, True, last="f32")
        else:
            x = ops.conv_chain(blocks, x, self.training, last="f32")
        return ops.bilinear_to_T(x, time_steps)


class _SeqEncoder(nn.Module):
    def __init__(self, input_channels, p=0, groups=1, kernel_size=None, stride=None):
        super().__init__()
        kw = dict(type='1d', leaky=True, downsample=False, kernel_size=kernel_size, stride=stride, p=p, groups=groups)
        chans = [input_channels, 64, 64, 128, 128, 256, 256]
        self.conv = nn.ModuleList([ConvNormRelu(chans[i], chans[i + 1], **kw) for i in range(6)])

    def forward(self, x, time_steps=None):
        return _run(self.conv, x)


class PoseEncoder(_SeqEncoder):
    """reference layers.py:201-240."""

    def __init__(self, output_feats=64, input_channels=96, kernel_size=None, stride=None, p=0, groups=1):
        super().__init__(input_channels, p=p, groups=groups, kernel_size=kernel_size, stride=stride)


class TextEncoder1D(_SeqEncoder):
    """reference layers.py:339-373.  Parameters exist for state_dict compatibility; the text
    modalities are outside the accelerated path and raise."""

    def __init__(self, output_feats=64, input_channels=300, kernel_size=None, stride=None, p=0, groups=1):
        super().__init__(input_channels, p=p, groups=groups, kernel_size=kernel_size, stride=stride)

    def forward(self, x, time_steps=None, **kwargs):
        raise NotImplementedError("mixstage_b200: text modalities are not on the accelerated path")


class PoseStyleEncoder(nn.Module):
    """reference layers.py:246-289.  (B,1,T,P) -> (B,S)."""

    def __init__(self, output_feats=64, input_channels=96, kernel_size=None, stride=None, p=0, groups=1, num_speakers=4):
        super().__init__()
        kw = dict(type='1d', leaky=True, kernel_size=kernel_size, stride=stride, p=p, groups=groups)
        chans = [input_channels, 64, 64, 128, 128, 256, 256, num_speakers]
        self.conv = nn.ModuleList([ConvNormRelu(chans[0], chans[1], downsample=False, **kw)])
        for i in range(1, 7):
            self.conv.append(ConvNormRelu(chans[i], chans[i + 1], downsample=True, **kw))

    def forward(self, x, time_steps=None):
        x = _run(self.conv, x)
        B, _, L, C = x.shape
        return ops.mean_rows(x.view(B, L, C))


class ClusterClassify(nn.Module):
    """reference layers.py:446-467.  (B,1,T,C_in) -> (B,1,T,num_clusters)."""

    def __init__(self, num_clusters=8, kernel_size=None, stride=None, p=0, groups=1, input_channels=256):
        super().__init__()
        kw = dict(type='1d', leaky=True, downsample=False, kernel_size=kernel_size, stride=stride, p=p, groups=groups)
        self.conv = nn.ModuleList([ConvNormRelu(input_channels, 256, **kw)])
        self.conv += nn.ModuleList([ConvNormRelu(256, 256, **kw) for _ in range(5)])
        self.logits = nn.Conv1d(256 * groups, num_clusters * groups, kernel_size=1, stride=1, groups=groups)
        self._logits = PlainConv(self.logits)

    def forward(self, x, time_steps=None):
        x = _run(self.conv, x)
        return self._logits(self.logits, x)


class EmbLin(nn.Module):
    """reference layers.py:652-663 -- parameter holder; the lookup is fused with the concat
    (ops.style_concat)."""

    def __init__(self, num_embeddings, embedding_dim):
        super().__init__()
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.emb = nn.Embedding(num_embeddings, embedding_dim)


class Group(nn.Module):
    """reference layers.py:593-650.  Constructed (aliases style_dec in the state_dict) but never
    called on the hot path."""

    def __init__(self, models, groups=1, dim=1):
        super().__init__()
        if not isinstance(models, list):
            models = [models]
        self.models = nn.ModuleList(models)
        self.groups = groups
        self.dim = dim

    def forward(self, *a, **k):
        raise NotImplementedError("mixstage_b200: Group is not on the accelerated path")


class Curriculum():
    """reference layers.py:677-696 (plain host object, not in the state_dict)."""

    def __init__(self, start, end, num_iters):
        self.start, self.end, self.num_iters = start, end, num_iters
        self.iters = 0
        self.diff = (end - start) / num_iters
        self.value = start

    def step(self, flag=True):
        if not flag:
            return self.value
        if self.iters < self.num_iters:
            v = self.value
            self.value += self.diff
            self.iters += 1
            return v
        return self.end
