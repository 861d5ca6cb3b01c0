"""Pose discriminator, mirror of reference src/model/speech2gesture.py:67-100."""
import torch
import torch.nn as nn

from . import ops
from .layers import ConvNormRelu, PlainConv
from ._lib import MixStageError


class Speech2Gesture_D(nn.Module):
    '''
    input_shape:  (N, time, pose_feats)
    output_shape: (N, *) ## discriminator scores
    '''

    def __init__(self, in_channels=104, out_channels=64, n_downsampling=2, p=0, groups=1, **kwargs):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv1d(in_channels * groups, out_channels * groups, 4, 2, padding=1, groups=groups),
                                   nn.LeakyReLU(negative_slope=0.2))
        conv2 = []
        ch_mul = 1
        for n in range(1, n_downsampling):
            ch_mul = min(2 ** n, 8)
            conv2.append(ConvNormRelu(out_channels, out_channels * ch_mul, type='1d', downsample=True, leaky=True,
                                      p=p, groups=groups))
        self.conv2 = nn.Sequential(*conv2)
        ch_mul_new = min(2 ** n_downsampling, 8)
        self.conv3 = ConvNormRelu(out_channels * ch_mul, out_channels * ch_mul_new, type='1d', leaky=True,
                                  kernel_size=4, stride=1, p=p, groups=groups)
        out_shape = 1 if 'out_shape' not in kwargs else kwargs['out_shape']
        self.logits = nn.Conv1d(out_channels * ch_mul_new * groups, out_shape * groups, kernel_size=4, stride=1,
                                groups=groups)
        self._conv1 = PlainConv(self.conv1[0], slope=0.2)
        self._logits = PlainConv(self.logits)
        self.precision = kwargs.get('precision', None)

    def forward(self, x):
        """x: (B, T, P) in the caller's dtype -> (scores (B, L'), [])."""
        with ops.precision_scope(self.precision):
            return self._forward(x)

    def _forward(self, x):
        ops._need_cuda(x)
        dtype = x.dtype
        B, T, P = x.shape
        h = ops.cast(x, torch.float32).contiguous().view(B, 1, T, P)
        h = self._conv1(self.conv1[0], h)
        for blk in self.conv2:
            h = blk(h)
        h = self.conv3(h)
        h = self._logits(self.logits, h)                  # (B,1,L',out_shape)
        out = h.reshape(B, h.shape[2], h.shape[3])
        if out.shape[-1] == 1:
            out = out.squeeze(-1)
        return ops.cast(out, dtype), []
