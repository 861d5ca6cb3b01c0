"""Speech2Gesture baseline generator and pose discriminator, mirrors of reference src/model/speech2gesture.py:13-100."""
import torch
import torch.nn as nn

from . import ops
from .layers import AudioEncoder, ConvNormRelu, PlainConv, UNet1D, _run
from ._lib import MixStageError


class Speech2Gesture_G(nn.Module):
    '''
    Baseline: http://people.eecs.berkeley.edu/~shiry/projects/speech2gesture/ (reference speech2gesture.py:13-40;
    SURVEY.md §8f row 4).  Same constructor, state_dict keys and forward(x, y, time_steps=None, **kwargs) as the reference;
    runs on the kernels of the Mix-StAGE generator (audio encoder, UNet, four ConvNormRelu blocks, 1x1 logits).

    input_shape:  (N, time, frequency)
    output_shape: (N, time, pose_feats)
    '''

    def __init__(self, time_steps=64, in_channels=256, out_feats=104, p=0, **kwargs):
        super().__init__()
        self.audio_encoder = AudioEncoder(output_feats=time_steps, p=p)
        self.unet = UNet1D(input_channels=in_channels, output_channels=in_channels, p=p)
        self.decoder = nn.Sequential(*[ConvNormRelu(in_channels, in_channels, type='1d', leaky=True, downsample=False, p=p)
                                       for _ in range(4)])
        self.logits = nn.Conv1d(in_channels, out_feats, kernel_size=1, stride=1)
        self._logits = PlainConv(self.logits)
        self.precision = kwargs.get('precision', None)

    def forward(self, x, y=None, time_steps=None, **kwargs):
        """x: audio (B,T,F) or (B,1,T,F) (a list [audio, ...] as the trainer passes it is accepted too) ->
        (pose (B,T,P) in the dtype of the input, [])."""
        with ops.precision_scope(self.precision):
            if isinstance(x, (list, tuple)):
                x = x[0]
            ops._need_cuda(x)
            dtype = x.dtype
            if x.dim() == 4:
                x = x.squeeze(1)
            B, T, Fm = x.shape
            a = ops.cast(x, torch.float32).contiguous().view(B, T, Fm, 1)          # NHWC with C = 1
            h = self.audio_encoder(a, time_steps if time_steps is not None else T)
            h = self.unet(h)
            h = _run(self.decoder, h, last="f32")
            z = self._logits(self.logits, h)                                        # (B,1,T,P)
            return ops.cast(z.view(B, z.shape[2], z.shape[3]), dtype), []


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
        h = _run(list(self.conv2) + [self.conv3], h, last="f32")
        h = self._logits(self.logits, h)                  # (B,1,L',out_shape)
        out = h.reshape(B, h.shape[2], h.shape[3])
        if out.shape[-1] == 1:
            out = out.squeeze(-1)
        return ops.cast(out, dtype), []
