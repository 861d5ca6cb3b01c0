"""Tensor-level wrappers and autograd nodes over the C-ABI kernels.

Every function here takes CUDA tensors, allocates outputs with torch (device memory
plumbing only) and launches hand-written kernels from libmixstage_b200.so on the current
torch stream.  Activations are channels-last fp32: (B, L, C) or (B, H, W, C).
There is no CPU path: CPU tensors raise MixStageError."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ConvDesc, MixStageError, call, dt_code, ptr, stream

LEAKY_SLOPE = 0.2


def _need_cuda(t):
    if not t.is_cuda:
        raise MixStageError("mixstage_b200 runs on CUDA tensors only (got %s); there is no CPU fallback" % t.device)


def _f32c(t):
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise MixStageError("internal: expected contiguous fp32 tensor, got %s %s" % (t.dtype, t.stride()))
    return t


def conv_out(n, k, s, p):
    return (n + 2 * p - k) // s + 1


def make_desc(x_shape, Cout, kh, kw, sh, sw, ph, pw, groups):
    B, H, W, Cin = x_shape
    return ConvDesc(B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups, conv_out(H, kh, sh, ph), conv_out(W, kw, sw, pw))


# ---------------------------------------------------------------------------- casts
def cast_raw(src, dtype):
    _need_cuda(src)
    src = src.contiguous()
    dst = torch.empty(src.shape, dtype=dtype, device=src.device)
    if src.numel():
        call("ms_cast", ptr(src), dt_code(src.dtype), ptr(dst), dt_code(dtype), src.numel(), stream())
    return dst


class _Cast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src_dtype = x.dtype
        return cast_raw(x, dtype)

    @staticmethod
    def backward(ctx, g):
        return cast_raw(g, ctx.src_dtype), None


def cast(x, dtype):
    if x.dtype == dtype:
        return x
    return _Cast.apply(x, dtype)


# ---------------------------------------------------------------------------- conv block
class PackedWeight:
    """fp32 re-tiled copies of a conv weight, refreshed when the parameter changes."""

    def __init__(self):
        self.key = None
        self.wf = None
        self.wt = None
        self.bias = None
        self.bias_key = None

    def get(self, weight, desc):
        key = (weight.data_ptr(), weight._version, weight.dtype, weight.device)
        if key != self.key:
            w = weight.detach()
            if not w.is_contiguous():
                w = w.contiguous()
            n = w.numel()
            if self.wf is None or self.wf.numel() != n or self.wf.device != w.device:
                self.wf = torch.empty(n, dtype=torch.float32, device=w.device)
                self.wt = torch.empty(n, dtype=torch.float32, device=w.device)
            call("ms_pack_conv_weight_f32", ptr(w), dt_code(w.dtype), desc, ptr(self.wf), ptr(self.wt), stream())
            self.key = key
        return self.wf, self.wt

    def get_bias(self, bias):
        if bias is None:
            return None
        key = (bias.data_ptr(), bias._version, bias.dtype, bias.device)
        if key != self.bias_key:
            self.bias = cast_raw(bias.detach(), torch.float32)
            self.bias_key = key
        return self.bias


class ConvCfg:
    """Static geometry + behaviour of one conv block (mirrors ConvNormRelu's ctor, layers.py:32-76)."""

    def __init__(self, kh, kw, sh, sw, ph, pw, groups, slope, has_bn, act):
        self.kh, self.kw, self.sh, self.sw, self.ph, self.pw = kh, kw, sh, sw, ph, pw
        self.groups = groups
        self.slope = slope          # LeakyReLU slope (0 = ReLU)
        self.has_bn = has_bn
        self.act = act              # activation present
        self.momentum = 0.1
        self.eps = 1e-5


class _ConvBlock(torch.autograd.Function):
    """y = act(bn(conv(x) + b)) [+ upsample2(y) + residual].  x: (B,H,W,Cin) fp32 channels-last."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, residual, cfg, packed, bn_buffers, training, up2):
        _need_cuda(x)
        x = _f32c(x)
        B, H, W, Cin = x.shape
        Cout = weight.shape[0]
        desc = make_desc(x.shape, Cout, cfg.kh, cfg.kw, cfg.sh, cfg.sw, cfg.ph, cfg.pw, cfg.groups)
        if weight.shape[1] * cfg.groups != Cin:
            raise MixStageError("conv: input has %d channels, weight expects %d" % (Cin, weight.shape[1] * cfg.groups))
        wf, wt = packed.get(weight, desc)
        b32 = packed.get_bias(bias)
        st = stream()
        dev = x.device
        rows = B * desc.Ho * desc.Wo
        z = torch.empty((B, desc.Ho, desc.Wo, Cout), dtype=torch.float32, device=dev)
        fuse_act = (not cfg.has_bn) and cfg.act
        call("ms_conv_fwd_f32", ptr(x), ptr(wf), ptr(b32), ptr(z), desc, 1 if fuse_act else 0, cfg.slope, st)
        ctx.cfg, ctx.desc, ctx.wt, ctx.training, ctx.up2 = cfg, desc, wt, training, up2
        ctx.rows, ctx.Cout = rows, Cout
        ctx.param_dtypes = (weight.dtype, None if bias is None else bias.dtype)
        ctx.has_res = residual is not None
        if not cfg.has_bn:
            ctx.save_for_backward(x, z)
            return z
        rm, rv, nbt = bn_buffers
        stats = torch.zeros(2, Cout, dtype=torch.float64, device=dev) if training else None
        if training:
            call("ms_col_stats_f32", ptr(z), rows, Cout, ptr(stats[0]), ptr(stats[1]), st)
        ss = torch.empty(4, Cout, dtype=torch.float32, device=dev)     # scale, shift, mean, rstd
        call("ms_bn_finalize", ptr(stats[0]) if training else None, ptr(stats[1]) if training else None, rows, Cout,
             ptr(gamma), ptr(beta), ptr(rm), ptr(rv), dt_code(gamma.dtype), 1 if training else 0,
             cfg.momentum, cfg.eps, ptr(ss[0]), ptr(ss[1]), ptr(ss[2]), ptr(ss[3]), st)
        if training:
            nbt.add_(1)
        slope = cfg.slope if cfg.act else 1.0
        if up2:
            if desc.Ho != 1:
                raise MixStageError("upsample+skip fusion is 1-D only")
            y = torch.empty((B, 1, 2 * desc.Wo, Cout), dtype=torch.float32, device=dev)
            res = _f32c(residual)
            if res.shape != y.shape:
                raise MixStageError("skip tensor shape %s != %s" % (tuple(res.shape), tuple(y.shape)))
        else:
            y = torch.empty_like(z)
            res = None
        call("ms_bn_act_fwd_f32", ptr(z), ptr(ss[0]), ptr(ss[1]), slope, rows, Cout, ptr(y), ptr(res),
             1 if up2 else 0, desc.Wo, st)
        ctx.gamma_dtype = gamma.dtype
        ctx.save_for_backward(x, z, ss)
        return y

    @staticmethod
    def backward(ctx, dy):
        cfg, desc, st = ctx.cfg, ctx.desc, stream()
        rows, Cout = ctx.rows, ctx.Cout
        dy = dy.contiguous()
        dev = dy.device
        need_x, need_w, need_b, need_g, need_be, need_res = ctx.needs_input_grad[:6]
        dgamma = dbeta = dres = dbias = dw = dx = None
        if cfg.has_bn:
            x, z, ss = ctx.saved_tensors
            slope = cfg.slope if cfg.act else 1.0
            red = torch.zeros(2, Cout, dtype=torch.float64, device=dev)
            if ctx.training or need_g or need_be:
                call("ms_bn_act_bwd_reduce_f32", ptr(dy), ptr(z), ptr(ss[0]), ptr(ss[1]), ptr(ss[2]), ptr(ss[3]), slope,
                     rows, Cout, 1 if ctx.up2 else 0, desc.Wo, ptr(red[0]), ptr(red[1]), st)
            dz = torch.empty_like(z)
            call("ms_bn_act_bwd_apply_f32", ptr(dy), ptr(z), ptr(ss[0]), ptr(ss[1]), ptr(ss[2]), ptr(ss[3]), slope,
                 rows, Cout, 1 if ctx.up2 else 0, desc.Wo, ptr(red[0]), ptr(red[1]), 1 if ctx.training else 0, ptr(dz), st)
            if need_g:
                dgamma = torch.empty(Cout, dtype=ctx.gamma_dtype, device=dev)
                call("ms_store_param_grad", ptr(red[0]), Cout, ptr(dgamma), dt_code(ctx.gamma_dtype), st)
            if need_be:
                dbeta = torch.empty(Cout, dtype=ctx.gamma_dtype, device=dev)
                call("ms_store_param_grad", ptr(red[1]), Cout, ptr(dbeta), dt_code(ctx.gamma_dtype), st)
            if ctx.has_res and need_res:
                dres = dy
        else:
            x, z = ctx.saved_tensors
            if cfg.act:
                dz = torch.empty_like(z)
                call("ms_lrelu_bwd_f32", ptr(dy), ptr(z), cfg.slope, z.numel(), ptr(dz), st)
            else:
                dz = dy
        wdt, bdt = ctx.param_dtypes
        if need_b and bdt is not None:
            if cfg.has_bn and ctx.training:
                # batch-stat BN removes any per-channel constant: d(bias) is identically zero
                dbias = torch.zeros(Cout, dtype=bdt, device=dev)
            else:
                acc = torch.zeros(Cout, dtype=torch.float64, device=dev)
                call("ms_col_stats_f32", ptr(dz), rows, Cout, ptr(acc), None, st)
                dbias = torch.empty(Cout, dtype=bdt, device=dev)
                call("ms_store_param_grad", ptr(acc), Cout, ptr(dbias), dt_code(bdt), st)
        if need_w:
            n = Cout * (desc.Cin // desc.groups) * desc.kh * desc.kw
            dwf = torch.empty(n, dtype=torch.float32, device=dev)
            call("ms_conv_wgrad_f32", ptr(x), ptr(dz), ptr(dwf), desc, st)
            dw = torch.empty((Cout, desc.Cin // desc.groups, desc.kh, desc.kw), dtype=wdt, device=dev)
            call("ms_unpack_conv_wgrad", ptr(dwf), desc, ptr(dw), dt_code(wdt), st)
        if need_x:
            dx = torch.empty_like(x)
            call("ms_conv_dgrad_f32", ptr(dz), ptr(ctx.wt), ptr(dx), desc, st)
        return dx, dw, dbias, dgamma, dbeta, dres, None, None, None, None, None


def conv_block(x, weight, bias, gamma, beta, cfg, packed, bn_buffers, training, residual=None, up2=False):
    """weight is the caller's parameter in its own layout/dtype: (Cout, Cin/g, k) or (Cout, Cin/g, kh, kw)."""
    if weight.dim() == 3:
        weight = weight.unsqueeze(2)        # view: (Cout, Cin/g, 1, k); autograd maps the grad back
    return _ConvBlock.apply(x, weight, bias, gamma, beta, residual, cfg, packed, bn_buffers, training, up2)


# ---------------------------------------------------------------------------- bilinear time resize
class _BilinearT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, T):
        x = _f32c(x)
        B, Hi, Wi, C = x.shape
        y = torch.empty((B, 1, T, C), dtype=torch.float32, device=x.device)
        call("ms_bilinear_to_T_fwd_f32", ptr(x), B, Hi, Wi, C, T, ptr(y), stream())
        ctx.shape = (B, Hi, Wi, C, T)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, Hi, Wi, C, T = ctx.shape
        dy = dy.contiguous()
        dx = torch.empty((B, Hi, Wi, C), dtype=torch.float32, device=dy.device)
        call("ms_bilinear_to_T_bwd_f32", ptr(dy), B, Hi, Wi, C, T, ptr(dx), stream())
        return dx, None


def bilinear_to_T(x, T):
    """(B,Hi,Wi,C) -> (B,1,T,C): F.interpolate(size=(T,1), mode='bilinear') + squeeze (layers.py:197-198)."""
    return _BilinearT.apply(x, T)


# ---------------------------------------------------------------------------- style embedding + concat
class _StyleConcat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx, soft, emb, rep):
        x = _f32c(x)
        rows, C = x.numel() // x.shape[-1], x.shape[-1]
        S, sd = emb.shape
        out = torch.empty(x.shape[:-1] + (C + sd,), dtype=torch.float32, device=x.device)
        e = emb.detach().contiguous()
        if idx is not None:
            idx = idx.contiguous()
            if idx.dtype != torch.int64:
                raise MixStageError("style index must be int64")
            if idx.numel() * rep != rows:
                raise MixStageError("style index count %d * %d != rows %d" % (idx.numel(), rep, rows))
        else:
            soft = _f32c(soft)
            if soft.numel() // S * rep != rows:
                raise MixStageError("soft style rows mismatch")
        call("ms_style_concat_fwd_f32", ptr(x), rows, C, ptr(idx), ptr(soft), rep, ptr(e), dt_code(e.dtype), S, sd,
             ptr(out), stream())
        ctx.save_for_backward(idx, soft, e)
        ctx.meta = (rows, C, S, sd, rep, x.shape, emb.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        idx, soft, e = ctx.saved_tensors
        rows, C, S, sd, rep, xshape, edt = ctx.meta
        dout = dout.contiguous()
        dev = dout.device
        dx = torch.empty(xshape, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        demb32 = torch.zeros((S, sd), dtype=torch.float32, device=dev) if ctx.needs_input_grad[3] else None
        dsoft = torch.empty_like(soft) if (soft is not None and ctx.needs_input_grad[2]) else None
        call("ms_style_concat_bwd_f32", ptr(dout), rows, C, ptr(idx), ptr(soft), rep, ptr(e), dt_code(e.dtype), S, sd,
             ptr(dx), ptr(demb32), ptr(dsoft), stream())
        demb = None if demb32 is None else cast_raw(demb32, edt)
        return dx, None, dsoft, demb, None


def style_concat(x, emb_weight, idx=None, soft=None, rep=1):
    """cat([x, style_emb(pose_style)], -1) (jlcss.py:175-180).  idx int64 ('emb') or soft fp32 ('lin');
    one style row per `rep` consecutive rows of x."""
    return _StyleConcat.apply(x, idx, soft, emb_weight, rep)


# ---------------------------------------------------------------------------- softmax + CE + argmax
class _SoftmaxCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, score, target, trep):
        score = _f32c(score)
        K = score.shape[-1]
        rows = score.numel() // K
        dev = score.device
        soft = torch.empty_like(score)
        amax = torch.empty(score.shape[:-1], dtype=torch.int64, device=dev)
        st = stream()
        if target is not None:
            target = target.contiguous()
            if target.dtype != torch.int64 or target.numel() * trep != rows:
                raise MixStageError("CE target must be int64 with rows/trep entries")
            acc = torch.zeros(1, dtype=torch.float64, device=dev)
            call("ms_softmax_ce_fwd_f32", ptr(score), rows, K, ptr(target), trep, ptr(soft), ptr(amax), ptr(acc), st)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            call("ms_scalar_finish", ptr(acc), 1.0 / rows, ptr(loss), st)
        else:
            call("ms_softmax_ce_fwd_f32", ptr(score), rows, K, None, 1, ptr(soft), ptr(amax), None, st)
            loss = torch.zeros((), dtype=torch.float32, device=dev)
        ctx.save_for_backward(soft, target)
        ctx.meta = (rows, K, trep)
        ctx.mark_non_differentiable(amax)
        return soft, loss, amax

    @staticmethod
    def backward(ctx, dsoft, dloss, _damax):
        soft, target = ctx.saved_tensors
        rows, K, trep = ctx.meta
        dscore = torch.empty_like(soft)
        g = None
        if target is not None and dloss is not None:
            g = dloss.to(torch.float32).contiguous()
        ds = None
        if dsoft is not None:
            ds = dsoft.contiguous()
        call("ms_softmax_ce_bwd_f32", ptr(soft), rows, K, ptr(target), trep, ptr(g), ptr(ds), ptr(dscore), stream())
        return dscore, None, None


def softmax_ce(score, target=None, trep=1):
    """Returns (softmax(score), mean CE vs target, argmax).  jlcss.py:183-187 / :159-165 / :203."""
    return _SoftmaxCE.apply(score, target, trep)


# ---------------------------------------------------------------------------- mixture
class _Mixture(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, w):
        z, w = _f32c(z), _f32c(w)
        K = w.shape[-1]
        rows = w.numel() // K
        P = z.shape[-1] // K
        if z.numel() != rows * K * P:
            raise MixStageError("mixture: z %s vs weights %s" % (tuple(z.shape), tuple(w.shape)))
        out = torch.empty(w.shape[:-1] + (P,), dtype=torch.float32, device=z.device)
        call("ms_mixture_fwd_f32", ptr(z), ptr(w), rows, K, P, ptr(out), stream())
        ctx.save_for_backward(z, w)
        ctx.meta = (rows, K, P)
        return out

    @staticmethod
    def backward(ctx, dout):
        z, w = ctx.saved_tensors
        rows, K, P = ctx.meta
        dout = dout.contiguous()
        dz = torch.empty_like(z)
        dw = torch.empty_like(w)
        call("ms_mixture_bwd_f32", ptr(dout), ptr(z), ptr(w), rows, K, P, ptr(dz), ptr(dw), stream())
        return dz, dw


def mixture(z, w):
    """index_select_outputs (jlcss.py:106-115): z (...,K*P), w (...,K) -> (...,P)."""
    return _Mixture.apply(z, w)


class _MeanRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _f32c(x)
        B, L, C = x.shape
        y = torch.empty((B, C), dtype=torch.float32, device=x.device)
        call("ms_mean_rows_fwd_f32", ptr(x), B, L, C, ptr(y), stream())
        ctx.meta = (B, L, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, L, C = ctx.meta
        dx = torch.empty((B, L, C), dtype=torch.float32, device=dy.device)
        call("ms_mean_rows_bwd_f32", ptr(dy.contiguous()), B, L, C, ptr(dx), stream())
        return dx


def mean_rows(x):
    return _MeanRows.apply(x)


# ---------------------------------------------------------------------------- GAN pieces
class _Velocity(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _f32c(x)
        B, T, P = x.shape
        v = torch.empty_like(x)
        call("ms_velocity_fwd_f32", ptr(x), B, T, P, ptr(v), stream())
        ctx.meta = (B, T, P)
        return v

    @staticmethod
    def backward(ctx, dv):
        B, T, P = ctx.meta
        dx = torch.empty((B, T, P), dtype=torch.float32, device=dv.device)
        call("ms_velocity_bwd_f32", ptr(dv.contiguous()), B, T, P, ptr(dx), stream())
        return dx


def velocity(x):
    """GAN.get_velocity, joint=False (gan.py:47-52)."""
    return _Velocity.apply(x)


class _L1Mean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, const):
        a = _f32c(a)
        if b is not None:
            b = _f32c(b)
            if b.shape != a.shape:
                raise MixStageError("l1: shape mismatch")
        n = a.numel()
        dev = a.device
        acc = torch.zeros(1, dtype=torch.float64, device=dev)
        sgn = torch.empty_like(a)
        st = stream()
        call("ms_l1_fwd_f32", ptr(a), ptr(b), float(const), n, ptr(acc), ptr(sgn), st)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        call("ms_scalar_finish", ptr(acc), 1.0 / n, ptr(loss), st)
        ctx.save_for_backward(sgn)
        ctx.n = n
        return loss

    @staticmethod
    def backward(ctx, g):
        (sgn,) = ctx.saved_tensors
        da = torch.empty_like(sgn)
        call("ms_l1_bwd_f32", ptr(sgn), ptr(g.to(torch.float32).contiguous()), ctx.n, ptr(da), stream())
        db = -da if ctx.needs_input_grad[1] else None
        return da, db, None


def l1_mean(a, b=None, const=0.0):
    """mean |a - b| (b tensor) or mean |a - const|: L1Loss(reduction='none') + mean (gan.py:64-75)."""
    return _L1Mean.apply(a, b, const)
