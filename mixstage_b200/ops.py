"""Tensor-level wrappers and autograd nodes over the C-ABI kernels.

Every function here takes CUDA tensors, allocates outputs with torch (device memory
plumbing only) and launches hand-written kernels from libmixstage_b200.so on the current
torch stream.  Activations are channels-last fp32: (B, L, C) or (B, H, W, C).
There is no CPU path: CPU tensors raise MixStageError."""
from __future__ import annotations

import ctypes

import torch

from . import _lib, igemm
from ._lib import MS_BF16, MS_BF16X2, MS_F32, ConvDesc, MixStageError, call, dt_code, ptr, stream

LEAKY_SLOPE = 0.2

# ---------------------------------------------------------------------------- precision
# "fp32"   : every convolution on the exact-fp32 CUDA-core kernels (csrc/conv_simt.cu)
# "bf16x3" : tcgen05 tensor cores with split-bf16 operands (hi + lo planes, 3 MMA passes, fp32 accumulate):
#            ~16 mantissa bits per operand, meets the fp32 parity tolerance
# "bf16"   : tcgen05 tensor cores with plain bf16 operands, fp32 accumulate (the fast mode)
PRECISIONS = ("fp32", "bf16x3", "bf16")
_precision = "fp32"
FORCE_REPACK = False        # set while a CUDA graph is being captured: weights change without a version bump
_weight_epoch = 0            # bumped when parameters change without a torch version bump (graph replays)
_stats_epoch = 0             # bumped whenever a training-mode BatchNorm updates running statistics in place
last_gemm_flops = 0.0       # algorithmic FLOPs (2*MAC, unpadded) of the GEMM launch that follows; read by bench.py


def set_precision(p):
    global _precision
    if p not in PRECISIONS:
        raise MixStageError("precision must be one of %s" % (PRECISIONS,))
    _precision = p


def get_precision():
    return _precision


def bump_weight_epoch():
    """Invalidate every packed-weight cache: parameters were updated by a replayed CUDA graph."""
    global _weight_epoch
    _weight_epoch += 1


class precision_scope:
    """with precision_scope("bf16x3"): ...   (None keeps the current setting)"""

    def __init__(self, p):
        self.p, self.old = p, None

    def __enter__(self):
        if self.p is not None:
            self.old = get_precision()
            set_precision(self.p)

    def __exit__(self, *a):
        if self.old is not None:
            set_precision(self.old)


# ---------------------------------------------------------------------------- step-scoped scratch
class _Arena:
    """Zero-initialised fp64 accumulators (BatchNorm sums, reduction slots, tickets) carved from ONE buffer that the
    train step clears with a single memset, instead of one torch.zeros launch per layer.  Outside a step
    (`active` False) every request falls back to torch.zeros."""

    def __init__(self):
        self.buf, self.keep = None, []
        self.off = self.used = self.need = 0
        self.buf32 = None           # second pool: zero-filled fp32 (split-K outputs that k-slices reduce into)
        self.off32 = self.used32 = self.need32 = 0
        self.active = False

    def begin(self, device):
        if self.buf is None or self.buf.device != device or self.buf.numel() < self.need:
            if self.buf is not None:
                self.keep.append(self.buf)          # captured graphs may still point into the old buffer
            self.buf = torch.zeros(max(self.need, 1 << 15), dtype=torch.float64, device=device)
        else:
            self.buf.zero_()
        if self.need32:
            if self.buf32 is None or self.buf32.device != device or self.buf32.numel() < self.need32:
                if self.buf32 is not None:
                    self.keep.append(self.buf32)
                self.buf32 = torch.zeros(self.need32, dtype=torch.float32, device=device)
            else:
                self.buf32.zero_()
        self.off = self.used = 0
        self.off32 = self.used32 = 0
        self.active = True

    def end(self):
        self.need = max(self.need, self.used)
        self.need32 = max(self.need32, self.used32)
        self.active = False

    def take_f32(self, shape, device):
        """Zero-filled fp32 tensor (an output that split-K slices reduce into): from the per-step pool inside a train
        step (one memset per step for all of them), else torch.zeros."""
        n = 1
        for d_ in shape:
            n *= d_
        n_al = (n + 63) // 64 * 64             # 256-byte granularity: TMA / vector accesses stay aligned
        self.used32 += n_al
        if not self.active or self.buf32 is None or self.buf32.device != device or self.off32 + n_al > self.buf32.numel():
            return torch.zeros(shape, dtype=torch.float32, device=device)
        t = self.buf32[self.off32:self.off32 + n].view(shape)
        self.off32 += n_al
        return t

    def take(self, shape, device):
        n = 1
        for d_ in shape:
            n *= d_
        n_al = (n + 1) // 2 * 2
        self.used += n_al
        if not self.active or self.buf.device != device or self.off + n_al > self.buf.numel():
            return torch.zeros(shape, dtype=torch.float64, device=device)
        t = self.buf[self.off:self.off + n].view(shape)
        self.off += n_al
        return t


arena = _Arena()
# When True (set by TrainStep around its body) parameter gradients are accumulated by the kernels straight into the
# parameters' existing .grad buffers (views of the flat gradient buffer) and autograd receives None for them.
DIRECT_GRADS = False


class SideWork:
    """A second stream for the weight-gradient GEMMs of a train step: they depend only on (x planes, dz planes) and feed
    nothing but the flat gradient buffer, so they run beside the input-gradient chain instead of inside it (also when the
    step is captured into a CUDA graph: the fork/join becomes graph edges).  Tensors handed to the side stream are kept
    referenced until join() so the allocator cannot recycle them early."""

    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self.keep = []
        self.forked = False

    def fork(self, *tensors):
        self.stream.wait_stream(torch.cuda.current_stream())
        self.keep.extend(tensors)
        self.forked = True
        return torch.cuda.stream(self.stream)

    def join(self):
        if self.forked:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.keep.clear()
        self.forked = False


SIDE = None                  # SideWork of the running train step (set by TrainStep), else None

import os as _os
# Fused blocks (csrc/conv_train.cu): training-mode ConvNormRelu forward / backward as ONE cooperative launch each, and the
# small-batch inference block with split-K.  MS_FUSED_BLOCKS=0 keeps the three-kernel path (A/B timing, debugging).
FUSED_BLOCKS = _os.environ.get("MS_FUSED_BLOCKS", "1") != "0"


class WgradAccum:
    """Persistent fp32 weight-gradient accumulators of one kind of train step (ms_wgrad_bf16_acc adds every pixel slice's
    tile in place) and the list of (accumulator -> parameter gradient) conversions that ONE ms_unpack_wgrad_multi launch
    performs at the end of backward.  Accumulators are carved from 32 MB chunks and never move (captured graphs hold
    their addresses); the step zero-fills the chunks before backward."""
    CHUNK = 8 << 20            # elements

    def __init__(self):
        self.chunks, self.off = [], 0
        self.slots = {}         # id(PackedWeight) -> accumulator
        self.where = {}         # id(PackedWeight) -> (chunk index, begin, end) incl. alignment padding
        self.touched = []       # accumulators used so far in the running step, in order (data-parallel exchange buckets)
        self.entries = {}       # (acc ptr, sink ptr) -> WgradEntry fields

    def acc_for(self, packed, n, device):
        t = self.slots.get(id(packed))
        if t is not None and t.numel() == n and t.device == device:
            if id(packed) not in self._touched_set:
                self._touched_set.add(id(packed))
                self.touched.append(id(packed))
            return t
        n_al = (n + 63) // 64 * 64
        if not self.chunks or self.chunks[-1].device != device or self.off + n_al > self.chunks[-1].numel():
            self.chunks.append(torch.zeros(max(self.CHUNK, n_al), dtype=torch.float32, device=device))
            self.off = 0
        t = self.chunks[-1][self.off:self.off + n]
        self.where[id(packed)] = (len(self.chunks) - 1, self.off, self.off + n_al)
        self.off += n_al
        self.slots[id(packed)] = t
        self._touched_set.add(id(packed))
        self.touched.append(id(packed))
        return t

    def begin_step(self):
        self.touched, self._touched_set = [], set()

    def ranges(self, keys):
        """Contiguous slices of the chunks covering the accumulators `keys` (adjacent allocations merged)."""
        spans = sorted(self.where[k] for k in keys)
        out = []
        for ci, b, e in spans:
            if out and out[-1][0] == ci and out[-1][2] == b:
                out[-1][2] = e
            else:
                out.append([ci, b, e])
        return [self.chunks[ci][b:e] for ci, b, e in out]

    def note(self, acc, sink, Cout, Cin_g, taps, kpad, dtype):
        # accumulate = 0: the conversion launch is the ONLY writer of a conv weight's gradient in a step (every use of the weight
        # adds into the same accumulator; the flat buffer is zero-filled before the step), so it stores instead of read-add-store
        self.entries[(acc.data_ptr(), sink.data_ptr())] = (acc.data_ptr(), sink.data_ptr(), dt_code(dtype), Cout, Cin_g, taps, kpad, 0)

    _touched_set = frozenset()

    def zero(self):
        for c in self.chunks:
            c.zero_()
        self.begin_step()


WACC = None                  # WgradAccum of the running train step (set by TrainStep), else None: per-slice partials


def _sink(p):
    """The .grad buffer of parameter p when direct accumulation applies, else None."""
    if not DIRECT_GRADS or p is None or not p.requires_grad:
        return None
    g = p.grad
    if g is None or not g.is_contiguous() or g.dtype != p.dtype:
        return None
    return g


def _fmt(precision):
    return MS_BF16X2 if precision == "bf16x3" else MS_BF16


class Planes:
    """bf16 tensor-core operand planes of an activation: hi plane [rows*rs] and, in split mode, the lo plane
    (residual of the bf16 rounding) `ps` elements later.  rs = row stride in elements (channels padded to 8)."""
    __slots__ = ("t", "fmt", "rs", "ps")

    def __init__(self, t, fmt, rs, ps):
        self.t, self.fmt, self.rs, self.ps = t, fmt, rs, ps


def alloc_planes(rows, rs, fmt, device):
    ps = (rows * rs + 7) // 8 * 8
    t = torch.empty((2 if fmt == MS_BF16X2 else 1) * ps, dtype=torch.bfloat16, device=device)
    return Planes(t, fmt, rs, ps)


def pad8(c):
    return (c + 7) // 8 * 8


def planes_of(x, fmt, rs):
    """Planes of a channels-last fp32 activation: the producer's side output when it made one, else a cast."""
    pl = getattr(x, "_ms_planes", None)
    if pl is not None and pl.fmt == fmt and pl.rs == rs:
        return pl
    if x.dtype == torch.bfloat16:
        x = as_f32(x)           # a planes-only activation in another operand format: go through fp32
    x = _f32c(x)
    C = x.shape[-1]
    rows = x.numel() // C
    pl = alloc_planes(rows, rs, fmt, x.device)
    call("ms_to_planes", ptr(x), rows, C, rs, ptr(pl.t), fmt, pl.ps, stream())
    try:
        x._ms_planes = pl         # a second consumer of the same tensor object reuses the cast
    except AttributeError:
        pass
    return pl


def attach_planes(y, pl):
    y._ms_planes = pl
    return y


def planes_view(pl, shape):
    """A planes-only activation (inference fast path): the hi plane seen as a bf16 tensor of the activation's shape,
    carrying the Planes object.  Consumers that need fp32 go through as_f32()."""
    n = 1
    for d_ in shape:
        n *= d_
    if pl.rs != shape[-1]:
        raise MixStageError("internal: planes_view needs an unpadded row stride")
    t = pl.t[:n].view(shape)
    t._ms_planes = pl
    return t


def as_f32(x):
    """fp32 contents of an activation: itself, or hi (+ lo) of a planes-only activation."""
    if x.dtype == torch.float32:
        return x
    pl = getattr(x, "_ms_planes", None)
    if x.dtype != torch.bfloat16 or pl is None:
        raise MixStageError("internal: expected an fp32 activation or a planes-only one, got %s" % x.dtype)
    C = x.shape[-1]
    rows = x.numel() // C
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    call("ms_planes_to_f32", ptr(pl.t), pl.fmt, pl.ps, rows, C, pl.rs, ptr(out), stream())
    out._ms_planes = pl
    return out


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


CAST_CACHE = None            # TrainStep sets a fresh dict for the duration of one step body (also while it is being captured)


def cast(x, dtype):
    """x in another floating dtype.  Inside a TrainStep body an input that takes no gradient (audio, ground-truth pose: read
    by several consumers of the step) is narrowed once: the copy is kept for the rest of that body (and of that capture)."""
    if x.dtype == dtype:
        return x
    if CAST_CACHE is not None and dtype == torch.float32 and not x.requires_grad:
        key = (x.data_ptr(), x._version, tuple(x.shape), x.dtype)
        y = CAST_CACHE.get(key)
        if y is None:
            y = CAST_CACHE[key] = cast_raw(x, dtype)
        return y
    return _Cast.apply(x, dtype)


def f32_of(x):
    """fp32 form of a module output: the fp32 tensor it was widened from when the producer left it on the object (the
    generator's pose, still connected to the autograd graph), else a cast."""
    src = getattr(x, "_ms_f32", None)
    if src is not None and src.shape == x.shape:
        return src
    return cast(x, torch.float32)


# ---------------------------------------------------------------------------- conv block
class PackedWeight:
    """fp32 re-tiled copies of a conv weight, refreshed when the parameter changes."""

    def __init__(self):
        self.key = None
        self.wf = None
        self.wt = None
        self.bias = None
        self.bias_key = None
        self.stats_epoch = 0    # bumped when this block's BatchNorm updates its running statistics in place
        self._src = {}          # slot -> arguments of the getter that filled it (refresh() replays them)

    def refresh(self, entries=None):
        """Re-derive every cached copy from the current parameters, in place (same buffers: captured CUDA graphs keep
        pointing at them).  Tensor-core re-tilings are appended to `entries` (for ONE ms_pack_igemm_weight_multi launch)
        when a list is given."""
        global FORCE_REPACK
        old, FORCE_REPACK = FORCE_REPACK, True
        try:
            for slot, args in list(self._src.items()):
                if slot == "f32":
                    self.get(*args)
                elif slot == "bias":
                    self.get_bias(*args)
                elif slot == "ss":
                    self.get_eval_ss(*args)
                else:
                    self.get_tc(*args, collect=entries)
        finally:
            FORCE_REPACK = old

    def get(self, weight, desc):
        key = (weight.data_ptr(), weight._version, _weight_epoch, weight.dtype, weight.device)
        # detached: a stored view with a grad_fn would keep the parameter's AccumulateGrad node (and the stream it was
        # created on) alive across iterations, which breaks CUDA-graph capture of the next backward
        self._src["f32"] = (weight.detach(), desc)
        if key != self.key or FORCE_REPACK:
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

    # ---- tensor-core path: cached descriptors per input shape, packed bf16 (hi[, lo]) weights per version
    def tc_plans(self, x_shape, Cout, cfg, rs, need_dgrad, npass=1):
        key = (tuple(x_shape), Cout, rs, need_dgrad, npass)
        pl = getattr(self, "_plans", None)
        if pl is None:
            pl = self._plans = {}
        if key not in pl:
            B, H, W, Cin = x_shape
            Ho, Wo = conv_out(H, cfg.kh, cfg.sh, cfg.ph), conv_out(W, cfg.kw, cfg.sw, cfg.pw)
            geo = (B, H, W, Cin, Cout, cfg.kh, cfg.kw, cfg.sh, cfg.sw, cfg.ph, cfg.pw, cfg.groups, Ho, Wo)
            pf = igemm.make_fwd(*geo, a_row_stride=rs if rs != Cin else None)
            pd = igemm.make_dgrad(*geo, out_row_stride=rs if rs != Cin else None) if need_dgrad else None
            for p_ in (pf, pd):
                if p_ is not None:
                    p_.desc.block_n = p_.block_n0 = igemm.pick_block_n(p_.desc, npass)
            pl[key] = (pf, pd)
        return pl[key]

    def get_tc(self, weight, plan, fmt, groups, collect=None):
        """Packed weight planes for `plan` (forward or dgrad tiling).  Returns (tensor, plane stride)."""
        slot = "_tcw%d" % plan.mode
        key = (weight.data_ptr(), weight._version, _weight_epoch, weight.dtype, weight.device, fmt, plan.wp_numel,
               tuple(plan.srctap))
        self._src[slot] = (weight.detach(), plan, fmt, groups)
        cur = getattr(self, slot, None)
        if cur is not None and cur[0] == key and not FORCE_REPACK:
            return cur[1], cur[2]
        w = weight.detach()
        if not w.is_contiguous():
            w = w.contiguous()
        ps = (plan.wp_numel + 7) // 8 * 8
        if cur is not None and cur[1].numel() == (2 if fmt == MS_BF16X2 else 1) * ps and cur[1].device == w.device:
            buf = cur[1]
        else:
            buf = torch.empty((2 if fmt == MS_BF16X2 else 1) * ps, dtype=torch.bfloat16, device=w.device)
        d = plan.desc
        Cout, Cin_g = w.shape[0], w.shape[1]
        taps_total = w.shape[2] * w.shape[3]
        lo = buf.data_ptr() + 2 * ps if fmt == MS_BF16X2 else None
        if collect is not None:
            e = _lib.PackEntry()
            e.w, e.wp, e.wp_lo = w.data_ptr(), buf.data_ptr(), lo
            e.pdt, e.Cout, e.Cin_g, e.taps_total, e.groups, e.mode = dt_code(w.dtype), Cout, Cin_g, taps_total, groups, plan.mode
            e.num_classes, e.class_n, e.ntaps, e.kpad = d.num_classes, d.class_n, d.ntaps, plan.kpad
            for i_, t_ in enumerate(plan.srctap):
                e.srctap[i_] = t_
            collect.append(e)
        else:
            call("ms_pack_igemm_weight_bf16", ptr(w), dt_code(w.dtype), Cout, Cin_g, taps_total, groups, plan.mode,
                 d.num_classes, d.class_n, d.ntaps, plan.kpad, plan.srctap_c, ptr(buf), lo, stream())
        setattr(self, slot, (key, buf, ps))
        return buf, ps

    def get_eval_ss(self, gamma, beta, cbias, bn_buffers, cfg):
        """Eval-mode BatchNorm folded to per-channel (scale, shift) with the conv bias inside the shift
        (GEMM outputs exclude it); recomputed when a parameter, the running statistics or the epochs change."""
        rm, rv, _ = bn_buffers
        key = tuple((t.data_ptr(), t._version) for t in (gamma, beta, rm, rv) + ((cbias,) if cbias is not None else ())) + (
            _weight_epoch, self.stats_epoch, gamma.dtype)
        self._src["ss"] = (gamma.detach(), beta.detach(), None if cbias is None else cbias.detach(),
                           tuple(None if b is None else b.detach() for b in bn_buffers), cfg)
        cur = getattr(self, "_eval_ss", None)
        if cur is None or cur[0] != key or FORCE_REPACK:
            C = gamma.numel()
            ss = cur[1] if cur is not None else torch.empty(4, C, dtype=torch.float32, device=gamma.device)
            call("ms_bn_finalize", None, None, 1, C, ptr(gamma), ptr(beta), ptr(cbias), ptr(rm), ptr(rv), dt_code(gamma.dtype),
                 0, cfg.momentum, cfg.eps, ptr(ss[0]), ptr(ss[1]), ptr(ss[2]), ptr(ss[3]), stream())
            cur = self._eval_ss = (key, ss)
        return cur[1]

    def get_bias(self, bias):
        if bias is None:
            return None
        key = (bias.data_ptr(), bias._version, _weight_epoch, bias.dtype, bias.device)
        self._src["bias"] = (bias.detach(),)
        if key != self.bias_key or FORCE_REPACK:
            b = bias.detach().contiguous()
            if self.bias is None or self.bias.numel() != b.numel() or self.bias.device != b.device:
                self.bias = torch.empty(b.shape, dtype=torch.float32, device=b.device)
            call("ms_cast", ptr(b), dt_code(b.dtype), ptr(self.bias), MS_F32, b.numel(), stream())     # in place: graphs keep the pointer
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


# ---- pieces shared by the CUDA-core and tensor-core variants of the block
def _bn_scale_shift(z, rows, Cout, gamma, beta, cbias, bn_buffers, training, cfg, packed=None):
    """Batch statistics + finalize (training: one fused launch that also advances num_batches_tracked) or the
    running-statistics fold (eval).  Returns ss = [scale, shift, mean, rstd] (4, Cout) fp32."""
    rm, rv, nbt = bn_buffers
    st, dev = stream(), z.device
    ss = torch.empty(4, Cout, dtype=torch.float32, device=dev)
    pdt = dt_code(gamma.dtype)
    if cbias is not None and cbias.dtype != gamma.dtype:
        raise MixStageError("conv bias and BatchNorm parameters must share a dtype")
    if training:
        global _stats_epoch
        _stats_epoch += 1                                      # running statistics change in place (no torch version bump)
        if packed is not None:
            packed.stats_epoch += 1
        acc = arena.take((2 * Cout + 2,), dev)                 # sum, sumsq, ticket
        base = acc.data_ptr()
        call("ms_bn_stats_finalize", ptr(z), rows, Cout, base, base + 8 * Cout, base + 16 * Cout, ptr(gamma), ptr(beta),
             ptr(cbias), ptr(rm), ptr(rv), ptr(nbt), pdt, cfg.momentum, cfg.eps, ptr(ss[0]), ptr(ss[1]), ptr(ss[2]),
             ptr(ss[3]), st)
    else:
        call("ms_bn_finalize", None, None, rows, Cout, ptr(gamma), ptr(beta), ptr(cbias), ptr(rm), ptr(rv), pdt, 0,
             cfg.momentum, cfg.eps, ptr(ss[0]), ptr(ss[1]), ptr(ss[2]), ptr(ss[3]), st)
    return ss


def _bn_act(z, ss, cfg, rows, Cout, desc, residual, up2, fmt, want_planes):
    """y = act(z*scale + shift) [upsampled x2 + residual], as fp32 and (optionally) as bf16 operand planes."""
    B = z.shape[0]
    dev = z.device
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
    yp = alloc_planes(y.numel() // Cout, Cout, fmt, dev) if (want_planes and Cout % 8 == 0) else None
    call("ms_bn_act_fwd_f32", ptr(z), ptr(ss[0]), ptr(ss[1]), slope, rows, Cout, ptr(y), ptr(res),
         1 if up2 else 0, desc.Wo, ptr(yp.t) if yp else None, yp.fmt if yp else 0, yp.ps if yp else 0, stream())
    return y, yp


def _bn_act_backward(ctx, dy, z, ss, want_f32, dzp):
    """dz = d(act(bn(z)))/dz . dy as fp32 (want_f32) and/or planes (dzp); affine-parameter gradients go to their
    sinks when direct accumulation is on, else come back as tensors.  Returns (dz, dgamma, dbeta)."""
    cfg, desc, rows, Cout = ctx.cfg, ctx.desc, ctx.rows, ctx.Cout
    st, dev = stream(), dy.device
    need_g, need_be = ctx.needs_input_grad[3], ctx.needs_input_grad[4]
    slope = cfg.slope if cfg.act else 1.0
    red = arena.take((2, Cout), dev)
    if ctx.training or need_g or need_be:
        call("ms_bn_act_bwd_reduce_f32", ptr(dy), ptr(z), ptr(ss[0]), ptr(ss[1]), ptr(ss[2]), ptr(ss[3]), slope,
             rows, Cout, 1 if ctx.up2 else 0, desc.Wo, ptr(red[0]), ptr(red[1]), st)
    sg = ctx.sinks[2] if need_g else None
    sb = ctx.sinks[3] if need_be else None
    direct = sg is not None and sb is not None and need_g and need_be
    dz = torch.empty_like(z) if want_f32 else None
    call("ms_bn_act_bwd_apply_f32", ptr(dy), ptr(z), ptr(ss[0]), ptr(ss[1]), ptr(ss[2]), ptr(ss[3]), slope,
         rows, Cout, 1 if ctx.up2 else 0, desc.Wo, ptr(red[0]), ptr(red[1]), 1 if ctx.training else 0, ptr(dz),
         ptr(dzp.t) if dzp else None, dzp.fmt if dzp else 0, dzp.ps if dzp else 0,
         ptr(sg) if direct else None, ptr(sb) if direct else None, dt_code(ctx.gamma_dtype), st)
    dgamma = dbeta = None
    if not direct:
        if need_g:
            dgamma = torch.empty(Cout, dtype=ctx.gamma_dtype, device=dev)
            call("ms_store_param_grad", ptr(red[0]), Cout, ptr(dgamma), dt_code(ctx.gamma_dtype), 0, st)
        if need_be:
            dbeta = torch.empty(Cout, dtype=ctx.gamma_dtype, device=dev)
            call("ms_store_param_grad", ptr(red[1]), Cout, ptr(dbeta), dt_code(ctx.gamma_dtype), 0, st)
    return dz, dgamma, dbeta


def _bias_grad(ctx, dz):
    """d(bias): identically zero under batch-statistics BatchNorm (any per-channel constant is removed), else the
    column sums of dz.  Returns a tensor for autograd, or None after accumulating into the sink."""
    cfg, rows, Cout = ctx.cfg, ctx.rows, ctx.Cout
    bdt = ctx.param_dtypes[1]
    if not ctx.needs_input_grad[2] or bdt is None:
        return None
    sink = ctx.sinks[1]
    dev = dz.device if dz is not None else None
    if cfg.has_bn and ctx.training:
        return None if sink is not None else torch.zeros(Cout, dtype=bdt, device=ctx.dev)
    acc = arena.take((Cout,), dev)
    call("ms_col_stats_f32", ptr(dz), rows, Cout, ptr(acc), None, stream())
    if sink is not None:
        call("ms_store_param_grad", ptr(acc), Cout, ptr(sink), dt_code(bdt), 1, stream())
        return None
    dbias = torch.empty(Cout, dtype=bdt, device=dev)
    call("ms_store_param_grad", ptr(acc), Cout, ptr(dbias), dt_code(bdt), 0, stream())
    return dbias


class _ConvBlock(torch.autograd.Function):
    """y = act(bn(conv(x) + b)) [+ upsample2(y) + residual].  x: (B,H,W,Cin) fp32 channels-last."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, residual, cfg, packed, bn_buffers, training, up2, precision="fp32",
                carrier=None):
        _need_cuda(x)
        if x.dtype != torch.bfloat16:          # bf16 = planes-only activation from the inference fast path (tc consumers only)
            x = _f32c(x)
        B, H, W, Cin = x.shape
        Cout = weight.shape[0]
        desc = make_desc(x.shape, Cout, cfg.kh, cfg.kw, cfg.sh, cfg.sw, cfg.ph, cfg.pw, cfg.groups)
        if weight.shape[1] * cfg.groups != Cin:
            raise MixStageError("conv: input has %d channels, weight expects %d" % (Cin, weight.shape[1] * cfg.groups))
        ctx.tc = False
        ctx.sinks = carrier.sinks if carrier is not None else (None, None, None, None)
        ctx.dev = x.device
        if precision != "fp32" and tc_eligible(cfg, B, H, W, Cin, Cout, ctx.needs_input_grad[0]):
            return _tc_forward(ctx, x, weight, bias, gamma, beta, residual, cfg, packed, bn_buffers, training, up2,
                               desc, _fmt(precision), carrier)
        x = _f32c(x)
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
        ss = _bn_scale_shift(z, rows, Cout, gamma, beta, None, bn_buffers, training, cfg, packed)
        tc_next = precision != "fp32" and carrier is not None         # the consumer may be a tensor-core layer
        y, yp = _bn_act(z, ss, cfg, rows, Cout, desc, residual, up2, _fmt(precision), tc_next)
        if yp is not None:
            carrier.planes = yp
        ctx.gamma_dtype = gamma.dtype
        ctx.save_for_backward(x, z, ss)
        return y

    @staticmethod
    def backward(ctx, dy):
        if ctx.tc:
            return _tc_backward(ctx, dy)
        cfg, desc, st = ctx.cfg, ctx.desc, stream()
        rows, Cout = ctx.rows, ctx.Cout
        dy = dy.contiguous()
        dev = dy.device
        need_x, need_w, need_b, need_g, need_be, need_res = ctx.needs_input_grad[:6]
        dgamma = dbeta = dres = dw = dx = None
        if cfg.has_bn:
            x, z, ss = ctx.saved_tensors
            dz, dgamma, dbeta = _bn_act_backward(ctx, dy, z, ss, True, None)
            if ctx.has_res and need_res:
                dres = dy
        else:
            x, z = ctx.saved_tensors
            if cfg.act:
                dz = torch.empty_like(z)
                call("ms_lrelu_bwd_f32", ptr(dy), ptr(z), cfg.slope, z.numel(), ptr(dz), None, 0, 0, st)
            else:
                dz = dy
        dbias = _bias_grad(ctx, dz)
        wdt = ctx.param_dtypes[0]
        if need_w:
            n = Cout * (desc.Cin // desc.groups) * desc.kh * desc.kw
            dwf = torch.empty(n, dtype=torch.float32, device=dev)
            call("ms_conv_wgrad_f32", ptr(x), ptr(dz), ptr(dwf), desc, st)
            sink = ctx.sinks[0]
            if sink is not None:
                call("ms_unpack_conv_wgrad", ptr(dwf), desc, ptr(sink), dt_code(wdt), 1, st)
            else:
                dw = torch.empty((Cout, desc.Cin // desc.groups, desc.kh, desc.kw), dtype=wdt, device=dev)
                call("ms_unpack_conv_wgrad", ptr(dwf), desc, ptr(dw), dt_code(wdt), 0, st)
        if need_x:
            dx = torch.empty_like(x)
            call("ms_conv_dgrad_f32", ptr(dz), ptr(ctx.wt), ptr(dx), desc, st)
        return dx, dw, dbias, dgamma, dbeta, dres, None, None, None, None, None, None, None


def tc_eligible(cfg, B, H, W, Cin, Cout, need_dgrad):
    """Geometries the tcgen05 implicit-GEMM kernels take; the rest (C_in = 1, N < 16, ...) stay on the
    CUDA-core kernels."""
    rs = pad8(Cin)
    if cfg.groups > 1 and rs != Cin:
        return False
    if Cin // cfg.groups < 8:
        return False            # C_in = 1 (audio_encoder.conv.0): direct kernel in csrc/conv_small.cu
    if not igemm.fwd_supported(Cin, Cout, cfg.groups, cfg.sh, cfg.sw, H, W, rs):
        return False
    if cfg.kh * cfg.kw > igemm.MAX_TAPS:
        return False
    if need_dgrad:
        if not igemm.dgrad_supported(rs if cfg.groups == 1 else Cin, Cout, cfg.groups, cfg.sh, cfg.sw, H, W, cfg.kh, cfg.kw):
            return False
    elif Cout % 8:
        return False
    return True


def _tc_forward(ctx, x, weight, bias, gamma, beta, residual, cfg, packed, bn_buffers, training, up2, desc, fmt, carrier):
    """Tensor-core variant of the block: z = igemm(x planes, W planes) in fp32 (conv bias folded into the BN
    finalize), then the same statistics / normalise kernels; the activation leaves both as fp32 (autograd,
    residual adds, non-GEMM consumers) and as bf16 operand planes for the next GEMM."""
    B, H, W, Cin = x.shape
    Cout = weight.shape[0]
    st, dev = stream(), x.device
    rs = pad8(Cin)
    need_dx = ctx.needs_input_grad[0]
    split = fmt == MS_BF16X2
    xp = planes_of(x, fmt, rs)
    pf, pd = packed.tc_plans(x.shape, Cout, cfg, rs, need_dx, 3 if split else 1)
    wp, wps = packed.get_tc(weight, pf, fmt, cfg.groups)
    rows = B * desc.Ho * desc.Wo
    d = pf.desc
    igemm.set_planes(pf, split, xp.ps, wps, 0)
    d.out_dtype = _lib.MS_F32
    fuse_act = (not cfg.has_bn) and cfg.act
    d.epilogue, d.slope = (2 if fuse_act else 0), cfg.slope
    b32 = None if cfg.has_bn else packed.get_bias(bias)
    d.split_k = 1 if fuse_act else igemm.igemm_split(d, 3 if split else 1)
    global last_gemm_flops
    last_gemm_flops = ctx.flops = 2.0 * rows * Cout * (Cin // cfg.groups) * cfg.kh * cfg.kw
    # training-mode BatchNorm blocks: GEMM + statistics + normalise in ONE cooperative launch (csrc/conv_train.cu)
    fused = FUSED_BLOCKS and cfg.has_bn and training and (Cout // cfg.groups) % 32 == 0 and Cout <= 8192
    zshape = (B, desc.Ho, desc.Wo, Cout)
    if fused:
        d.block_n, d.split_k = _block_plan(pf, 3 if split else 1, True)
        z = arena.take_f32(zshape, dev) if d.split_k > 1 else torch.empty(zshape, dtype=torch.float32, device=dev)
    else:
        z = torch.empty(zshape, dtype=torch.float32, device=dev)
        d.block_n = pf.block_n0
        d.out_numel = z.numel()
        call("ms_igemm_bf16", d, ptr(xp.t), ptr(wp), ptr(b32), None, None, ptr(z), st)
    ctx.tc, ctx.fmt = True, fmt
    ctx.cfg, ctx.desc, ctx.training, ctx.up2 = cfg, desc, training, up2
    ctx.rows, ctx.Cout, ctx.rs = rows, Cout, rs
    ctx.plans = (pf, pd)
    ctx.wt = packed.get_tc(weight, pd, fmt, cfg.groups) if pd is not None else None
    ctx.x_shape, ctx.xp_meta = tuple(x.shape), (xp.rs, xp.ps)
    ctx.param_dtypes = (weight.dtype, None if bias is None else bias.dtype)
    ctx.has_res = residual is not None
    ctx.packed = packed
    ctx.fused = False
    if not cfg.has_bn:
        ctx.save_for_backward(xp.t, z)
        return z
    if fused:
        ss, y, yp = _fused_block_fwd(d, xp, wp, z, rows, Cout, gamma, beta, bias, bn_buffers, cfg, packed, desc, residual,
                                     up2, fmt)
        ctx.fused = True
    else:
        ss = _bn_scale_shift(z, rows, Cout, gamma, beta, bias, bn_buffers, training, cfg, packed)
        y, yp = _bn_act(z, ss, cfg, rows, Cout, desc, residual, up2, fmt, True)
    ctx.gamma_dtype = gamma.dtype
    ctx.save_for_backward(xp.t, z, ss)
    if yp is not None and carrier is not None:
        carrier.planes = yp           # attached to the output by conv_block (autograd returns a fresh tensor object)
    return y


def _block_plan(plan, npass, stats):
    """(block_n, split_k) of a fused-block GEMM phase, cached on the plan."""
    key = ("_bp", npass, stats)
    cache = plan.__dict__.setdefault("_bp_cache", {})
    if key not in cache:
        cache[key] = igemm.block_plan(plan.desc, npass, stats)
    return cache[key]


def _block_bn(C, gamma, beta, cbias, bn_buffers, cfg, training, sums, ss):
    b = _lib.BlockBn()
    b.C, b.pdt, b.training = C, dt_code(gamma.dtype), 1 if training else 0
    b.momentum, b.eps, b.slope = cfg.momentum, cfg.eps, (cfg.slope if cfg.act else 1.0)
    b.gamma, b.beta, b.conv_bias = ptr(gamma), ptr(beta), ptr(cbias)
    if bn_buffers is not None:
        b.running_mean, b.running_var, b.num_batches_tracked = ptr(bn_buffers[0]), ptr(bn_buffers[1]), ptr(bn_buffers[2])
    b.sums, b.ss = ptr(sums), ptr(ss)
    return b


def _fused_block_fwd(d, xp, wp, z, rows, Cout, gamma, beta, cbias, bn_buffers, cfg, packed, desc, residual, up2, fmt):
    """z = igemm, batch statistics, finalize, y = act(bn(z)) [upsample x2 + skip] -> fp32 + operand planes: one launch."""
    global _stats_epoch
    dev = z.device
    if cbias is not None and cbias.dtype != gamma.dtype:
        raise MixStageError("conv bias and BatchNorm parameters must share a dtype")
    _stats_epoch += 1
    packed.stats_epoch += 1
    B = z.shape[0]
    if up2:
        if desc.Ho != 1:
            raise MixStageError("upsample+skip fusion is 1-D only")
        oshape = (B, 1, 2 * desc.Wo, Cout)
        res = _f32c(residual)
        if tuple(res.shape) != oshape:
            raise MixStageError("skip tensor shape %s != %s" % (tuple(res.shape), oshape))
    else:
        oshape, res = tuple(z.shape), None
    y = torch.empty(oshape, dtype=torch.float32, device=dev)
    yp = alloc_planes(y.numel() // Cout, Cout, fmt, dev)
    ss = torch.empty(4, Cout, dtype=torch.float32, device=dev)
    acc = arena.take((2 * Cout + 2,), dev)             # sum, sumsq, barrier counter
    bn = _block_bn(Cout, gamma, beta, cbias, bn_buffers, cfg, True, acc, ss)
    call("ms_conv_block_train_fwd", d, ptr(xp.t), ptr(wp), ptr(z), bn, ptr(y), ptr(yp.t), yp.fmt, yp.ps, ptr(res), None, 0, 0,
         1 if up2 else 0, acc.data_ptr() + 16 * Cout, stream())
    return ss, y, yp


def _tc_backward(ctx, dy):
    cfg, desc, st, fmt = ctx.cfg, ctx.desc, stream(), ctx.fmt
    rows, Cout = ctx.rows, ctx.Cout
    dy = dy.contiguous()
    dev = dy.device
    need_x, need_w, need_b, need_g, need_be, need_res = ctx.needs_input_grad[:6]
    dgamma = dbeta = dres = dbias = dw = dx = None
    split = fmt == MS_BF16X2
    dzp = alloc_planes(rows, Cout, fmt, dev)
    wdt, bdt = ctx.param_dtypes
    pf, pd = ctx.plans
    xrs, xps = ctx.xp_meta
    B, H, W, Cin = ctx.x_shape
    fused_dx = None
    if cfg.has_bn and ctx.fused and ctx.training and need_g and need_be and (ctx.sinks[2] is None) == (ctx.sinks[3] is None):
        # BatchNorm backward (reductions + apply -> dz planes, affine gradients) and the input-gradient GEMM: one launch
        xpt, z, ss = ctx.saved_tensors
        sg, sb = ctx.sinks[2], ctx.sinks[3]
        if sg is None:
            dgamma = torch.zeros(Cout, dtype=ctx.gamma_dtype, device=dev)
            dbeta = torch.zeros(Cout, dtype=ctx.gamma_dtype, device=dev)
            sg, sb = dgamma, dbeta
        red = arena.take((2 * Cout + 2,), dev)         # dgamma, dbeta, barrier counter
        bn = _lib.BlockBn()
        bn.C, bn.pdt, bn.training = Cout, dt_code(ctx.gamma_dtype), 1
        bn.momentum, bn.eps, bn.slope = cfg.momentum, cfg.eps, (cfg.slope if cfg.act else 1.0)
        bn.sums, bn.ss = ptr(red), ptr(ss)
        bn.gamma = bn.beta = ptr(ss)                    # not read by the backward; non-NULL for the argument check
        dgd, wtp, dxf = None, None, None
        global last_gemm_flops
        last_gemm_flops = ctx.flops if need_x else 0.0
        if need_x:
            wt, wtps = ctx.wt
            igemm.set_planes(pd, split, dzp.ps, wtps, 0)
            pd.desc.block_n, pd.desc.split_k = _block_plan(pd, 3 if split else 1, False)
            dshape = (B, H, W, xrs)
            dxf = arena.take_f32(dshape, dev) if pd.desc.split_k > 1 else torch.empty(dshape, dtype=torch.float32, device=dev)
            dgd, wtp = pd.desc, wt
        call("ms_conv_block_train_bwd", dgd, ptr(dy), ptr(z), bn, rows, 1 if ctx.up2 else 0, desc.Wo, ptr(dzp.t), fmt, dzp.ps,
             ptr(sg), ptr(sb), dt_code(ctx.gamma_dtype), ptr(wtp), ptr(dxf), red.data_ptr() + 16 * Cout, st)
        if need_x:
            fused_dx = dxf if xrs == Cin else dxf[..., :Cin]
        dz = None
        if ctx.has_res and need_res:
            dres = dy
    elif cfg.has_bn:
        xpt, z, ss = ctx.saved_tensors
        need_f32 = need_b and bdt is not None and not ctx.training
        dz, dgamma, dbeta = _bn_act_backward(ctx, dy, z, ss, need_f32, dzp)
        if ctx.has_res and need_res:
            dres = dy
    else:
        xpt, z = ctx.saved_tensors
        if cfg.act:
            dz = torch.empty_like(z)
            call("ms_lrelu_bwd_f32", ptr(dy), ptr(z), cfg.slope, z.numel(), ptr(dz), ptr(dzp.t), fmt, dzp.ps, st)
        else:
            dz = dy
            call("ms_to_planes", ptr(dy), rows, Cout, Cout, ptr(dzp.t), fmt, dzp.ps, st)
    dbias = _bias_grad(ctx, dz)
    Cin_g = Cin // cfg.groups
    last_gemm_flops = ctx.flops
    if need_w:
        igemm.set_planes(pf, split, xps, 0, dzp.ps)
        nsplit, pf.desc.wgrad_c_tile = igemm.wgrad_split(pf.desc, npass=3 if split else 1)
        pf.desc.split_k = nsplit
        sink = ctx.sinks[0]
        if sink is not None and WACC is not None and FUSED_BLOCKS:
            # every pixel slice adds into ONE persistent accumulator; the step converts all of them in one launch
            acc = WACC.acc_for(ctx.packed, pf.wp_numel, dev)
            WACC.note(acc, sink, Cout, Cin_g, cfg.kh * cfg.kw, pf.kpad, wdt)
            if SIDE is not None:
                with SIDE.fork(xpt, dzp.t):
                    call("ms_wgrad_bf16_acc", pf.desc, ptr(xpt), ptr(dzp.t), ptr(acc), stream())
            else:
                call("ms_wgrad_bf16_acc", pf.desc, ptr(xpt), ptr(dzp.t), ptr(acc), st)
            dwp = None
        elif sink is not None and SIDE is not None:
            with SIDE.fork(xpt, dzp.t):
                dwp = torch.empty(nsplit * pf.wp_numel, dtype=torch.float32, device=dev)
                call("ms_wgrad_bf16", pf.desc, ptr(xpt), ptr(dzp.t), ptr(dwp), stream())
                call("ms_unpack_igemm_wgrad", ptr(dwp), Cout, Cin_g, cfg.kh * cfg.kw, pf.desc.ntaps, pf.kpad, ptr(sink),
                     dt_code(wdt), nsplit, 1, stream())
                SIDE.keep.append(dwp)
            dwp = None
        else:
            dwp = torch.empty(nsplit * pf.wp_numel, dtype=torch.float32, device=dev)     # one partial per row slice
            call("ms_wgrad_bf16", pf.desc, ptr(xpt), ptr(dzp.t), ptr(dwp), st)
        if dwp is None:
            pass
        elif sink is not None:
            call("ms_unpack_igemm_wgrad", ptr(dwp), Cout, Cin_g, cfg.kh * cfg.kw, pf.desc.ntaps, pf.kpad, ptr(sink),
                 dt_code(wdt), nsplit, 1, st)
        else:
            dw = torch.empty((Cout, Cin_g, cfg.kh, cfg.kw), dtype=wdt, device=dev)
            call("ms_unpack_igemm_wgrad", ptr(dwp), Cout, Cin_g, cfg.kh * cfg.kw, pf.desc.ntaps, pf.kpad, ptr(dw),
                 dt_code(wdt), nsplit, 0, st)
    if fused_dx is not None:
        dx = fused_dx
    elif need_x:
        wt, wtps = ctx.wt
        igemm.set_planes(pd, split, dzp.ps, wtps, 0)
        dxf = torch.empty((B, H, W, xrs), dtype=torch.float32, device=dev)
        pd.desc.block_n = pd.block_n0
        pd.desc.split_k = igemm.igemm_split(pd.desc, 3 if split else 1)
        pd.desc.out_numel = dxf.numel()
        call("ms_igemm_bf16", pd.desc, ptr(dzp.t), ptr(wt), None, None, None, ptr(dxf), st)
        dx = dxf if xrs == Cin else dxf[..., :Cin]
    return dx, dw, dbias, dgamma, dbeta, dres, None, None, None, None, None, None, None


def _tc_eval(x, weight, bias, gamma, beta, residual, cfg, packed, bn_buffers, up2, fmt, want, row_w=None):
    """Inference fast path of the block (eval mode, no autograd): ONE tcgen05 launch whose epilogue applies the folded
    BatchNorm + LeakyReLU (+ UNet upsample-and-skip) and writes the next layer's bf16 operand planes directly
    (want "planes"), fp32 (want "f32") or both -- no fp32 activation round trip through HBM."""
    B, H, W, Cin = x.shape
    Cout = weight.shape[0]
    st, dev = stream(), x.device
    rs = pad8(Cin)
    split = fmt == MS_BF16X2
    xp = planes_of(x, fmt, rs)
    pf, _ = packed.tc_plans(x.shape, Cout, cfg, rs, False, 3 if split else 1)
    wp, wps = packed.get_tc(weight, pf, fmt, cfg.groups)
    Ho, Wo = conv_out(H, cfg.kh, cfg.sh, cfg.ph), conv_out(W, cfg.kw, cfg.sw, cfg.pw)
    rows = B * Ho * Wo
    d = pf.desc
    if cfg.has_bn:
        ss = packed.get_eval_ss(gamma, beta, bias, bn_buffers, cfg)
        d.epilogue, d.slope = 1, (cfg.slope if cfg.act else 1.0)
        b32, scale, shift = None, ptr(ss[0]), ptr(ss[1])
    else:
        d.epilogue, d.slope = (2 if cfg.act else 0), cfg.slope
        b32, scale, shift = packed.get_bias(bias), None, None
    d.split_k, d.out_numel = 1, 0
    oshape = (B, 1, 2 * Wo, Cout) if up2 else (B, Ho, Wo, Cout)
    rows_out = 2 * rows if up2 else rows
    res = None
    if up2:
        if Ho != 1:
            raise MixStageError("upsample+skip fusion is 1-D only")
        if tuple(residual.shape) != oshape:
            raise MixStageError("skip tensor shape %s != %s" % (tuple(residual.shape), oshape))
        res = planes_of(residual, fmt, Cout)
    y32 = torch.empty(oshape, dtype=torch.float32, device=dev) if want != "planes" else None
    yp = alloc_planes(rows_out, Cout, fmt, dev) if want != "f32" else None
    global last_gemm_flops
    small = False
    if FUSED_BLOCKS and cfg.has_bn and row_w is None and Cout <= 8192 and (Cout // cfg.groups) % 32 == 0:
        bn_, ksplit = _block_plan(pf, 3 if split else 1, True)
        small = ksplit > 1 or (igemm.block_resident(d, bn_) and igemm._tiles_m(d) * d.num_classes * ((d.class_n + bn_ - 1) // bn_) <= 148)
    if small:
        # small batch: at most one tile per SM.  Full-K tiles normalise straight out of TMEM (no barrier at all); long-K
        # layers with few tiles run k-slices over the whole machine, reduce into an fp32 tile in L2, one device-wide barrier,
        # then the folded BatchNorm + LeakyReLU (+ upsample/skip) -> planes / fp32: ONE launch either way (conv_train.cu)
        d.epilogue, d.block_n, d.split_k, d.out_dtype = 0, bn_, ksplit, _lib.MS_F32
        igemm.set_planes(pf, split, xp.ps, wps, 0)
        z = arena.take_f32((B, Ho, Wo, Cout), dev) if ksplit > 1 else None
        sync = arena.take((2,), dev)
        bn = _block_bn(Cout, gamma, beta, None, None, cfg, False, None, ss)
        bn.sums = ptr(sync)                  # unused in inference form; non-NULL for the argument check
        last_gemm_flops = 2.0 * rows * Cout * (Cin // cfg.groups) * cfg.kh * cfg.kw
        call("ms_conv_block_train_fwd", d, ptr(xp.t), ptr(wp), ptr(z), bn, ptr(y32), ptr(yp.t) if yp is not None else None,
             fmt, yp.ps if yp is not None else 0, None, ptr(res.t) if res is not None else None, fmt,
             res.ps if res is not None else 0, 1 if up2 else 0, ptr(sync), st)
        if y32 is not None:
            if yp is not None:
                y32._ms_planes = yp
            return y32
        return planes_view(yp, oshape)
    igemm.set_planes(pf, split, xp.ps, wps, yp.ps if yp is not None else 0)
    d.out_dtype = fmt if yp is not None else _lib.MS_F32
    d.block_n = pf.block_n0
    last_gemm_flops = 2.0 * rows * Cout * (Cin // cfg.groups) * cfg.kh * cfg.kw
    if row_w is not None:
        # soft cluster weight of every row applied in the epilogue (classes = clusters): the mixture moves into the GEMMs
        if up2 or row_w.dtype != torch.float32 or not row_w.is_contiguous() or row_w.numel() != rows * d.num_classes:
            raise MixStageError("row weights must be contiguous fp32 (rows, groups)")
        call("ms_igemm_bf16_mix", d, ptr(xp.t), ptr(wp), ptr(b32), scale, shift, ptr(yp.t) if yp is not None else ptr(y32),
             ptr(y32) if yp is not None else None, ptr(row_w), d.num_classes, 1, 0, st)
    else:
        call("ms_igemm_bf16_fused", d, ptr(xp.t), ptr(wp), ptr(b32), scale, shift, ptr(yp.t) if yp is not None else ptr(y32),
             ptr(y32) if yp is not None else None, ptr(res.t) if res is not None else None,
             (2 if split else 1) if res is not None else 0, res.ps if res is not None else 0, 1 if up2 else 0, st)
    if y32 is not None:
        if yp is not None:
            y32._ms_planes = yp
        return y32
    return planes_view(yp, oshape)


def fast_eval(precision=None):
    """True when blocks run on the inference fast path: tensor-core arithmetic, no autograd."""
    return (precision or _precision) != "fp32" and not torch.is_grad_enabled()


class MixedLogits:
    """The grouped 1x1 `logits` convolution + index_select_outputs (jlcss.py:83,106-115,194) as ONE dense GEMM on the
    inference fast path.  With the last sub-decoder block's epilogue having multiplied every row of cluster k's 256
    channels by its soft weight w_k (conv_block(row_w=...)),
        pose[row, p] = sum_k sum_c W[k*P + p, c] * (w_k * a_k[row, c]) + sum_k w_k * b[k*P + p]
    is a (rows, K*256) x (K*256, P) product accumulated in TMEM; the (B,T,K*P) per-cluster outputs are never written."""

    def __init__(self):
        self.packed = PackedWeight()
        self.key = None
        self.dense = None
        self._w = None              # detached view of the grouped weight the dense copy was derived from

    def refresh_dense(self):
        """Re-derive the dense (P, K*Cg) copy from the grouped weight, in place (TrainStep: the parameters were updated by
        a kernel that does not bump torch versions; captured graphs keep pointing at the same buffer)."""
        if self._w is None or self.dense is None:
            return
        KP, Cg = self._w.shape[0], self._w.shape[1]
        P = self.dense.shape[0]
        K = KP // P
        self.dense.copy_(self._w.view(K, P, Cg).permute(1, 0, 2).reshape(P, K * Cg, 1, 1))
        self.key = None

    @staticmethod
    def eligible(P, K, Cg):
        return P % 16 == 0 and 16 <= P <= 128 and K <= 16 and (K * Cg) % 64 == 0

    def __call__(self, x, weight, bias, soft):
        """x: planes-only (B,1,T,K*Cg) already row-weighted; weight (K*P, Cg, 1); bias (K*P,); soft (rows, K) fp32."""
        B, H, W, Ct = x.shape
        KP, Cg = weight.shape[0], weight.shape[1]
        K = Ct // Cg
        P = KP // K
        key = (weight.data_ptr(), weight._version, _weight_epoch, weight.dtype)
        self._w = weight.detach()
        if key != self.key or FORCE_REPACK:
            dense = self._w.view(K, P, Cg).permute(1, 0, 2).reshape(P, K * Cg, 1, 1)
            if self.dense is not None and self.dense.shape == dense.shape and self.dense.dtype == dense.dtype:
                self.dense.copy_(dense)        # same buffer: captured graphs / cached re-tilings keep pointing at it
            else:
                self.dense = dense.contiguous()
            self.key = key             # (the copy bumps self.dense's version: get_tc re-tiles it)
        fmt = x._ms_planes.fmt
        split = fmt == MS_BF16X2
        cfg = ConvCfg(1, 1, 1, 1, 0, 0, 1, 0.0, has_bn=False, act=False)
        xp = planes_of(x, fmt, Ct)
        pf, _ = self.packed.tc_plans((B, H, W, Ct), P, cfg, Ct, False, 3 if split else 1)
        pf.desc.block_n = P
        wp, wps = self.packed.get_tc(self.dense, pf, fmt, 1)
        b32 = self.packed.get_bias(bias)
        d = pf.desc
        d.epilogue, d.slope, d.split_k, d.out_numel = 0, 1.0, 1, 0
        igemm.set_planes(pf, split, xp.ps, wps, 0)
        d.out_dtype = _lib.MS_F32
        rows = B * H * W
        if soft.dtype != torch.float32 or not soft.is_contiguous() or soft.numel() != rows * K:
            raise MixStageError("mixture weights must be contiguous fp32 (rows, K)")
        y = torch.empty((B, H, W, P), dtype=torch.float32, device=x.device)
        global last_gemm_flops
        last_gemm_flops = 2.0 * rows * KP * Cg
        call("ms_igemm_bf16_mix", d, ptr(xp.t), ptr(wp), ptr(b32), None, None, ptr(y), None, ptr(soft), K, 2, K, stream())
        return y


def _cin1_eval(x, weight, bias, gamma, beta, cfg, packed, bn_buffers, fmt, want):
    """Inference fast path of the C_in = 1 block (audio_encoder.conv.0): conv + folded BatchNorm + LeakyReLU in one
    streaming kernel that writes the next layer's operand planes (and fp32 only when asked)."""
    x = _f32c(x)
    B, H, W, _ = x.shape
    w4 = weight.unsqueeze(2) if weight.dim() == 3 else weight
    Cout = w4.shape[0]
    desc = make_desc(x.shape, Cout, cfg.kh, cfg.kw, cfg.sh, cfg.sw, cfg.ph, cfg.pw, 1)
    wf, _ = packed.get(w4, desc)
    ss = packed.get_eval_ss(gamma, beta, bias, bn_buffers, cfg)
    oshape = (B, desc.Ho, desc.Wo, Cout)
    rows = B * desc.Ho * desc.Wo
    y32 = torch.empty(oshape, dtype=torch.float32, device=x.device) if want != "planes" else None
    yp = alloc_planes(rows, Cout, fmt, x.device) if want != "f32" else None
    call("ms_conv_cin1_bnact", ptr(x), ptr(wf), ptr(ss[0]), ptr(ss[1]), cfg.slope if cfg.act else 1.0, desc, ptr(y32),
         ptr(yp.t) if yp is not None else None, fmt, yp.ps if yp is not None else 0, stream())
    if y32 is not None:
        if yp is not None:
            y32._ms_planes = yp
        return y32
    return planes_view(yp, oshape)


def conv_block(x, weight, bias, gamma, beta, cfg, packed, bn_buffers, training, residual=None, up2=False, precision=None,
               want="both", row_w=None):
    """weight is the caller's parameter in its own layout/dtype: (Cout, Cin/g, k) or (Cout, Cin/g, kh, kw).
    want ("planes" | "f32" | "both") only matters on the inference fast path (eval mode under torch.no_grad() with a
    tensor-core precision): which forms of the activation the consumers need."""
    prec = precision or _precision
    _need_cuda(x)
    tc_ok = (prec != "fp32" and x.dim() == 4 and weight.shape[1] * cfg.groups == x.shape[3] and
             tc_eligible(cfg, x.shape[0], x.shape[1], x.shape[2], x.shape[3], weight.shape[0], False))
    if tc_ok and not training and not torch.is_grad_enabled():
        w4 = weight.detach()
        return _tc_eval(x, w4.unsqueeze(2) if w4.dim() == 3 else w4, bias, gamma, beta, residual, cfg, packed, bn_buffers,
                        up2, _fmt(prec), want, row_w)
    if row_w is not None:
        raise MixStageError("row weights are an inference-fast-path epilogue (eval mode, no autograd, tensor-core precision)")
    if (prec != "fp32" and not training and not torch.is_grad_enabled() and cfg.has_bn and x.dim() == 4 and x.shape[3] == 1
            and cfg.groups == 1 and weight.shape[0] % 8 == 0 and x.dtype == torch.float32):
        return _cin1_eval(x, weight.detach(), bias, gamma, beta, cfg, packed, bn_buffers, _fmt(prec), want)
    if x.dtype == torch.bfloat16 and not tc_ok:
        x = as_f32(x)                       # planes-only activation feeding a CUDA-core layer
    if residual is not None and residual.dtype == torch.bfloat16:
        residual = as_f32(residual)
    carrier = _Carrier()
    carrier.sinks = (_sink(weight), _sink(bias), _sink(gamma), _sink(beta))
    if weight.dim() == 3:
        weight = weight.unsqueeze(2)        # view: (Cout, Cin/g, 1, k); autograd maps the grad back
    y = _ConvBlock.apply(x, weight, bias, gamma, beta, residual, cfg, packed, bn_buffers, training, up2,
                         (precision or _precision), carrier)
    if carrier.planes is not None:
        y._ms_planes = carrier.planes
    return y


class _Carrier:
    """Hands the bf16 operand planes produced inside an autograd node to the tensor object the caller receives."""
    __slots__ = ("planes", "sinks")

    def __init__(self):
        self.planes = None
        self.sinks = (None, None, None, None)     # .grad buffers of (weight, bias, gamma, beta) for direct accumulation


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
    def forward(ctx, x, idx, soft, emb, rep, carrier=None):
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
        if C % 128 == 0 and sd <= 32 and (C + sd) % 2 == 0:
            # warp-per-row kernel; in the tensor-core modes it also emits the operand planes the conv stacks read
            pl = None
            rs = pad8(C + sd)
            if carrier is not None and _precision != "fp32" and rs - C <= 32:
                pl = alloc_planes(rows, rs, _fmt(_precision), x.device)
                carrier.planes = pl
            call("ms_style_concat_planes_fwd_f32", ptr(x), rows, C, ptr(idx), ptr(soft), rep, ptr(e), dt_code(e.dtype), S, sd,
                 ptr(out), ptr(pl.t) if pl else None, pl.fmt if pl else 0, pl.ps if pl else 0, rs, stream())
        else:
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
        return dx, None, dsoft, demb, None, None


def style_concat(x, emb_weight, idx=None, soft=None, rep=1):
    """cat([x, style_emb(pose_style)], -1) (jlcss.py:175-180).  idx int64 ('emb') or soft fp32 ('lin');
    one style row per `rep` consecutive rows of x."""
    carrier = _Carrier()
    out = _StyleConcat.apply(x, idx, soft, emb_weight, rep, carrier)
    if carrier.planes is not None:
        out._ms_planes = carrier.planes
    return out


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
            acc = arena.take((1,), dev)
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
        acc = arena.take((1,), dev)
        st = stream()
        call("ms_l1_fwd_f32", ptr(a), ptr(b), float(const), n, ptr(acc), None, st)       # no sign tensor: backward re-derives it
        loss = torch.empty((), dtype=torch.float32, device=dev)
        call("ms_scalar_finish", ptr(acc), 1.0 / n, ptr(loss), st)
        ctx.save_for_backward(a, b)
        ctx.n, ctx.const = n, float(const)
        return loss

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = torch.empty_like(a)
        call("ms_l1_bwd_ab_f32", ptr(a), ptr(b), ctx.const, ptr(g.to(torch.float32).contiguous()), ctx.n, ptr(da), stream())
        db = -da if ctx.needs_input_grad[1] else None
        return da, db, None


def l1_mean(a, b=None, const=0.0):
    """mean |a - b| (b tensor) or mean |a - const|: L1Loss(reduction='none') + mean (gan.py:64-75)."""
    return _L1Mean.apply(a, b, const)


# ---------------------------------------------------------------------------- loss bookkeeping of a train step
class LossTerm:
    """A loss of the step before scaling and widening: the fp32 scalar (connected to the autograd graph), a host weight
    and, optionally, the index of a device-resident lambda (TrainStep.lambda_dev) it is multiplied by."""
    __slots__ = ("value", "weight", "lam")

    def __init__(self, value, weight=1.0, lam=None):
        self.value, self.weight, self.lam = value, float(weight), lam


RAW_LOSSES = False           # TrainStep: the modules hand back LossTerm objects; ONE launch scales, widens and sums them


def loss_term(value32, dtype, weight=None, lam_dev=None, lam=None):
    """A module's loss entry.  Default: the reference's tensor -- `value` in the model dtype, times `weight` (a host number)
    or times lam_dev[lam] (a device scalar).  Under TrainStep (RAW_LOSSES): the unscaled fp32 scalar and how to scale it."""
    if RAW_LOSSES:
        return LossTerm(value32, 1.0 if weight is None else weight, lam if lam_dev is not None else None)
    out = cast(value32, dtype)
    if lam_dev is not None:
        return lam_dev[lam] * out
    if weight is not None:
        return out * weight
    return out


class _CombineLosses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, *values):
        weights, lam_idx, lam_dev = spec
        n = len(values)
        vals = [_f32c(v) for v in values]
        dev = vals[0].device
        P = (ctypes.c_void_p * n)(*[v.data_ptr() for v in vals])
        W = (ctypes.c_double * n)(*weights)
        L = (ctypes.c_int * n)(*lam_idx)
        total = torch.empty((), dtype=torch.float32, device=dev)
        report = torch.empty(n, dtype=torch.float64, device=dev)
        call("ms_loss_combine", P, W, L, n, ptr(lam_dev), ptr(total), ptr(report), stream())
        ctx.spec = (W, L, n, lam_dev)
        ctx.mark_non_differentiable(report)
        return total, report

    @staticmethod
    def backward(ctx, gtotal, _greport):
        W, L, n, lam_dev = ctx.spec
        g = torch.empty(n, dtype=torch.float32, device=gtotal.device)
        call("ms_loss_combine_bwd", ptr(gtotal.to(torch.float32).contiguous()), W, L, n, ptr(lam_dev), ptr(g), stream())
        return (None,) + tuple(g[i] if ctx.needs_input_grad[i + 1] else None for i in range(n))


def combine_losses(terms, lam_dev=None):
    """(sum_i w_i * l_i as an fp32 scalar to call backward() on, the n scaled terms as a detached fp64 vector)."""
    if not terms or len(terms) > _lib.LOSS_MAX_TERMS:
        raise MixStageError("combine_losses takes 1..%d terms" % _lib.LOSS_MAX_TERMS)
    for t in terms:
        if not isinstance(t, LossTerm):
            raise MixStageError("combine_losses: a module returned a finished loss tensor while RAW_LOSSES is set")
    if any(t.lam is not None for t in terms) and (lam_dev is None or lam_dev.dtype != torch.float64):
        raise MixStageError("combine_losses: device-resident lambdas must be an fp64 tensor")
    spec = ([t.weight for t in terms], [-1 if t.lam is None else int(t.lam) for t in terms], lam_dev)
    return _CombineLosses.apply(spec, *[t.value for t in terms])


# ---------------------------------------------------------------------------- chains of blocks in one launch
# At batch 16 the train step is bound by the number of dependent launches, not by their arithmetic: a conv stack (UNet1D,
# ClusterClassify, the grouped sub-decoders, AudioEncoder.conv.1-7, PoseStyleEncoder) runs as ONE cooperative launch per
# direction (csrc/conv_train.cu: conv_chain_fwd_kernel / conv_chain_bwd_kernel) and its weight gradients as one more.
CHAINS = _os.environ.get("MS_CHAINS", "1") != "0"


class ChainBlock:
    """What a chain needs of one ConvNormRelu: parameters, static geometry, packed-weight cache, BatchNorm buffers."""
    __slots__ = ("weight", "bias", "gamma", "beta", "cfg", "packed", "bn_buffers")

    def __init__(self, weight, bias, gamma, beta, cfg, packed, bn_buffers):
        self.weight, self.bias, self.gamma, self.beta = weight, bias, gamma, beta
        self.cfg, self.packed, self.bn_buffers = cfg, packed, bn_buffers


class _ChainSpec:
    __slots__ = ("blocks", "res_from", "fmt", "sinks", "out_planes")


def _chain_shapes(blocks, x_shape, res_from):
    """Input shape of every block and the output shape of the chain; None when a block cannot take the tensor-core path."""
    shapes = []
    cur = tuple(x_shape)
    for i, b in enumerate(blocks):
        cfg = b.cfg
        B, H, W, Cin = cur
        w = b.weight
        Cout = w.shape[0]
        if w.shape[1] * cfg.groups != Cin or not cfg.has_bn:
            return None
        if not tc_eligible(cfg, B, H, W, Cin, Cout, True):
            return None
        if (Cout // cfg.groups) % 32 or Cout > 8192:
            return None
        if i > 0 and pad8(Cin) != Cin:
            return None
        Ho, Wo = conv_out(H, cfg.kh, cfg.sh, cfg.ph), conv_out(W, cfg.kw, cfg.sw, cfg.pw)
        shapes.append(cur)
        if res_from[i] is not None:
            if Ho != 1:
                return None
            cur = (B, 1, 2 * Wo, Cout)
        else:
            cur = (B, Ho, Wo, Cout)
    shapes.append(cur)
    return shapes


class _ConvChain(torch.autograd.Function):
    """y = block_{n-1}(... block_0(x)) for training-mode ConvNormRelu blocks; block i with res_from[i] = j produces
    upsample2(act(bn(conv(.)))) + output of block j (UNet1D, layers.py:150-152).  One launch forward, one backward (+ one
    for all weight gradients)."""

    @staticmethod
    def forward(ctx, x, spec, *params):
        blocks, res_from, fmt = spec.blocks, spec.res_from, spec.fmt
        n = len(blocks)
        split = fmt == MS_BF16X2
        npass = 3 if split else 1
        dev = x.device
        if x.dtype != torch.bfloat16:
            x = _f32c(x)
        need_dx0 = ctx.needs_input_grad[0]
        cur = tuple(x.shape)
        xp = planes_of(x, fmt, pad8(cur[3]))
        layers = (_lib.ChainFwdLayer * n)()
        keep, rec = [], []
        flops = 0.0
        global _stats_epoch, last_gemm_flops
        for i, b in enumerate(blocks):
            weight = params[4 * i]
            cfg, packed = b.cfg, b.packed
            B, H, W, Cin = cur
            Cout = weight.shape[0]
            rs = pad8(Cin)
            desc = make_desc(cur, Cout, cfg.kh, cfg.kw, cfg.sh, cfg.sw, cfg.ph, cfg.pw, cfg.groups)
            pf, pd = packed.tc_plans(cur, Cout, cfg, rs, need_dx0 if i == 0 else True, npass)
            wp, wps = packed.get_tc(weight, pf, fmt, cfg.groups)
            wt = packed.get_tc(weight, pd, fmt, cfg.groups) if pd is not None else None
            d = pf.desc
            igemm.set_planes(pf, split, xp.ps, wps, 0)
            d.out_dtype, d.epilogue, d.slope, d.out_numel = _lib.MS_F32, 0, cfg.slope, 0
            d.block_n, d.split_k = _block_plan(pf, npass, True)
            rows = B * desc.Ho * desc.Wo
            zshape = (B, desc.Ho, desc.Wo, Cout)
            z = arena.take_f32(zshape, dev) if d.split_k > 1 else torch.empty(zshape, dtype=torch.float32, device=dev)
            up2 = res_from[i] is not None
            oshape = (B, 1, 2 * desc.Wo, Cout) if up2 else zshape
            rows_out = 2 * rows if up2 else rows
            y = torch.empty(oshape, dtype=torch.float32, device=dev) if i == n - 1 else None
            yp = alloc_planes(rows_out, Cout, fmt, dev)
            ss = torch.empty(4, Cout, dtype=torch.float32, device=dev)
            acc = arena.take((2 * Cout + 2,), dev)
            gamma, beta, cbias = params[4 * i + 2], params[4 * i + 3], params[4 * i + 1]
            if cbias is not None and cbias.dtype != gamma.dtype:
                raise MixStageError("conv bias and BatchNorm parameters must share a dtype")
            bn = _block_bn(Cout, gamma, beta, cbias, b.bn_buffers, cfg, True, acc, ss)
            _stats_epoch += 1
            packed.stats_epoch += 1
            L = layers[i]
            L.d, L.a, L.w, L.z = ctypes.pointer(d), ptr(xp.t), ptr(wp), ptr(z)
            L.bn, L.y, L.planes, L.pfmt, L.pstride = ctypes.pointer(bn), ptr(y), ptr(yp.t), yp.fmt, yp.ps
            L.up2 = 1 if up2 else 0
            if up2:
                rp = rec[res_from[i]]["yp"]
                if rec[res_from[i]]["oshape"] != oshape:
                    raise MixStageError("skip tensor shape %s != %s" % (rec[res_from[i]]["oshape"], oshape))
                L.res, L.res_planes, L.res_pfmt, L.res_pstride = None, ptr(rp.t), rp.fmt, rp.ps
            keep.append((bn, acc))
            fl = 2.0 * rows * Cout * (Cin // cfg.groups) * cfg.kh * cfg.kw
            flops += fl
            rec.append(dict(xp=xp, z=z, ss=ss, desc=desc, pf=pf, pd=pd, wt=wt, rows=rows, Cout=Cout, oshape=oshape, yp=yp,
                            in_shape=cur, rs=rs, up2=up2, flops=fl, gamma_dtype=gamma.dtype,
                            wdt=weight.dtype, bdt=None if cbias is None else cbias.dtype))
            xp, cur = yp, oshape
        sync = arena.take((2,), dev)
        last_gemm_flops = flops
        call("ms_conv_chain_fwd", ctypes.cast(layers, ctypes.c_void_p), n, ptr(sync), stream())
        ctx.spec, ctx.rec, ctx.n = spec, rec, n
        ctx.flops = flops
        spec.out_planes = rec[-1]["yp"]
        # tensors the backward reads stay referenced through ctx.rec (they are not autograd inputs / outputs)
        return y

    @staticmethod
    def backward(ctx, dy):
        spec, rec, n = ctx.spec, ctx.rec, ctx.n
        blocks, res_from, fmt = spec.blocks, spec.res_from, spec.fmt
        split = fmt == MS_BF16X2
        npass = 3 if split else 1
        dy = dy.contiguous()
        dev = dy.device
        st = stream()
        need = ctx.needs_input_grad
        layers = (_lib.ChainBwdLayer * n)()
        keep = []
        grads = [None] * (4 * n)
        dyin = [None] * n                  # incoming gradient of every block's OUTPUT (its main consumer's input gradient)
        dxs = [None] * n
        consumers = {}                     # block j -> blocks whose skip tensor is block j's output
        for i, j in enumerate(res_from):
            if j is not None:
                consumers.setdefault(j, []).append(i)
        global last_gemm_flops
        flops = 0.0
        wg_items = []
        for k, i in enumerate(range(n - 1, -1, -1)):
            r, b = rec[i], blocks[i]
            cfg = b.cfg
            Cout, rows = r["Cout"], r["rows"]
            need_w, need_g, need_be = need[2 + 4 * i], need[2 + 4 * i + 2], need[2 + 4 * i + 3]
            sinks = spec.sinks[i]
            dyin[i] = dy if i == n - 1 else dxs[i + 1]
            dy2 = None
            cons = consumers.get(i, [])
            if len(cons) > 1:
                raise MixStageError("internal: a block's output feeds more than one skip connection")
            if cons:
                dy2 = dyin[cons[0]]
            dzp = alloc_planes(rows, Cout, fmt, dev)
            red = arena.take((2 * Cout + 2,), dev)
            bn = _lib.BlockBn()
            bn.C, bn.pdt, bn.training = Cout, dt_code(r["gamma_dtype"]), 1
            bn.momentum, bn.eps, bn.slope = cfg.momentum, cfg.eps, (cfg.slope if cfg.act else 1.0)
            bn.sums, bn.ss = ptr(red), ptr(r["ss"])
            bn.gamma = bn.beta = ptr(r["ss"])              # not read by the backward; non-NULL for the argument check
            sg = sb = None
            if need_g and need_be:
                sg, sb = sinks[2], sinks[3]
                if sg is None or sb is None:
                    sg = torch.zeros(Cout, dtype=r["gamma_dtype"], device=dev)
                    sb = torch.zeros(Cout, dtype=r["gamma_dtype"], device=dev)
                    grads[4 * i + 2], grads[4 * i + 3] = sg, sb
            elif need_g or need_be:
                raise MixStageError("chain: BatchNorm weight and bias must both (or neither) require gradients")
            B, H, W, Cin = r["in_shape"]
            xrs = r["rs"]
            need_dx = (i > 0) or need[0]
            L = layers[k]
            L.dy, L.dy2, L.z, L.bn = ptr(dyin[i]), ptr(dy2), ptr(r["z"]), ctypes.pointer(bn)
            L.rows, L.up2, L.rows_per_seq = rows, 1 if r["up2"] else 0, r["desc"].Wo
            L.dz_planes, L.pfmt, L.pstride = ptr(dzp.t), fmt, dzp.ps
            L.grad_gamma, L.grad_beta, L.gdt = ptr(sg), ptr(sb), dt_code(r["gamma_dtype"])
            if need_dx:
                pd = r["pd"]
                wt, wtps = r["wt"]
                igemm.set_planes(pd, split, dzp.ps, wtps, 0)
                pd.desc.out_numel = 0
                pd.desc.block_n, pd.desc.split_k = _block_plan(pd, npass, False)
                dshape = (B, H, W, xrs)
                dxf = arena.take_f32(dshape, dev) if pd.desc.split_k > 1 else torch.empty(dshape, dtype=torch.float32, device=dev)
                dxs[i] = dxf
                L.dg, L.wt, L.dx = ctypes.pointer(pd.desc), ptr(wt), ptr(dxf)
                flops += r["flops"]
            keep.append((bn, red, dzp))
            r["dzp"] = dzp
            # d(conv bias) is identically zero under batch-statistics BatchNorm
            if need[2 + 4 * i + 1] and r["bdt"] is not None and sinks[1] is None:
                grads[4 * i + 1] = torch.zeros(Cout, dtype=r["bdt"], device=dev)
            if need_w:
                wg_items.append(i)
        sync = arena.take((2,), dev)
        last_gemm_flops = flops
        call("ms_conv_chain_bwd", ctypes.cast(layers, ctypes.c_void_p), n, ptr(sync), st)
        # ---- weight gradients: dW = dz^T x, every block of the chain in one launch beside the main stream
        direct = [i for i in wg_items if spec.sinks[i][0] is not None and WACC is not None]
        if direct:
            items = (_lib.WgradItem * len(direct))()
            fl = 0.0
            # the blocks share ONE launch: pixel slices chosen so that ~two waves of CTAs carry equal operand bytes (slicing a
            # block as if it had the machine to itself multiplies the 128 KB-per-CTA accumulator reductions instead)
            splits = igemm.wgrad_multi_splits([rec[i]["pf"].desc for i in direct])
            for k, i in enumerate(direct):
                r, b = rec[i], blocks[i]
                pf = r["pf"]
                B, H, W, Cin = r["in_shape"]
                igemm.set_planes(pf, split, r["xp"].ps, 0, r["dzp"].ps)
                nsplit, pf.desc.wgrad_c_tile = splits[k]
                pf.desc.split_k = nsplit
                acc = WACC.acc_for(b.packed, pf.wp_numel, dev)
                WACC.note(acc, spec.sinks[i][0], r["Cout"], Cin // b.cfg.groups, b.cfg.kh * b.cfg.kw, pf.kpad, r["wdt"])
                it = items[k]
                # the descriptor is shared, mutable plan state: the launch below copies it, so one COPY per item
                dcopy = _lib.IgemmDesc.from_buffer_copy(pf.desc)
                keep.append(dcopy)
                it.d, it.x, it.dz, it.acc = ctypes.pointer(dcopy), ptr(r["xp"].t), ptr(r["dzp"].t), ptr(acc)
                fl += r["flops"]
            last_gemm_flops = fl
            if SIDE is not None:
                with SIDE.fork(*([rec[i]["xp"].t for i in direct] + [rec[i]["dzp"].t for i in direct])):
                    call("ms_wgrad_bf16_acc_multi", ctypes.cast(items, ctypes.c_void_p), len(direct), stream())
            else:
                call("ms_wgrad_bf16_acc_multi", ctypes.cast(items, ctypes.c_void_p), len(direct), st)
        for i in wg_items:
            if i in direct:
                continue
            r, b = rec[i], blocks[i]
            pf = r["pf"]
            B, H, W, Cin = r["in_shape"]
            Cin_g = Cin // b.cfg.groups
            igemm.set_planes(pf, split, r["xp"].ps, 0, r["dzp"].ps)
            nsplit, pf.desc.wgrad_c_tile = igemm.wgrad_split(pf.desc, npass=npass)
            pf.desc.split_k = nsplit
            last_gemm_flops = r["flops"]
            dwp = torch.empty(nsplit * pf.wp_numel, dtype=torch.float32, device=dev)
            call("ms_wgrad_bf16", pf.desc, ptr(r["xp"].t), ptr(r["dzp"].t), ptr(dwp), st)
            sink = spec.sinks[i][0]
            taps = b.cfg.kh * b.cfg.kw
            if sink is not None:
                call("ms_unpack_igemm_wgrad", ptr(dwp), r["Cout"], Cin_g, taps, pf.desc.ntaps, pf.kpad, ptr(sink),
                     dt_code(r["wdt"]), nsplit, 1, st)
            else:
                dw = torch.empty((r["Cout"], Cin_g, b.cfg.kh, b.cfg.kw), dtype=r["wdt"], device=dev)
                call("ms_unpack_igemm_wgrad", ptr(dwp), r["Cout"], Cin_g, taps, pf.desc.ntaps, pf.kpad, ptr(dw),
                     dt_code(r["wdt"]), nsplit, 0, st)
                grads[4 * i] = dw
        dx0 = None
        if need[0]:
            Cin0 = rec[0]["in_shape"][3]
            dx0 = dxs[0] if rec[0]["rs"] == Cin0 else dxs[0][..., :Cin0]
        ctx.rec = None
        return (dx0, None) + tuple(grads)


def _chain_eval(blocks, x, shapes, res_from, fmt, last):
    """Inference form of a chain (eval mode, no autograd, small batch): BatchNorm folded from the running statistics inside
    the launch, no statistics, accumulators normalised straight out of TMEM."""
    n = len(blocks)
    split = fmt == MS_BF16X2
    npass = 3 if split else 1
    dev = x.device
    cur = tuple(x.shape)
    xp = planes_of(x, fmt, pad8(cur[3]))
    layers = (_lib.ChainFwdLayer * n)()
    keep, outs = [], []
    flops = 0.0
    for i, b in enumerate(blocks):
        cfg, packed = b.cfg, b.packed
        w4 = b.weight.detach()
        if w4.dim() == 3:
            w4 = w4.unsqueeze(2)
        B, H, W, Cin = cur
        Cout = w4.shape[0]
        pf, _ = packed.tc_plans(cur, Cout, cfg, pad8(Cin), False, npass)
        wp, wps = packed.get_tc(w4, pf, fmt, cfg.groups)
        d = pf.desc
        igemm.set_planes(pf, split, xp.ps, wps, 0)
        d.out_dtype, d.epilogue, d.slope, d.out_numel = _lib.MS_F32, 0, cfg.slope, 0
        d.block_n, d.split_k = _block_plan(pf, npass, True)
        Ho, Wo = conv_out(H, cfg.kh, cfg.sh, cfg.ph), conv_out(W, cfg.kw, cfg.sw, cfg.pw)
        rows = B * Ho * Wo
        up2 = res_from[i] is not None
        oshape = (B, 1, 2 * Wo, Cout) if up2 else (B, Ho, Wo, Cout)
        z = arena.take_f32((B, Ho, Wo, Cout), dev) if d.split_k > 1 else None
        is_last = i == n - 1
        y = torch.empty(oshape, dtype=torch.float32, device=dev) if (is_last and last != "planes") else None
        yp = alloc_planes(2 * rows if up2 else rows, Cout, fmt, dev) if not (is_last and last == "f32") else None
        rm, rv, _ = b.bn_buffers
        bn = _lib.BlockBn()
        bn.C, bn.pdt, bn.training = Cout, dt_code(b.gamma.dtype), 2
        bn.momentum, bn.eps, bn.slope = cfg.momentum, cfg.eps, (cfg.slope if cfg.act else 1.0)
        bn.gamma, bn.beta, bn.conv_bias = ptr(b.gamma), ptr(b.beta), ptr(b.bias)
        bn.running_mean, bn.running_var = ptr(rm), ptr(rv)
        L = layers[i]
        L.d, L.a, L.w, L.z, L.bn = ctypes.pointer(d), ptr(xp.t), ptr(wp), ptr(z), ctypes.pointer(bn)
        L.y, L.planes, L.pfmt, L.pstride = ptr(y), (ptr(yp.t) if yp is not None else None), fmt, (yp.ps if yp is not None else 0)
        L.up2 = 1 if up2 else 0
        if up2:
            rp = outs[res_from[i]][1]
            L.res, L.res_planes, L.res_pfmt, L.res_pstride = None, ptr(rp.t), rp.fmt, rp.ps
        keep.append((bn, z))
        outs.append((y, yp, oshape))
        flops += 2.0 * rows * Cout * (Cin // cfg.groups) * cfg.kh * cfg.kw
        xp, cur = yp, oshape
    sync = arena.take((2,), dev)
    global last_gemm_flops
    last_gemm_flops = flops
    call("ms_conv_chain_fwd", ctypes.cast(layers, ctypes.c_void_p), n, ptr(sync), stream())
    y, yp, oshape = outs[-1]
    if y is not None:
        if yp is not None:
            y._ms_planes = yp
        return y
    return planes_view(yp, oshape)


def _chain_small(blocks, shapes, fmt):
    """True when every block of an inference chain fits the small-batch form (all tiles of a launch resident in TMEM, or
    split-K)."""
    npass = 3 if fmt == MS_BF16X2 else 1
    for b, shp in zip(blocks, shapes[:-1]):
        w = b.weight
        pf, _ = b.packed.tc_plans(shp, w.shape[0], b.cfg, pad8(shp[3]), False, npass)
        bn_, ks = _block_plan(pf, npass, True)
        d = pf.desc
        tiles = igemm._tiles_m(d) * d.num_classes * ((d.class_n + bn_ - 1) // bn_)
        if ks == 1 and not (tiles <= 148 and igemm.block_resident(d, bn_)):
            return False
    return True


def _block_chainable(b, shape, first):
    cfg, w = b.cfg, b.weight
    B, H, W, Cin = shape
    Cout = w.shape[0]
    if not cfg.has_bn or w.shape[1] * cfg.groups != Cin:
        return False
    if not tc_eligible(cfg, B, H, W, Cin, Cout, True):
        return False
    if (Cout // cfg.groups) % 32 or Cout > 8192:
        return False
    return first or pad8(Cin) == Cin


def _out_shape(b, shape):
    cfg = b.cfg
    B, H, W, _ = shape
    return (B, conv_out(H, cfg.kh, cfg.sh, cfg.ph), conv_out(W, cfg.kw, cfg.sw, cfg.pw), b.weight.shape[0])


def _run_chain(blocks, x, training, res_from, fmt, last):
    """One launch for `blocks` (all chainable); None when the inference form does not fit (caller goes block by block)."""
    if training:
        spec = _ChainSpec()
        spec.blocks, spec.res_from, spec.fmt = blocks, res_from, fmt
        spec.sinks = [(_sink(b.weight), _sink(b.bias), _sink(b.gamma), _sink(b.beta)) for b in blocks]
        spec.out_planes = None
        params = []
        for b in blocks:
            w = b.weight
            params += [w.unsqueeze(2) if w.dim() == 3 else w, b.bias, b.gamma, b.beta]
        y = _ConvChain.apply(x, spec, *params)
        if spec.out_planes is not None:
            y._ms_planes = spec.out_planes
        return y
    shapes = _chain_shapes(blocks, x.shape, res_from)
    if shapes is not None and not torch.is_grad_enabled() and _chain_small(blocks, shapes, fmt):
        return _chain_eval(blocks, x, shapes, res_from, fmt, last)
    return None


def conv_chain(blocks, x, training, res_from=None, last="f32", precision=None):
    """Run ChainBlocks one after the other.  Maximal runs of blocks that qualify (tensor-core precision and geometry;
    training-mode BatchNorm, or small-batch inference under no_grad) go out as ONE launch per direction, the rest block by
    block through conv_block.  res_from[i] = j: block i is a UNet decoder step, upsample2(.) + output of block j.
    `last`: form of the final activation on the inference fast path ("planes" | "f32" | "both")."""
    prec = precision or _precision
    n = len(blocks)
    skips = res_from is not None and any(r is not None for r in res_from)
    res_from = list(res_from) if res_from is not None else [None] * n
    _need_cuda(x)
    can_chain = CHAINS and FUSED_BLOCKS and prec != "fp32" and x.dim() == 4 and (training or not torch.is_grad_enabled())
    fmt = _fmt(prec) if prec != "fp32" else None
    if can_chain and skips:
        if n <= _lib.CHAIN_MAX and _chain_shapes(blocks, x.shape, res_from) is not None:
            y = _run_chain(blocks, x, training, res_from, fmt, last)
            if y is not None:
                return y
        can_chain = False
    outs = []
    i = 0
    while i < n:
        j = i
        if can_chain:
            # longest run of chainable blocks starting at i
            shape = tuple(x.shape)
            while j < n and j - i < _lib.CHAIN_MAX and _block_chainable(blocks[j], shape, j == i):
                shape = _out_shape(blocks[j], shape)
                j += 1
        if j - i >= 2:
            seg_last = last
            if j < n:
                # the next block runs on its own: operand planes if it is a tensor-core block, fp32 if it is a CUDA-core one
                nb = blocks[j]
                B_, H_, W_, C_ = shape
                seg_last = "planes" if tc_eligible(nb.cfg, B_, H_, W_, C_, nb.weight.shape[0], False) else "f32"
            y = _run_chain(blocks[i:j], x, training, [None] * (j - i), fmt, seg_last)
            if y is not None:
                x = y
                outs += [None] * (j - i - 1) + [x]
                i = j
                continue
        b = blocks[i]
        want = "planes" if i < n - 1 else last
        if res_from[i] is not None:
            x = conv_block(x, b.weight, b.bias, b.gamma, b.beta, b.cfg, b.packed, b.bn_buffers, training,
                           residual=outs[res_from[i]], up2=True, precision=prec, want=want)
        else:
            x = conv_block(x, b.weight, b.bias, b.gamma, b.beta, b.cfg, b.packed, b.bn_buffers, training, precision=prec, want=want)
        outs.append(x)
        i += 1
    return x
