"""Host-side descriptors for the tcgen05 implicit-GEMM kernel (csrc/conv_tc.cu, ms_igemm_bf16).

A convolution (forward) or its input gradient is described by
  * a 5-D view (channel, w-like, h-parity, h-like, batch) of the bf16 activation,
  * a tap table: per tap the channel offset, w shift, h-parity coordinate and h shift of the box,
  * "classes" of output columns (conv groups, or output parities of a strided conv's input gradient),
  * the weight re-tiling recipe (which source tap feeds each (class, tap)).
Everything here is pure Python on small integers; tests/test_igemm_desc_cpu.py checks the
descriptors against F.conv2d through the CPU specification of the kernel."""
from __future__ import annotations

import ctypes
import os as _os

from ._lib import MS_BF16, MS_F32, MixStageError

from ._lib import IgemmDesc, MAX_CLASSES, MAX_TAPS      # noqa: E402

BLOCK_M, BLOCK_K = 128, 64


def _pow2ceil(n):
    p = 1
    while p < n:
        p <<= 1
    return p


def _boxes(Wo, Ho, rows=BLOCK_M):
    bw = min(rows, _pow2ceil(Wo))
    bh = min(rows // bw, _pow2ceil(Ho))
    bb = rows // (bw * bh)
    return bw, bh, bb


class Plan:
    """A filled descriptor plus the weight re-tiling recipe that goes with it."""

    def __init__(self, desc, mode, srctap, kpad, wp_rows):
        self.desc = desc
        self.mode = mode              # 0 forward, 1 dgrad
        self.srctap = srctap          # list[int]
        self.kpad = kpad
        self.wp_rows = wp_rows        # num_classes * class_n
        self.srctap_c = (ctypes.c_int16 * len(srctap))(*srctap)

    @property
    def wp_numel(self):
        return self.wp_rows * self.desc.ntaps * self.kpad


def fwd_supported(Cin, Cout, groups, sh, sw, H, W, a_row_stride=None):
    cg, ng = Cin // groups, Cout // groups
    if (a_row_stride or Cin) % 8 or ng % 16 or ng < 16:
        return False
    if sh not in (1, 2) or sw not in (1, 2):
        return False
    if (sh == 2 and H % 2) or (sw == 2 and W % 2):
        return False
    if groups > MAX_CLASSES:
        return False
    return True


def dgrad_supported(Cin, Cout, groups, sh, sw, H, W, kh, kw):
    cg, ng = Cin // groups, Cout // groups
    if Cout % 8 or cg % 16 or cg < 16:
        return False
    if sh not in (1, 2) or sw not in (1, 2):
        return False
    if (sh == 2 and (H % 2 or kh % 2)) or (sw == 2 and (W % 2 or kw % 2)):
        return False
    if groups > 1 and (sh != 1 or sw != 1):
        return False
    if groups > MAX_CLASSES:
        return False
    return True


def _split(e, s):
    """input coordinate offset e (in input pixels) -> (shift in s-strided units, parity)."""
    if s == 1:
        return e, 0
    return e // 2, e % 2          # floor semantics


def make_fwd(B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups, Ho, Wo, out_dtype=MS_F32, epilogue=0, slope=1.0,
             a_row_stride=None):
    """Forward conv: A = x (B,H,W,Cin[row stride a_row_stride]) bf16, out (B,Ho,Wo,Cout)."""
    if not fwd_supported(Cin, Cout, groups, sh, sw, H, W, a_row_stride):
        raise MixStageError("igemm forward: unsupported geometry")
    C = a_row_stride or Cin        # elements between consecutive pixels
    cg, ng = Cin // groups, Cout // groups
    d = IgemmDesc()
    Wd, Hd = W // sw, H // sh
    d.a_dims[:] = [Cin * sw if sw == 2 else Cin, Wd, sh, Hd, B]
    if sw == 2:
        d.a_dims[0] = C + Cin      # parity 1 channels start at +C; valid extent C + Cin
    row = C * W                    # elements per image row
    d.a_strides[:] = [1, C * sw, row, row * sh, row * H]
    taps, srctap = [], []
    for th in range(kh):
        fh, par_h = _split(th - ph, sh)
        for tw in range(kw):
            fw, par_w = _split(tw - pw, sw)
            taps.append((par_w * C, fw, par_h, fh))
            srctap.append(th * kw + tw)
    ntaps = len(taps)
    if ntaps > MAX_TAPS:
        raise MixStageError("too many taps")
    for i, t in enumerate(taps):
        d.taps[i][:] = t
    d.ntaps, d.shared_taps = ntaps, 1
    d.cchunks = (cg + BLOCK_K - 1) // BLOCK_K
    d.num_classes, d.class_n = groups, ng
    d.block_n = min(256, ng)
    for g in range(groups):
        d.a_chan_base[g] = g * cg
        d.out_off[g] = g * ng
    d.out_dims[:] = [Wo, Ho, B]
    d.out_strides[:] = [Cout, Wo * Cout, Ho * Wo * Cout]
    bw, bh, bb = _boxes(Wo, Ho)
    d.box[:] = [BLOCK_K, bw, 1, bh, bb]
    d.out_dtype, d.epilogue, d.slope = out_dtype, epilogue, slope
    d.planes = 1
    return Plan(d, 0, srctap, d.cchunks * BLOCK_K, groups * ng)


def make_dgrad(B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups, Ho, Wo, out_row_stride=None):
    """Input gradient: A = dz (B,Ho,Wo,Cout) bf16, out = dx (B,H,W,out_row_stride) fp32.
    With groups == 1 and out_row_stride > Cin the extra columns are produced as zeros
    (their weight rows are zero), e.g. the 266 -> 272 channel padding of the style concat."""
    Co = out_row_stride or Cin
    if groups == 1 and Co != Cin:
        Cin_eff = Co
    else:
        Cin_eff = Cin
    if not dgrad_supported(Cin_eff, Cout, groups, sh, sw, H, W, kh, kw):
        raise MixStageError("igemm dgrad: unsupported geometry")
    cg, ng = Cin_eff // groups, Cout // groups
    d = IgemmDesc()
    d.a_dims[:] = [Cout, Wo, 1, Ho, B]
    d.a_strides[:] = [1, Cout, Cout * Wo, Cout * Wo, Cout * Wo * Ho]

    def dim_taps(k, s, p):
        """per output parity r: list of (source tap j, shift of the dz coordinate)."""
        out = []
        for r in range(s):
            lst = [(j, (r + p - j) // s) for j in range(k) if (r + p - j) % s == 0]
            out.append(lst)
        return out

    th_l, tw_l = dim_taps(kh, sh, ph), dim_taps(kw, sw, pw)
    nt = len(th_l[0]) * len(tw_l[0])
    for lst in th_l:
        assert len(lst) == len(th_l[0])
    for lst in tw_l:
        assert len(lst) == len(tw_l[0])
    parity_classes = [(rh, rw) for rh in range(sh) for rw in range(sw)]
    if groups > 1:
        classes = [(g, 0, 0) for g in range(groups)]
    else:
        classes = [(0, rh, rw) for rh, rw in parity_classes]
    if len(classes) > MAX_CLASSES or (1 if groups > 1 else len(classes)) * nt > MAX_TAPS:
        raise MixStageError("igemm dgrad: tap table too large")
    srctap = []
    shared = 1 if groups > 1 else 0
    rows = []
    for qi, (g, rh, rw) in enumerate(classes):
        tl = [(jh * kw + jw, dw_, dh_) for (jh, dh_) in th_l[rh] for (jw, dw_) in tw_l[rw]]
        for ti, (src, dw_, dh_) in enumerate(tl):
            srctap.append(src)
            if not shared or qi == 0:
                d.taps[(0 if shared else qi * nt) + ti][:] = (0, dw_, 0, dh_)
        d.a_chan_base[qi] = g * ng
        d.out_off[qi] = g * cg + rh * W * Co + rw * Co
    d.ntaps, d.shared_taps = nt, shared
    d.cchunks = (ng + BLOCK_K - 1) // BLOCK_K
    d.num_classes, d.class_n = len(classes), cg
    d.block_n = min(256, cg)
    Wd, Hd = W // sw, H // sh
    d.out_dims[:] = [Wd, Hd, B]
    d.out_strides[:] = [Co * sw, Co * W * sh, Co * W * H]
    bw, bh, bb = _boxes(Wd, Hd)
    d.box[:] = [BLOCK_K, bw, 1, bh, bb]
    d.out_dtype, d.epilogue, d.slope = MS_F32, 0, 1.0
    d.planes = 1
    return Plan(d, 1, srctap, d.cchunks * BLOCK_K, len(classes) * cg)


def pick_block_n(d, npass=1, sms=148):
    """Output-column tile.  Wide tiles re-read the activation box least (A is 16 KB per k-step whatever the tile
    width), so: the widest tile that fills the machine on its own; else the widest tile for which split-K
    (slices of >= 4 k-steps) still reaches ~one CTA per SM; else the narrowest."""
    tiles_m = 1
    for dim, box in ((d.out_dims[0], d.box[1]), (d.out_dims[1], d.box[3]), (d.out_dims[2], d.box[4])):
        tiles_m *= (dim + box - 1) // box
    cands = [min(256, d.class_n)] + [b for b in (128, 64, 32) if b < min(256, d.class_n)]
    num_k = d.ntaps * d.cchunks * npass
    for bn in cands:
        ctas = tiles_m * d.num_classes * ((d.class_n + bn - 1) // bn)
        if ctas >= sms or ctas * max(1, num_k // 4) >= (3 * sms) // 4:
            return bn
    return cands[-1]


def _tiles_m(d, box=None):
    bw, bh, bb = box or (d.box[1], d.box[3], d.box[4])
    return ((d.out_dims[0] + bw - 1) // bw) * ((d.out_dims[1] + bh - 1) // bh) * ((d.out_dims[2] + bb - 1) // bb)


def _normalise_split(total, split):
    split = max(1, min(split, total))
    per = (total + split - 1) // split
    return (total + per - 1) // per


def igemm_split(d, npass, sms=148):
    """Split-K factor of a forward / input-gradient launch: slices of >= 4 k-steps until ~one CTA per SM."""
    ctas = _tiles_m(d) * d.num_classes * ((d.class_n + d.block_n - 1) // d.block_n)
    num_k = d.ntaps * d.cchunks * npass
    if ctas * 2 > sms or num_k < 8:
        return 1
    return _normalise_split(num_k, min(sms // ctas, num_k // 4))          # one wave: at most one CTA per SM


_PLAN_RATE = float(_os.environ.get("MS_PLAN_RATE", "40"))
_PLAN_SPLIT_COST = float(_os.environ.get("MS_PLAN_SPLIT_COST", "4500"))


def block_plan(d, npass, stats, sms=148):
    """(block_n, split_k) of the GEMM phase of a fused block (csrc/conv_train.cu) by a cycle estimate.

    One k-step stages A (16 KB per plane) and W (block_n * 128 B per plane) once for all MMA passes, so it is bound by
    max(tensor time 2 * block_n * npass clk, operand bytes / (L2 -> SM rate)); the rate is ~6300 B/clk chip-wide
    (B300_MICROARCH.md) and ~40 B/clk for one SM (measured).  Full-K tiles (split 1) cost one pipeline fill per tile and ONE
    device barrier (statistics come out of the accumulators); split-K costs the reductions into z, a second barrier and
    the statistics pass over z, and pays when few tiles face a long K (audio encoder tail, UNet bottleneck).
    stats: forward of a training block (full-K needs 32-column chunks); False for the input-gradient GEMM, whose split
    form has no statistics pass."""
    tiles_m = _tiles_m(d)
    k_steps = d.ntaps * d.cchunks
    planes = 2 if npass > 1 else 1
    gran = 32 if stats else 16
    force = _os.environ.get("MS_BLOCK_FORCE", "")        # experiments: "full" / "split" forward form whatever the estimate says
    cands = [b for b in (256, 128, 64, 32) if b <= d.class_n and b % gran == 0]
    if min(256, d.class_n) not in cands and d.class_n % gran == 0 and d.class_n <= 256:
        cands.insert(0, d.class_n)
    if not cands:
        cands = [min(256, d.class_n)]
    best = None
    for bn in cands:
        tiles = tiles_m * d.num_classes * ((d.class_n + bn - 1) // bn)
        stage = planes * (16384 + bn * 128)
        mma = 2 * bn * npass
        for split in sorted({1} | {_normalise_split(k_steps, s) for s in (2, 3, 4, 6, 8, 12, 16, 24) if tiles * s <= 2 * sms}):
            items = tiles * split
            active = min(items, sms)
            rate = min(_PLAN_RATE, 6300.0 / active)      # measured: one SM's TMA path sustains ~40 B/clk (tools/chain_phases.py)
            kclk = max(mma, stage / rate)
            if 196608 // stage < 3:
                kclk *= 1.4                       # a two-stage ring does not cover the TMA latency
            waves = (items + sms - 1) // sms
            per = (k_steps + split - 1) // split
            cost = waves * (per * kclk + 2500)
            if split > 1:
                cost += waves * (128 * bn * 4 / 40.0)
                cost += _PLAN_SPLIT_COST if stats else 1500
            if stats and force == "full" and split > 1:
                continue
            if stats and force == "split" and split == 1 and k_steps >= 4:
                cost += 1e9
            if best is None or cost < best[0]:
                best = (cost, bn, split)
    return best[1], best[2]


def block_resident(d, block_n, sms=148):
    """True when every CTA of a full-K fused launch keeps all of its tiles in TMEM (<= 512 columns, <= 16 tiles)."""
    tiles = _tiles_m(d) * d.num_classes * ((d.class_n + block_n - 1) // block_n)
    per_cta = (tiles + min(tiles, sms) - 1) // min(tiles, sms)
    return per_cta <= 16 and per_cta * block_n <= 512


def wgrad_tiles(d, c_tile=256):
    """(128 output channels) x (c_tile input channels) tiles per (class, tap) of a weight gradient."""
    kpad = d.cchunks * BLOCK_K
    return ((d.class_n + 127) // 128) * ((kpad + c_tile - 1) // c_tile) * d.num_classes * d.ntaps


def wgrad_row_tiles(d):
    """64-pixel-row tiles (the k-steps) of a weight gradient."""
    bw, bh, bb = d.box[1], d.box[3], d.box[4]
    if bb > 1:
        bb //= 2
    elif bh > 1:
        bh //= 2
    else:
        bw //= 2
    return _tiles_m(d, (bw, bh, bb))


def wgrad_multi_splits(descs, sms=148):
    """Pixel-slice counts (and c_tile = 256) for the weight gradients of a chain that share ONE launch: about two waves of
    CTAs in total, every CTA with roughly the same operand bytes to stream, at least two k-steps per CTA (each CTA ends with
    a 128 x c_tile fp32 reduction into the accumulator, which more slicing only multiplies)."""
    per_tile = []
    for d in descs:
        kpad = d.cchunks * BLOCK_K
        ct = min(256, kpad)
        rt = wgrad_row_tiles(d)
        per_tile.append((wgrad_tiles(d, 256), rt, rt * (128 + ct) * 128.0))
    total = sum(t * w for t, _, w in per_tile)
    target = max(total / (2.0 * sms), 1.0)
    out = []
    for t, rt, w in per_tile:
        split = int(round(w / target))
        split = max(1, min(split, max(1, rt // 2)))
        out.append((_normalise_split(rt, split), 256))
    return out


def wgrad_split(d, sms=148, npass=1):
    """(split, c_tile) of a weight-gradient launch.  Tiles are (128 output channels) x (c_tile input channels) per
    (class, tap); the 64-pixel-row tiles are sliced `split` ways and every slice writes its own partial dWp.
    Chosen by a cycle estimate: a k-step (64 pixel rows) is bound by the L2 -> SM operand traffic, (128 + c_tile) * 128
    bytes at ~42 B/clk/SM (B300_MICROARCH.md: ~6300 B/clk LTS cap over 148 SMs), so wide tiles do 2x the work per byte
    of narrow ones; CTAs run in waves of `sms`; every partial costs one more pass of the summing kernel."""
    bw, bh, bb = d.box[1], d.box[3], d.box[4]
    if bb > 1:
        bb //= 2
    elif bh > 1:
        bh //= 2
    else:
        bw //= 2
    total_rt = _tiles_m(d, (bw, bh, bb))
    kpad = d.cchunks * BLOCK_K
    wp_numel = d.num_classes * d.class_n * d.ntaps * kpad
    best = None
    for c_tile in (256, 128, 64):
        if c_tile > kpad and c_tile != 64 and kpad <= c_tile // 2:
            continue
        tiles = ((d.class_n + 127) // 128) * ((kpad + c_tile - 1) // c_tile) * d.num_classes * d.ntaps
        cands = {1}
        for waves in (1, 2):
            cands.add(_normalise_split(total_rt, max(1, (waves * sms) // tiles)))
        for split in sorted(cands):
            ksteps = ((total_rt + split - 1) // split) * npass
            waves = (tiles * split + sms - 1) // sms
            kstep_clk = max(2 * min(c_tile, kpad), 3.05 * (128 + min(c_tile, kpad)))
            cost = waves * (ksteps * kstep_clk + 6000) + split * wp_numel * 4 / 1500.0
            if best is None or cost < best[0]:
                best = (cost, split, c_tile)
    return best[1], best[2]


def set_planes(plan, split, a_plane_stride=0, w_plane_stride=0, out_plane_stride=0):
    """Operand format of a plan: split=False plain bf16, split=True split-bf16 (hi/lo planes, 3 MMA passes)."""
    d = plan.desc
    d.planes = 2 if split else 1
    d.a_plane_stride, d.w_plane_stride, d.out_plane_stride = a_plane_stride, w_plane_stride, out_plane_stride
    return plan
