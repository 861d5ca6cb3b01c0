"""Kernel-level parity on the GPU: every C-ABI entry point against its torch-CPU
specification (tests/cpu_emu.py) on the shapes the generator uses,
including the ragged ones (C_in=1, C_in=266, N=1, N=25, L=1, 3x8 kernel, stride 2)."""
import numpy as np
import pytest
import torch

import cpu_emu
from mixstage_b200 import _lib, ops
from mixstage_b200._lib import ConvDesc, ptr

pytestmark = pytest.mark.gpu
DEV = "cuda"


def run_both(name, args):
    """args: list of python scalars / ConvDesc / (tensor, 'in'|'out'|'inout') tuples.  Runs the CUDA
    entry point on device copies and the CPU spec on host copies; returns ([gpu outs], [cpu outs])."""
    gpu_args, cpu_args, outs_g, outs_c = [], [], [], []
    keep = []
    for a in args:
        if isinstance(a, tuple):
            t, role = a
            if t is None:
                gpu_args.append(None)
                cpu_args.append(None)
                continue
            tc = t.clone().contiguous()
            tg = t.clone().contiguous().to(DEV)
            keep += [tc, tg]
            gpu_args.append(ptr(tg))
            cpu_args.append(ptr(tc))
            if role != "in":
                outs_g.append(tg)
                outs_c.append(tc)
        else:
            gpu_args.append(a)
            cpu_args.append(a)
    st = torch.cuda.current_stream().cuda_stream
    _lib.call(name, *gpu_args, st)
    getattr(cpu_emu, name)(*cpu_args, None)
    torch.cuda.synchronize()
    return [t.cpu() for t in outs_g], outs_c


def close(g, c, tol=2e-5):
    scale = float(c.double().abs().max()) + 1e-30
    err = float((g.double() - c.double()).abs().max())
    assert err <= tol * scale + 1e-7, (err, scale)


CONVS = [
    # B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups
    (2, 16, 64, 1, 64, 3, 3, 1, 1, 1, 1, 1),        # audio_encoder.conv.0
    (2, 16, 32, 64, 64, 4, 4, 2, 2, 1, 1, 1),       # conv.1 (stride 2)
    (2, 8, 8, 256, 256, 3, 8, 1, 1, 1, 3, 1),       # conv.7 (3x8, pad (1,3))
    (3, 1, 64, 256, 256, 3, 1 and 3, 1, 1, 0, 1, 1),  # placeholder replaced below
]
CONVS[3] = (3, 1, 64, 256, 256, 1, 3, 1, 1, 0, 1, 1)      # unet / classify k3
CONVS += [
    (3, 1, 64, 256, 256, 1, 4, 1, 2, 0, 1, 1),      # unet down k4 s2
    (2, 1, 2, 256, 25, 1, 4, 1, 2, 0, 1, 1),        # pose_style_encoder.conv.6 (L 2 -> 1, N=25)
    (2, 1, 64, 266, 2048, 1, 3, 1, 1, 0, 1, 1),     # decoder.0 as dense conv
    (2, 1, 32, 2048, 2048, 1, 3, 1, 1, 0, 1, 8),    # decoder.1-3 grouped
    (2, 1, 64, 2048, 768, 1, 1, 1, 1, 0, 0, 8),     # logits grouped 1x1, N=96 per group
    (2, 1, 64, 256, 8, 1, 1, 1, 1, 0, 0, 1),        # classify logits N=8
    (2, 1, 15, 256, 1, 1, 4, 1, 1, 0, 0, 1),        # D.logits N=1, p=0
    (2, 1, 16, 128, 256, 1, 4, 1, 1, 0, 1, 1),      # D.conv3 k4 s1 p1 (16 -> 15)
    (2, 1, 64, 96, 64, 1, 4, 1, 2, 0, 1, 1),        # D.conv1
    (3, 5, 7, 1, 16, 3, 3, 1, 1, 1, 1, 1),          # C_in=1 3x3 row kernel: odd width, short height, 2 channel groups
    (1, 1, 300, 256, 5, 1, 1, 1, 1, 0, 0, 1),       # 1x1 over 256 channels, N=5 (< 8), rows not a multiple of the grid
    (3, 5, 37, 1, 64, 3, 3, 1, 1, 1, 1, 1),         # C_in=1 3x3, 64 channels (register-accumulator wgrad): ragged 16-pixel segments
    (2, 1, 16, 1, 64, 3, 3, 1, 1, 1, 1, 1),         # the same with a single row (no row above or below)
]


def _desc(c):
    B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, g = c
    return ConvDesc(B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, g, ops.conv_out(H, kh, sh, ph), ops.conv_out(W, kw, sw, pw))


@pytest.mark.parametrize("c", CONVS)
def test_conv_fwd_dgrad_wgrad(c):
    torch.manual_seed(0)
    d = _desc(c)
    B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, g = c
    x = torch.randn(B, H, W, Cin)
    w = torch.randn(Cout, Cin // g, kh, kw, dtype=torch.float64) / (Cin // g * kh * kw) ** 0.5
    bias = torch.randn(Cout)
    n = w.numel()
    (wf, wt), (wf_c, wt_c) = run_both("ms_pack_conv_weight_f32", [(w, "in"), 1, d, (torch.zeros(n), "out"), (torch.zeros(n), "out")])
    assert torch.equal(wf, wf_c) and torch.equal(wt, wt_c)
    y = torch.zeros(B, d.Ho, d.Wo, Cout)
    for act in (0, 1):
        (yg,), (yc,) = run_both("ms_conv_fwd_f32", [(x, "in"), (wf, "in"), (bias, "in"), (y, "out"), d, act, 0.2])
        close(yg, yc)
    dy = torch.randn(B, d.Ho, d.Wo, Cout)
    (dxg,), (dxc,) = run_both("ms_conv_dgrad_f32", [(dy, "in"), (wt, "in"), (torch.zeros_like(x), "out"), d])
    close(dxg, dxc)
    (dwg,), (dwc,) = run_both("ms_conv_wgrad_f32", [(x, "in"), (dy, "in"), (torch.zeros(n), "out"), d])
    close(dwg, dwc, 5e-5)
    (g64,), (c64,) = run_both("ms_unpack_conv_wgrad", [(dwc, "in"), d, (torch.zeros(n, dtype=torch.float64), "out"), 1, 0])
    (g64a,), (c64a,) = run_both("ms_unpack_conv_wgrad", [(dwc, "in"), d, (torch.ones(n, dtype=torch.float64), "inout"), 1, 1])
    close(g64a, c64a, 1e-9)
    close(g64a - 1.0, c64, 1e-6)
    assert torch.equal(g64, c64)


@pytest.mark.parametrize("c,fmt", [((2, 16, 64, 1, 64, 3, 3, 1, 1, 1, 1, 1), 2), ((3, 5, 7, 1, 16, 3, 3, 1, 1, 1, 1, 1), 3),
                                   ((2, 6, 10, 1, 24, 3, 3, 1, 1, 1, 1, 1), 3)])
def test_conv_cin1_bnact(c, fmt):
    """audio_encoder.conv.0 in eval mode: conv + folded BatchNorm + LeakyReLU in one kernel, fp32 and bf16 planes out
    (fmt 2 = MS_BF16, 3 = MS_BF16X2); the third case (24 channels = 3 groups) takes the generic C_in=1 kernel."""
    torch.manual_seed(4)
    d = _desc(c)
    B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, g = c
    assert fmt in (_lib.MS_BF16, _lib.MS_BF16X2)
    x = torch.randn(B, H, W, 1)
    w = torch.randn(Cout, 1, kh, kw, dtype=torch.float64) / 3.0
    n = w.numel()
    (wf, wt), _ = run_both("ms_pack_conv_weight_f32", [(w, "in"), 1, d, (torch.zeros(n), "out"), (torch.zeros(n), "out")])
    scale, shift = torch.rand(Cout) + 0.5, torch.randn(Cout) * 0.2
    numel = B * H * W * Cout
    ps = (numel + 7) // 8 * 8
    planes = torch.zeros(2 * ps, dtype=torch.bfloat16)
    (yg, pg), (yc, pc) = run_both("ms_conv_cin1_bnact", [(x, "in"), (wf, "in"), (scale, "in"), (shift, "in"), 0.2, d,
                                                         (torch.zeros(numel), "out"), (planes, "out"), fmt, ps])
    close(yg, yc)
    # planes: hi within one bf16 ulp of the CPU spec (the fp32 values differ in the last bits), hi + lo ~ fp32 value
    assert float((pg[:numel].float() - pc[:numel].float()).abs().max()) <= 2 ** -7 * float(yc.abs().max())
    if fmt == _lib.MS_BF16X2:
        rec = pg[:numel].float() + pg[ps:ps + numel].float()
        assert float((rec - yg).abs().max()) <= 2 ** -15 * float(yc.abs().max()) + 1e-7


@pytest.mark.parametrize("rows,C,L,up2", [(1024, 256, 64, 0), (64, 256, 2, 1), (32, 25, 1, 0), (2048, 2048, 64, 0), (8192, 64, 64, 0),
                                               (4100, 128, 4100, 0), (4097, 256, 4097, 0)])
def test_batchnorm_chain(rows, C, L, up2):
    torch.manual_seed(1)
    x = torch.randn(rows, C) * 1.7 + 0.3
    s, ss = torch.zeros(C, dtype=torch.float64), torch.zeros(C, dtype=torch.float64)
    (sg, ssg), (sc, ssc) = run_both("ms_col_stats_f32", [(x, "in"), rows, C, (s, "inout"), (ss, "inout")])
    close(sg, sc, 1e-9)
    close(ssg, ssc, 1e-9)
    for pdt, dt in ((0, torch.float32), (1, torch.float64)):
        gamma, beta = torch.rand(C, dtype=dt) + 0.5, torch.randn(C, dtype=dt) * 0.1
        rm, rv = torch.randn(C, dtype=dt) * 0.1, torch.rand(C, dtype=dt) + 0.5
        cbias = torch.randn(C, dtype=dt) * 0.2
        for training in (1, 0):
            for cb in (None, cbias):
                outs = [(torch.zeros(C), "out") for _ in range(4)]
                g, c = run_both("ms_bn_finalize", [(sc, "in"), (ssc, "in"), rows, C, (gamma, "in"), (beta, "in"), (cb, "in"),
                                                   (rm, "inout"), (rv, "inout"), pdt, training, 0.1, 1e-5] + outs)
                for a, b in zip(g, c):
                    close(a, b, 1e-6)
    outs = [(torch.zeros(C), "out") for _ in range(4)]
    g, c = run_both("ms_bn_finalize", [(sc, "in"), (ssc, "in"), rows, C, (gamma, "in"), (beta, "in"), (None, "in"),
                                       (rm, "inout"), (rv, "inout"), 1, 1, 0.1, 1e-5] + outs)
    scale, shift, mean, rstd = c[2], c[3], c[4], c[5]
    # statistics + finalize + batch counter in one launch == the two-kernel sequence
    s0, ss0 = torch.zeros(C, dtype=torch.float64), torch.zeros(C, dtype=torch.float64)
    outs = [(torch.zeros(C), "out") for _ in range(4)]
    g2, c2 = run_both("ms_bn_stats_finalize", [(x, "in"), rows, C, (s0, "inout"), (ss0, "inout"), (torch.zeros(2, dtype=torch.int32), "inout"),
                                               (gamma, "in"), (beta, "in"), (cbias, "in"), (rm, "inout"), (rv, "inout"),
                                               (torch.full((1,), 5, dtype=torch.int64), "inout"), 1, 0.1, 1e-5] + outs)
    assert int(g2[5]) == 6
    for a, b in zip(g2[6:], c2[6:]):
        close(a, b, 1e-6)
    close(g2[3], c2[3], 1e-6)
    close(g2[4], c2[4], 1e-6)
    res = torch.randn(rows * (2 if up2 else 1), C)
    y = torch.zeros_like(res)
    (yg,), (yc,) = run_both("ms_bn_act_fwd_f32", [(x, "in"), (scale, "in"), (shift, "in"), 0.2, rows, C, (y, "out"),
                                                  (res if up2 else None, "in"), up2, L, (None, "in"), 0, 0])
    close(yg, yc, 1e-6)
    # bf16 operand planes as a second output (single plane and hi/lo split): bit-exact vs the specification
    n = y.numel()
    ps = (n + 7) // 8 * 8
    for pfmt in (2, 3):
        pl = torch.zeros(2 * ps, dtype=torch.bfloat16)
        (yg2, pg), (yc2, pc) = run_both("ms_bn_act_fwd_f32", [(x, "in"), (scale, "in"), (shift, "in"), 0.2, rows, C, (y, "out"),
                                                              (res if up2 else None, "in"), up2, L, (pl, "out"), pfmt, ps])
        hi_g, hi_c = pg[:n].float(), pc[:n].float()
        assert float((hi_g - yg2.reshape(-1)).abs().max()) <= 2 ** -8 * float(yg2.abs().max())
        if pfmt == 3:
            rec = hi_g + pg[ps:ps + n].float()
            assert float((rec - yg2.reshape(-1)).abs().max()) <= 2 ** -15 * float(yg2.abs().max())
        close(hi_g, hi_c, 1e-2)
    if C % 8 == 0:
        pl = torch.zeros(2 * ps, dtype=torch.bfloat16)
        (pg,), (pc,) = run_both("ms_to_planes", [(res, "in"), res.shape[0], C, C, (pl, "out"), 3, ps])
        assert torch.equal(pg, pc)
        rs = C + 8
        ps2 = res.shape[0] * rs
        pl = torch.full((2 * ps2,), 7.0, dtype=torch.bfloat16)
        (pg,), (pc,) = run_both("ms_to_planes", [(res, "in"), res.shape[0], C, rs, (pl, "out"), 3, ps2])
        assert torch.equal(pg, pc)
        assert float(pg[:ps2].view(-1, rs)[:, C:].abs().max()) == 0.0
    dy = torch.randn_like(res)
    z2 = [(torch.zeros(C, dtype=torch.float64), "inout") for _ in range(2)]
    (dgg, dbg), (dgc, dbc) = run_both("ms_bn_act_bwd_reduce_f32", [(dy, "in"), (x, "in"), (scale, "in"), (shift, "in"), (mean, "in"),
                                                                     (rstd, "in"), 0.2, rows, C, up2, L] + z2)
    close(dgg, dgc, 1e-5)
    close(dbg, dbc, 1e-5)
    for training in (1, 0):
        (dxg,), (dxc,) = run_both("ms_bn_act_bwd_apply_f32", [(dy, "in"), (x, "in"), (scale, "in"), (shift, "in"), (mean, "in"),
                                                                (rstd, "in"), 0.2, rows, C, up2, L, (dgc, "in"), (dbc, "in"),
                                                                training, (torch.zeros(rows, C), "out"), (None, "in"), 0, 0,
                                                                (None, "in"), (None, "in"), 0])
        close(dxg, dxc, 1e-5)
        psd = (rows * C + 7) // 8 * 8
        (dxg2, pg, gg, gb), (dxc2, pc, cg_, cb_) = run_both("ms_bn_act_bwd_apply_f32", [(dy, "in"), (x, "in"), (scale, "in"), (shift, "in"), (mean, "in"),
                                                                       (rstd, "in"), 0.2, rows, C, up2, L, (dgc, "in"), (dbc, "in"),
                                                                       training, (torch.zeros(rows, C), "out"),
                                                                       (torch.zeros(2 * psd, dtype=torch.bfloat16), "out"), 3, psd,
                                                                       (torch.ones(C, dtype=torch.float64), "inout"),
                                                                       (torch.ones(C, dtype=torch.float64), "inout"), 1])
        close(gg, 1.0 + dgc, 1e-9)           # affine-parameter gradients accumulated into their buffers
        close(gb, 1.0 + dbc, 1e-9)
        close(gg, cg_, 1e-9)
        rec = pg[:rows * C].float() + pg[psd:psd + rows * C].float()
        assert float((rec - dxg2.reshape(-1)).abs().max()) <= 2 ** -15 * float(dxg2.abs().max()) + 1e-30


def test_small_ops():
    torch.manual_seed(2)
    B, T, P, K, S, sd, C = 4, 64, 96, 8, 25, 10, 256
    # bilinear (8,7)->(64,1) and (32,7)->(256,1)
    for Hi, Tt in ((8, 64), (32, 256), (3, 64)):
        x = torch.randn(2, Hi, 7, 64)
        (yg,), (yc,) = run_both("ms_bilinear_to_T_fwd_f32", [(x, "in"), 2, Hi, 7, 64, Tt, (torch.zeros(2, Tt, 64), "out")])
        close(yg, yc, 1e-6)
        dy = torch.randn(2, Tt, 64)
        (dg,), (dc,) = run_both("ms_bilinear_to_T_bwd_f32", [(dy, "in"), 2, Hi, 7, 64, Tt, (torch.zeros_like(x), "out")])
        close(dg, dc, 1e-5)
    # style concat, emb and lin modes, fp64 table
    rows = B * T
    x = torch.randn(rows, C)
    emb = torch.randn(S, sd, dtype=torch.float64)
    idx = torch.randint(0, S, (B,))
    soft = torch.softmax(torch.randn(B, S), -1)
    for (i, s_) in ((idx, None), (None, soft)):
        (og,), (oc,) = run_both("ms_style_concat_fwd_f32", [(x, "in"), rows, C, (i, "in"), (s_, "in"), T, (emb, "in"), 1, S, sd,
                                                            (torch.zeros(rows, C + sd), "out")])
        close(og, oc, 1e-6)
        dout = torch.randn(rows, C + sd)
        g, c = run_both("ms_style_concat_bwd_f32", [(dout, "in"), rows, C, (i, "in"), (s_, "in"), T, (emb, "in"), 1, S, sd,
                                                    (torch.zeros(rows, C), "out"), (torch.zeros(S, sd), "inout"),
                                                    (torch.zeros(B, S) if s_ is not None else None, "out")])
        for a, b in zip(g, c):
            close(a, b, 1e-5)
    # per-frame style index (rep=1)
    idx2 = torch.randint(0, S, (rows,))
    (og,), (oc,) = run_both("ms_style_concat_fwd_f32", [(x, "in"), rows, C, (idx2, "in"), (None, "in"), 1, (emb, "in"), 1, S, sd,
                                                        (torch.zeros(rows, C + sd), "out")])
    close(og, oc, 1e-6)
    # warp-per-row form with operand planes (row stride 272 for the 266 concatenated channels), both plane formats
    if C % 128 == 0:
        rs = (C + sd + 7) // 8 * 8
        ps = rows * rs
        for (i, s_) in ((idx, None), (None, soft)):
            for pfmt in (2, 3):
                g, c = run_both("ms_style_concat_planes_fwd_f32",
                                [(x, "in"), rows, C, (i, "in"), (s_, "in"), T, (emb, "in"), 1, S, sd, (torch.zeros(rows, C + sd), "out"),
                                 (torch.full(((2 if pfmt == 3 else 1) * ps,), 7.0, dtype=torch.bfloat16), "out"), pfmt, ps, rs])
                close(g[0], c[0], 1e-6)
                assert torch.equal(g[1].float(), c[1].float())            # planes: same roundings, zero-filled padding columns
    # softmax / CE / argmax
    for Kk, r, trep in ((8, rows, 1), (25, B, 1), (4, 64, 16)):
        score = torch.randn(r, Kk) * 3
        tgt = torch.randint(0, Kk, (r // trep,))
        g, c = run_both("ms_softmax_ce_fwd_f32", [(score, "in"), r, Kk, (tgt, "in"), trep, (torch.zeros(r, Kk), "out"),
                                                  (torch.zeros(r, dtype=torch.int64), "out"), (torch.zeros(1, dtype=torch.float64), "inout")])
        close(g[0], c[0], 1e-6)
        assert torch.equal(g[1], c[1])
        close(g[2], c[2], 1e-6)
        gce, dsoft = torch.tensor([0.7]), torch.randn(r, Kk)
        (dg,), (dc,) = run_both("ms_softmax_ce_bwd_f32", [(c[0], "in"), r, Kk, (tgt, "in"), trep, (gce, "in"), (dsoft, "in"),
                                                          (torch.zeros(r, Kk), "out")])
        close(dg, dc, 1e-5)
    # mixture
    z, w, dout = torch.randn(rows, K * P), torch.softmax(torch.randn(rows, K), -1), torch.randn(rows, P)
    (og,), (oc,) = run_both("ms_mixture_fwd_f32", [(z, "in"), (w, "in"), rows, K, P, (torch.zeros(rows, P), "out")])
    close(og, oc, 1e-6)
    g, c = run_both("ms_mixture_bwd_f32", [(dout, "in"), (z, "in"), (w, "in"), rows, K, P, (torch.zeros(rows, K * P), "out"),
                                           (torch.zeros(rows, K), "out")])
    close(g[0], c[0], 1e-6)
    close(g[1], c[1], 1e-5)
    # mean rows, velocity, l1
    x3 = torch.randn(B, 4, 25)
    (og,), (oc,) = run_both("ms_mean_rows_fwd_f32", [(x3, "in"), B, 4, 25, (torch.zeros(B, 25), "out")])
    close(og, oc, 1e-6)
    (og,), (oc,) = run_both("ms_mean_rows_bwd_f32", [(torch.randn(B, 25), "in"), B, 4, 25, (torch.zeros(B, 4, 25), "out")])
    close(og, oc, 1e-6)
    xp = torch.randn(B, T, P)
    (og,), (oc,) = run_both("ms_velocity_fwd_f32", [(xp, "in"), B, T, P, (torch.zeros(B, T, P), "out")])
    assert torch.equal(og, oc)
    (og,), (oc,) = run_both("ms_velocity_bwd_f32", [(xp, "in"), B, T, P, (torch.zeros(B, T, P), "out")])
    assert torch.equal(og, oc)
    n = B * T * P
    for b_, cst in ((torch.randn(n), 0.0), (None, 1.0)):
        g, c = run_both("ms_l1_fwd_f32", [(xp.reshape(-1), "in"), (b_, "in"), cst, n, (torch.zeros(1, dtype=torch.float64), "inout"),
                                          (torch.zeros(n), "out")])
        close(g[0], c[0], 1e-9)
        assert torch.equal(g[1], c[1])
    (og,), (oc,) = run_both("ms_l1_bwd_f32", [(c[1], "in"), (torch.tensor([0.3]), "in"), n, (torch.zeros(n), "out")])
    close(og, oc, 1e-6)
    # forward without the sign tensor (vector loads) and the backward that re-derives the sign from the operands
    for b_, cst in ((torch.randn(n), 0.0), (None, 1.0)):
        g, c = run_both("ms_l1_fwd_f32", [(xp.reshape(-1), "in"), (b_, "in"), cst, n, (torch.zeros(1, dtype=torch.float64), "inout"),
                                          (None, "out")])
        close(g[0], c[0], 1e-9)
        (og,), (oc,) = run_both("ms_l1_bwd_ab_f32", [(xp.reshape(-1), "in"), (b_, "in"), cst, (torch.tensor([0.3]), "in"), n,
                                                     (torch.zeros(n), "out")])
        assert torch.equal(og, oc)
    # casts
    src = torch.randn(1000, dtype=torch.float64)
    (og,), (oc,) = run_both("ms_cast", [(src, "in"), 1, (torch.zeros(1000), "out"), 0, 1000])
    assert torch.equal(og, oc)


@pytest.mark.gpu
@pytest.mark.parametrize("dt,sdt", [(1, 1), (1, 0), (0, 0)])
def test_clip_adam_mixed(dt, sdt):
    """The vectorised update (optionally fp32 moments beside fp64 parameters) against the host specification and, with
    native moments, against the scalar entry point; a length that leaves a tail after the 4-wide body."""
    torch.manual_seed(5)
    n = 4 * 50_000 + 3
    T = torch.float64 if dt == 1 else torch.float32
    S = torch.float64 if sdt == 1 else torch.float32
    p, g = torch.randn(n, dtype=T), torch.randn(n, dtype=T) * 1e-2
    m, v = (torch.randn(n, dtype=torch.float64) * 1e-3).to(S), (torch.rand(n, dtype=torch.float64) * 1e-5).to(S)
    sq = (g.double() ** 2).sum().reshape(1)
    step = torch.tensor([3], dtype=torch.int64)
    lr = torch.tensor([2e-4], dtype=torch.float64)
    args = [(p, "inout"), (g, "in"), (m, "inout"), (v, "inout"), dt, sdt, n, (sq, "in"), (step, "in"), 1e-4, 0.9, 0.999, 1e-8, 1.0,
            (lr, "in")]
    gpu, cpu = run_both("ms_clip_adam_mixed", args)
    tol = 1e-12 if dt == 1 else 1e-6
    close(gpu[0], cpu[0], tol)
    close(gpu[1], cpu[1], 1e-12 if sdt == 1 else 1e-6)
    close(gpu[2], cpu[2], 1e-12 if sdt == 1 else 1e-6)
    assert float((gpu[0] - p).abs().max()) > 1e-5            # it moved
    if dt == sdt:
        old, _ = run_both("ms_clip_adam", [(p, "inout"), (g, "in"), (m, "inout"), (v, "inout"), dt, n, (sq, "in"), (step, "in"), 1e-4, 0.9,
                                           0.999, 1e-8, 1.0, (lr, "in")])
        for a, b in zip(gpu, old):
            close(a, b, 1e-14 if dt == 1 else 1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [0, 1])
def test_grad_sqnorm(dt):
    """Fixed-order sum of squares (4 elements per thread + tail) and the step counter it advances; two launches agree bit
    for bit (the data-parallel replicas rely on that)."""
    torch.manual_seed(6)
    n = 4 * 300_000 + 1
    g = torch.randn(n, dtype=torch.float64 if dt == 1 else torch.float32)
    args = [(g, "in"), dt, n, (torch.zeros(1, dtype=torch.float64), "out"), (torch.tensor([4], dtype=torch.int64), "inout")]
    gpu, cpu = run_both("ms_grad_sqnorm", args)
    close(gpu[0], cpu[0], 1e-12)
    assert int(gpu[1]) == 5 and int(cpu[1]) == 5
    gpu2, _ = run_both("ms_grad_sqnorm", args)
    assert torch.equal(gpu[0], gpu2[0])


@pytest.mark.gpu
def test_unpack_wgrad_multi_layouts():
    """The conversion launch on a table that mixes every store path: fp64 store-only sinks with an even row (the 16-byte
    path; taps 1, 3, 16, 24), an odd row (scalar path), an accumulating sink, an fp32 sink, and a row too long for shared
    memory."""
    import ctypes
    torch.manual_seed(8)
    #        Cout Cin_g taps kpad pdt acc
    cases = [(256, 256, 3, 256, 1, 0), (64, 64, 16, 64, 1, 0), (256, 256, 24, 256, 1, 0), (96, 256, 1, 256, 1, 0),
             (40, 33, 3, 64, 1, 0), (64, 64, 3, 64, 1, 1), (128, 128, 3, 128, 0, 0), (32, 266, 3, 320, 1, 0), (8, 512, 24, 512, 1, 0)]
    accs = [torch.randn(co * t * kp) for co, ci, t, kp, _, _ in cases]
    sinks = [torch.randn(co * ci * t, dtype=torch.float64 if pdt == 1 else torch.float32) for co, ci, t, _, pdt, _ in cases]
    outs = []
    for dev in ("cuda", "cpu"):
        a_ = [a.to(dev).contiguous() for a in accs]
        s_ = [s.clone().to(dev).contiguous() for s in sinks]
        arr = (_lib.WgradEntry * len(cases))()
        for i, (co, ci, t, kp, pdt, ac) in enumerate(cases):
            arr[i].acc, arr[i].dw = a_[i].data_ptr(), s_[i].data_ptr()
            arr[i].pdt, arr[i].Cout, arr[i].Cin_g, arr[i].taps, arr[i].kpad, arr[i].accumulate = pdt, co, ci, t, kp, ac
        tab = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        if dev == "cuda":
            tab = tab.cuda()
            _lib.call("ms_unpack_wgrad_multi", ptr(tab), len(cases), 0, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
        else:
            cpu_emu.ms_unpack_wgrad_multi(tab.data_ptr(), len(cases), 0, None)
        outs.append([s.cpu() for s in s_])
    for i, (g, c) in enumerate(zip(*outs)):
        assert torch.equal(g, c), cases[i]


@pytest.mark.gpu
def test_loss_combine_and_backward():
    """ops.combine_losses on the device against the host specification: scaled terms, their sum, the seeds of backward,
    with a device-resident lambda that changes between two calls (what a replayed graph sees)."""
    torch.manual_seed(9)
    vals = [torch.randn((), device=DEV, requires_grad=(i != 3)) for i in range(5)]
    lam = torch.tensor([0.25, 3.0], dtype=torch.float64, device=DEV)
    terms = [ops.LossTerm(vals[0]), ops.LossTerm(vals[1], 1.0, 1), ops.LossTerm(vals[2]), ops.LossTerm(vals[3], 0.1),
             ops.LossTerm(vals[4], 0.1, 0)]
    for lam_now in ([0.25, 3.0], [0.5, 7.0]):
        lam.copy_(torch.tensor(lam_now, dtype=torch.float64))
        total, report = ops.combine_losses(terms, lam)
        w = [1.0, lam_now[1], 1.0, 0.1, 0.1 * lam_now[0]]
        want = torch.tensor([wi * float(v) for wi, v in zip(w, vals)], dtype=torch.float64)
        assert report.dtype == torch.float64 and not report.requires_grad
        assert torch.allclose(report.cpu(), want, rtol=1e-12, atol=0)
        assert abs(float(total) - float(want.sum())) <= 1e-6 * max(1.0, float(want.abs().sum()))
        for v in vals:
            v.grad = None
        (2.0 * total).backward()
        for i, v in enumerate(vals):
            if i == 3:
                assert v.grad is None
            else:
                assert abs(float(v.grad) - 2.0 * w[i]) <= 1e-6 * max(1.0, abs(2.0 * w[i])), i
    with pytest.raises(ops.MixStageError):
        ops.combine_losses([torch.zeros((), device=DEV)], lam)
