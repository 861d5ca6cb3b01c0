"""tcgen05 implicit-GEMM kernel (csrc/conv_tc.cu) on the GPU against its CPU specification and
against F.conv2d, over the generator's geometries: forward and input gradient, stride 1/2, 1-D/2-D,
grouped, ragged channel counts, both output dtypes and both epilogues."""
import pytest
import torch
import torch.nn.functional as F

import cpu_emu
from mixstage_b200 import _lib, igemm
from mixstage_b200._lib import ptr
from test_igemm_desc_cpu import GEOMS, _conv_out

pytestmark = pytest.mark.gpu

BIG = [
    (16, 1, 64, 256, 256, 1, 3, 1, 1, 0, 1, 1),      # the most repeated GEMM (M=1024, N=256, K=768)
    (16, 1, 64, 2048, 2048, 1, 3, 1, 1, 0, 1, 8),    # decoder.1-3
    (4, 64, 64, 64, 64, 4, 4, 2, 2, 1, 1, 1),        # audio_encoder.conv.1
    (4, 8, 8, 256, 256, 3, 8, 1, 1, 1, 3, 1),        # audio_encoder.conv.7
    (16, 1, 2, 256, 256, 1, 3, 1, 1, 0, 1, 1),       # unet bottom: 32 rows in a 128-row tile
    (130, 1, 2, 256, 256, 1, 3, 1, 1, 0, 1, 1),      # ragged batch tail
]


def _sync_ok():
    torch.cuda.synchronize()


def _gpu_igemm(plan, a, w_src, mode, Cout, Cin_g, taps_total, groups, bias=None, scale=None, shift=None, out=None):
    st = torch.cuda.current_stream().cuda_stream
    wp = torch.zeros(plan.wp_numel, dtype=torch.bfloat16, device="cuda")
    d = plan.desc
    _lib.call("ms_pack_igemm_weight_bf16", ptr(w_src), _lib.dt_code(w_src.dtype), Cout, Cin_g, taps_total, groups, mode,
              d.num_classes, d.class_n, d.ntaps, plan.kpad, plan.srctap_c, ptr(wp), None, st)
    _lib.call("ms_igemm_bf16", d, ptr(a), ptr(wp), ptr(bias), ptr(scale), ptr(shift), ptr(out), st)
    torch.cuda.synchronize()
    return wp


@pytest.mark.parametrize("g", GEOMS + BIG)
def test_igemm_fwd_dgrad(g):
    torch.manual_seed(0)
    B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups = g
    Ho, Wo = _conv_out(H, kh, sh, ph), _conv_out(W, kw, sw, pw)
    x = torch.randn(B, H, W, Cin).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin // groups, kh, kw, dtype=torch.float64) / (Cin // groups * kh * kw) ** 0.5)
    w = w.to(torch.bfloat16).double()
    bias = torch.randn(Cout)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w, bias.double(), stride=(sh, sw), padding=(ph, pw), groups=groups).permute(0, 2, 3, 1)
    plan = igemm.make_fwd(B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups, Ho, Wo)
    out = torch.full((B, Ho, Wo, Cout), float("nan"), device="cuda")
    wp = _gpu_igemm(plan, x.cuda(), w.cuda(), 0, Cout, Cin // groups, kh * kw, groups, bias=bias.cuda(), out=out)
    # weight re-tiling is bit-exact vs the spec
    wp_c = torch.zeros(plan.wp_numel, dtype=torch.bfloat16)
    cpu_emu.ms_pack_igemm_weight_bf16(ptr(w), 1, Cout, Cin // groups, kh * kw, groups, 0, plan.desc.num_classes,
                                      plan.desc.class_n, plan.desc.ntaps, plan.kpad, plan.srctap, ptr(wp_c), None, None)
    assert torch.equal(wp.cpu(), wp_c)
    err = float((out.cpu().double() - ref).abs().max())
    assert err < 1e-3 * float(ref.abs().max()), (err, float(ref.abs().max()))
    # fused scale/shift + LeakyReLU epilogue with bf16 output
    scale, shift = torch.rand(Cout) + 0.5, torch.randn(Cout) * 0.1
    plan2 = igemm.make_fwd(B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups, Ho, Wo, out_dtype=_lib.MS_BF16, epilogue=1, slope=0.2)
    out2 = torch.zeros((B, Ho, Wo, Cout), dtype=torch.bfloat16, device="cuda")
    _gpu_igemm(plan2, x.cuda(), w.cuda(), 0, Cout, Cin // groups, kh * kw, groups, scale=scale.cuda(), shift=shift.cuda(), out=out2)
    ref2 = F.leaky_relu((ref - bias.double()) * scale.double() + shift.double(), 0.2)
    err2 = float((out2.cpu().double() - ref2).abs().max())
    assert err2 < 1e-2 * float(ref2.abs().max()), (err2,)
    # input gradient
    dz = torch.randn(B, Ho, Wo, Cout).to(torch.bfloat16)
    dref = torch.nn.grad.conv2d_input((B, Cin, H, W), w, dz.double().permute(0, 3, 1, 2), stride=(sh, sw), padding=(ph, pw),
                                      groups=groups).permute(0, 2, 3, 1)
    plan3 = igemm.make_dgrad(B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups, Ho, Wo)
    dx = torch.full((B, H, W, Cin), float("nan"), device="cuda")
    _gpu_igemm(plan3, dz.cuda(), w.cuda(), 1, Cout, Cin // groups, kh * kw, groups, out=dx)
    err3 = float((dx.cpu().double() - dref).abs().max())
    assert err3 < 1e-3 * float(dref.abs().max()), (err3, float(dref.abs().max()))


@pytest.mark.parametrize("g", GEOMS + BIG)
def test_wgrad_tc(g):
    torch.manual_seed(3)
    B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups = g
    Ho, Wo = _conv_out(H, kh, sh, ph), _conv_out(W, kw, sw, pw)
    x = torch.randn(B, H, W, Cin).to(torch.bfloat16)
    dz = torch.randn(B, Ho, Wo, Cout).to(torch.bfloat16)
    wref = torch.nn.grad.conv2d_weight(x.double().permute(0, 3, 1, 2), (Cout, Cin // groups, kh, kw), dz.double().permute(0, 3, 1, 2),
                                       stride=(sh, sw), padding=(ph, pw), groups=groups)
    plan = igemm.make_fwd(B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups, Ho, Wo)
    st = torch.cuda.current_stream().cuda_stream
    dwp = torch.full((plan.wp_numel,), float("nan"), device="cuda")
    xg, dzg = x.cuda(), dz.cuda()          # keep the device copies alive across the launch
    _lib.call("ms_wgrad_bf16", plan.desc, ptr(xg), ptr(dzg), ptr(dwp), st)
    dw = torch.zeros(Cout, Cin // groups, kh, kw, dtype=torch.float64, device="cuda")
    _lib.call("ms_unpack_igemm_wgrad", ptr(dwp), Cout, Cin // groups, kh * kw, plan.desc.ntaps, plan.kpad, ptr(dw), 1, 1, 0, st)
    torch.cuda.synchronize()
    err = float((dw.cpu() - wref).abs().max())
    assert err < 1e-3 * float(wref.abs().max()), (err, float(wref.abs().max()))


SPLIT = [GEOMS[0], GEOMS[1], GEOMS[3], GEOMS[6], GEOMS[9], BIG[1], BIG[3], BIG[4]]


def _planes(t, rs=None):
    """fp32 (…, C) CUDA tensor -> (hi/lo planes tensor, plane stride) via ms_to_planes."""
    C = t.shape[-1]
    rs = rs or C
    rows = t.numel() // C
    ps = (rows * rs + 7) // 8 * 8
    pl = torch.zeros(2 * ps, dtype=torch.bfloat16, device="cuda")
    _lib.call("ms_to_planes", ptr(t), rows, C, rs, ptr(pl), 3, ps, torch.cuda.current_stream().cuda_stream)
    return pl, ps


@pytest.mark.parametrize("g", SPLIT)
def test_split_bf16_fwd_dgrad_wgrad(g):
    """bf16x3: hi/lo operand planes, three MMA passes -> ~fp32 accuracy (error bound 3e-5 of the output scale)."""
    torch.manual_seed(5)
    B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups = g
    Ho, Wo = _conv_out(H, kh, sh, ph), _conv_out(W, kw, sw, pw)
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(B, H, W, Cin)
    w = torch.randn(Cout, Cin // groups, kh, kw, dtype=torch.float64) / (Cin // groups * kh * kw) ** 0.5
    dz = torch.randn(B, Ho, Wo, Cout)
    xd, wd, dzd = x.double().permute(0, 3, 1, 2), w, dz.double().permute(0, 3, 1, 2)
    ref = F.conv2d(xd, wd, None, stride=(sh, sw), padding=(ph, pw), groups=groups).permute(0, 2, 3, 1)
    dref = torch.nn.grad.conv2d_input((B, Cin, H, W), wd, dzd, stride=(sh, sw), padding=(ph, pw), groups=groups).permute(0, 2, 3, 1)
    wref = torch.nn.grad.conv2d_weight(xd, (Cout, Cin // groups, kh, kw), dzd, stride=(sh, sw), padding=(ph, pw), groups=groups)
    xg, wg, dzg = x.cuda(), w.cuda(), dz.cuda()
    xp, xps = _planes(xg)
    dzp, dzps = _planes(dzg)

    def packed(plan, mode):
        ps = (plan.wp_numel + 7) // 8 * 8
        wp = torch.zeros(2 * ps, dtype=torch.bfloat16, device="cuda")
        d = plan.desc
        _lib.call("ms_pack_igemm_weight_bf16", ptr(wg), 1, Cout, Cin // groups, kh * kw, groups, mode, d.num_classes, d.class_n,
                  d.ntaps, plan.kpad, plan.srctap_c, ptr(wp), wp.data_ptr() + 2 * ps, st)
        return wp, ps

    pf = igemm.make_fwd(B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups, Ho, Wo)
    pf.desc.block_n = igemm.pick_block_n(pf.desc, 3)
    wp, wps = packed(pf, 0)
    igemm.set_planes(pf, True, xps, wps, 0)
    out = torch.full((B, Ho, Wo, Cout), float("nan"), device="cuda")
    pf.desc.split_k, pf.desc.out_numel = igemm.igemm_split(pf.desc, 3), out.numel()      # split-K where the grid is small
    _lib.call("ms_igemm_bf16", pf.desc, ptr(xp), ptr(wp), None, None, None, ptr(out), st)
    torch.cuda.synchronize()
    err = float((out.cpu().double() - ref).abs().max())
    assert err < 3e-5 * float(ref.abs().max()), (err, float(ref.abs().max()))

    pd = igemm.make_dgrad(B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups, Ho, Wo)
    pd.desc.block_n = igemm.pick_block_n(pd.desc, 3)
    wt, wtps = packed(pd, 1)
    igemm.set_planes(pd, True, dzps, wtps, 0)
    dx = torch.full((B, H, W, Cin), float("nan"), device="cuda")
    pd.desc.split_k, pd.desc.out_numel = igemm.igemm_split(pd.desc, 3), dx.numel()
    _lib.call("ms_igemm_bf16", pd.desc, ptr(dzp), ptr(wt), None, None, None, ptr(dx), st)
    torch.cuda.synchronize()
    err = float((dx.cpu().double() - dref).abs().max())
    assert err < 3e-5 * float(dref.abs().max()), (err, float(dref.abs().max()))

    igemm.set_planes(pf, True, xps, 0, dzps)
    nsplit, pf.desc.wgrad_c_tile = igemm.wgrad_split(pf.desc, npass=3)
    pf.desc.split_k = nsplit
    dwp = torch.full((nsplit * pf.wp_numel,), float("nan"), device="cuda")
    _lib.call("ms_wgrad_bf16", pf.desc, ptr(xp), ptr(dzp), ptr(dwp), st)
    dw = torch.zeros(Cout, Cin // groups, kh, kw, dtype=torch.float64, device="cuda")
    _lib.call("ms_unpack_igemm_wgrad", ptr(dwp), Cout, Cin // groups, kh * kw, pf.desc.ntaps, pf.kpad, ptr(dw), 1, nsplit, 0, st)
    torch.cuda.synchronize()
    err = float((dw.cpu() - wref).abs().max())
    assert err < 3e-5 * float(wref.abs().max()), (err, float(wref.abs().max()))


# ---- persistent kernel: CTAs that walk several tiles (accumulator double buffering, ring across tile boundaries)
MANY = [
    (640, 1, 64, 256, 512, 1, 3, 1, 1, 0, 1, 1),     # 320 row tiles x 2 column tiles = 640 tiles on 148 CTAs
    (96, 1, 64, 1024, 1024, 1, 3, 1, 1, 0, 1, 4),    # grouped: 48 row tiles x 4 classes
    (40, 32, 32, 64, 64, 4, 4, 2, 2, 1, 1, 1),       # 2-D stride 2, N = 64
]


@pytest.mark.parametrize("g", MANY)
@pytest.mark.parametrize("planes", [1, 2])
def test_persistent_many_tiles(g, planes):
    torch.manual_seed(11)
    B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups = g
    Ho, Wo = _conv_out(H, kh, sh, ph), _conv_out(W, kw, sw, pw)
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(B, H, W, Cin, device="cuda")
    w = torch.randn(Cout, Cin // groups, kh, kw, dtype=torch.float64, device="cuda") / (Cin // groups * kh * kw) ** 0.5
    if planes == 1:
        x = x.to(torch.bfloat16).float()
        w = w.to(torch.bfloat16).double()
    scale, shift = torch.rand(Cout, device="cuda") + 0.5, torch.randn(Cout, device="cuda") * 0.1
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w, None, stride=(sh, sw), padding=(ph, pw), groups=groups).permute(0, 2, 3, 1)
    ref = F.leaky_relu(ref * scale.double() + shift.double(), 0.2)
    plan = igemm.make_fwd(B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups, Ho, Wo, out_dtype=_lib.MS_F32, epilogue=1, slope=0.2)
    d = plan.desc
    ps = (plan.wp_numel + 7) // 8 * 8
    wp = torch.zeros(2 * ps, dtype=torch.bfloat16, device="cuda")
    _lib.call("ms_pack_igemm_weight_bf16", ptr(w), 1, Cout, Cin // groups, kh * kw, groups, 0, d.num_classes, d.class_n,
              d.ntaps, plan.kpad, plan.srctap_c, ptr(wp), wp.data_ptr() + 2 * ps, st)
    xp, xps = _planes(x)
    igemm.set_planes(plan, planes == 2, xps, ps, 0)
    out = torch.full((B, Ho, Wo, Cout), float("nan"), device="cuda")
    _lib.call("ms_igemm_bf16", d, ptr(xp), ptr(wp), None, ptr(scale), ptr(shift), ptr(out), st)
    torch.cuda.synchronize()
    err = float((out.double() - ref).abs().max())
    assert err < (1e-3 if planes == 1 else 3e-5) * float(ref.abs().max()), (err, float(ref.abs().max()))
    # planes + fp32 outputs of the fused form agree with each other and with the plain launch
    d.out_dtype = _lib.MS_BF16X2 if planes == 2 else _lib.MS_BF16
    ops_ = (out.numel() + 7) // 8 * 8
    d.out_plane_stride = ops_
    yp = torch.zeros(2 * ops_, dtype=torch.bfloat16, device="cuda")
    y32 = torch.full_like(out, float("nan"))
    _lib.call("ms_igemm_bf16_fused", d, ptr(xp), ptr(wp), None, ptr(scale), ptr(shift), ptr(yp), ptr(y32), None, 0, 0, 0, st)
    torch.cuda.synchronize()
    assert torch.equal(y32, out)
    got = yp[:out.numel()].float() + (yp[ops_:ops_ + out.numel()].float() if planes == 2 else 0)
    tol = 2 ** -8 if planes == 1 else 2 ** -15
    assert float((got - out.reshape(-1)).abs().max()) <= tol * float(out.abs().max())


@pytest.mark.parametrize("planes", [1, 2])
def test_mixture_inside_the_gemms(planes):
    """row_w_mode 1 (cluster weight per row and class in the grouped block's epilogue) and row_w_mode 2 (dense logits GEMM over
    the K weighted channel groups + mixed bias) against index_select_outputs' definition (jlcss.py:106-115)."""
    torch.manual_seed(13)
    K, Cg, P, B, T = 8, 256, 96, 40, 64
    st = torch.cuda.current_stream().cuda_stream
    rows = B * T
    x = torch.randn(B, 1, T, K * Cg, device="cuda")
    w1 = torch.randn(K * Cg, Cg, 1, 3, dtype=torch.float64, device="cuda") / (Cg * 3) ** 0.5
    wl = torch.randn(K * P, Cg, 1, 1, dtype=torch.float64, device="cuda") / Cg ** 0.5
    bl = torch.randn(K * P, device="cuda")
    soft = torch.softmax(torch.randn(rows, K, device="cuda"), -1).contiguous()
    if planes == 1:
        x, w1, wl = x.to(torch.bfloat16).float(), w1.to(torch.bfloat16).double(), wl.to(torch.bfloat16).double()
    scale, shift = torch.rand(K * Cg, device="cuda") + 0.5, torch.randn(K * Cg, device="cuda") * 0.1
    h = F.conv2d(x.double().permute(0, 3, 1, 2), w1, None, padding=(0, 1), groups=K).permute(0, 2, 3, 1)
    h = F.leaky_relu(h * scale.double() + shift.double(), 0.2)                                   # (B,1,T,K*Cg)
    hw = (h.reshape(rows, K, Cg) * soft.double().unsqueeze(-1)).reshape(B, 1, T, K * Cg)
    # ---- mode 1
    plan = igemm.make_fwd(B, 1, T, K * Cg, K * Cg, 1, 3, 1, 1, 0, 1, K, 1, T, epilogue=1, slope=0.2)
    d = plan.desc
    fmt = _lib.MS_BF16X2 if planes == 2 else _lib.MS_BF16
    ps = (plan.wp_numel + 7) // 8 * 8
    wp = torch.zeros(2 * ps, dtype=torch.bfloat16, device="cuda")
    _lib.call("ms_pack_igemm_weight_bf16", ptr(w1), 1, K * Cg, Cg, 3, K, 0, d.num_classes, d.class_n, d.ntaps, plan.kpad,
              plan.srctap_c, ptr(wp), wp.data_ptr() + 2 * ps, st)
    xp, xps = _planes(x)
    n = rows * K * Cg
    ops_ = (n + 7) // 8 * 8
    igemm.set_planes(plan, planes == 2, xps, ps, ops_)
    d.out_dtype = fmt
    yp = torch.zeros(2 * ops_, dtype=torch.bfloat16, device="cuda")
    y32 = torch.full((B, 1, T, K * Cg), float("nan"), device="cuda")
    _lib.call("ms_igemm_bf16_mix", d, ptr(xp), ptr(wp), None, ptr(scale), ptr(shift), ptr(yp), ptr(y32), ptr(soft), K, 1, 0, st)
    torch.cuda.synchronize()
    tol = 1e-3 if planes == 1 else 3e-5
    assert float((y32.double() - hw).abs().max()) < tol * float(hw.abs().max())
    # ---- mode 2 on the kernel's own planes
    A = yp[:n].double() + (yp[ops_:ops_ + n].double() if planes == 2 else 0)                     # the weighted planes, exactly
    ref = torch.einsum("rkc,kpc->rp", A.view(rows, K, Cg), wl.view(K, P, Cg)) + soft.double() @ bl.double().view(K, P)
    dense = wl.view(K, P, Cg).permute(1, 0, 2).reshape(P, K * Cg, 1, 1).contiguous()
    p2 = igemm.make_fwd(B, 1, T, K * Cg, P, 1, 1, 1, 1, 0, 0, 1, 1, T)
    d2 = p2.desc
    d2.block_n = P
    ps2 = (p2.wp_numel + 7) // 8 * 8
    wp2 = torch.zeros(2 * ps2, dtype=torch.bfloat16, device="cuda")
    _lib.call("ms_pack_igemm_weight_bf16", ptr(dense), 1, P, K * Cg, 1, 1, 0, d2.num_classes, d2.class_n, d2.ntaps, p2.kpad,
              p2.srctap_c, ptr(wp2), wp2.data_ptr() + 2 * ps2, st)
    igemm.set_planes(p2, planes == 2, ops_, ps2, 0)
    pose = torch.full((B, 1, T, P), float("nan"), device="cuda")
    _lib.call("ms_igemm_bf16_mix", d2, ptr(yp), ptr(wp2), ptr(bl), None, None, ptr(pose), None, ptr(soft), K, 2, K, st)
    torch.cuda.synchronize()
    err = float((pose.double().view(rows, P) - ref).abs().max())
    assert err < (1e-3 if planes == 1 else 3e-5) * float(ref.abs().max()), (err, float(ref.abs().max()))
