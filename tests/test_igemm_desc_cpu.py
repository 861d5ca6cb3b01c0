"""Host logic of the tcgen05 implicit-GEMM descriptors (mixstage_b200/igemm.py): tap tables,
space-to-depth views, parity classes and weight re-tiling, checked against F.conv2d and its
gradients through the CPU specification of the kernel (tests/cpu_emu.py).  No GPU."""
import pytest
import torch
import torch.nn.functional as F

import cpu_emu
from mixstage_b200 import igemm
from mixstage_b200._lib import ptr

# B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups
GEOMS = [
    (2, 1, 64, 256, 256, 1, 3, 1, 1, 0, 1, 1),       # k3 s1
    (2, 1, 64, 256, 256, 1, 4, 1, 2, 0, 1, 1),       # k4 s2 (1-D)
    (3, 1, 4, 256, 256, 1, 4, 1, 2, 0, 1, 1),        # unet bottleneck 4 -> 2
    (2, 16, 16, 64, 64, 4, 4, 2, 2, 1, 1, 1),        # 2-D 4x4 s2
    (2, 8, 16, 64, 128, 3, 3, 1, 1, 1, 1, 1),        # 2-D 3x3
    (2, 8, 8, 128, 64, 3, 8, 1, 1, 1, 3, 1),         # 3x8 pad (1,3): W 8 -> 7
    (2, 1, 32, 512, 512, 1, 3, 1, 1, 0, 1, 4),       # grouped
    (2, 1, 32, 512, 384, 1, 1, 1, 1, 0, 0, 4),       # grouped 1x1, N=96 per group
    (2, 1, 16, 128, 256, 1, 4, 1, 1, 0, 1, 1),       # D.conv3 (16 -> 15)
    (2, 1, 64, 96, 64, 1, 4, 1, 2, 0, 1, 1),         # C_in = 96 (not a multiple of 64)
]


def _conv_out(n, k, s, p):
    return (n + 2 * p - k) // s + 1


@pytest.mark.parametrize("g", GEOMS)
def test_fwd_and_dgrad_descriptors_match_conv2d(g):
    torch.manual_seed(0)
    B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups = g
    Ho, Wo = _conv_out(H, kh, sh, ph), _conv_out(W, kw, sw, pw)
    x = torch.randn(B, H, W, Cin).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin // groups, kh, kw) / (Cin // groups * kh * kw) ** 0.5).to(torch.bfloat16).float()
    bias = torch.randn(Cout)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w, bias, stride=(sh, sw), padding=(ph, pw), groups=groups).permute(0, 2, 3, 1)
    plan = igemm.make_fwd(B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups, Ho, Wo)
    wp = torch.zeros(plan.wp_numel, dtype=torch.bfloat16)
    cpu_emu.ms_pack_igemm_weight_bf16(ptr(w), 0, Cout, Cin // groups, kh * kw, groups, 0, plan.desc.num_classes,
                                      plan.desc.class_n, plan.desc.ntaps, plan.kpad, plan.srctap, ptr(wp), None, None)
    out = torch.full((B, Ho, Wo, Cout), float("nan"))
    cpu_emu.ms_igemm_bf16(plan.desc, ptr(x), ptr(wp), ptr(bias), None, None, ptr(out), None)
    assert float((out - ref).abs().max()) < 2e-4 * float(ref.abs().max())
    # input gradient
    dz = torch.randn(B, Ho, Wo, Cout).to(torch.bfloat16)
    dref = torch.nn.grad.conv2d_input((B, Cin, H, W), w, dz.float().permute(0, 3, 1, 2), stride=(sh, sw), padding=(ph, pw),
                                      groups=groups).permute(0, 2, 3, 1)
    plan = igemm.make_dgrad(B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups, Ho, Wo)
    wp = torch.zeros(plan.wp_numel, dtype=torch.bfloat16)
    cpu_emu.ms_pack_igemm_weight_bf16(ptr(w), 0, Cout, Cin // groups, kh * kw, groups, 1, plan.desc.num_classes,
                                      plan.desc.class_n, plan.desc.ntaps, plan.kpad, plan.srctap, ptr(wp), None, None)
    dx = torch.full((B, H, W, Cin), float("nan"))
    cpu_emu.ms_igemm_bf16(plan.desc, ptr(dz), ptr(wp), None, None, None, ptr(dx), None)
    assert float((dx - dref).abs().max()) < 2e-4 * float(dref.abs().max())
    # weight gradient (uses the forward descriptor)
    wref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin // groups, kh, kw), dz.float().permute(0, 3, 1, 2),
                                       stride=(sh, sw), padding=(ph, pw), groups=groups)
    plan = igemm.make_fwd(B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups, Ho, Wo)
    dwp = torch.full((plan.wp_numel,), float("nan"))
    cpu_emu.ms_wgrad_bf16(plan.desc, ptr(x), ptr(dz), ptr(dwp), None)
    dw = torch.zeros(Cout, Cin // groups, kh, kw)
    cpu_emu.ms_unpack_igemm_wgrad(ptr(dwp), Cout, Cin // groups, kh * kw, plan.desc.ntaps, plan.kpad, ptr(dw), 0, 1, 0, None)
    assert float((dw - wref).abs().max()) < 2e-4 * float(wref.abs().max())


def test_padded_channel_rows_266():
    """C_in = 266 stored with a 272-element row stride (style concat buffer)."""
    torch.manual_seed(1)
    B, L, Cin, Cp, Cout = 2, 32, 266, 272, 256
    x = torch.zeros(B, 1, L, Cp, dtype=torch.bfloat16)
    x[..., :Cin] = torch.randn(B, 1, L, Cin).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 1, 3) / 28.0).to(torch.bfloat16).float()
    ref = F.conv2d(x[..., :Cin].float().permute(0, 3, 1, 2), w, None, padding=(0, 1)).permute(0, 2, 3, 1)
    plan = igemm.make_fwd(B, 1, L, Cin, Cout, 1, 3, 1, 1, 0, 1, 1, 1, L, a_row_stride=Cp)
    wp = torch.zeros(plan.wp_numel, dtype=torch.bfloat16)
    cpu_emu.ms_pack_igemm_weight_bf16(ptr(w), 0, Cout, Cin, 3, 1, 0, 1, Cout, 3, plan.kpad, plan.srctap, ptr(wp), None, None)
    out = torch.zeros(B, 1, L, Cout)
    cpu_emu.ms_igemm_bf16(plan.desc, ptr(x), ptr(wp), None, None, None, ptr(out), None)
    assert float((out - ref).abs().max()) < 2e-4 * float(ref.abs().max())
    dz = torch.randn(B, 1, L, Cout).to(torch.bfloat16)
    dref = torch.nn.grad.conv2d_input((B, Cin, 1, L), w, dz.float().permute(0, 3, 1, 2), padding=(0, 1)).permute(0, 2, 3, 1)
    plan = igemm.make_dgrad(B, 1, L, Cin, Cout, 1, 3, 1, 1, 0, 1, 1, 1, L, out_row_stride=Cp)
    wp = torch.zeros(plan.wp_numel, dtype=torch.bfloat16)
    cpu_emu.ms_pack_igemm_weight_bf16(ptr(w), 0, Cout, Cin, 3, 1, 1, plan.desc.num_classes, plan.desc.class_n,
                                      plan.desc.ntaps, plan.kpad, plan.srctap, ptr(wp), None, None)
    dx = torch.full((B, 1, L, Cp), float("nan"))
    cpu_emu.ms_igemm_bf16(plan.desc, ptr(dz), ptr(wp), None, None, None, ptr(dx), None)
    assert float((dx[..., :Cin] - dref).abs().max()) < 2e-4 * float(dref.abs().max())
    assert float(dx[..., Cin:].abs().max()) == 0.0
