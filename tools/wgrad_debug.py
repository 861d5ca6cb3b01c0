"""GPU debug helper: error structure of ms_wgrad_bf16 per tap / n-chunk / c-chunk."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
from mixstage_b200 import _lib, igemm
from mixstage_b200._lib import ptr

GEOMS = {
    "g3": (2, 16, 16, 64, 64, 4, 4, 2, 2, 1, 1, 1),
    "g12": (4, 64, 64, 64, 64, 4, 4, 2, 2, 1, 1, 1),
    "g0": (2, 1, 64, 256, 256, 1, 3, 1, 1, 0, 1, 1),
    "s1": (1, 1, 64, 64, 64, 1, 1, 1, 1, 0, 0, 1),
    "s2": (1, 1, 64, 128, 64, 1, 1, 1, 1, 0, 0, 1),
    "s3": (1, 1, 64, 64, 128, 1, 1, 1, 1, 0, 0, 1),
    "s4": (2, 1, 64, 64, 64, 1, 1, 1, 1, 0, 0, 1),
    "s5": (1, 1, 128, 64, 64, 1, 1, 1, 1, 0, 0, 1),
}


def co(n, k, s, p):
    return (n + 2 * p - k) // s + 1


def run(name, mode):
    B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups = GEOMS[name]
    Ho, Wo = co(H, kh, sh, ph), co(W, kw, sw, pw)
    torch.manual_seed(3)
    if mode == "rand":
        x = torch.randn(B, H, W, Cin)
        dz = torch.randn(B, Ho, Wo, Cout)
    elif mode == "ones":
        x = torch.ones(B, H, W, Cin)
        dz = torch.ones(B, Ho, Wo, Cout)
    elif mode == "xchan":      # x = channel index / 64, z = 1 -> dw[n][c] = rows * c/64
        x = (torch.arange(Cin).float() / 64).expand(B, H, W, Cin).clone()
        dz = torch.ones(B, Ho, Wo, Cout)
    elif mode == "zchan":
        x = torch.ones(B, H, W, Cin)
        dz = (torch.arange(Cout).float() / 64).expand(B, Ho, Wo, Cout).clone()
    x, dz = x.to(torch.bfloat16), dz.to(torch.bfloat16)
    wref = torch.nn.grad.conv2d_weight(x.double().permute(0, 3, 1, 2), (Cout, Cin // groups, kh, kw), dz.double().permute(0, 3, 1, 2),
                                       stride=(sh, sw), padding=(ph, pw), groups=groups)
    plan = igemm.make_fwd(B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups, Ho, Wo)
    st = torch.cuda.current_stream().cuda_stream
    dwp = torch.full((plan.wp_numel,), float("nan"), device="cuda")
    xg, dzg = x.cuda(), dz.cuda()
    _lib.call("ms_wgrad_bf16", plan.desc, ptr(xg), ptr(dzg), ptr(dwp), st)
    dw = torch.zeros(Cout, Cin // groups, kh, kw, dtype=torch.float64, device="cuda")
    _lib.call("ms_unpack_igemm_wgrad", ptr(dwp), Cout, Cin // groups, kh * kw, plan.desc.ntaps, plan.kpad, ptr(dw), 1, 1, 0, st)
    torch.cuda.synchronize()
    dw = dw.cpu()
    err = (dw - wref).abs()
    print("== %s %s box=%s err=%.4g max=%.4g nan=%d" % (name, mode, list(plan.desc.box), float(err.nan_to_num(1e9).max()),
                                                 float(wref.abs().max()), int(torch.isnan(dw).sum())))
    for t in range(kh * kw):
        e = err[:, :, t // kw, t % kw]
        parts = []
        for n0 in range(0, Cout, 64):
            for c0 in range(0, Cin // groups, 64):
                parts.append("n%d c%d: %.3g" % (n0, c0, float(e[n0:n0 + 64, c0:c0 + 64].nan_to_num(1e9).max())))
        print("  tap %d: %s" % (t, "; ".join(parts[:8])))
    if mode != "rand":
        print("  got[0:3,0:4,0,0]", dw[0:3, 0:4, 0, 0].tolist(), "ref", wref[0:3, 0:4, 0, 0].tolist())
        print("  got[64:66,64:68,0,0]", dw[64:66, 64:68, 0, 0].tolist() if Cout > 64 and Cin > 64 else None)


if __name__ == "__main__":
    for n in sys.argv[1].split(","):
        for m in sys.argv[2].split(","):
            run(n, m)
