"""Parity of the CUDA path (through the reference-facing classes and the C-ABI) against the
oracle run live on the CPU and against the golden vectors produced by the reference."""
import numpy as np
import pytest
import torch

from model_cases import run_case
from oracle_cases import CASES, load_golden, run_oracle

pytestmark = pytest.mark.gpu

OUT_TOL = 1e-3          # fp32 mode: relative Frobenius error of pose / labels_cap_soft (north_star: 1e-3 fp32)
LOSS_TOL = 1e-3
# Gradients: relative Frobenius error per parameter tensor.  The reference's own fp32 run differs from
# its fp64 run by 2.5e-3 on cfg2_gstep (a few LeakyReLU masks flip where |z|~1e-7, each worth
# ~1/sqrt(rows) of a tensor).  Worse, the GAN term is a piecewise-linear net under an L1 loss: in the
# fp64 oracle itself, perturbing the generator output of cfg2_gstep by 1e-5 (relative) moves
# d(G_gan)/d(pose) by 2.8e-2 because one discriminator pre-activation sits within 1e-6 of the kink
# (tools/grad_debug.py, tools/d_debug.py; the kernels themselves agree with their CPU spec to 1e-6 in
# tests/test_kernels_gpu.py).  Bound: 5e-2 (1e-1 for the 2-sequence stress case).
GRAD_TOL = {"cfg5_stress_small": 1e-1}


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("name", list(CASES))
def test_cuda_path_matches_oracle_and_golden(golden_dir, name):
    got = run_case(name, "cuda", torch.float64)
    ref = run_oracle(name)
    gold = load_golden(golden_dir, name)
    pose = got["pose"].cpu()
    assert pose.dtype == torch.float64 and tuple(pose.shape) == tuple(gold["pose"].shape)
    assert _rel(pose, ref["pose"]) < OUT_TOL
    assert _rel(pose, torch.from_numpy(gold["pose"])) < OUT_TOL
    np.testing.assert_allclose(got["losses"], gold["losses"], rtol=LOSS_TOL, atol=1e-5)
    soft = got["labels_cap_soft"].cpu()
    gsoft = torch.from_numpy(gold["labels_cap_soft"]).reshape(soft.shape)
    assert _rel(soft, gsoft) < OUT_TOL
    # cluster assignment: bit-exact wherever the reference's top-2 margin exceeds the fp32 noise floor
    if gsoft.shape[-1] > 1:
        top2 = torch.topk(gsoft.double(), 2, dim=-1).values
        safe = (top2[..., 0] - top2[..., 1]) > 1e-4
    else:                                   # StAGE variant: a single cluster, weight 1
        safe = torch.ones(gsoft.shape[:-1], dtype=torch.bool)
    am = soft.argmax(-1)
    gam = torch.from_numpy(gold["cluster_argmax"].astype(np.int64)).reshape(am.shape)
    assert bool((am[safe] == gam[safe]).all())
    assert float((am == gam).double().mean()) >= 0.995
    kind, kw = CASES[name][3], CASES[name][4]
    if kind == "gan" and kw["step"] != "eval":
        tol = GRAD_TOL.get(name, 5e-2)
        sd, sdd = ref["sd"], ref["sdd"]
        gscale = max([float(v.grad.norm()) for v in sd.values() if v.requires_grad and v.grad is not None] + [0.0])
        for n, p in got["G"].named_parameters():
            r = sd[n].grad
            if r is None or float(r.abs().max()) == 0.0:
                assert p.grad is None or float(p.grad.abs().max()) <= 1e-6 * gscale, n
                continue
            assert p.grad is not None and p.grad.dtype == p.dtype, n
            err = float((p.grad.cpu().double() - r).norm())
            assert err <= tol * float(r.norm()) + 1e-6 * gscale, (n, err, float(r.norm()))
        for n, p in got["D"].named_parameters():
            r = sdd[n].grad
            err = float((p.grad.cpu().double() - r).norm())
            assert err <= tol * float(r.norm()) + 1e-7, (n, err, float(r.norm()))
        # golden gradient norms / selected tensors straight from the reference
        gp = dict(got["G"].named_parameters())
        for n, v in zip(gold["g_grad_names"], gold["g_grad_norms"]):
            if v > 1e-8:
                assert abs(float(gp[str(n)].grad.norm()) - v) <= tol * v + 1e-6 * gscale, n
        for k in gold:
            if k.startswith("ggrad/"):
                g = torch.from_numpy(gold[k]).double()
                assert float((gp[k[6:]].grad.cpu().double() - g).norm()) <= tol * float(g.norm()) + 1e-6 * gscale, k
        gsd, dsd = got["G"].state_dict(), got["D"].state_dict()
        for k, v in ref["log_g"].updates.items():
            assert float((gsd[k].cpu().double() - v).abs().max()) < 1e-4, k
        for blk, cnt in ref["log_g"].counts.items():
            assert int(gsd[blk + ".norm.num_batches_tracked"]) == cnt, blk
        for k, v in ref["log_d"].updates.items():
            assert float((dsd[k].cpu().double() - v).abs().max()) < 1e-4, k
        for k in gold:
            if k.startswith("gstat/"):
                assert float((gsd[k[6:]].cpu().double() - torch.from_numpy(gold[k]).double()).abs().max()) < 1e-4, k


def test_style_index_bit_exact():
    """argmax=1: the integer style index fed to the embedding must equal the oracle's exactly."""
    got = run_case("cfg1_train_fwd", "cuda", torch.float64)
    ref = run_oracle("cfg1_train_fwd")
    idx = got["G"].style_index.cpu()
    want = ref["aux"]["style_index"][:, 0]
    assert torch.equal(idx, want)


def test_fp32_parameters_and_inputs():
    """The module also runs with fp32 master parameters / fp32 inputs (no .double())."""
    got = run_case("cfg2_gstep", "cuda", torch.float32)
    ref = run_oracle("cfg2_gstep")
    assert got["pose"].dtype == torch.float32
    assert _rel(got["pose"].cpu(), ref["pose"]) < OUT_TOL
    for n, p in got["G"].named_parameters():
        if p.grad is not None:
            assert p.grad.dtype == torch.float32


def test_cpu_tensors_are_rejected():
    import mixstage_b200 as M
    D = M.Speech2Gesture_D(in_channels=96)
    with pytest.raises(M.MixStageError):
        D(torch.randn(2, 64, 96))


# ---- tensor-core modes (tcgen05 implicit GEMM).  bf16x3 = split-bf16 operands, must meet the fp32 bar (1e-3);
# bf16 = plain bf16 operands: 2e-2 in eval mode, 6e-2 in train mode (batch-stat BatchNorm with random-init weights
# amplifies operand rounding: the reference itself shows 3.5e-2 under bf16 operand emulation).
TC_CASES = [(n, "bf16x3", 1e-3) for n in CASES] + [("cfg1_eval_sample", "bf16", 2e-2), ("cfg2_eval", "bf16", 2e-2),
                                                    ("sample_long", "bf16", 2e-2), ("cfg2_gstep", "bf16", 6e-2),
                                                    ("cfg2_dstep", "bf16", 2e-2)]


@pytest.mark.parametrize("name,precision,tol", TC_CASES)
def test_tensor_core_path_matches_oracle_and_golden(golden_dir, name, precision, tol):
    got = run_case(name, "cuda", torch.float64, precision=precision)
    ref = run_oracle(name)
    gold = load_golden(golden_dir, name)
    pose = got["pose"].cpu()
    assert _rel(pose, ref["pose"]) < tol
    assert _rel(pose, torch.from_numpy(gold["pose"])) < tol
    np.testing.assert_allclose(got["losses"], gold["losses"], rtol=tol, atol=max(1e-5, tol * 1e-2))
    soft = got["labels_cap_soft"].cpu()
    gsoft = torch.from_numpy(gold["labels_cap_soft"]).reshape(soft.shape)
    assert _rel(soft, gsoft) < tol
    if gsoft.shape[-1] > 1:
        top2 = torch.topk(gsoft.double(), 2, dim=-1).values
        safe = (top2[..., 0] - top2[..., 1]) > (1e-3 if precision == "bf16x3" else 5e-2)
    else:                                   # StAGE variant: a single cluster
        safe = torch.ones(gsoft.shape[:-1], dtype=torch.bool)
    am = soft.argmax(-1)
    gam = torch.from_numpy(gold["cluster_argmax"].astype(np.int64)).reshape(am.shape)
    assert bool((am[safe] == gam[safe]).all())
    kind, kw = CASES[name][3], CASES[name][4]
    if kind == "gan" and kw["step"] != "eval" and precision == "bf16x3":
        gtol = GRAD_TOL.get(name, 5e-2)
        sd, sdd = ref["sd"], ref["sdd"]
        gscale = max([float(v.grad.norm()) for v in sd.values() if v.requires_grad and v.grad is not None] + [0.0])
        for n, p in got["G"].named_parameters():
            r = sd[n].grad
            if r is None or float(r.abs().max()) == 0.0:
                assert p.grad is None or float(p.grad.abs().max()) <= 1e-5 * gscale, n
                continue
            assert p.grad is not None and p.grad.dtype == p.dtype, n
            err = float((p.grad.cpu().double() - r).norm())
            assert err <= gtol * float(r.norm()) + 1e-5 * gscale, (n, err, float(r.norm()))
        for n, p in got["D"].named_parameters():
            r = sdd[n].grad
            err = float((p.grad.cpu().double() - r).norm())
            assert err <= gtol * float(r.norm()) + 1e-6, (n, err, float(r.norm()))
        gsd = got["G"].state_dict()
        for k, v in ref["log_g"].updates.items():
            assert float((gsd[k].cpu().double() - v).abs().max()) < 1e-3, k
        for blk, cnt in ref["log_g"].counts.items():
            assert int(gsd[blk + ".norm.num_batches_tracked"]) == cnt, blk


def test_tensor_core_kernels_are_launched(monkeypatch):
    """The bf16x3 run must really go through the tcgen05 entry points (no silent fp32 route)."""
    from mixstage_b200 import _lib, ops
    seen = set()
    orig = ops.call

    def spy(name, *a):
        seen.add(name)
        return orig(name, *a)

    monkeypatch.setattr(ops, "call", spy)
    run_case("cfg2_gstep", "cuda", torch.float64, precision="bf16x3")
    assert {"ms_igemm_bf16", "ms_wgrad_bf16", "ms_pack_igemm_weight_bf16"} <= seen

