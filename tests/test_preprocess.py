"""Pose preprocessing in front of the hot path: the oracle's restatement of KMeans.predict /
ZNorm / RemoveJoints against golden vectors made by the reference's own function bodies (oracle/make_golden.py run_prep),
and the fused CUDA kernel (ms_pose_prepare through mixstage_b200.PosePreprocessor) against both."""
import numpy as np
import pytest
import torch

import cpu_emu
import mixstage_oracle as O
from oracle_cases import load_golden

MASK = [0, 7, 8, 9]
CASES = {"prep_pvs_k8": (["pose", "velocity", "speed"], 8), "prep_pva_k16": (["pose", "velocity", "acceleration"], 16)}


def _oracle(name):
    feats, K = CASES[name]
    x, mean, var, centers = O.synth_prep(4, 64, K, feats)
    xr = O.remove_joints(x, MASK)
    labels = O.kmeans_predict(xr, centers, feats)
    soft = O.kmeans_predict(xr, centers, feats, soft_labels=True)
    zn = O.znorm(x, mean.view(1, 1, -1), var.view(1, 1, -1))
    y = O.remove_joints(zn, MASK)
    inv = O.inv_znorm(zn[..., 4:5].expand(-1, -1, x.shape[-1]).contiguous(), mean.view(1, 1, -1), var.abs().view(1, 1, -1))
    return x, mean, var, centers, labels, soft, y, zn, inv


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(golden_dir, name):
    gold = load_golden(golden_dir, name)
    _, _, _, _, labels, soft, y, _, inv = _oracle(name)
    assert (labels.numpy() == gold["labels"]).all()                   # bit-exact cluster assignment
    np.testing.assert_array_equal(y.numpy(), gold["y"])               # same fp64 operations in the same order
    np.testing.assert_allclose(soft.numpy(), gold["soft"], rtol=1e-14, atol=0)
    np.testing.assert_array_equal(inv[:1].numpy(), gold["inv"])
    assert len(np.unique(gold["labels"])) >= 3                         # the case is not degenerate


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_pose_prepare_matches_oracle_and_golden(golden_dir, name):
    import mixstage_b200 as M
    feats, K = CASES[name]
    gold = load_golden(golden_dir, name)
    x, mean, var, centers, labels, soft, y, zn, inv = _oracle(name)
    pp = M.PosePreprocessor(num_joints=52, mask=MASK, muvar=(mean, var), centers=centers, feats=feats)
    xg = x.cuda()
    yg, lg = pp(xg)
    assert lg.dtype == torch.int64 and tuple(lg.shape) == (4, 64) and yg.dtype == torch.float64
    # labels: exact wherever the best two distances differ by more than fp64 summation-order noise
    f = O.kmeans_feats(O.remove_joints(x, MASK), feats).view(-1, 1, centers.shape[1])
    mse = ((centers.view(1, *centers.shape) - f) ** 2).sum(-1)
    top2 = torch.topk(mse, 2, dim=-1, largest=False).values
    safe = ((top2[:, 1] - top2[:, 0]) > 1e-9 * top2[:, 1]).view(4, 64)
    assert float(safe.double().mean()) > 0.99
    assert bool((lg.cpu()[safe] == labels[safe]).all())
    assert bool((lg.cpu()[safe].numpy() == gold["labels"][safe.numpy()]).all())
    # ZNorm + RemoveJoints: identical operations per element (subtract, divide)
    assert float(((yg.cpu() - y).abs() / (y.abs() + 1e-300)).max()) < 1e-14
    np.testing.assert_allclose(yg.cpu().numpy(), gold["y"], rtol=1e-14, atol=0)
    # soft labels
    _, sg = pp(xg, soft_labels=True)
    np.testing.assert_allclose(sg.cpu().numpy(), gold["soft"], rtol=1e-9, atol=1e-12)
    # the separate entry points agree with the fused call
    assert torch.equal(pp.predict(xg), lg)
    assert torch.equal(pp.normalize(xg), yg)
    pp2 = M.PosePreprocessor(num_joints=52, mask=MASK, muvar=(mean, var.abs()), centers=centers, feats=feats)
    src = zn[..., 4:5].expand(-1, -1, x.shape[-1]).contiguous()
    ig = pp2.inv_znorm(src.cuda())
    # x * sqrt(var) + mean cancels: a last-bit difference of the product (torch evaluates var ** 0.5 through pow) is amplified
    # by |x * sd| / |result|; bound the error against the magnitude of the terms instead of the result
    mag = (src[:1].abs() * var.abs().sqrt() + mean.abs()).numpy()
    assert float((np.abs(ig.cpu()[:1].numpy() - gold["inv"]) / mag).max()) < 1e-15


@pytest.mark.gpu
def test_ragged_and_edge_cases():
    """T = 1 (velocity and acceleration are all zero), a single frame batch, and an empty joint mask."""
    import mixstage_b200 as M
    feats = ["pose", "velocity", "acceleration", "speed"]
    g = torch.Generator().manual_seed(5)
    for (B, T, mask) in [(1, 1, MASK), (3, 2, []), (2, 5, [51])]:
        x = torch.randn(B, T, 104, generator=g, dtype=torch.float64) * 30
        J = 52 - len(mask)
        D = 2 * J * 3 + J
        centers = torch.randn(5, D, generator=g, dtype=torch.float64) * 30
        mean, var = torch.randn(104, generator=g, dtype=torch.float64), torch.rand(104, generator=g, dtype=torch.float64) + 0.1
        pp = M.PosePreprocessor(num_joints=52, mask=mask, muvar=(mean, var), centers=centers, feats=feats)
        y, lab = pp(x.cuda())
        xr = O.remove_joints(x, mask)
        assert torch.equal(lab.cpu(), O.kmeans_predict(xr, centers, feats))
        ref = O.remove_joints(O.znorm(x, mean.view(1, 1, -1), var.view(1, 1, -1)), mask)
        assert float((y.cpu() - ref).abs().max()) < 1e-12


def test_cpu_tensor_is_rejected():
    import mixstage_b200 as M
    if not torch.cuda.is_available():
        pytest.skip("constructor places its tables on the device")
    pp = M.PosePreprocessor(muvar=(torch.zeros(104), torch.ones(104)), centers=torch.zeros(2, 240))
    with pytest.raises(M.MixStageError):
        pp(torch.zeros(1, 4, 104, dtype=torch.float64))


# ---------------------------------------------------------------------------- evaluation metrics
def _metric_dicts(golden_dir):
    gold = load_golden(golden_dir, "metrics_l1_vel_pck")
    gd = dict(zip([str(k) for k in gold["keys"]], gold["values"]))
    mean, var, batches = O.synth_metric_batches()
    return gd, mean, var, batches


def _compare_metrics(av, gd):
    assert set(av) == set(gd) and len(gd) == 2 + 2 * 52 + 2 + 1
    for k, v in gd.items():
        if k.endswith("_L1") or k.endswith("_VelL1"):
            assert abs(av[k] - v) <= 1e-12 * abs(v), (k, av[k], v)
        else:                       # the reference keeps PCK in fp32 (metrics.py:274): its running sums round at 6e-8
            assert abs(av[k] - v) <= 2e-7, (k, av[k], v)


def test_metrics_oracle_matches_reference_golden(golden_dir):
    gd, mean, var, batches = _metric_dicts(golden_dir)
    _compare_metrics(O.pose_metrics(batches, mean, var, MASK), gd)
    assert 0.5 < gd["test_pck"] < 0.99 and gd["test_pck_0.1"] < gd["test_pck_0.2"]      # thresholds discriminate


@pytest.mark.gpu
def test_cuda_pose_metrics_match_oracle_and_golden(golden_dir):
    import mixstage_b200 as M
    gd, mean, var, batches = _metric_dicts(golden_dir)
    pm = M.PoseMetrics((mean, var), num_joints=52, mask=MASK, alphas=(0.1, 0.2))
    for y, g in batches:
        pm(y.cuda(), g.cuda())
    av = pm.get_averages("test")
    _compare_metrics(av, gd)
    ref = O.pose_metrics(batches, mean, var, MASK)
    for k in ref:
        assert abs(av[k] - ref[k]) <= 2e-7, k
    pm.reset()
    pm(batches[1][0].cuda(), batches[1][1].cuda())
    one = O.pose_metrics(batches[1:], mean, var, MASK)
    assert abs(pm.get_averages("test")["test_L1"] - one["test_L1"]) < 1e-12


# ---------------------------------------------------------------------------- host logic on the CPU kernel specification
@pytest.mark.parametrize("name", list(CASES))
def test_preprocessor_host_logic_on_cpu_spec(monkeypatch, golden_dir, name):
    """Column gather built from the joint mask, feature coding, output shapes/dtypes and the separate entry points of
    PosePreprocessor, with ms_pose_prepare routed to tests/cpu_emu.py."""
    import mixstage_b200 as M
    cpu_emu.install(monkeypatch)
    feats, K = CASES[name]
    gold = load_golden(golden_dir, name)
    x, mean, var, centers, labels, soft, y, zn, inv = _oracle(name)
    pp = M.PosePreprocessor(num_joints=52, mask=MASK, muvar=(mean, var), centers=centers, feats=feats, device="cpu")
    assert pp.P == 96 and pp.D == centers.shape[1] and pp.cols.tolist()[:3] == [1, 2, 3] and pp.cols.tolist()[48] == 53
    yg, lg = pp(x)
    assert (lg.numpy() == gold["labels"]).all()
    np.testing.assert_allclose(yg.numpy(), gold["y"], rtol=1e-14, atol=0)
    _, sg = pp(x, soft_labels=True)
    np.testing.assert_allclose(sg.numpy(), gold["soft"], rtol=1e-12, atol=1e-15)
    assert torch.equal(pp.predict(x), lg) and torch.equal(pp.normalize(x), yg)
    with pytest.raises(M.MixStageError):
        pp(x[..., :100])
    with pytest.raises(M.MixStageError):
        M.PosePreprocessor(feats=["pose", "spatial"], device="cpu")
    with pytest.raises(M.MixStageError):
        M.PosePreprocessor(centers=torch.zeros(8, 7), device="cpu")


def test_metrics_host_logic_on_cpu_spec(monkeypatch, golden_dir):
    """AverageMeter-style weighting of PoseMetrics over two batches of different size, kernel routed to tests/cpu_emu.py."""
    import mixstage_b200 as M
    cpu_emu.install(monkeypatch)
    gd, mean, var, batches = _metric_dicts(golden_dir)
    pm = M.PoseMetrics((mean, var), num_joints=52, mask=MASK, alphas=(0.1, 0.2), device="cpu")
    for y, g in batches:
        pm(y, g)
    _compare_metrics(pm.get_averages("test"), gd)
    with pytest.raises(M.MixStageError):
        pm(batches[0][0], batches[1][1])
