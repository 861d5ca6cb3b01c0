"""Host-side planners of the fused blocks (mixstage_b200/igemm.py) on the geometries the model really runs: whatever the
cost model prefers, the kernels' structural requirements must hold -- tile widths the epilogue can walk, k-slices that
all own at least one k-step, TMEM residency only when the accumulators fit, weight-gradient slices of >= 2 row tiles."""
import itertools

import pytest

from mixstage_b200 import igemm

# (B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, groups): audio encoder 1-7, UNet k3 / k4s2, classifier, grouped decoder, D
LAYERS = [
    (16, 64, 64, 64, 64, 4, 4, 2, 2, 1, 1, 1), (16, 32, 32, 64, 128, 3, 3, 1, 1, 1, 1, 1), (16, 32, 32, 128, 128, 4, 4, 2, 2, 1, 1, 1),
    (16, 16, 16, 128, 256, 3, 3, 1, 1, 1, 1, 1), (16, 16, 16, 256, 256, 4, 4, 2, 2, 1, 1, 1), (16, 8, 8, 256, 256, 3, 3, 1, 1, 1, 1, 1),
    (16, 8, 8, 256, 256, 3, 8, 1, 1, 1, 3, 1), (16, 1, 64, 256, 256, 1, 3, 1, 1, 0, 1, 1), (16, 1, 64, 256, 256, 1, 4, 1, 2, 0, 1, 1),
    (16, 1, 2, 256, 256, 1, 3, 1, 1, 0, 1, 1), (16, 1, 64, 272, 256, 1, 3, 1, 1, 0, 1, 1), (16, 1, 64, 272, 2048, 1, 3, 1, 1, 0, 1, 1),
    (16, 1, 64, 2048, 2048, 1, 3, 1, 1, 0, 1, 8), (16, 1, 64, 96, 64, 1, 4, 1, 2, 0, 1, 1), (16, 1, 32, 64, 128, 1, 4, 1, 2, 0, 1, 1),
]


def _plans(c, B):
    _, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, g = c
    Ho, Wo = (H + 2 * ph - kh) // sh + 1, (W + 2 * pw - kw) // sw + 1
    geo = (B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, g, Ho, Wo)
    if not igemm.fwd_supported(Cin, Cout, g, sh, sw, H, W):
        return None, None
    pf = igemm.make_fwd(*geo)
    pd = igemm.make_dgrad(*geo) if igemm.dgrad_supported(Cin, Cout, g, sh, sw, H, W, kh, kw) else None
    return pf, pd


@pytest.mark.parametrize("c,B,npass", list(itertools.product(LAYERS, (2, 16, 128), (1, 3))))
def test_block_plan_is_runnable(c, B, npass):
    pf, pd = _plans(c, B)
    if pf is None:
        pytest.skip("layer not on the tensor-core path")
    for plan, stats in ((pf, True), (pd, False)):
        if plan is None:
            continue
        d = plan.desc
        bn, split = igemm.block_plan(d, npass, stats)
        k_steps = d.ntaps * d.cchunks
        assert bn % (32 if stats else 16) == 0 or bn == d.class_n
        assert 16 <= bn <= 256 and bn <= max(d.class_n, 16)
        assert 1 <= split <= k_steps
        per = -(-k_steps // split)
        assert (split - 1) * per < k_steps                      # every k-slice owns at least one k-step
        stage = (2 if npass > 1 else 1) * (16384 + bn * 128)
        assert 196608 // stage >= 2                             # the ring holds at least two stages
        if split == 1 and igemm.block_resident(d, bn):
            tiles = igemm._tiles_m(d) * d.num_classes * (-(-d.class_n // bn))
            per_cta = -(-tiles // min(tiles, 148))
            assert per_cta * bn <= 512 and per_cta <= 16        # TMEM columns / accumulator slots


@pytest.mark.parametrize("B", [2, 16, 128])
def test_wgrad_multi_splits(B):
    descs = [p.desc for p in (_plans(c, B)[0] for c in LAYERS) if p is not None]
    out = igemm.wgrad_multi_splits(descs)
    assert len(out) == len(descs)
    ctas = 0
    for d, (split, ct) in zip(descs, out):
        rt = igemm.wgrad_row_tiles(d)
        assert ct == 256 and 1 <= split <= max(1, rt // 2)      # >= 2 k-steps per CTA
        per = -(-rt // split)
        assert (split - 1) * per < rt
        ctas += igemm.wgrad_tiles(d, ct) * split
    assert ctas >= min(148, sum(igemm.wgrad_tiles(d, 256) for d in descs))


def test_normalise_split():
    for total in range(1, 70):
        for s in range(1, 40):
            n = igemm._normalise_split(total, s)
            per = -(-total // n)
            assert 1 <= n <= max(1, min(s, total)) and (n - 1) * per < total
