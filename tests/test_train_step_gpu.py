"""TrainStep on the GPU: the CUDA-graph replay must reproduce the eager step (same kernels), and one graphed step must
match the oracle's train-loop body (forward, backward, clip, Adam) from the same state."""
import pytest
import torch

import mixstage_b200 as M
import mixstage_oracle as O
from model_cases import build
from oracle_cases import D_SEED, G_SEED, leafify

pytestmark = pytest.mark.gpu


def _mk(precision, graphs, B=8, T=64):
    spec = O.Spec(num_speakers=4)
    M.set_precision(precision)
    G, D, gan = build(spec, T, "cuda", torch.float64)
    G.thresh.value, G.thresh.iters = 1.0, 1000
    ts = M.TrainStep(gan, use_graphs=graphs)
    batch = [t.cuda() for t in O.synth_inputs(B, T, spec)]
    audio, pose, labels, style = batch
    return spec, G, D, ts, (audio, labels, pose, style)


@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
def test_graph_replay_equals_eager(precision):
    try:
        _, Gg, Dg, tsg, batch = _mk(precision, True)
        _, Ge, De, tse, _ = _mk(precision, False)
        kinds = ["G", "D", "G", "G", "D", "D"]
        for it, kind in enumerate(kinds):
            # same starting state for both (Adam's sign-like first updates make free-running trajectories diverge
            # chaotically from 1e-7 differences, see test_train_step_cpu.py): copy eager -> graphed before each step
            with torch.no_grad():
                for fa, fb in ((tsg.fG, tse.fG), (tsg.fD, tse.fD)):
                    for name in ("p", "m", "v", "step_count"):
                        getattr(fa, name).copy_(getattr(fb, name))
                for (ma, mb) in ((Gg, Ge), (Dg, De)):
                    for ba, bb in zip(ma.buffers(), mb.buffers()):
                        ba.copy_(bb)
            fg, lg = tsg.step(*batch, kind=kind)
            fe, le = tse.step(*batch, kind=kind)
            torch.cuda.synchronize()
            # identical kernels; only fp32 reductions (split-K tiles) reorder
            assert float((fg - fe).norm() / fe.norm()) < 1e-4, (it, kind)
            assert torch.allclose(lg, le, rtol=1e-4, atol=1e-5), (it, kind, lg, le)
            diff = (tsg.fG.p - tse.fG.p).abs() if kind == "G" else (tsg.fD.p - tse.fD.p).abs()
            assert float((diff > 2e-5).double().mean()) < 1e-2, (it, kind)     # lr*sign(g) flips of near-zero gradients
        assert tsg.replays == len(kinds) and len(tsg.graphs) == 2
        assert int(tsg.fG.step_count) == 3 and int(tsg.fD.step_count) == 3
        for (k, a), (_, b) in zip(Gg.state_dict().items(), Ge.state_dict().items()):
            if k.endswith("num_batches_tracked"):
                assert int(a) == int(b), k
    finally:
        M.set_precision("fp32")


def test_graphed_step_matches_oracle():
    try:
        spec, G, D, ts, batch = _mk("bf16x3", True, B=16)
        audio, labels, pose, style = [t.cpu() for t in batch]
        sd = leafify(O.synth_state(O.g_state_shapes(spec), G_SEED))
        sdd = leafify(O.synth_state(O.d_state_shapes(spec.out_feats), D_SEED))
        p0 = ts.fG.p.clone()
        fake, losses = ts.step(*batch, kind="G")
        torch.cuda.synchronize()
        lg, ld = O.BNLog(), O.BNLog()
        f2, l2, _ = O.gan_forward(sd, sdd, spec, audio, labels, pose, style, step="G", log_g=lg, log_d=ld)
        sum(l2).backward()
        assert float((fake.cpu() - f2.detach()).norm() / f2.detach().norm()) < 1e-3
        for a, b in zip(losses.tolist(), l2):
            assert abs(a - float(b.detach())) < 1e-3 * max(1.0, abs(float(b.detach())))
        names = [n for n, p in G.named_parameters() if p.requires_grad]
        ps = [sd[n].detach().clone() for n in names]
        gs = [sd[n].grad if sd[n].grad is not None else torch.zeros_like(sd[n]) for n in names]
        ms, vs = [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
        O.clip_and_adam(ps, gs, ms, vs, 1)
        gp = dict(G.named_parameters())
        bad = tot = 0
        for n, p1 in zip(names, ps):
            bad += int(((gp[n].detach().cpu() - p1).abs() > 2e-5).sum())
            tot += p1.numel()
        assert bad <= 1e-2 * tot, (bad, tot)
        assert float((ts.fG.p - p0).abs().max()) <= 1.0001e-4          # |Adam step 1| == lr
        gsd = G.state_dict()
        for k, v in lg.updates.items():
            assert float((gsd[k].cpu().double() - v).abs().max()) < 1e-3, k
        for blk, cnt in lg.counts.items():
            assert int(gsd[blk + ".norm.num_batches_tracked"]) == cnt, blk
    finally:
        M.set_precision("fp32")
