"""Speech2Gesture_G baseline (reference speech2gesture.py:13-40): the oracle restatement against
golden vectors of the executed reference, the mixstage_b200 module graph on the CPU kernel specification, and the CUDA
path on the GPU."""
import numpy as np
import pytest
import torch

import cpu_emu
import mixstage_b200 as M
import mixstage_oracle as O
from oracle_cases import leafify, load_golden

S2G_SEED, P, B, T = 9, 96, 8, 64
TOL = 2e-6


def _oracle(training):
    sd = leafify(O.synth_state(O.s2g_state_shapes(P), S2G_SEED))
    spec = O.Spec(num_speakers=4, out_feats=P)
    audio, pose, _, _ = O.synth_inputs(B, T, spec)
    log = O.BNLog()
    if training:
        out, _ = O.s2g_forward(sd, audio, T, True, log)
        loss = (out - pose).abs().mean()
        loss.backward()
    else:
        with torch.no_grad():
            out, _ = O.s2g_forward(sd, audio, T, False, log)
        loss = (out - pose).abs().mean()
    return sd, log, out.detach(), float(loss.detach())


@pytest.mark.parametrize("name,training", [("s2g_eval", False), ("s2g_train", True)])
def test_oracle_matches_reference_golden(golden_dir, name, training):
    gold = load_golden(golden_dir, name)
    sd, log, out, loss = _oracle(training)
    np.testing.assert_allclose(out.numpy(), gold["pose"], rtol=0, atol=TOL)
    assert abs(loss - float(gold["losses"][0])) < 1e-10
    if training:
        for n, v in zip(gold["grad_names"], gold["grad_norms"]):
            assert abs(float(sd[str(n)].grad.norm()) - v) <= 1e-9 * max(1.0, v), n
        for k in gold:
            if k.startswith("grad/"):
                np.testing.assert_allclose(sd[k[5:]].grad.numpy(), gold[k], rtol=0, atol=TOL * max(1.0, float(np.abs(gold[k]).max())))
            if k.startswith("gstat/"):
                np.testing.assert_allclose(log.updates[k[6:]].numpy(), gold[k], rtol=0, atol=TOL)
        for n, inc in zip(gold["nbt_names"], gold["nbt_incr"]):
            assert log.counts.get(str(n)[: -len(".norm.num_batches_tracked")], 0) == int(inc), n


def test_state_dict_contract():
    G = M.Speech2Gesture_G(time_steps=T, out_feats=P)
    want = {k: v[0] for k, v in O.s2g_state_shapes(P).items()}
    got = {k: tuple(v.shape) for k, v in G.state_dict().items()}
    assert got == want


def _run_module(device, training, precision="fp32"):
    from mixstage_b200 import ops
    old = ops.get_precision()
    ops.set_precision(precision)
    try:
        G = M.Speech2Gesture_G(time_steps=T, out_feats=P)
        G.load_state_dict(O.synth_state(O.s2g_state_shapes(P), S2G_SEED, torch.float32))
        G = G.to(device=device, dtype=torch.float64)
        spec = O.Spec(num_speakers=4, out_feats=P)
        audio, pose, _, _ = (t.to(device) for t in O.synth_inputs(B, T, spec))
        G.train(training)
        if training:
            out, il = G(audio, pose)
            loss = (out - pose).abs().mean()
            loss.backward()
        else:
            with torch.no_grad():
                out, il = G([audio], pose, input_modalities=["audio/log_mel_400"])      # the trainer passes a list
            loss = (out - pose).abs().mean()
        assert il == [] and out.dtype == torch.float64
        return G, out.detach().cpu(), float(loss.detach())
    finally:
        ops.set_precision(old)


def _check(G, out, loss, training, tol, gtol):
    sd, log, ref, rloss = _oracle(training)
    assert float((out - ref).norm() / ref.norm()) < tol
    assert abs(loss - rloss) < tol * max(1.0, abs(rloss))
    if training:
        for n, p in G.named_parameters():
            r = sd[n].grad
            err = float((p.grad.cpu().double() - r).norm())
            assert err <= gtol * float(r.norm()) + 1e-9, (n, err, float(r.norm()))
        gsd = G.state_dict()
        for k, v in log.updates.items():
            assert float((gsd[k].cpu().double() - v).abs().max()) < 1e-3 if tol > 1e-4 else 1e-5, k
        for blk, cnt in log.counts.items():
            assert int(gsd[blk + ".norm.num_batches_tracked"]) == cnt, blk


@pytest.mark.parametrize("training", [False, True])
def test_module_graph_on_cpu_spec(monkeypatch, training):
    cpu_emu.install(monkeypatch)
    G, out, loss = _run_module("cpu", training)
    _check(G, out, loss, training, 2e-5, 1e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("training,precision,tol", [(False, "fp32", 1e-3), (True, "fp32", 1e-3), (True, "bf16x3", 1e-3),
                                                    (False, "bf16", 2e-2)])
def test_cuda_path_matches_oracle(training, precision, tol):
    G, out, loss = _run_module("cuda", training, precision)
    _check(G, out, loss, training, tol, 5e-2)
