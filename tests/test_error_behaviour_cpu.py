"""Error behaviour at the drop-in boundary (SURVEY.md section 8b "Errors"): the reference raises Python exceptions / asserts;
the B200 classes raise at the same places, and raise -- never fall back -- for what they do not accelerate."""
import pytest
import torch

import cpu_emu
import mixstage_b200 as M
import mixstage_oracle as O
from model_cases import MOD, build


def _inputs(spec, B=2, T=64):
    audio, pose, labels, style = O.synth_inputs(B, T, spec)
    return audio, pose, labels, style


def test_cpu_tensors_are_rejected_without_a_fallback():
    spec = O.Spec(num_speakers=2)
    G, D, gan = build(spec, 64, "cpu", torch.float64)
    audio, pose, labels, style = _inputs(spec)
    G.eval()
    with pytest.raises(M.MixStageError):
        with torch.no_grad():
            G([audio, labels], pose, input_modalities=MOD, style=style, sample_flag=1, description="test")
    with pytest.raises(M.MixStageError):
        D(pose)


def test_unet_length_asserts_like_the_reference(monkeypatch):
    """layers.py:136-138: T >= 16 and T % 32 == 0."""
    cpu_emu.install(monkeypatch)
    spec = O.Spec(num_speakers=2)
    G, D, gan = build(spec, 48, "cpu", torch.float64)
    audio, pose, labels, style = _inputs(spec, T=48)
    G.eval()
    with pytest.raises(AssertionError):
        with torch.no_grad():
            G([audio, labels], pose, input_modalities=MOD, style=style, sample_flag=1, description="test")


def test_unaccelerated_options_raise(monkeypatch):
    cpu_emu.install(monkeypatch)
    spec = O.Spec(num_speakers=2)
    # dropout p > 0 (reference ConvNormRelu applies nn.Dropout; jobs use p = 0)
    with pytest.raises((AssertionError, M.MixStageError, NotImplementedError)):
        M.JointLateClusterSoftStyle4_G(time_steps=64, out_feats=96, p=0.5, style_dict={0: 0, 1: 1}, shape={MOD[0]: [64, 64]})
    G, D, gan = build(spec, 64, "cpu", torch.float64)
    audio, pose, labels, style = _inputs(spec)
    G.eval()
    # text modalities are outside the accelerated path
    with pytest.raises((NotImplementedError, M.MixStageError)):
        with torch.no_grad():
            G([audio, audio, labels], pose, input_modalities=["text/w2v", MOD[0]], style=style, sample_flag=1, description="test")
    # a criterion other than L1Loss
    with pytest.raises((NotImplementedError, M.MixStageError, ValueError)):
        M.GAN(G, D, criterion="MSELoss", no_grad=0, input_modalities=MOD)
    # style indices that do not tile (B, T)
    with pytest.raises(M.MixStageError):
        with torch.no_grad():
            G([audio, labels], pose, input_modalities=MOD, style=style[:, :7].contiguous(), sample_flag=1, description="test")
    # unused parameter holders exist for state_dict compatibility but must not be called
    with pytest.raises(Exception):
        G.style_dec(torch.zeros(2, 1, 64, 256))


def test_state_dict_round_trip_with_the_oracle_state():
    spec = O.Spec(num_speakers=4)
    G, D, gan = build(spec, 64, "cpu", torch.float64)
    sd = G.state_dict()
    want = O.synth_state(O.g_state_shapes(spec), 7, torch.float64)
    assert set(sd) == set(want)
    for k in want:
        assert tuple(sd[k].shape) == tuple(want[k].shape), k
        assert torch.allclose(sd[k].double(), want[k].double(), atol=1e-6), k
    G2, _, _ = build(spec, 64, "cpu", torch.float64)
    missing = G2.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys


def test_train_step_on_cpu_spec_never_uses_graphs(monkeypatch):
    cpu_emu.install(monkeypatch)
    from mixstage_b200 import train_step
    monkeypatch.setattr(train_step, "call", M._lib.call)
    monkeypatch.setattr(train_step, "stream", lambda: None)
    spec = O.Spec(num_speakers=2)
    G, D, gan = build(spec, 64, "cpu", torch.float64)
    ts = M.TrainStep(gan, use_graphs=False)
    assert ts.use_graphs is False                      # CPU: graphs are never used
    audio, pose, labels, style = _inputs(spec)
    G.thresh.value, G.thresh.iters = 1.0, 1000
    fake, losses = ts.step(audio, labels, pose, style, kind="D")
    assert tuple(fake.shape) == (2, 64, 96) and tuple(losses.shape) == (5,)
