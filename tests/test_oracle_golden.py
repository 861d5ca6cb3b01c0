"""Pins oracle/mixstage_oracle.py against the golden vectors produced by the unmodified
reference (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import mixstage_oracle as O
from oracle_cases import CASES, load_golden, run_oracle

TOL = 2e-6   # golden tensors are stored as fp32 of an fp64 run


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(golden_dir, name):
    gold = load_golden(golden_dir, name)
    res = run_oracle(name)
    np.testing.assert_allclose(res["pose"].numpy(), gold["pose"], rtol=0, atol=TOL)
    np.testing.assert_allclose(np.array(res["losses"]), gold["losses"], rtol=1e-10, atol=1e-12)
    soft = res["aux"]["labels_cap_soft"].detach().numpy().reshape(gold["labels_cap_soft"].shape)
    np.testing.assert_allclose(soft, gold["labels_cap_soft"], rtol=0, atol=TOL)
    assert (soft.argmax(-1) == gold["cluster_argmax"]).all()       # bit-exact cluster assignment
    if "g_grad_names" in gold:
        sd, sdd = res["sd"], res["sdd"]
        for n, v in zip(gold["g_grad_names"], gold["g_grad_norms"]):
            g = sd[str(n)].grad
            got = 0.0 if g is None else float(g.norm())
            assert abs(got - v) <= 1e-9 * max(1.0, v) + 1e-12, (n, got, v)
        for n, v in zip(gold["d_grad_names"], gold["d_grad_norms"]):
            assert abs(float(sdd[str(n)].grad.norm()) - v) <= 1e-9 * max(1.0, v) + 1e-12, n
        # parameters the reference leaves without a gradient must have none (or zero) here
        with_grad = set(map(str, gold["g_grad_names"]))
        for k, v in sd.items():
            if v.requires_grad and k not in with_grad and not k.startswith("style_dec_gr."):
                assert v.grad is None or float(v.grad.abs().max()) == 0.0, k
        for k in gold:
            if k.startswith("ggrad/"):
                np.testing.assert_allclose(sd[k[6:]].grad.numpy(), gold[k], rtol=0,
                                           atol=TOL * max(1.0, float(np.abs(gold[k]).max())))
            if k.startswith("dgrad/"):
                np.testing.assert_allclose(sdd[k[6:]].grad.numpy(), gold[k], rtol=0,
                                           atol=TOL * max(1.0, float(np.abs(gold[k]).max())))
            if k.startswith("gstat/"):
                got = res["log_g"].updates.get(k[6:], sd[k[6:]])
                np.testing.assert_allclose(got.numpy(), gold[k], rtol=0, atol=TOL)
            if k.startswith("dstat/"):
                got = res["log_d"].updates.get(k[6:], sdd[k[6:]])
                np.testing.assert_allclose(got.numpy(), gold[k], rtol=0, atol=TOL)
        for n, inc in zip(gold["nbt_names"], gold["nbt_incr"]):
            blk = str(n)[: -len(".norm.num_batches_tracked")]
            if blk.startswith("style_dec_gr.models.0."):
                blk = "style_dec." + blk[len("style_dec_gr.models.0."):]
            assert res["log_g"].counts.get(blk, 0) == int(inc), n
        for n, inc in zip(gold["d_nbt_names"], gold["d_nbt_incr"]):
            assert res["log_d"].counts.get(str(n)[: -len(".norm.num_batches_tracked")], 0) == int(inc), n


def test_state_dict_contract(golden_dir):
    """Key names/shapes of the oracle's parameter tables == the reference's state_dict."""
    import os
    want = {"G": {}, "D": {}}
    with open(os.path.join(golden_dir, "state_dict_keys.txt")) as f:
        for line in f:
            which, key, shape, dtype = line.split()
            want[which][key] = tuple(int(s) for s in shape.split("x")) if shape != "-" else ()
    from oracle_cases import CFG2
    got_g = {k: v[0] for k, v in O.g_state_shapes(CFG2).items()}
    got_d = {k: v[0] for k, v in O.d_state_shapes(96).items()}
    assert got_g == want["G"]
    assert got_d == want["D"]
    assert len(got_g) == 391


def test_algorithmic_flops_match_survey():
    from oracle_cases import CFG2, CFG5
    assert abs(O.flops_per_sequence(CFG2, 64, False) / 2e6 - 1041.60) < 0.01
    assert abs(O.flops_per_sequence(CFG2, 64, True) / 2e6 - 1052.50) < 0.01
    assert abs(O.flops_per_sequence(CFG5, 256, False) / 2e6 - 5843.58) < 0.01
    assert abs(O.flops_per_sequence(CFG5, 256, True) / 2e6 - 5887.76) < 0.01
