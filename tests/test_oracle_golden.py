"""The oracle (oracle/reasoning_core.py) against the golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py) and, when /root/reference is present, against the live reference itself."""
import pytest
import torch

import parity
from oracle import reasoning_core as rc
from oracle import ref_import


@pytest.mark.parametrize("name", parity.GOLDEN_CASES)
def test_oracle_matches_golden(name):
    z, meta = parity.load_golden(name)
    seed = None if meta["train_seed"] < 0 else int(meta["train_seed"])
    sd, (v, q, a), out = parity.oracle_case(meta["model"], int(meta["B"]), int(meta["num_ans"]), train_seed=seed,
                                            weight_seed=int(meta["weight_seed"]), input_seed=int(meta["input_seed"]),
                                            want_input_grads=True)
    assert parity.rel_err(out["logits"], z["logits"]) <= 1e-6
    assert abs(out["loss"].item() - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    for k, t in parity.flatten_alpha(out["alpha_dict"]).items():
        assert parity.rel_err(t, z["alpha." + k]) <= 1e-6, k
    floor = parity.golden_grad_floor(z)
    grads = dict(out["grads"])
    grads["__dv"], grads["__dq"] = out["dv"], out["dq"]
    for n in parity.golden_grad_names(z):
        if n.endswith("conv_att.conv.bias"):
            continue      # analytically zero: sum_i dz[i,g] = 0; pure rounding noise on both sides
        assert parity.compare_grad_to_golden(z, n, grads[n], floor) <= 1e-5, n


def test_param_table_matches_reference_counts():
    assert len(rc.param_shapes("ODA", 3000)) == 38
    assert len(rc.param_shapes("CoR2", 2000)) == 62
    n = lambda m, c: sum(int(torch.Size(s).numel()) for _, s in rc.param_shapes(m, c))
    assert n("ODA", 3000) == 7348434
    assert n("CoR2", 2000) == 11940244


@pytest.mark.skipif(not ref_import.available(), reason="reference sources only exist in the authoring container")
@pytest.mark.parametrize("model,C", [("ODA", 3000), ("CoR2", 2000)])
@pytest.mark.parametrize("seed", [None, 5])
def test_oracle_matches_live_reference(model, C, seed):
    sd = rc.synth_state_dict(model, C, seed=3)
    v, q, a = rc.synth_inputs(3, 36, C, seed=99)
    ref = ref_import.reference_step(model, sd, v, q, a, train_seed=seed, want_input_grads=True)
    drop = rc.no_drop if seed is None else rc.PhiloxDrop(seed)
    out = rc.step(model, sd, v, q, a, drop=drop, want_input_grads=True)
    assert parity.rel_err(out["logits"], ref["logits"]) <= 1e-6
    gmax = max(g.abs().max().item() for g in ref["grads"].values())
    for k, g in ref["grads"].items():
        if k.endswith("conv_att.conv.bias"):
            assert out["grads"][k].abs().max().item() <= 1e-6 * gmax
            continue
        assert parity.rel_err(out["grads"][k], g, 1e-6 * gmax) <= 1e-5, k
    assert parity.rel_err(out["dv"], ref["dv"]) <= 1e-5
    assert parity.rel_err(out["dq"], ref["dq"]) <= 1e-5
    # state_dict keys, shapes and ORDER are the reference's
    _, m = ref_import.build_model(model, C)
    ref_keys = [(k, tuple(p.shape)) for k, p in m.state_dict().items() if not k.startswith("seq2vec")]
    assert ref_keys == rc.param_shapes(model, C)


def test_n_step_chain_oracle_reduces_to_the_pinned_two_step_model():
    """oracle.cor_forward (the reference's blocks composed for any number of steps; UNPINNED beyond two) with steps = 2
    is the pinned cor2_forward, bit for bit, in eval and in train mode."""
    sd = rc.synth_state_dict("CoR2", 120, seed=2)
    v, q, _ = rc.synth_inputs(3, 36, 120, seed=8)
    for drop in (rc.no_drop, rc.PhiloxDrop(17)):
        y2, a2 = rc.cor_forward(sd, v, q, drop, 36, 2)
        y, a = rc.cor2_forward(sd, v, q, drop, 36)
        assert torch.equal(y2, y) and set(a2) == set(a)
        assert all(torch.equal(x, z) for x, z in zip(a2["alpha2"], a["alpha2"])) and torch.equal(a2["feature"], a["feature"])
    assert rc.cor_layers(2) == rc.COR2_LAYERS
