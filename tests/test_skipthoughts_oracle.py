"""CPU: the question-encoder oracle (oracle/skipthoughts.py) against the fixtures made by the reference's own
BayesianGRU + nn.Embedding (oracle/make_golden_gru.py), and against those classes live when /root/reference exists."""
import os

import numpy as np
import pytest
import torch

from oracle import make_golden_gru as mg, ref_import, skipthoughts as st

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,seed,af", mg.CASES)
def test_oracle_matches_reference_fixture(name, seed, af):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    idx, dx = torch.from_numpy(z["idx"]), torch.from_numpy(z["dx"])
    i2, d2 = mg.inputs()
    assert torch.equal(idx, i2) and torch.equal(dx, d2)           # the committed inputs are the generator's
    sd = st.synth_state_dict(mg.V, seed=10, I=mg.I, H=mg.H)
    masks = st.seq_masks(seed, mg.B, mg.I, mg.H, mg.P) if seed is not None else None
    out = st.step(sd, idx, dx, af, masks)
    assert np.abs(out["x"].numpy() - z["x"]).max() <= 1e-6
    assert np.abs(out["hs"].numpy() - z["hs"]).max() <= 1e-6
    for k, g in out["grads"].items():
        ref = z["grad." + k]
        assert np.abs(g.numpy() - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max()), k
    assert np.abs(z["grad.embedding.weight"][0]).max() == 0.0     # padding_idx row
    assert np.abs(out["x"][3].numpy() - out["hs"][3, -1].numpy()).max() == 0.0   # all-PAD question: lengths-1 = -1 wraps


@pytest.mark.skipif(not ref_import.available(), reason="reference sources only exist in the authoring container")
@pytest.mark.parametrize("name,seed,af", mg.CASES)
def test_oracle_matches_live_reference(name, seed, af):
    idx, dx = mg.inputs()
    sd = st.synth_state_dict(mg.V, seed=10, I=mg.I, H=mg.H)
    masks = st.seq_masks(seed, mg.B, mg.I, mg.H, mg.P) if seed is not None else None
    x, hs, grads = mg.reference_step(sd, idx, dx, af, masks)
    out = st.step(sd, idx, dx, af, masks)
    assert (out["x"] - x).abs().max() <= 1e-6 and (out["hs"] - hs).abs().max() <= 1e-6
    for k, g in grads.items():
        assert (out["grads"][k] - g).abs().max() <= 1e-6 * max(1.0, g.abs().max().item()), k


def test_state_dict_keys_match_the_reference_encoder():
    """seq2vec.* keys a reference checkpoint holds (SkipThoughts.load_bayesiangru_state_dict, :957-973)."""
    import vqa_playground_pytorch_b200  # noqa: F401
    from vqa_playground_pytorch_b200 import blocks
    m = blocks.SkipThoughts(["PAD", "UNK", "what", "is"], af="relu")
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert shapes == st.param_shapes(4)
    assert sum(int(np.prod(s)) for s in shapes.values()) - 4 * 620 == 21_751_200     # 21.75 M
