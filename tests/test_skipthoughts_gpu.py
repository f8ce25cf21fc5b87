"""GPU: blocks.SkipThoughts (ops.BayesianGruFn -> libvqacore) against the oracle at the real sizes (620 -> 2400)."""
import pytest
import torch

import parity
from oracle import skipthoughts as st

pytestmark = pytest.mark.gpu


@pytest.fixture
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda", 0)


def _run(dev, af, seed, precision, B=5, T=7, V=40):
    from vqa_playground_pytorch_b200 import blocks
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(1, V, (B, T), generator=g)
    for b, ln in enumerate([T, 3, 1, 0, 5][:B]):
        idx[b, ln:] = 0
    dx = torch.randn(B, 2400, generator=g)
    sd = st.synth_state_dict(V, seed=10)
    m = blocks.SkipThoughts(["w%d" % i for i in range(V)], af=af, precision=precision).to(dev)
    m.load_state_dict(sd)
    m.train(seed is not None)
    m.fixed_seed = seed
    x = m(idx.to(dev))
    x.backward(dx.to(dev))
    masks = st.seq_masks(seed, B, 620, 2400, 0.25) if seed is not None else None
    ref = st.step(sd, idx, dx, af, masks)
    return m, x, ref


@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
@pytest.mark.parametrize("af,seed", [("relu", None), ("relu", 17), ("tanh", 23)])
def test_encoder_matches_oracle(cuda, af, seed, precision):
    m, x, ref = _run(cuda, af, seed, precision)
    assert parity.rel_err(x.detach().cpu(), ref["x"]) <= 1e-4
    grads = {k: p.grad.detach().cpu() for k, p in m.named_parameters()}
    gmax = max(g.abs().max().item() for g in ref["grads"].values())
    for k, g in ref["grads"].items():
        assert parity.rel_err(grads[k], g, 1e-6 * gmax) <= 2e-4, k
    assert grads["embedding.weight"][0].abs().max().item() == 0.0


def test_encoder_feeds_the_core(cuda):
    """seq2vec=SkipThoughts in front of the CoR2 core: one backward reaches the embedding table."""
    from vqa_playground_pytorch_b200 import blocks, ops
    import importlib
    cf = importlib.import_module("vqa_playground_pytorch_b200.config.CoR2")
    V, B, T = 30, 4, 6
    enc = blocks.SkipThoughts(["w%d" % i for i in range(V)], af="relu")
    model = cf.Model(None, 50, seq2vec=enc).to(cuda).train()
    g = torch.Generator().manual_seed(5)
    idx = torch.randint(1, V, (B, T), generator=g).to(cuda)
    v = torch.randn(B, 36, 2048, generator=g).abs().to(cuda)
    a = torch.softmax(torch.randn(B, 50, generator=g), 1).to(cuda)
    loss = ops.kld_loss(model({"v": v, "q_idxes": idx}), a)
    loss.backward()
    assert torch.isfinite(loss).item()
    gemb = model.seq2vec.embedding.weight.grad
    assert gemb is not None and gemb.abs().max().item() > 0.0 and gemb[0].abs().max().item() == 0.0


@pytest.mark.parametrize("model,C", [("CoR2", 2000), ("ODA", 3000)])
@pytest.mark.parametrize("seed", [None, 9])
def test_core_returns_the_question_embedding_gradient(cuda, model, C, seed):
    """dq of vqa_cor2_bwd / vqa_oda_bwd (vqa_model_bwd_params.dq) against autograd through the oracle."""
    import importlib
    from oracle import reasoning_core as rc
    from vqa_playground_pytorch_b200 import ops
    sd, (v, q, a), _ = parity.oracle_case(model, 6, C, train_seed=seed, run=False)
    cf = importlib.import_module("vqa_playground_pytorch_b200.config." + model)
    m = cf.Model(None, C)
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda).train(seed is not None)
    m.fixed_seed = seed
    qd = q.to(cuda).requires_grad_(True)
    ops.kld_loss(m({"v": v.to(cuda), "q_idxes": qd}), a.to(cuda)).backward()
    masks = {name: (ops.stash_tensor(name) > 0).cpu() for name in parity.RELU_STASHES[model]}
    drop = rc.no_drop if seed is None else rc.PhiloxDrop(seed)
    ref = rc.step(model, sd, v, q, a, drop=drop, want_input_grads=True, ties=rc.ReluTies(masks=masks))
    assert parity.rel_err(qd.grad.cpu(), ref["dq"]) <= 1e-4
