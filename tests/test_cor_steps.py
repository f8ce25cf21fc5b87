"""CoR with a configurable number of attention steps (config.CoR2.Model(steps=...)).
steps = 2 composed block by block must equal the fused plan (which is pinned to the reference's goldens);
steps = 3 is compared with the reference's own blocks composed once more (oracle.reasoning_core.cor_forward) —
UNPINNED: the reference ships no 3-step model."""
import pytest
import torch

import parity

pytestmark = pytest.mark.gpu


def _grads(m):
    return {n: p.grad.detach().cpu() for n, p in m.named_parameters()}


@pytest.mark.parametrize("seed", [None, 7])
def test_composed_two_step_chain_equals_the_fused_plan(cuda, seed):
    from oracle import reasoning_core as rc
    from vqa_playground_pytorch_b200 import ops
    from vqa_playground_pytorch_b200.config import CoR2
    C, B = 2000, 5
    sd = rc.synth_state_dict("CoR2", C, seed=4)
    v, q, a = (t.cuda() for t in rc.synth_inputs(B, 36, C, seed=2))
    outs = []
    for compose in (False, True):
        m = CoR2.Model(None, C, precision="tf32x3", compose=compose)
        m.load_state_dict(sd)
        m = m.cuda().train(seed is not None)
        m.fixed_seed = seed
        y = m({"v": v, "q_idxes": q})
        ops.kld_loss(y, a).backward()
        outs.append((y.detach().cpu(), _grads(m), parity.flatten_alpha(m.alpha_dict)))
    (y0, g0, a0), (y1, g1, a1) = outs
    assert parity.rel_err(y1, y0) <= 2e-5
    assert set(a0) == set(a1)
    for k in a0:
        assert parity.rel_err(a1[k].cpu(), a0[k].cpu()) <= 2e-5, k
    gmax = max(t.abs().max().item() for t in g0.values())
    for k in g0:
        if k.endswith("conv_att.conv.bias"):
            continue
        assert parity.rel_err(g1[k], g0[k], 1e-6 * gmax) <= 1e-4, k


@pytest.mark.parametrize("seed", [None, 11])
def test_three_step_chain_matches_the_composed_reference_blocks(cuda, seed):
    from oracle import reasoning_core as rc
    from vqa_playground_pytorch_b200 import ops
    from vqa_playground_pytorch_b200.config import CoR2
    C, B, steps = 300, 4, 3
    m = CoR2.Model(None, C, precision="tf32x3", steps=steps)
    assert len(m.state_dict()) == 90 and CoR2.layers_for(3) == rc.cor_layers(3)
    torch.manual_seed(0)
    sd = {k: t.detach().clone() for k, t in m.state_dict().items()}
    v, q, a = rc.synth_inputs(B, 36, C, seed=5)
    m = m.cuda().train(seed is not None)
    m.fixed_seed = seed
    y = m({"v": v.cuda(), "q_idxes": q.cuda()})
    ops.kld_loss(y, a.cuda()).backward()
    assert set(m.alpha_dict) == {"alpha1", "alpha2", "alpha3", "feature", "v3_feature"}
    # oracle: the reference's blocks composed once more, same Philox masks
    leaves = {k: t.clone().requires_grad_() for k, t in sd.items()}
    drop = rc.no_drop if seed is None else rc.PhiloxDrop(seed)
    logits, alpha = rc.cor_forward(leaves, v, q, drop, 36, steps)
    rc.kld_loss(logits, a).backward()
    assert parity.rel_err(y, logits) <= parity.FP32_TOL
    fa, fb = parity.flatten_alpha(m.alpha_dict), parity.flatten_alpha(alpha)
    for k in fb:
        assert parity.rel_err(fa[k].cpu(), fb[k]) <= parity.FP32_TOL, k
    got = _grads(m)
    gmax = max(t.grad.abs().max().item() for t in leaves.values())
    # ReLU ties are not replayed here: allow the few rows a rounding-level tie moves, as the un-replayed test does
    bad = [(parity.rel_err(got[k], t.grad, 1e-6 * gmax), k) for k, t in leaves.items() if not k.endswith("conv_att.conv.bias")]
    worst = max(bad)
    print("3-step chain, seed %s: logits %.2e, worst gradient %.2e (%s)" % (seed, parity.rel_err(y, logits), worst[0], worst[1]))
    assert worst[0] <= 5e-2 and sum(e > parity.FP32_TOL for e, _ in bad) <= 3, sorted(bad)[-4:]
