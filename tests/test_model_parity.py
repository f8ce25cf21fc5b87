"""Parity proper: the CUDA path, called through the product's public API (config.<M>.Model -> C ABI),
against (i) the golden vectors made by the unmodified reference and (ii) the oracle on seeded inputs."""
import pytest
import torch

import parity

pytestmark = pytest.mark.gpu


def _check_against_oracle(out, ref, tol=parity.FP32_TOL):
    assert parity.rel_err(out["logits"], ref["logits"]) <= tol
    assert abs(out["loss"].item() - ref["loss"].item()) <= tol * abs(ref["loss"].item())
    fa, fb = parity.flatten_alpha(out["alpha_dict"]), parity.flatten_alpha(ref["alpha_dict"])
    assert set(fa) == set(fb)
    for k in fb:
        assert parity.rel_err(fa[k], fb[k]) <= tol, k
    gmax = max(g.abs().max().item() for g in ref["grads"].values())
    for k, g in ref["grads"].items():
        if k.endswith("conv_att.conv.bias"):      # analytically zero on both sides
            assert out["grads"][k].abs().max().item() <= 1e-6 * gmax, k
            continue
        assert parity.rel_err(out["grads"][k], g, 1e-6 * gmax) <= tol, k


@pytest.mark.parametrize("name", parity.GOLDEN_CASES)
def test_cuda_matches_reference_golden(cuda, name):
    from oracle import reasoning_core as rc
    z, meta = parity.load_golden(name)
    seed = None if meta["train_seed"] < 0 else int(meta["train_seed"])
    model, B, C = meta["model"], int(meta["B"]), int(meta["num_ans"])
    sd = rc.synth_state_dict(model, C, seed=int(meta["weight_seed"]))
    v, q, a = rc.synth_inputs(B, 36, C, seed=int(meta["input_seed"]))
    out = parity.run_cuda_model(model, sd, v, q, a, train_seed=seed)
    assert parity.rel_err(out["logits"], z["logits"]) <= parity.FP32_TOL
    assert abs(out["loss"].item() - float(z["loss"])) <= parity.FP32_TOL * abs(float(z["loss"]))
    for k, t in parity.flatten_alpha(out["alpha_dict"]).items():
        assert parity.rel_err(t, z["alpha." + k]) <= parity.FP32_TOL, k
    floor = parity.golden_grad_floor(z)
    for n in parity.golden_grad_names(z):
        if n.startswith("__") or n.endswith("conv_att.conv.bias"):
            continue
        assert parity.compare_grad_to_golden(z, n, out["grads"][n], floor) <= parity.FP32_TOL, n


@pytest.mark.parametrize("model,C", [("CoR2", 2000), ("ODA", 3000)])
@pytest.mark.parametrize("B,N,seed", [(2, 36, None), (7, 36, 11), (3, 10, None), (3, 10, 5), (2, 100, None),
                                      (2, 100, 9), (33, 36, 3)])
def test_cuda_matches_oracle(cuda, model, C, B, N, seed):
    sd, (v, q, a), _ = parity.oracle_case(model, B, C, N=N, weight_seed=21, input_seed=B * 1000 + N, run=False)
    out = parity.run_cuda_model(model, sd, v, q, a, N=N, train_seed=seed)
    ref = parity.oracle_with_same_relu_pattern(model, sd, v, q, a, out, N=N, train_seed=seed, tie_tol=1e-5)
    _check_against_oracle(out, ref)


@pytest.mark.parametrize("model,C", [("CoR2", 2000), ("ODA", 3000)])
@pytest.mark.parametrize("B,N,seed", [(5, 36, None), (9, 36, 13), (3, 100, 2), (130, 36, 21)])
def test_tensor_core_fp32_parity_mode(cuda, model, C, B, N, seed):
    """precision='tf32x3' (tcgen05, error-compensated 3xTF32, fp32 accumulate in TMEM) must meet the same 1e-4
    bound as the CUDA-core fp32 path (BASELINE.json: fp32 mode max relative error <= 1e-4)."""
    sd, (v, q, a), _ = parity.oracle_case(model, B, C, N=N, weight_seed=31, input_seed=B * 77 + N, run=False)
    out = parity.run_cuda_model(model, sd, v, q, a, N=N, train_seed=seed, precision="tf32x3")
    # gradients are compared on the SAME ReLU activation pattern; the patterns may differ only at rounding-level
    # ties (|z| <= 1e-4 max|z|, the forward tolerance) — one tie alone moves a wgrad row by ~1/sqrt(rows)
    ref = parity.oracle_with_same_relu_pattern(model, sd, v, q, a, out, N=N, train_seed=seed, tie_tol=1e-4)
    _check_against_oracle(out, ref)


@pytest.mark.parametrize("model,C", [("CoR2", 2000), ("ODA", 3000)])
def test_tf32_throughput_mode(cuda, model, C):
    """precision='tf32' (single-pass TF32 tensor cores): the reduced-precision bound of BASELINE.json, <= 2e-2 on
    logits (fp32 reference), attention weights likewise."""
    sd, (v, q, a), ref = parity.oracle_case(model, 16, C, train_seed=4, weight_seed=5, input_seed=6)
    out = parity.run_cuda_model(model, sd, v, q, a, train_seed=4, precision="tf32")
    assert parity.rel_err(out["logits"], ref["logits"]) <= 2e-2
    fa, fb = parity.flatten_alpha(out["alpha_dict"]), parity.flatten_alpha(ref["alpha_dict"])
    for k in fb:
        assert parity.rel_err(fa[k], fb[k]) <= 2e-2, k


@pytest.mark.parametrize("model,C", [("CoR2", 2000), ("ODA", 3000)])
def test_batch_of_one(cuda, model, C):
    """The reference crashes at B=1 (e.squeeze() drops the batch dim); parity is taken from
    the B=2 oracle run with the row duplicated."""
    sd, (v, q, a), ref = parity.oracle_case(model, 2, C, weight_seed=4, input_seed=8)
    v2, q2, a2 = v[:1].repeat(2, 1, 1), q[:1].repeat(2, 1), a[:1].repeat(2, 1)
    from oracle import reasoning_core as rc
    ref2 = rc.step(model, sd, v2, q2, a2)
    out = parity.run_cuda_model(model, sd, v[:1], q[:1], a[:1])
    assert parity.rel_err(out["logits"], ref2["logits"][:1]) <= parity.FP32_TOL
    gmax = max(g.abs().max().item() for g in ref2["grads"].values())
    for k, g in ref2["grads"].items():
        if k.endswith("conv_att.conv.bias"):
            continue
        assert parity.rel_err(out["grads"][k] * 2.0, g, 1e-6 * gmax) <= parity.FP32_TOL, k


def test_full_size_properties(cuda):
    """BASELINE config (CoR2, B=256, N=36): size-independent properties instead of an oracle run:
    every alpha column sums to 1, batch rows are independent (a row's logits do not depend on its
    neighbours), eval is deterministic, the train mask depends only on (seed, index)."""
    from oracle import reasoning_core as rc
    from vqa_playground_pytorch_b200.config import CoR2
    B, N, C = 256, 36, 2000
    sd = rc.synth_state_dict("CoR2", C, seed=10)
    m = CoR2.Model(None, C, precision="fp32")      # CUDA-core path: bitwise reproducible (no split-K atomics)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(1234)
    v = torch.relu(torch.randn(B, N, 2048, device="cuda", generator=g))
    q = 0.1 * torch.relu(torch.randn(B, 2400, device="cuda", generator=g))
    y1 = m({"v": v, "q_idxes": q})
    a1 = torch.cat(m.alpha_dict["alpha1"], 2)
    assert torch.allclose(a1.sum(1), torch.ones(B, 4, device="cuda"), atol=1e-5)
    y2 = m({"v": v, "q_idxes": q})
    assert torch.equal(y1, y2)
    ys = m({"v": v[100:132], "q_idxes": q[100:132]})
    assert parity.rel_err(ys, y1[100:132]) <= 1e-5
    m.train()
    m.fixed_seed = 99
    t1 = m({"v": v, "q_idxes": q})
    t2 = m({"v": v, "q_idxes": q})
    assert torch.equal(t1, t2)
    m.fixed_seed = 100
    t3 = m({"v": v, "q_idxes": q})
    assert not torch.equal(t1, t3)
    ts = m({"v": v[:32], "q_idxes": q[:32]})          # rows 0..31 keep their mask indices
    m.fixed_seed = 99
    ts = m({"v": v[:32], "q_idxes": q[:32]})
    assert parity.rel_err(ts, t1[:32]) <= 1e-5


@pytest.mark.parametrize("name,C", [("CoR2", 2000), ("ODA", 3000)])
def test_full_size_train_step_is_reproducible_on_tensor_cores(cuda, name, C):
    """B=256, N=36, 3xTF32, train mode: the forward/backward plans run on two lanes and read their dropout masks from
    a per-step cache — a missing dependency between the lanes (a kernel reading masks or packed weights before they
    are written) shows up as run-to-run differences far above the split-K rounding noise.  The batch-slice check ties
    the cached masks to the (seed, index) contract: rows 0..31 alone must give the rows 0..31 of the full batch."""
    import importlib
    from oracle import reasoning_core as rc
    cf = importlib.import_module("vqa_playground_pytorch_b200.config." + name)
    B, N = 256, 36
    m = cf.Model(None, C, precision="tf32x3")
    m.load_state_dict(rc.synth_state_dict(name, C, seed=10))
    m = m.cuda().train()
    g = torch.Generator(device="cuda").manual_seed(77)
    v = torch.relu(torch.randn(B, N, 2048, device="cuda", generator=g))
    q = 0.1 * torch.relu(torch.randn(B, 2400, device="cuda", generator=g))
    a = torch.softmax(torch.randn(B, C, device="cuda", generator=g), 1)
    from vqa_playground_pytorch_b200 import ops
    runs = []
    for _ in range(3):
        m.fixed_seed = 4321
        for p in m.parameters():
            p.grad = None
        out = m({"v": v, "q_idxes": q})
        ops.kld_loss(out, a).backward()
        torch.cuda.synchronize()
        runs.append((out.detach().clone(), [p.grad.detach().clone() for p in m.core_parameters()]))
    for out, grads in runs[1:]:
        assert parity.rel_err(out, runs[0][0]) <= 1e-5
        gmax = max(t.abs().max().item() for t in runs[0][1])
        names = [k for k, _ in m.named_parameters()]
        worst = max((parity.rel_err(t, t0, 1e-6 * gmax), names[i]) for i, (t, t0) in enumerate(zip(grads, runs[0][1]))
                    if not names[i].endswith("conv_att.conv.bias"))      # analytically zero: rounding noise only
        assert worst[0] <= 1e-4, worst
    m.fixed_seed = 4321
    part = m({"v": v[:32], "q_idxes": q[:32]})
    assert parity.rel_err(part, runs[0][0][:32]) <= 1e-5


def test_cuda_graph_step_matches_eager(cuda):
    """engine.GraphedStep (fwd+loss+bwd captured once, replayed) gives the eager step's loss and gradients for the
    same device-resident Philox key, and draws a fresh dropout mask on every replay."""
    import torch
    from oracle import reasoning_core as rc
    from vqa_playground_pytorch_b200 import ops
    from vqa_playground_pytorch_b200.config import CoR2
    from vqa_playground_pytorch_b200.engine import GraphedStep, kld_loss
    B, C = 8, 2000
    sd = rc.synth_state_dict("CoR2", C, seed=3)
    v, q, a = (t.cuda() for t in rc.synth_inputs(B, 36, C, seed=9))
    m = CoR2.Model(None, C, precision="tf32x3")
    m.load_state_dict(sd)
    m = m.cuda().train()
    sample = {"v": v, "q_idxes": q, "a": a}
    step = GraphedStep(m, sample, warmup=2, seed=1000)
    l1 = step(sample).item()
    key1 = int(m.seed_device.item())
    g1 = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    l2 = step(sample).item()
    assert int(m.seed_device.item()) == key1 + 1
    assert l1 != l2                                   # new key -> new mask -> different loss
    # eager step with the key of the first replay
    m2 = CoR2.Model(None, C, precision="tf32x3")
    m2.load_state_dict(sd)
    m2 = m2.cuda().train()
    m2.fixed_seed = key1
    loss = kld_loss(m2(sample), a)
    loss.backward()
    assert abs(loss.item() - l1) <= 1e-5 * abs(l1)
    gmax = max(p.grad.abs().max().item() for p in m2.parameters())
    for n, p in m2.named_parameters():
        if n.endswith("conv_att.conv.bias"):          # analytically zero, rounding noise on both sides
            continue
        assert parity.rel_err(g1[n], p.grad, 1e-6 * gmax) <= 1e-4, n
