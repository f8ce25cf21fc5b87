"""Fused clip + Adam (vqa_clip_adam_step) against nn.utils.clip_grad_norm_ + torch.optim.Adam on the same gradients
(the reference's optimizer step, train.py:82-86 / :292)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("clip", [0.25, None])
def test_fused_clip_adam_matches_torch(cuda, clip):
    from oracle import reasoning_core as rc
    from vqa_playground_pytorch_b200 import ops
    from vqa_playground_pytorch_b200.config import ODA
    from vqa_playground_pytorch_b200.optim import FusedClipAdam
    from vqa_playground_pytorch_b200.parallel import GradSink
    C, B, N = 50, 6, 36
    sd = rc.synth_state_dict("ODA", C, seed=3)
    model = ODA.Model(None, C, precision="fp32")
    model.load_state_dict(sd)
    model = model.cuda().train()
    sink = GradSink(model.core_parameters(), "ODA")
    model.grad_sink = sink
    opt = FusedClipAdam(sink, lr=1e-3, clip_grad=clip)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, 0.5 ** (1 / 50000))

    twins = [p.detach().clone().requires_grad_() for p in model.core_parameters()]
    ref = torch.optim.Adam(twins, lr=1e-3)
    ref_sched = torch.optim.lr_scheduler.ExponentialLR(ref, 0.5 ** (1 / 50000))
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(3):
        v = torch.randn(B, N, 2048, device="cuda", generator=g).relu_()
        q = 0.1 * torch.randn(B, 2400, device="cuda", generator=g).relu_()
        a = torch.softmax(torch.randn(B, C, device="cuda", generator=g), 1)
        ops.manual_seed(100 + step)
        loss = ops.kld_loss(model({"v": v, "q_idxes": q}), a)
        loss.backward()
        for t, p in zip(twins, model.core_parameters()):
            t.grad = p.grad.detach().clone()
        if clip:
            total = torch.nn.utils.clip_grad_norm_(twins, clip)
        ref.step(); ref_sched.step()
        opt.step(); sched.step()
        if clip:
            assert abs(opt.grad_norm().item() - total.item()) <= 1e-5 * total.item()
            for t, p in zip(twins, model.core_parameters()):      # clip_grad_norm_ scales the gradients in place
                assert torch.allclose(p.grad, t.grad, rtol=1e-5, atol=1e-12)
        for t, p in zip(twins, model.core_parameters()):
            scale = max(t.abs().max().item(), 1e-12)
            assert (p.detach() - t.detach()).abs().max().item() <= 2e-6 * scale, step
    assert opt.param_groups[0]["lr"] == pytest.approx(ref.param_groups[0]["lr"])


def _small_oda(C=50):
    from oracle import reasoning_core as rc
    from vqa_playground_pytorch_b200.config import ODA
    from vqa_playground_pytorch_b200.parallel import GradSink
    model = ODA.Model(None, C, precision="fp32")
    model.load_state_dict(rc.synth_state_dict("ODA", C, seed=3))
    model = model.cuda().train()
    sink = GradSink(model.core_parameters(), "ODA")
    model.grad_sink = sink
    return model, sink


def _batch(g, B, C, N=36):
    v = torch.randn(B, N, 2048, device="cuda", generator=g).relu_()
    q = 0.1 * torch.randn(B, 2400, device="cuda", generator=g).relu_()
    a = torch.softmax(torch.randn(B, C, device="cuda", generator=g), 1)
    return {"v": v, "q_idxes": q, "a": a}


def test_zero_grad_set_to_none_does_not_stop_training(cuda):
    """The reference's step calls optimizer.zero_grad() (train.py:77); current torch sets p.grad = None there.  With a
    gradient sink the kernels write into the flat buffer regardless — the sink must re-bind p.grad after every backward,
    or torch.optim.Adam / clip_grad_norm_ silently skip every parameter."""
    from vqa_playground_pytorch_b200 import ops
    model, sink = _small_oda()
    opt = torch.optim.Adam(model.core_parameters(), lr=1e-3)
    g = torch.Generator(device="cuda").manual_seed(2)
    before = [p.detach().clone() for p in model.core_parameters()]
    for _ in range(2):
        s = _batch(g, 4, 50)
        opt.zero_grad()                                   # set_to_none=True is the default
        assert all(p.grad is None for p in model.core_parameters())
        ops.kld_loss(model(s), s["a"]).backward()
        assert all(p.grad is not None and p.grad.data_ptr() == sl.data_ptr()
                   for p, sl in zip(model.core_parameters(), sink.slices))
        assert torch.nn.utils.clip_grad_norm_(model.core_parameters(), 0.25).item() > 0
        opt.step()
    moved = sum(int((p.detach() != b).any()) for p, b in zip(model.core_parameters(), before))
    assert moved >= len(before) - 2          # everything but the analytically-zero conv_att bias gradient moves


def test_optimizer_state_dict_round_trip_and_torch_layout(cuda):
    """save -> load -> step equals stepping on (train.py:280,684 checkpoint the optimizer); the state has
    torch.optim.Adam's layout, so it loads into a stock Adam over the same parameters and vice versa."""
    from vqa_playground_pytorch_b200 import ops
    from vqa_playground_pytorch_b200.optim import FusedClipAdam
    g = torch.Generator(device="cuda").manual_seed(3)
    batches = [_batch(g, 4, 50) for _ in range(3)]

    def run(model, opt, batch, seed):
        ops.manual_seed(seed)
        ops.kld_loss(model(batch), batch["a"]).backward()
        opt.step()

    m1, s1 = _small_oda()
    o1 = FusedClipAdam(s1, lr=1e-3, clip_grad=0.25)
    run(m1, o1, batches[0], 1); run(m1, o1, batches[1], 2)
    sd_opt, sd_model = o1.state_dict(), {k: t.clone() for k, t in m1.state_dict().items()}
    assert set(sd_opt["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and float(sd_opt["state"][0]["step"]) == 2.0
    run(m1, o1, batches[2], 3)

    m2, s2 = _small_oda()
    m2.load_state_dict(sd_model)
    o2 = FusedClipAdam(s2, lr=5e-2, clip_grad=0.25)            # lr comes back from the checkpoint
    o2.load_state_dict(sd_opt)
    assert o2.step_count == 2 and o2.param_groups[0]["lr"] == 1e-3
    run(m2, o2, batches[2], 3)
    # same state, same gradients up to the ordering of atomic additions; Adam turns gradient noise into a fraction of lr
    for (n, p), r in zip(m2.named_parameters(), m1.core_parameters()):
        if n.endswith("conv_att.conv.bias"):          # analytically zero gradient: its update is sign(noise) * lr
            continue
        assert (p.detach() - r.detach()).abs().max().item() <= 0.05 * 1e-3, n          # << one update (lr = 1e-3)
    # torch.optim.Adam accepts the same dict, and its own state_dict loads here
    twins = [p.detach().clone().requires_grad_() for p in m2.core_parameters()]
    stock = torch.optim.Adam(twins, lr=1e-3)
    stock.load_state_dict(sd_opt)
    assert float(stock.state[twins[0]]["step"]) == 2.0
    o3 = FusedClipAdam(s2, lr=1e-3)
    o3.load_state_dict(stock.state_dict())
    lo, hi = s2.offsets[0]
    assert torch.equal(o3.exp_avg[lo:hi].view_as(twins[0]), stock.state[twins[0]]["exp_avg"])


def test_graph_captured_optimizer_step_matches_the_reference_order(cuda):
    """engine.GraphedStep(optimizer=FusedClipAdam(device_clock=True, lr_gamma=...)) replays fwd + loss + bwd + clip +
    Adam with the learning rate decayed BEFORE the update, the reference's scheduler-before-optimizer order
    (train.py:75-76, ExponentialLR(0.5 ** (1 / 50000)) at :296), and a re-created optimizer restarts the clock, which
    is the per-epoch reset of train.py:726-729.  Compared with eager steps driven by torch.optim.Adam + ExponentialLR
    on the same gradients (same Philox keys)."""
    from vqa_playground_pytorch_b200 import ops
    from vqa_playground_pytorch_b200.engine import GraphedStep
    from vqa_playground_pytorch_b200.optim import FusedClipAdam
    gamma = 0.5 ** (1 / 5.0)                                 # a fast decay makes an ordering mistake visible
    g = torch.Generator(device="cuda").manual_seed(4)
    batches = [_batch(g, 4, 50) for _ in range(4)]
    m1, s1 = _small_oda()
    opt = FusedClipAdam(s1, lr=1e-3, clip_grad=0.25, device_clock=True, lr_gamma=gamma)
    before = [p.detach().clone() for p in m1.core_parameters()]
    step = GraphedStep(m1, batches[0], optimizer=opt, warmup=2, seed=500)
    for p, b in zip(m1.core_parameters(), before):            # warm-up and capture leave no trace
        assert torch.equal(p, b)
    assert opt.steps_done() == 0
    for s in batches:
        step(s)
    assert opt.steps_done() == 4
    assert opt.lr_dev.item() == pytest.approx(1e-3 * gamma ** 4, rel=1e-12)

    m2, s2 = _small_oda()
    ref = torch.optim.Adam(m2.core_parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.ExponentialLR(ref, gamma)
    for i, s in enumerate(batches):
        m2.fixed_seed = 500 + i + 1                          # the key replay i draws
        loss = ops.kld_loss(m2(s), s["a"])
        sched.step()                                          # train.py:76, before the optimizer
        ref.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m2.core_parameters(), 0.25)
        ref.step()
    # The two sides compute their gradients separately (atomic ordering differs at the 1e-7 level) and Adam turns
    # noise-level gradients into a fraction of lr, so the bound is in units of lr: 5 % of one update.  Stepping the
    # scheduler AFTER the optimizer instead would change every update by 1 - gamma = 13 % (0.4 lr over the four steps).
    for (n, p), r in zip(m1.named_parameters(), m2.core_parameters()):
        if n.endswith("conv_att.conv.bias"):          # analytically zero gradient: its update is sign(noise) * lr
            continue
        assert (p.detach() - r.detach()).abs().max().item() <= 0.05 * 1e-3, n
    fresh = FusedClipAdam(s1, lr=1e-3, clip_grad=0.25, device_clock=True, lr_gamma=gamma)     # epoch reset
    assert fresh.steps_done() == 0 and fresh.lr_dev.item() == 1e-3 and not fresh.exp_avg.any()
