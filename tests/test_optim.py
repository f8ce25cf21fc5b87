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
