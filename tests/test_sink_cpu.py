"""GradSink re-binds p.grad after a backward (CPU tensors: the binding logic needs no GPU)."""
import torch

from oracle import reasoning_core as rc


def test_sink_rebinds_grads_after_zero_grad_set_to_none():
    from vqa_playground_pytorch_b200.parallel import GradSink
    params = [torch.nn.Parameter(torch.zeros(s)) for _, s in rc.param_shapes("ODA", 20)]
    sink = GradSink(params, "ODA")
    opt = torch.optim.Adam(params, lr=1e-3)
    opt.zero_grad()                                  # torch default: set_to_none=True
    assert all(p.grad is None for p in params)
    sink.flat.fill_(1.0)                             # what the backward kernels do: write the flat buffer
    sink.after_backward()
    assert all(p.grad is not None and p.grad.data_ptr() == s.data_ptr() for p, s in zip(params, sink.slices))
    opt.step()
    assert all(bool((p != 0).all()) for p in params)
