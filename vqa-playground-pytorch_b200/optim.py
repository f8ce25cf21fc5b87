"""Fused optimizer step for the reasoning core: global-norm clipping + Adam over the flat gradient buffer.

Replaces, in the reference's train step (train.py:82-86, optimizer built at train.py:292),

    nn.utils.clip_grad_norm_(model.parameters(), 0.25)
    optimizer.step()                                   # torch.optim.Adam(params, lr=cf.lr)

by two kernel launches (`vqa_clip_adam_step`).  It needs the gradients in ONE flat buffer, which is what
`parallel.GradSink` / `DataParallelEngine` give (p.grad of every core parameter is a view into it).  The class is a
`torch.optim.Optimizer`, so `lr_scheduler.ExponentialLR(optimizer, 0.5 ** (1 / 50000))` (train.py:296) and the
reference's scheduler-before-optimizer order keep working unchanged; re-creating it resets the moments, which is the
reference's per-epoch optimizer reset (train.py:726-729).
"""
import ctypes as C

import torch

from . import _lib


class FusedClipAdam(torch.optim.Optimizer):
    """Adam(lr, betas, eps) with optional clip_grad_norm_(max_norm) folded in.

        sink = DataParallelEngine(model)            # or parallel.GradSink(model.core_parameters(), model.MODEL)
        opt = FusedClipAdam(sink, lr=1e-4, clip_grad=0.25)
        loss.backward(); sink.wait(); opt.step()

    Only the parameters of the sink are updated (the reasoning core); a model with an external `seq2vec` keeps its
    own optimizer for that module."""

    def __init__(self, sink, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, clip_grad=None, write_clipped_grads=True):
        if not sink.flat.is_cuda:
            raise ValueError("FusedClipAdam runs on the GPU only (there is no CPU path)")
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("FusedClipAdam: bad hyper-parameter")
        self.sink = sink
        params = [p for p in sink.params]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.clip_grad = float(clip_grad) if clip_grad else 0.0
        self.write_clipped_grads = bool(write_clipped_grads)
        self.exp_avg = torch.zeros_like(sink.flat)
        self.exp_avg_sq = torch.zeros_like(sink.flat)
        self.scratch = torch.zeros(1, device=sink.flat.device, dtype=torch.float32)
        self.step_count = 0
        self._segs = (_lib.ParamSegment * len(params))()
        for i, p in enumerate(params):
            lo, hi = sink.offsets[i]
            if not p.is_contiguous():
                raise ValueError("FusedClipAdam: parameters must be contiguous")
            self._segs[i].param, self._segs[i].offset, self._segs[i].numel = p.data_ptr(), lo, hi - lo

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        group = self.param_groups[0]
        self.step_count += 1
        for i, p in enumerate(self.sink.params):          # parameters may have been re-allocated (.to(), load_state_dict)
            self._segs[i].param = p.data_ptr()
        pr = _lib.ClipAdam()
        pr.nsegs, pr.segs = len(self._segs), self._segs
        pr.grads_flat, pr.exp_avg, pr.exp_avg_sq = self.sink.flat.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        pr.total = self.sink.flat.numel()
        pr.lr, (pr.beta1, pr.beta2), pr.eps = float(group["lr"]), group["betas"], float(group["eps"])
        pr.step, pr.max_norm, pr.write_clipped_grads = self.step_count, self.clip_grad, int(self.write_clipped_grads)
        pr.scratch = self.scratch.data_ptr()
        stream = C.c_void_p(torch.cuda.current_stream(self.sink.flat.device).cuda_stream)
        _lib.check(_lib.lib().vqa_clip_adam_step(C.byref(pr), stream), "vqa_clip_adam_step")
        return loss

    def zero_grad(self, set_to_none=False):
        """The backward plan zero-fills the flat buffer itself (one memset) — nothing to do between steps."""
        return None

    def grad_norm(self):
        """||g||_2 of the last step (device tensor; only meaningful when clipping is on)."""
        return self.scratch.sqrt()
