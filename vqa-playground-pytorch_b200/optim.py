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
    own optimizer for that module.

    device_clock=True keeps the step counter (and, with lr_gamma, the learning rate of the reference's per-iteration
    ExponentialLR(0.5 ** (1 / 50000)), train.py:296, stepped BEFORE the optimizer as train.py:75-76 does) in device
    memory, which makes step() the same launch every iteration: `engine.GraphedStep(..., optimizer=opt)` captures it
    together with forward, loss and backward.

    state_dict() / load_state_dict() use torch.optim.Adam's layout (per-parameter 'step', 'exp_avg', 'exp_avg_sq' in
    the parameter order of the sink == the reference's state_dict order), so the reference's ckpt_optim.pth.tar
    (train.py:280, :684) loads here and a checkpoint written here loads into torch.optim.Adam."""

    def __init__(self, sink, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, clip_grad=None, write_clipped_grads=True,
                 device_clock=False, lr_gamma=None):
        if not sink.flat.is_cuda:
            raise ValueError("FusedClipAdam runs on the GPU only (there is no CPU path)")
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("FusedClipAdam: bad hyper-parameter")
        if lr_gamma is not None and not device_clock:
            raise ValueError("FusedClipAdam: lr_gamma (in-kernel ExponentialLR) needs device_clock=True; without it "
                             "use torch.optim.lr_scheduler.ExponentialLR on this optimizer")
        self.sink = sink
        params = [p for p in sink.params]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.clip_grad = float(clip_grad) if clip_grad else 0.0
        self.write_clipped_grads = bool(write_clipped_grads)
        dev = sink.flat.device
        self.exp_avg = torch.zeros_like(sink.flat)
        self.exp_avg_sq = torch.zeros_like(sink.flat)
        self.scratch = torch.zeros(1024, device=dev, dtype=torch.float32)      # VQA_CLIP_SCRATCH_FLOATS
        self.step_count = 0
        self.device_clock = bool(device_clock)
        self.lr_gamma = float(lr_gamma) if lr_gamma else 0.0
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int64) if device_clock else None
        self.lr_dev = torch.full((1,), float(lr), device=dev, dtype=torch.float64) if device_clock else None
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
        if self.device_clock:
            pr.step_dev, pr.lr_dev, pr.lr_gamma = self.step_dev.data_ptr(), self.lr_dev.data_ptr(), self.lr_gamma
        stream = C.c_void_p(torch.cuda.current_stream(self.sink.flat.device).cuda_stream)
        _lib.check(_lib.lib().vqa_clip_adam_step(C.byref(pr), stream), "vqa_clip_adam_step")
        return loss

    def set_lr(self, lr):
        """Write the learning rate (host value and, with the device clock, the device copy a captured step reads)."""
        self.param_groups[0]["lr"] = float(lr)
        if self.device_clock:
            self.lr_dev.fill_(float(lr))

    def steps_done(self):
        return int(self.step_dev.item()) if self.device_clock else self.step_count

    def zero_grad(self, set_to_none=False):
        """The backward plan zero-fills the flat buffer itself (one memset) — nothing to do between steps."""
        return None

    def grad_norm(self):
        """||g||_2 of the last step (device tensor; only meaningful when clipping is on)."""
        return self.scratch[:1].sqrt()

    # ---- checkpointing (train.py:250-284 saves / restores optimizer.state_dict()) -------------------------------
    def state_dict(self):
        n = self.steps_done()
        state = {}
        if n > 0:
            for i, p in enumerate(self.sink.params):
                lo, hi = self.sink.offsets[i]
                state[i] = {"step": torch.tensor(float(n)),
                            "exp_avg": self.exp_avg[lo:hi].view(p.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[lo:hi].view(p.shape).clone()}
        g = dict(self.param_groups[0])
        if self.device_clock:
            g["lr"] = float(self.lr_dev.item())
        g["params"] = list(range(len(self.sink.params)))
        for k, dflt in (("weight_decay", 0), ("amsgrad", False), ("maximize", False), ("foreach", None),
                        ("capturable", False), ("differentiable", False), ("fused", None),
                        ("decoupled_weight_decay", False)):
            g.setdefault(k, dflt)                     # keys torch.optim.Adam.load_state_dict expects to find
        return {"state": state, "param_groups": [g]}

    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.sink.params):
            raise ValueError("FusedClipAdam.load_state_dict: expected one parameter group with %d parameters" %
                             len(self.sink.params))
        g = groups[0]
        if g.get("weight_decay", 0) or g.get("amsgrad", False) or g.get("maximize", False):
            raise ValueError("FusedClipAdam.load_state_dict: weight_decay / amsgrad / maximize are not supported "
                             "(the reference uses plain Adam, train.py:292)")
        self.param_groups[0].update(lr=float(g["lr"]), betas=tuple(g["betas"]), eps=float(g["eps"]))
        state, ids = state_dict["state"], g["params"]
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        n = 0
        steps = set()
        for i, pid in enumerate(ids):
            st = state.get(pid, state.get(str(pid)))
            if st is None:
                continue
            lo, hi = self.sink.offsets[i]
            self.exp_avg[lo:hi].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[lo:hi].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("FusedClipAdam.load_state_dict: parameters disagree on the step count %s" % sorted(steps))
        n = steps.pop() if steps else 0
        self.step_count = n
        if self.device_clock:
            self.step_dev.fill_(n)
            self.lr_dev.fill_(float(g["lr"]))
