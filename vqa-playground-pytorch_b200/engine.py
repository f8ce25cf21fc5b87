"""Step-level host logic around the core: the reference's train step (train.py:41-107) restated for the
one-process-per-GPU engine, and a pinned-memory input prefetcher that overlaps host->device copies of the
next batch with the current step (at >1e5 samples/s the reference's per-item Python loader and
synchronous `.cuda()` copies would starve the GPU).
"""
import torch

from . import ops


def kld_loss(logits, target):
    """KLDivLoss(size_average=False)(log_softmax(x), a) — train.py:536-544 — through the fused kernel."""
    return ops.kld_loss(logits, target)


def train_step(model, sample, optimizer=None, scheduler=None, engine=None, clip_grad=None):
    """One iteration of train.py:63-86: forward, KLD loss, (scheduler.step — the reference steps it BEFORE the
    optimizer, :75-76), backward, gradient all-reduce, optional clip_grad_norm_(0.25) (:82), optimizer step."""
    output = model(sample)
    loss = kld_loss(output, sample['a'])
    if scheduler is not None:
        scheduler.step()
    if optimizer is not None and (engine is None):
        optimizer.zero_grad()
    loss.backward()
    if engine is not None:
        engine.wait()
    fused = getattr(optimizer, "clip_grad", None) is not None      # optim.FusedClipAdam clips inside its step
    if clip_grad and not fused:
        torch.nn.utils.clip_grad_norm_(model.parameters(), clip_grad)
    if optimizer is not None:
        optimizer.step()
    return output, loss


@torch.no_grad()
def predict_answers(model, sample, a_vocab=None, eval_metric="OpenEnded"):
    """The body of the reference's eval loop for one batch (train.py:125-169): forward in eval mode, answer index per
    question on the GPU (ops.argmax_rows; MultipleChoice restricts it to sample['a_mc_idx'], train.py:153-164), ONE
    device->host copy of the B indices, then the id -> answer mapping.  Returns the list of
    {'question_id', 'answer'} items train.py:166-169 appends (answer = the index when no a_vocab is given), ready for
    official_test.test_local."""
    if eval_metric not in ("OpenEnded", "MultipleChoice"):
        raise ValueError("<train.py> %s is not allowed" % eval_metric)
    was_training = model.training
    model.eval()
    try:
        output = model(sample)
    finally:
        model.train(was_training)
    mc = sample.get("a_mc_idx") if eval_metric == "MultipleChoice" else None
    pred = ops.argmax_rows(output, mc).cpu().tolist()
    q_id = sample.get("q_id")
    q_ids = q_id.cpu().tolist() if torch.is_tensor(q_id) else (list(q_id) if q_id is not None else list(range(len(pred))))
    word = a_vocab.idx2word if a_vocab is not None else (lambda i: i)
    return [{"question_id": q_ids[j], "answer": word(int(pred[j]))} for j in range(len(pred))]


def pack_feature_shard(samples, keys=("v", "q_idxes", "a"), feature_key="v"):
    """Pre-packed input shard: the reference builds every sample in Python from h5py / redis
    (datasets.py:517-549, :905-970: `torch.Tensor(feature[idx])` per item, fp32) and copies it with a synchronous
    `.cuda()`.  Here a list of batches (dicts of CPU tensors) becomes PINNED host memory once, with the region
    features stored as bf16 — half the bytes over PCIe, widened exactly on the device (ops.cast_bf16_to_f32).
    A feature store that keeps bf16 on disk loses nothing: the model then sees exactly the stored values."""
    out = []
    for smp in samples:
        d = {}
        for k in keys:
            t = smp[k]
            if k == feature_key and t.dtype == torch.float32:
                t = t.to(torch.bfloat16)
            d[k] = t.contiguous().pin_memory()
        out.append(d)
    return out


class HostPrefetcher:
    """Iterates over host batches (dicts of PINNED tensors) yielding device dicts; the copy of batch i+1 runs
    on a side stream while batch i is being computed.  Two device staging slots, allocated ONCE per prefetcher and
    reused by every pass (`reset(batches)` starts another pass: building a new prefetcher per epoch would make the
    caching allocator cudaMalloc fresh buffers for the new copy stream, a device-wide synchronisation that cost 1 ms
    per step in the first version of the bench).  bf16 host tensors (pack_feature_shard) are copied as bf16 and
    widened to fp32 on the device."""

    def __init__(self, host_batches, device, keys=("v", "q_idxes", "a"), widen_into=None):
        """widen_into: dict of preallocated fp32 device tensors (e.g. GraphedStep.static): bf16 features are widened
        straight into them on the consumer's stream when the batch is handed out, instead of into a staging copy that
        the consumer would have to copy again."""
        self.batches = host_batches
        self.device = torch.device(device)
        self.keys = keys
        self.widen_into = widen_into
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [None, None]
        self.stage = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.bytes_per_batch = 0

    def reset(self, host_batches):
        """Another pass over (other) host batches of the same shapes with the same device buffers."""
        self.batches = host_batches
        return self

    def _allocate(self, slot, host):
        # on the CONSUMER's stream pool: the buffers live as long as the prefetcher and every cross-stream use is
        # ordered by the ready / consumed events
        widened = lambda k: self.widen_into is not None and k in self.widen_into
        self.slots[slot] = {k: torch.empty(host[k].shape, device=self.device, dtype=torch.float32 if
                                           host[k].dtype == torch.bfloat16 else host[k].dtype)
                            for k in self.keys if not (host[k].dtype == torch.bfloat16 and widened(k))}
        self.stage[slot] = {k: torch.empty(host[k].shape, device=self.device, dtype=torch.bfloat16)
                            for k in self.keys if host[k].dtype == torch.bfloat16}

    def _issue(self, i, slot):
        host = self.batches[i]
        if self.slots[slot] is None:
            self._allocate(slot, host)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])        # previous user of the slot has finished
            nbytes = 0
            for k in self.keys:
                if host[k].dtype == torch.bfloat16:
                    self.stage[slot][k].copy_(host[k], non_blocking=True)
                    if k in self.slots[slot]:
                        ops.cast_bf16_to_f32(self.stage[slot][k], out=self.slots[slot][k])
                else:
                    self.slots[slot][k].copy_(host[k], non_blocking=True)
                nbytes += host[k].numel() * host[k].element_size()
            self.bytes_per_batch = nbytes
            self.ready[slot].record(self.copy_stream)

    def __iter__(self):
        n = len(self.batches)
        if n == 0:
            return
        cur = torch.cuda.current_stream(self.device)
        for s in (0, 1):
            self.consumed[s].record(cur)
        self._issue(0, 0)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                self._issue(i + 1, slot ^ 1)
            cur.wait_event(self.ready[slot])
            out = dict(self.slots[slot])
            for k, st in self.stage[slot].items():
                if k not in out:
                    out[k] = ops.cast_bf16_to_f32(st, out=self.widen_into[k])
            yield out
            self.consumed[slot].record(cur)


class GraphedStep:
    """fwd + KLD loss + bwd (+ gradient all-reduce, + optimizer step) of one fixed-shape batch captured in a CUDA graph
    and replayed every step.

    Everything the step launches (the forward plan, the loss kernel, the backward plan — ~90 kernels and memsets)
    becomes one graph launch, which removes the per-launch CPU cost and most of the inter-kernel gaps.  The Philox
    dropout key lives in device memory (`model.seed_device`) and is advanced by a kernel captured at the head of
    the graph, so every replay draws a fresh mask; the model reads that key only inside this step (eager forwards
    keep drawing their own keys from ops.next_seed()).  The initial key comes from ops.next_seed() mixed with the
    rank, so `ops.manual_seed()` controls it and data-parallel ranks draw different masks.
    `optimizer` (an optim.FusedClipAdam with device_clock=True) puts clip_grad_norm_ + Adam of train.py:82-86 in
    the graph as well.

        step = GraphedStep(model, example_sample, engine)     # warms up, then captures
        loss = step(sample)                                   # copies the sample into the static buffers, replays
    """

    def __init__(self, model, example, engine=None, warmup=3, seed=None, capture_collectives=False, optimizer=None):
        dev = example["v"].device
        self.model, self.engine, self.optimizer = model, engine, optimizer
        if optimizer is not None and not getattr(optimizer, "device_clock", False):
            raise ValueError("GraphedStep: the optimizer must keep its step count on the device "
                             "(optim.FusedClipAdam(..., device_clock=True)) to be captured")
        self.static = {k: torch.empty_like(t) for k, t in example.items() if torch.is_tensor(t)}
        for k, t in self.static.items():
            t.copy_(example[k])
        if seed is None:
            rank = torch.distributed.get_rank() if (torch.distributed.is_available() and
                                                    torch.distributed.is_initialized()) else 0
            seed = (ops.next_seed() ^ (rank * 0x9E3779B97F4A7C15)) & 0x7FFFFFFFFFFFFFFF
        model.seed_device = torch.tensor([seed], dtype=torch.int64, device=dev)
        self._one = torch.ones((), dtype=torch.float32, device=dev)
        # capture_collectives: the bucketed NCCL all-reduces are captured INSIDE the graph, each forked off the backward
        # at its bucket's gradient-group events, so that communication overlaps the remaining backward kernels of the
        # same replay.  Otherwise the reduction runs after the replay (nothing to overlap with).
        self.capture_collectives = bool(capture_collectives) and engine is not None and engine.world_size > 1
        if optimizer is not None and engine is not None and engine.world_size > 1 and not self.capture_collectives:
            raise ValueError("GraphedStep: a captured optimizer step needs the all-reduce in the graph too "
                             "(capture_collectives=True)")
        if engine is not None:
            engine.defer = not self.capture_collectives
        # the warm-up steps update the parameters and the optimizer state when an optimizer is captured: restore after
        saved = None
        if optimizer is not None:
            saved = ([p.detach().clone() for p in model.parameters()], optimizer.exp_avg.clone(),
                     optimizer.exp_avg_sq.clone(), optimizer.step_dev.clone(), optimizer.lr_dev.clone())
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _lib
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.lib().vqa_launch_count()
        # capture on the SAME side stream the warm-up ran on: autograd caches each parameter's gradient-accumulator
        # stream at first use, and a capture that has to wait on another (uncaptured) stream is invalid
        # NCCL's watchdog thread queries events while we capture: keep the capture check thread-local in that case
        mode = {"capture_error_mode": "thread_local"} if self.capture_collectives else {}
        with torch.cuda.graph(self.graph, stream=side, **mode):
            self.loss = self._body()
        self.launches_per_replay = int(_lib.lib().vqa_launch_count() - n0)    # libvqacore kernels inside the graph
        if saved is not None:
            with torch.no_grad():
                for p, t in zip(model.parameters(), saved[0]):
                    p.copy_(t)
                optimizer.exp_avg.copy_(saved[1]); optimizer.exp_avg_sq.copy_(saved[2])
                optimizer.step_dev.copy_(saved[3]); optimizer.lr_dev.copy_(saved[4])
        model.seed_device.fill_(seed)            # replays continue from the initial key, whatever the warm-up drew

    def _body(self):
        self.model.use_seed_device = True
        try:
            if self.model.training:
                ops.seed_advance(self.model.seed_device)
            out = self.model(self.static)
            loss = kld_loss(out, self.static["a"])
            if self.engine is None:
                for p in self.model.parameters():
                    p.grad = None
            loss.backward(gradient=self._one)       # preallocated seed gradient: no fill kernel per step
            if self.capture_collectives:
                self.engine.wait()                  # joins the communication stream back into the capturing stream
            if self.optimizer is not None:
                self.optimizer.step()
        finally:
            self.model.use_seed_device = False
        self.logits = out
        return loss

    def __call__(self, sample=None):
        if sample is not None:
            for k, t in self.static.items():
                if sample[k].data_ptr() != t.data_ptr():
                    t.copy_(sample[k], non_blocking=True)
        self.graph.replay()
        if self.engine is not None and not self.capture_collectives:
            self.engine.reduce_all(overlapped=False)
            self.engine.wait()
        return self.loss
