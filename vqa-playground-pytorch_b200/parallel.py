"""Data-parallel training support: one process per GPU, gradients all-reduced over NVLink peer memory
(libvqacore's own kernel, vqa_peer_allreduce_f32) or over NCCL.

Replaces the reference's only multi-GPU mechanism, single-process `nn.DataParallel`
(train.py:517): every sample is independent through the whole forward, so the batch is
sharded across ranks and the ONE collective is a SUM all-reduce of the parameter gradients
(the loss is sum-reduced, train.py:541, so gradients are summed — not averaged — to match a
single-GPU run at the global batch).

Gradients live in one flat fp32 buffer laid out in BACKWARD-COMPLETION order; `p.grad` of every
core parameter is a view into it and the backward kernels write there directly
(ops.ModelCoreFn with a grad sink: no autograd copies).  The buffer is cut into a few
contiguous buckets; a bucket is all-reduced on a side stream as soon as the backward plan has
enqueued its last producer, so communication overlaps the remaining backward kernels.
The host logic (layout, bucketing, reduction semantics) also runs on CPU tensors with the
gloo backend, which is how tests/test_parallel.py covers world_size 2 without GPUs.

Two transports for the one collective:
  "peer"  the flat buffer is SYMMETRIC memory (torch.distributed._symmetric_memory: the same allocation on every
          GPU, mapped into every process); vqa_peer_allreduce_f32 reduces a bucket in place with loads / stores over
          NVLink, uses no shared memory (its CTAs sit next to the persistent GEMMs of the backward instead of taking
          their SMs) and leaves bit-identical sums on every rank;
  "nccl"  dist.all_reduce on the communication stream (any backend; the only choice on CPU / gloo).
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

# state_dict index groups in the order the backward plans (csrc/model.cu) finish them
_COR2_ORDER = [[52, 53], list(range(44, 52)), [42, 43], list(range(16, 24)), list(range(34, 42)), [32, 33],
               list(range(24, 32)), [2, 3], [56, 57, 60, 61], [54, 55, 58, 59], [14, 15], list(range(6, 14)), [0, 1],
               [4, 5]]
_ODA_ORDER = [[36, 37], list(range(16, 36)), list(range(6, 14)), [4, 5], [0, 1], [2, 3, 14, 15]]
COMPLETION_ORDER = {"CoR2": _COR2_ORDER, "ODA": _ODA_ORDER}


def plan_buckets(sizes_in_order, num_buckets, tail=0):
    """Cut a sequence of tensor sizes into <= num_buckets contiguous buckets of roughly equal bytes.
    tail > 0: the last `tail` entries form the final bucket on their own (the groups that complete at the very end of
    the backward: their all-reduce is the part nothing can hide, so it is kept as small as the plan allows) and the
    rest is cut into <= num_buckets - 1.  Returns a list of (first_index, last_index_exclusive)."""
    if not sizes_in_order:
        return []
    if tail > 0 and num_buckets > 1 and len(sizes_in_order) > tail:
        head = plan_buckets(sizes_in_order[:-tail], num_buckets - 1)
        return head + [(len(sizes_in_order) - tail, len(sizes_in_order))]
    total = sum(sizes_in_order)
    target = total / max(1, num_buckets)
    buckets, start, acc = [], 0, 0
    for i, s in enumerate(sizes_in_order):
        acc += s
        remaining_buckets = num_buckets - len(buckets) - 1
        if acc >= target and remaining_buckets > 0 and i + 1 < len(sizes_in_order):
            buckets.append((start, i + 1))
            start, acc = i + 1, 0
    buckets.append((start, len(sizes_in_order)))
    return buckets


TAIL_GROUPS = {"CoR2": 2, "ODA": 1}      # groups whose gradients appear only at the end of the backward plan


class GradSink:
    """Flat gradient buffer + its per-parameter views (see module docstring)."""

    def __init__(self, params, model_name, num_buckets=4, alloc=None):
        order = [i for grp in COMPLETION_ORDER[model_name] for i in grp]
        assert sorted(order) == list(range(len(params))), "completion order must cover every parameter once"
        self.order = order
        dev = params[0].device
        total = sum(p.numel() for p in params)
        # `alloc(n)` supplies the storage (symmetric memory for the peer transport); it may return more than n elements
        self.storage = alloc(total) if alloc is not None else torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat = self.storage[:total]
        self.slices = [None] * len(params)
        offsets, off = {}, 0
        for i in order:
            n = params[i].numel()
            self.slices[i] = self.flat[off:off + n].view(params[i].shape)
            offsets[i] = (off, off + n)
            off += n
        # buckets are cut at completion-group boundaries
        group_sizes = [sum(params[i].numel() for i in grp) for grp in COMPLETION_ORDER[model_name]]
        self.bucket_groups = plan_buckets(group_sizes, num_buckets, tail=TAIL_GROUPS.get(model_name, 0))
        self.bucket_ranges = []
        groups = COMPLETION_ORDER[model_name]
        for g0, g1 in self.bucket_groups:
            first, last = groups[g0][0], groups[g1 - 1][-1]
            self.bucket_ranges.append((offsets[first][0], offsets[last][1]))
        # 16-byte vector transport: interior boundaries move DOWN to a multiple of 4 floats (the <= 3 elements that
        # change sides wait for the later bucket, whose groups complete later), the end moves up into the padding
        self.vector_ranges = []
        for k, (lo, hi) in enumerate(self.bucket_ranges):
            lo4 = lo // 4 * 4
            hi4 = hi // 4 * 4 if k + 1 < len(self.bucket_ranges) else (hi + 3) // 4 * 4
            self.vector_ranges.append((lo4, hi4))
        self.accumulate = False
        self.params = list(params)
        self.offsets = [offsets[i] for i in range(len(params))]      # (first, last) element of each parameter's slice
        self.bind()

    def bind(self):
        """p.grad of every parameter is its slice of the flat buffer.  Re-done after every backward: the reference's
        step calls optimizer.zero_grad() (train.py:77), which in current torch sets p.grad = None — the kernels keep
        writing into the flat buffer regardless, but torch.optim.Adam / clip_grad_norm_ skip parameters without
        .grad, so a stale binding would silently stop training."""
        for p, s in zip(self.params, self.slices):
            if p.grad is not s:
                p.grad = s

    def after_backward(self):
        self.bind()


class DataParallelEngine(GradSink):
    """Wraps a config.<M>.Model for one-process-per-GPU data parallelism.

        engine = DataParallelEngine(model)            # after dist.init_process_group
        loss = ...; loss.backward(); engine.wait()    # grads now hold the global SUM
    """

    def __init__(self, model, num_buckets=4, process_group=None, allreduce="auto"):
        """allreduce: "peer" (NVLink peer-memory kernel; CUDA, one node, <= 8 ranks), "nccl" (dist.all_reduce), or
        "auto" = peer when it can be set up, else nccl (VQA_ALLREDUCE overrides "auto")."""
        self.model = model
        self.group = process_group
        params = model.core_parameters()
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        if allreduce == "auto":
            allreduce = os.environ.get("VQA_ALLREDUCE", "auto")
        if allreduce not in ("auto", "peer", "nccl"):
            raise ValueError("allreduce must be 'auto', 'peer' or 'nccl', got %r" % (allreduce,))
        self.peer = None
        alloc = None
        if self.world_size > 1 and params[0].is_cuda and allreduce in ("auto", "peer"):
            alloc = self._symmetric_alloc(params[0].device, strict=allreduce == "peer")
        num_buckets = int(os.environ.get("VQA_BUCKETS", num_buckets))
        super().__init__(params, model.MODEL, num_buckets, alloc=alloc)
        if alloc is not None and self._symm is not None:
            self._rendezvous_peers()
        self.transport = "peer" if self.peer is not None else ("nccl" if self.world_size > 1 else "none")
        self.nvls = bool(self.peer and self.peer.get("multicast"))
        self.is_cuda = self.flat.is_cuda
        self.comm_stream = torch.cuda.Stream(device=self.flat.device) if self.is_cuda else None
        self._pending = []
        self.defer = False          # True: the caller triggers the reduction itself (engine.GraphedStep)
        # One CUDA event per gradient group of the backward plan (include/vqacore.h: group_events): the plan records
        # event k when group k is complete, and bucket b is reduced on the communication stream as soon as the
        # events of ITS groups have fired - the rest of the backward keeps running underneath.
        self.group_events = None
        if self.is_cuda and self.world_size > 1:
            self.group_events = [torch.cuda.Event() for _ in COMPLETION_ORDER[model.MODEL]]
            for ev in self.group_events:
                ev.record()             # creates the handle (torch events are lazy)
        model.grad_sink = self

    # ---- peer transport ---------------------------------------------------------------------------------------
    def _symmetric_alloc(self, device, strict):
        """Returns alloc(n) -> zeroed symmetric fp32 storage of >= n elements (padded to 16 bytes), or None when
        symmetric memory is unavailable (strict: raise instead)."""
        self._symm = None
        try:
            import torch.distributed._symmetric_memory as symm
            if self.world_size > 8:
                raise RuntimeError("the peer transport handles up to 8 ranks of one node")
            self._symm = symm
        except Exception as exc:
            if strict:
                raise RuntimeError("allreduce='peer' requested but symmetric memory is unavailable: %s" % exc)
            return None

        def alloc(n):
            try:
                t = self._symm.empty((n + 3) // 4 * 4, dtype=torch.float32, device=device)
                t.zero_()
                return t
            except Exception as exc:
                if strict:
                    raise RuntimeError("allreduce='peer': symmetric allocation failed: %s" % exc)
                self._symm = None
                return torch.zeros(n, device=device, dtype=torch.float32)
        return alloc

    def _rendezvous_peers(self):
        """Exchange the mappings of the gradient and signal buffers (collective); falls back to NCCL on failure."""
        from . import _lib
        try:
            group = self.group if self.group is not None else dist.group.WORLD
            if hasattr(self._symm, "enable_symm_mem_for_group"):
                try:
                    self._symm.enable_symm_mem_for_group(group.group_name)
                except Exception:
                    pass
            nsig = int(_lib.lib().vqa_peer_allreduce_signal_bytes()) // 4
            self._signals = self._symm.empty(nsig, dtype=torch.int32, device=self.storage.device)
            self._signals.zero_()
            hb = self._symm.rendezvous(self.storage, group)
            hs = self._symm.rendezvous(self._signals, group)
            ok = torch.tensor([1], device=self.storage.device)
        except Exception as exc:
            if os.environ.get("VQA_ALLREDUCE") == "peer":
                raise
            import warnings
            warnings.warn("peer all-reduce unavailable (%s); using NCCL" % exc)
            ok = torch.tensor([0], device=self.storage.device)
            hb = hs = None
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)      # every rank or none
        torch.cuda.synchronize()
        if int(ok.item()) == 1:
            self._handles = (hb, hs)
            self.peer = {"buffers": [int(x) for x in hb.buffer_ptrs], "signals": [int(x) for x in hs.buffer_ptrs]}
            # NVLS: the multicast mapping of the gradient buffer, when the fabric offers one (all ranks or none)
            # measured (profiles/r2_experiments.md): the in-switch reduction wins at 8 ranks (half the NVLink bytes) and
            # loses to plain peer loads/stores at 2 (no traffic advantage): VQA_PEER_MC = auto (>= 4 ranks) | 1 | 0
            mc = 0
            want_mc = os.environ.get("VQA_PEER_MC", "auto")
            if want_mc == "1" or (want_mc == "auto" and self.world_size >= 4):
                try:
                    mc = int(hb.multicast_ptr or 0)
                except Exception:
                    mc = 0
            flag = torch.tensor([1 if mc else 0], device=self.storage.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            self.peer["multicast"] = mc if int(flag.item()) == 1 else 0
            self.peer_max_ctas = int(os.environ.get("VQA_PEER_CTAS", "0"))
            self.peer_spin_ms = int(os.environ.get("VQA_PEER_SPIN_MS", "20000"))

    def peer_error(self):
        """True when a peer all-reduce gave up waiting for another rank (vqa_peer_allreduce_params.spin_limit_ms)."""
        return self.peer is not None and int(self._signals[-1].item()) != 0

    def _peer_reduce(self, k):
        from . import _lib
        lo, hi = self.vector_ranges[k]
        pr = _lib.PeerAllreduce()
        pr.world, pr.rank = self.world_size, self.rank
        for r in range(self.world_size):
            pr.buffers[r], pr.signals[r] = self.peer["buffers"][r], self.peer["signals"][r]
        pr.offset, pr.count = lo, hi - lo
        pr.max_ctas, pr.spin_limit_ms = self.peer_max_ctas, self.peer_spin_ms
        pr.multicast = self.peer["multicast"] or None
        pr.cta_threads = int(os.environ.get("VQA_PEER_THREADS", "0"))
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().vqa_peer_allreduce_f32(C.byref(pr), stream), "vqa_peer_allreduce_f32")

    def broadcast_parameters(self, src=0):
        """Every rank starts from rank 0's weights (the reference's DataParallel replicates each step)."""
        if self.world_size > 1:
            for p in self.model.parameters():
                dist.broadcast(p.data, src, group=self.group)

    def reduce_bucket(self, k, overlapped=True):
        """All-reduce (SUM) bucket k on the communication stream: after the events of the bucket's gradient groups
        (overlapped with the rest of the backward), or after everything enqueued so far (overlapped=False)."""
        if self.world_size == 1:
            return
        lo, hi = self.bucket_ranges[k]
        chunk = self.flat[lo:hi]
        if self.is_cuda:
            if overlapped and self.group_events is not None:
                g0, g1 = self.bucket_groups[k]
                for g in range(g0, g1):
                    self.comm_stream.wait_event(self.group_events[g])
            else:
                self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                if self.peer is not None:
                    self._peer_reduce(k)
                else:
                    self._pending.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self._pending.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def reduce_all(self, overlapped=True):
        for k in range(len(self.bucket_ranges)):
            self.reduce_bucket(k, overlapped)

    def after_backward(self):
        # called by ops.ModelCoreFn.backward right after the backward plan has been enqueued
        self.bind()
        if not self.defer:
            self.reduce_all()

    def wait(self):
        for w in self._pending:
            w.wait()                # the current stream waits for the collective's end event (capturable)
        self._pending = []
        if self.is_cuda and self.world_size > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
