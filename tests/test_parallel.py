"""Host-side logic of the data-parallel engine on CPU tensors with the gloo backend (world_size 2):
flat gradient buffer layout, bucket plan, SUM all-reduce semantics."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import reasoning_core as rc


class FakeModel(torch.nn.Module):
    """Parameter container with the reference's state_dict layout (no compute)."""

    def __init__(self, name, C):
        super().__init__()
        self.MODEL = name
        self.plist = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(s)) for _, s in rc.param_shapes(name, C)])
        self.grad_sink = None

    def core_parameters(self):
        return list(self.plist)


def test_bucket_plan_covers_everything():
    from vqa_playground_pytorch_b200.parallel import COMPLETION_ORDER, GradSink, plan_buckets
    assert plan_buckets([5, 5, 5, 5], 2) == [(0, 2), (2, 4)]
    assert plan_buckets([100], 4) == [(0, 1)]
    assert plan_buckets([], 4) == []
    for name, C in (("CoR2", 2000), ("ODA", 3000)):
        n = len(rc.param_shapes(name, C))
        order = [i for g in COMPLETION_ORDER[name] for i in g]
        assert sorted(order) == list(range(n))
        m = FakeModel(name, C)
        sink = GradSink(m.core_parameters(), name, num_buckets=4)
        assert 1 <= len(sink.bucket_ranges) <= 4
        # buckets tile the flat buffer exactly, in order
        pos = 0
        for lo, hi in sink.bucket_ranges:
            assert lo == pos and hi > lo
            pos = hi
        assert pos == sink.flat.numel() == sum(p.numel() for p in m.core_parameters())
        # every p.grad is a view of the flat buffer with the parameter's shape
        for p in m.core_parameters():
            assert p.grad.shape == p.shape
            assert p.grad.untyped_storage().data_ptr() == sink.flat.untyped_storage().data_ptr()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, name, C):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vqa_playground_pytorch_b200.parallel import DataParallelEngine
    torch.manual_seed(rank)
    m = FakeModel(name, C)
    for p in m.parameters():
        p.data.normal_()
    eng = DataParallelEngine(m, num_buckets=3)
    eng.broadcast_parameters()
    ref0 = [p.data.clone() for p in m.parameters()]
    # "backward": each rank writes rank-dependent gradients straight into the sink slices
    for i, s in enumerate(eng.slices):
        s.fill_(float(rank + 1) * (i + 1))
    eng.after_backward()
    eng.wait()
    ok = True
    for i, p in enumerate(m.core_parameters()):
        expect = float(sum(r + 1 for r in range(world)) * (i + 1))        # SUM, not mean (train.py:541)
        ok &= bool(torch.all(p.grad == expect))
    # parameters identical across ranks after the broadcast
    for p, r in zip(m.parameters(), ref0):
        t = p.data.clone()
        dist.broadcast(t, 0)
        ok &= bool(torch.equal(t, p.data))
    flag = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    assert flag.item() == 1.0


@pytest.mark.parametrize("name,C", [("CoR2", 2000), ("ODA", 3000)])
def test_gloo_world_size_2(name, C):
    mp.spawn(_worker, args=(2, _free_port(), name, C), nprocs=2, join=True)


def test_completion_order_matches_the_library_groups():
    """parallel.COMPLETION_ORDER (flat-buffer layout, bucket cuts) must be the grouping the backward plans signal
    through vqa_model_bwd_params.group_events (vqa_grad_groups; pure host call, no GPU)."""
    import ctypes as C
    from vqa_playground_pytorch_b200 import _lib
    from vqa_playground_pytorch_b200.parallel import COMPLETION_ORDER
    L = _lib.lib()
    for model_id, name, n in ((0, "CoR2", 62), (1, "ODA", 38)):
        tab = (C.c_int * n)()
        groups = L.vqa_grad_groups(model_id, tab, n)
        assert groups == len(COMPLETION_ORDER[name])
        for g, members in enumerate(COMPLETION_ORDER[name]):
            assert all(tab[i] == g for i in members), (name, g)
    assert L.vqa_grad_groups(0, (C.c_int * 10)(), 10) == -1


def test_vector_ranges_tile_the_padded_buffer_in_16_byte_units():
    """The peer transport moves 16-byte vectors: bucket boundaries move DOWN to a multiple of 4 floats (the <= 3 elements
    that change sides belong to a parameter that is complete by then and simply travel with the later bucket), the end
    moves up into the padding; the ranges still tile the buffer, in order."""
    from vqa_playground_pytorch_b200.parallel import GradSink
    for name, C in (("CoR2", 2000), ("ODA", 3000)):
        m = FakeModel(name, C)
        sink = GradSink(m.core_parameters(), name)
        total = sink.flat.numel()
        pos = 0
        for (lo, hi), (lo4, hi4) in zip(sink.bucket_ranges, sink.vector_ranges):
            assert lo4 == pos and lo4 % 4 == 0 and hi4 % 4 == 0 and hi4 > lo4
            assert lo - 3 <= lo4 <= lo and (hi - 3 <= hi4 <= hi or hi == total)
            pos = hi4
        assert total <= pos <= total + 3
        # the late groups (TAIL_GROUPS) are alone in the last bucket
        from vqa_playground_pytorch_b200.parallel import COMPLETION_ORDER, TAIL_GROUPS
        g0, g1 = sink.bucket_groups[-1]
        assert g1 == len(COMPLETION_ORDER[name]) and g1 - g0 == TAIL_GROUPS[name]
