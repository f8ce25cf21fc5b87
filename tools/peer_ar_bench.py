"""Latency / bandwidth of vqa_peer_allreduce_f32 by message size (torchrun, N GPUs):
    torchrun --nproc-per-node N tools/peer_ar_bench.py"""
import ctypes as C, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm
from vqa_playground_pytorch_b200 import _lib
L = _lib.lib()
n = 12 << 20
buf = symm.empty(n, dtype=torch.float32, device=dev); buf.normal_()
sig = symm.empty(int(L.vqa_peer_allreduce_signal_bytes()) // 4, dtype=torch.int32, device=dev); sig.zero_()
hb, hs = symm.rendezvous(buf, dist.group.WORLD), symm.rendezvous(sig, dist.group.WORLD)
torch.cuda.synchronize(); dist.barrier()
mc = int(hb.multicast_ptr or 0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(count, use_mc, ctas, reps=20):
    pr = _lib.PeerAllreduce()
    pr.world, pr.rank = world, rank
    for r in range(world):
        pr.buffers[r], pr.signals[r] = int(hb.buffer_ptrs[r]), int(hs.buffer_ptrs[r])
    pr.offset, pr.count, pr.max_ctas, pr.spin_limit_ms = 0, count, ctas, 20000
    pr.multicast = mc if use_mc else None
    for _ in range(3):
        _lib.check(L.vqa_peer_allreduce_f32(C.byref(pr), st))
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        _lib.check(L.vqa_peer_allreduce_f32(C.byref(pr), st))
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


ref = torch.empty(1 << 20, device=dev)
for count in (1024, 1 << 18, 1 << 20, 1 << 21, 12 << 20):
    line = "count %9d floats (%7.2f MB):" % (count, count * 4 / 1e6)
    for use_mc in ([True, False] if mc else [False]):
        for ctas in (32, 128, 160):
            buf.mul_(0.01)
            us = run(count, use_mc, ctas)
            line += "  %s/%d %.1f us" % ("nvls" if use_mc else "p2p", ctas, us)
    cnt = min(count, 1 << 20)
    ref[:cnt].copy_(buf[:cnt]); 
    if rank == 0:
        print(line, flush=True)
t = torch.randn(12 << 20, device=dev)
for count in (1024, 1 << 20, 12 << 20):
    for _ in range(3): dist.all_reduce(t[:count])
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): dist.all_reduce(t[:count])
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print("nccl count %9d: %.1f us" % (count, e0.elapsed_time(e1) / 20 * 1e3), flush=True)
if rank == 0: print("error flag", int(sig[-1].item()))
dist.barrier(); os._exit(0)
