import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
from oracle import reasoning_core as rc
from vqa_playground_pytorch_b200.config import CoR2
from vqa_playground_pytorch_b200.engine import GraphedStep
from vqa_playground_pytorch_b200.optim import FusedClipAdam
from vqa_playground_pytorch_b200.parallel import DataParallelEngine
C, Bl, N = 2000, 8, 36
B = Bl * world
sd = rc.synth_state_dict("CoR2", C, seed=3)
v, q, a = (t.to(dev) for t in rc.synth_inputs(B, N, C, seed=9))
sl = slice(rank * Bl, (rank + 1) * Bl)
shard = {"v": v[sl].contiguous(), "q_idxes": q[sl].contiguous(), "a": a[sl].contiguous()}
def cmp(name, t):
    g = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(g, t.contiguous())
    d = (g[0] - g[1]).abs()
    if rank == 0:
        nz = (d > 0).nonzero().flatten()
        print(name, "max diff %.3e, #diff %d of %d, first idx %s last idx %s" % (d.max().item(), nz.numel(), d.numel(), nz[:3].tolist(), nz[-3:].tolist()), flush=True)
for mode in sys.argv[1:]:
    cap = "cap" in mode; use_opt = "opt" in mode; train = "train" in mode
    m = CoR2.Model(None, C, precision=CoR2.precision); m.load_state_dict(sd); m = m.to(dev).train(train)
    eng = DataParallelEngine(m); eng.broadcast_parameters()
    opt = FusedClipAdam(eng, lr=1e-3, clip_grad=0.25, device_clock=True, lr_gamma=0.5 ** (1 / 50000)) if use_opt else None
    step = GraphedStep(m, shard, eng, warmup=2, capture_collectives=cap, optimizer=opt)
    if rank == 0: print("== mode", mode, "transport", eng.transport, "vector ranges", eng.vector_ranges, "total", eng.flat.numel(), flush=True)
    for it in range(3):
        step(shard); torch.cuda.synchronize()
        cmp(" step %d grads" % it, eng.flat)
        cmp(" step %d params" % it, torch.cat([p.detach().reshape(-1) for p in m.core_parameters()]))
    print(rank, "peer_error", eng.peer_error(), flush=True)
dist.barrier(); os._exit(0)
