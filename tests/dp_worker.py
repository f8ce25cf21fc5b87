"""Worker of tests/test_dp_gpu.py: one process per GPU (torchrun), NCCL.  Checks, on real devices, what the
reference's nn.DataParallel (train.py:517) guarantees and the one-process-per-GPU engine must keep:
  1. the all-reduced gradients of a batch sharded over the ranks equal the 1-GPU gradients at the GLOBAL batch
     (the loss is a sum, train.py:541, so gradients are summed, not averaged) — through the CUDA-graph step with
     the bucketed all-reduce captured inside it, overlapped with the backward;
  2. after K optimizer steps every rank holds bit-identical parameters.
Prints 'DP_OK' on rank 0 when everything holds."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import parity
    from oracle import reasoning_core as rc
    from vqa_playground_pytorch_b200 import ops
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

    def fresh(train):
        m = CoR2.Model(None, C, precision=CoR2.precision)
        m.load_state_dict(sd)
        return m.to(dev).train(train)

    # ---- 0. the transport itself: a bucket reduced by the engine's all-reduce equals the NCCL sum, and the peer
    # kernel leaves BIT-identical results on every rank
    m = fresh(False)
    eng = DataParallelEngine(m)
    want_transport = os.environ.get("VQA_ALLREDUCE", "auto")
    if want_transport in ("peer", "nccl"):
        assert eng.transport == want_transport, (eng.transport, want_transport)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for rep in range(3):
        eng.flat.copy_(torch.randn(eng.flat.shape, device=dev, generator=g))
        ref = eng.flat.clone()
        dist.all_reduce(ref)
        eng.reduce_all(overlapped=False)
        eng.wait()
        torch.cuda.synchronize()
        assert not eng.peer_error(), "a peer all-reduce timed out"
        err = (eng.flat - ref).abs().max().item() / ref.abs().max().item()
        assert err <= 1e-6, ("engine all-reduce differs from NCCL", err, eng.transport)
        if eng.transport == "peer":
            both = [torch.empty_like(eng.flat) for _ in range(world)]
            dist.all_gather(both, eng.flat)
            assert all(torch.equal(both[0], t) for t in both[1:]), "peer all-reduce results differ between ranks"
    transport = eng.transport

    # ---- 1. gradient equality, eval mode (dropout masks are indexed per rank, so train mode has no 1-GPU twin)
    m = fresh(False)
    eng = DataParallelEngine(m)
    step = GraphedStep(m, shard, eng, warmup=2, capture_collectives=True)
    step(shard)
    torch.cuda.synchronize()
    got = [p.grad.detach().clone() for p in m.core_parameters()]
    ref_m = fresh(False)
    ops.kld_loss(ref_m({"v": v, "q_idxes": q}), a).backward()
    want = [p.grad for p in ref_m.core_parameters()]
    names = [n for n, _ in ref_m.named_parameters()]
    gmax = max(t.abs().max().item() for t in want)
    worst = max((parity.rel_err(g, w, 1e-6 * gmax), n) for g, w, n in zip(got, want, names)
                if not n.endswith("conv_att.conv.bias"))
    assert worst[0] <= 1e-4, ("all-reduced gradients differ from the global-batch gradients", worst)

    # ---- 2. train mode, K graph-captured steps with the fused optimizer: ranks stay bit-identical
    m = fresh(True)
    eng = DataParallelEngine(m)
    eng.broadcast_parameters()
    opt = FusedClipAdam(eng, lr=1e-3, clip_grad=0.25, device_clock=True, lr_gamma=0.5 ** (1 / 50000))
    step = GraphedStep(m, shard, eng, warmup=2, capture_collectives=True, optimizer=opt)
    for _ in range(3):
        step(shard)
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().reshape(-1) for p in m.core_parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    for r in range(1, world):
        assert torch.equal(gathered[0], gathered[r]), "parameters diverged between rank 0 and rank %d" % r
    moved = (flat != torch.cat([sd[n].reshape(-1) for n in names]).to(dev)).float().mean().item()
    assert moved > 0.9, "the optimizer did not move the parameters (%.3f)" % moved
    # the masks differ between ranks (different Philox keys), the reduced gradients do not
    keys = [torch.empty_like(m.seed_device) for _ in range(world)]
    dist.all_gather(keys, m.seed_device)
    assert len({int(k.item()) for k in keys}) == world, "ranks drew the same dropout key"
    if rank == 0:
        print("DP_OK transport %s%s, worst gradient error %.2e (%s)" % ((transport, " (NVLS)" if eng.nvls else "") + worst),
              flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)          # skip NCCL teardown (collectives captured in a graph; see bench.py)


if __name__ == "__main__":
    main()
