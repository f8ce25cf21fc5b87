"""Where the end-to-end step time goes: H2D copy alone, compute alone, and the two overlapped (GPU box).
    python tools/e2e_probe.py [--model CoR2] [--batch 256]"""
import argparse, importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="CoR2"); ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--regions", type=int, default=36); ap.add_argument("--steps", type=int, default=40)
    args = ap.parse_args()
    from vqa_playground_pytorch_b200 import ops
    from vqa_playground_pytorch_b200.parallel import DataParallelEngine
    from vqa_playground_pytorch_b200.engine import GraphedStep, HostPrefetcher, pack_feature_shard
    cf = importlib.import_module("vqa_playground_pytorch_b200.config." + args.model)
    dev = torch.device("cuda", 0)
    C, B, N = bench.NUM_ANS[args.model], args.batch, args.regions
    torch.manual_seed(10)
    model = cf.Model(None, C, num_regions=N).to(dev).train()
    engine = DataParallelEngine(model)
    gen = torch.Generator().manual_seed(1)
    host = [bench.make_batch(B, N, C, "cpu", gen) for _ in range(4)]
    resident = [tuple(t.to(dev) for t in b) for b in host]
    shard = pack_feature_shard([{"v": b[0], "q_idxes": b[1], "a": b[2]} for b in host])
    g = GraphedStep(model, {"v": resident[0][0].clone(), "q_idxes": resident[0][1].clone(), "a": resident[0][2].clone()}, engine)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, n=args.steps):
        fn(3); torch.cuda.synchronize(); e0.record(); t0 = time.perf_counter(); fn(n); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n

    dv = torch.empty_like(shard[0]["v"], device=dev)
    def copy_only(n):
        for i in range(n):
            dv.copy_(shard[i % 4]["v"], non_blocking=True)
    def compute_only(n):
        for i in range(n):
            g({"v": resident[i % 4][0], "q_idxes": resident[i % 4][1], "a": resident[i % 4][2]})
    def compute_item(n):
        for i in range(n):
            g({"v": resident[i % 4][0], "q_idxes": resident[i % 4][1], "a": resident[i % 4][2]}).item()
    def e2e(n, item=True):
        pf = HostPrefetcher([shard[i % 4] for i in range(n)], dev, widen_into=g.static)
        for s in pf:
            l = g(s)
            if item:
                l.item()
    cs = torch.cuda.Stream()
    def copy_and_compute(n):
        for i in range(n):
            with torch.cuda.stream(cs):
                dv.copy_(shard[i % 4]["v"], non_blocking=True)
            g({"v": resident[i % 4][0], "q_idxes": resident[i % 4][1], "a": resident[i % 4][2]})
        torch.cuda.current_stream().wait_stream(cs)
    nb = shard[0]["v"].numel() * 2
    for name, fn in (("copy_only(v bf16)", copy_only), ("compute_only", compute_only), ("compute+item", compute_item),
                     ("copy||compute (no deps)", copy_and_compute), ("e2e prefetcher + item", e2e),
                     ("e2e prefetcher, no item", lambda n: e2e(n, False))):
        dev_ms, wall_ms = timed(fn)
        print("%-28s device %.3f ms/step  wall %.3f ms/step%s" % (name, dev_ms, wall_ms,
              "  (%.1f GB/s)" % (nb / dev_ms / 1e6) if name.startswith("copy_only") else ""), flush=True)


if __name__ == "__main__":
    main()
