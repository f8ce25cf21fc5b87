"""One eager train step between cudaProfilerStart/Stop, for `ncu --profile-from-start off`.

    ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/step_metrics.csv \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        python tools/ncu_step.py --model CoR2

Numbers printed under ncu are never bench values; this only produces the per-launch list that profiles/ keeps.
"""
import argparse
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="CoR2")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--regions", type=int, default=36)
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--eval-mode", action="store_true")
    ap.add_argument("--steps", type=int, default=1)
    args = ap.parse_args()
    from vqa_playground_pytorch_b200 import ops
    from vqa_playground_pytorch_b200.parallel import DataParallelEngine
    cf = importlib.import_module("vqa_playground_pytorch_b200.config." + args.model)
    C = {"CoR2": 2000, "ODA": 3000}[args.model]
    dev = torch.device("cuda", 0)
    torch.manual_seed(10)
    model = cf.Model(None, C, num_regions=args.regions, precision=args.precision).to(dev)
    model.train(not args.eval_mode)
    ops.manual_seed(1234)
    DataParallelEngine(model)
    g = torch.Generator(device=dev).manual_seed(5)
    v = torch.randn(args.batch, args.regions, 2048, device=dev, generator=g).abs_()
    q = torch.randn(args.batch, 2400, device=dev, generator=g)
    a = torch.rand(args.batch, C, device=dev, generator=g)
    a = a / a.sum(1, keepdim=True)

    def step():
        out = model({"v": v, "q_idxes": q})
        if args.eval_mode:
            return out
        loss = ops.kld_loss(out, a)
        loss.backward()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
