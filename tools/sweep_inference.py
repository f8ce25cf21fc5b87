"""Inference-only sweep over the number of regions (BASELINE.json configs[4]): eval-mode forward of both models through
the public API (config.<M>.Model), batch 256, 3xTF32 and single-pass TF32, device-resident inputs rotating over 4
batches, CUDA-event timing.  Writes a markdown table (stdout) — GPU box only."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    B, steps, warm = 256, 20, 5
    dev = torch.device("cuda", 0)
    rows = []
    for name, C in (("CoR2", 2000), ("ODA", 3000)):
        cf = importlib.import_module("vqa_playground_pytorch_b200.config." + name)
        for N in (10, 20, 36, 50, 75, 100):
            for prec in ("tf32x3", "tf32"):
                torch.manual_seed(10)
                model = cf.Model(None, C, num_regions=N, precision=prec).to(dev).eval()
                g = torch.Generator(device=dev).manual_seed(N)
                batches = [{"v": torch.randn(B, N, 2048, device=dev, generator=g).relu_(),
                            "q_idxes": 0.1 * torch.randn(B, 2400, device=dev, generator=g).relu_()} for _ in range(4)]
                with torch.no_grad():
                    for i in range(warm):
                        model(batches[i % 4])
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for i in range(steps):
                        out = model(batches[i % 4])
                    e1.record()
                    torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                assert torch.isfinite(out).all()
                rows.append((name, N, prec, ms, B / ms * 1e3))
                del model
    print("| model | regions | precision | ms / batch of %d | samples/s |" % B)
    print("|---|---|---|---|---|")
    for name, N, prec, ms, rate in rows:
        print("| %s | %d | %s | %.3f | %.0f |" % (name, N, prec, ms, rate))


if __name__ == "__main__":
    main()
