"""Timing of the small-M (one row per sample) grouped linears through the C ABI (CUDA events). GPU box only.
Each shape is timed with a 256 MB L2 flush between calls so that weights come from HBM as in the real step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vqa_playground_pytorch_b200 import ops
from vqa_playground_pytorch_b200._lib import ACT_RELU

flush = torch.empty(64 << 20, device="cuda")

def t(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3

math = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
shapes = [("q_proj4", 4, 256, 2400, 310), ("gates", 2, 256, 310, 2048), ("classif", 1, 256, 510, 2000),
          ("glimpse", 4, 256, 2048, 155), ("fusion_final.h", 2, 256, 310, 510), ("compress", 1, 9216, 2048, 310)]
for name, g, M, K, N in shapes:
    xs = [torch.relu(torch.randn(M, K, device="cuda")) for _ in range(g)]
    ws = [torch.randn(N, K, device="cuda") / K ** 0.5 for _ in range(g)]
    bs = [torch.zeros(N, device="cuda") for _ in range(g)]
    dys = [torch.randn(M, N, device="cuda") for _ in range(g)]
    ys = ops.linear_forward(xs, ws, bs, ACT_RELU, 0.0, 1, [1] * g, math)
    f = t(lambda: ops.linear_forward(xs, ws, bs, ACT_RELU, 0.0, 1, [1] * g, math, outs=ys))
    wg = t(lambda: ops.linear_backward(xs, ws, ys, dys, ACT_RELU, 0.0, 1, [1] * g, False, math))
    bw = t(lambda: ops.linear_backward(xs, ws, ys, dys, ACT_RELU, 0.0, 1, [1] * g, True, math))
    wbytes = g * N * K * 4 / 1e6
    print("%-16s g=%d M=%d K=%d N=%d  weights %.1f MB | fwd %.1f us  wgrad(+dz) %.1f us  wgrad+dgrad %.1f us" %
          (name, g, M, K, N, wbytes, f, wg, bw), flush=True)
