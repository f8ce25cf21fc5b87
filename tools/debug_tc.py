"""Numerics report of the tcgen05 linear fwd / wgrad / dgrad against fp64 on the CPU. GPU box only.
    python tools/debug_tc.py [math ...]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import philox  # noqa: E402
from vqa_playground_pytorch_b200 import ops  # noqa: E402
from vqa_playground_pytorch_b200._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID  # noqa: E402


def err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def case(M, K, N, act, p, math, groups=1, seed=123, ldx=None, timing=False):
    g = torch.Generator().manual_seed(M * 7 + K * 3 + N)
    dev = "cuda"
    xs, ws, bs, dys = [], [], [], []
    for i in range(groups):
        x = torch.relu(torch.randn(M, K, generator=g))
        if ldx:
            xp = torch.zeros(M, ldx)
            xp[:, :K] = x
            xs.append(xp.to(dev)[:, :K])
        else:
            xs.append(x.to(dev))
        ws.append((torch.randn(N, K, generator=g) / K ** 0.5).to(dev))
        bs.append((torch.randn(N, generator=g) * 0.1).to(dev))
        dys.append(torch.randn(M, N, generator=g).to(dev))
    layers = list(range(5, 5 + groups))
    ys = ops.linear_forward(xs, ws, bs, act, p, seed, layers, math)
    dws, dbs, dxs = ops.linear_backward(xs, ws, ys, dys, act, p, seed, layers, True, math)
    torch.cuda.synchronize()
    out = []
    for i in range(groups):
        x = xs[i].double().cpu()
        if p > 0:
            m = torch.from_numpy(philox.dropout_mask(seed, layers[i], (M, K), p)).double() / (1 - p)
            x = x * m
        w, b, dy = ws[i].double().cpu(), bs[i].double().cpu(), dys[i].double().cpu()
        z = x @ w.t() + b
        y = {ACT_NONE: z, ACT_RELU: torch.relu(z), ACT_SIGMOID: torch.sigmoid(z)}[act]
        yg = ys[i].double().cpu()      # activation derivative from the GPU's own y (relu sign flips at |z|~ulp)
        dz = {ACT_NONE: dy, ACT_RELU: dy * (yg > 0), ACT_SIGMOID: dy * yg * (1 - yg)}[act]
        dw, db = dz.t() @ x, dz.sum(0)
        dx = dz @ w
        if p > 0:
            dx = dx * m
        out.append((err(ys[i], y), err(dws[i], dw), err(dbs[i], db), err(dxs[i], dx)))
    e = np.max(np.array(out), axis=0)
    msg = "M=%-6d K=%-5d N=%-5d act=%d p=%.1f g=%d math=%-7s  y %.2e  dW %.2e  db %.2e  dX %.2e" % (
        M, K, N, act, p, groups, math, e[0], e[1], e[2], e[3])
    if timing:
        for fn, name in ((lambda: ops.linear_forward(xs, ws, bs, act, p, seed, layers, math), "fwd"),
                         (lambda: ops.linear_backward(xs, ws, ys, dys, act, p, seed, layers, True, math), "bwd")):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                fn()
            torch.cuda.synchronize()
            msg += "  %s %.1f us" % (name, (time.perf_counter() - t0) / 10 * 1e6)
    print(msg, flush=True)


def main():
    maths = sys.argv[1:] or ["tf32x3", "tf32"]
    for math in maths:
        case(128, 32, 128, ACT_NONE, 0.0, math)
        case(128, 64, 160, ACT_NONE, 0.0, math)
        case(256, 256, 310, ACT_RELU, 0.0, math)
        case(200, 100, 70, ACT_SIGMOID, 0.0, math)
        case(256, 2400, 310, ACT_RELU, 0.5, math, groups=4)
        case(256, 310, 2048, ACT_SIGMOID, 0.5, math, groups=2, ldx=320)
        case(256, 510, 2000, ACT_NONE, 0.5, math, ldx=512)
        case(256, 2048, 155, ACT_RELU, 0.5, math, groups=4)
        case(9216, 2048, 310, ACT_RELU, 0.5, math, timing=True)
        case(9216, 2048, 310, ACT_RELU, 0.0, math, timing=True)
    case(9216, 2048, 310, ACT_RELU, 0.5, "fp32", timing=True)


if __name__ == "__main__":
    main()
