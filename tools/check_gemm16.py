"""Op-level check of the bf16-plane tcgen05 GEMM paths (math = bf16x3 / bf16, M >= 1024) against fp64 torch.
    python tools/check_gemm16.py [bf16x3|bf16]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import philox                                        # noqa: E402
from vqa_playground_pytorch_b200 import ops                      # noqa: E402
from vqa_playground_pytorch_b200._lib import ACT_RELU            # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def mask(seed, layer, shape):
    return torch.from_numpy(philox.dropout_mask(seed, layer, shape, 0.5)).cuda().double() * 2.0


def main():
    math = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
    tol = 1e-4 if math == "bf16x3" else 2e-2
    g = torch.Generator(device="cuda").manual_seed(0)
    worst = 0.0
    for (M, K, N, pdrop) in ((2304, 2048, 310, 0.5), (2304, 310, 2048, 0.5), (1100, 1240, 510, 0.0), (2304, 320, 155, 0.0)):
        x = torch.randn(M, K, device="cuda", generator=g).relu_()
        w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
        b = 0.1 * torch.randn(N, device="cuda", generator=g)
        dy = torch.randn(M, N, device="cuda", generator=g)
        seed, layer = 4242, 3
        y = ops.linear_forward([x], [w], [b], ACT_RELU, pdrop, seed, [layer], math)[0]
        dws, dbs, dxs = ops.linear_backward([x], [w], [y], [dy], ACT_RELU, pdrop, seed, [layer], True, math)
        torch.cuda.synchronize()
        xd = x.double() * (mask(seed, layer, (M, K)) if pdrop else 1.0)
        z = xd @ w.double().t() + b.double()
        dz = dy.double() * (y > 0)
        errs = {"y": rel(y, z.relu()), "dW": rel(dws[0], dz.t() @ xd), "db": rel(dbs[0], dz.sum(0)),
                "dX": rel(dxs[0], (dz @ w.double()) * (mask(seed, layer, (M, K)) if pdrop else 1.0))}
        print("linear %s M%d K%d N%d p=%.1f:" % (math, M, K, N, pdrop), " ".join("%s %.2e" % kv for kv in errs.items()), flush=True)
        worst = max(worst, *errs.values())
    # Mutan, x1 [B*36, 310] with x2 [B, 310] broadcast over regions
    B, R_, K1, K2, F, ranks = 64, 36, 310, 310, 510, 2
    x1 = torch.randn(B * R_, K1, device="cuda", generator=g, requires_grad=True)
    x2 = torch.randn(B, K2, device="cuda", generator=g, requires_grad=True)
    W1 = [(torch.randn(F, K1, device="cuda", generator=g) / K1 ** 0.5).requires_grad_() for _ in range(ranks)]
    W2 = [(torch.randn(F, K2, device="cuda", generator=g) / K2 ** 0.5).requires_grad_() for _ in range(ranks)]
    b1 = [(0.1 * torch.randn(F, device="cuda", generator=g)).requires_grad_() for _ in range(ranks)]
    b2 = [(0.1 * torch.randn(F, device="cuda", generator=g)).requires_grad_() for _ in range(ranks)]
    dy = torch.randn(B * R_, F, device="cuda", generator=g)
    leaves = [x1, x2] + W1 + b1 + W2 + b2
    d = lambda t: t.detach().double().requires_grad_()
    lv = [d(t) for t in leaves]
    X1, X2 = lv[0], lv[1]
    W1d, b1d, W2d, b2d = lv[2:2 + ranks], lv[2 + ranks:2 + 2 * ranks], lv[2 + 2 * ranks:2 + 3 * ranks], lv[2 + 3 * ranks:]
    ref = sum((X1 @ W1d[r].t() + b1d[r]) * (X2 @ W2d[r].t() + b2d[r]).repeat_interleave(R_, 0) for r in range(ranks))
    ref_grads = torch.autograd.grad(ref, lv, dy.double())
    wb = [t for r in range(ranks) for t in (W1[r], b1[r])] + [t for r in range(ranks) for t in (W2[r], b2[r])]
    y = ops.MutanFn.apply(x1, x2, math, ranks, *wb)
    grads = torch.autograd.grad(y, leaves, dy)
    torch.cuda.synchronize()
    errs = [rel(y, ref)] + [rel(a, b_) for a, b_ in zip(grads, ref_grads)]
    print("mutan %s:" % math, " ".join("%.2e" % e for e in errs), flush=True)
    worst = max(worst, *errs)
    print("WORST %.3e (tol %.0e) %s" % (worst, tol, "OK" if worst <= tol else "FAIL"))
    return 0 if worst <= tol else 1


if __name__ == "__main__":
    sys.exit(main())
