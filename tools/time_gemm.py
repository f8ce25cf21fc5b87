"""Timing of the dominant GEMM shapes through the C ABI (CUDA events). GPU box only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vqa_playground_pytorch_b200 import ops
from vqa_playground_pytorch_b200._lib import ACT_RELU

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

M, K, N = 9216, 2048, 310
x = torch.relu(torch.randn(M, K, device="cuda")); w = torch.randn(N, K, device="cuda") / K ** 0.5
b = torch.zeros(N, device="cuda"); dy = torch.randn(M, N, device="cuda")
tag = os.environ.get("VQA_TC_DEBUG", "0")
for math in sys.argv[1:] or ["tf32x3"]:
    for p in (0.0, 0.5):
        y = ops.linear_forward([x], [w], [b], ACT_RELU, p, 1, [1], math)
        f = t(lambda: ops.linear_forward([x], [w], [b], ACT_RELU, p, 1, [1], math, outs=y))
        wg = t(lambda: ops.linear_backward([x], [w], y, [dy], ACT_RELU, p, 1, [1], False, math))
        bw = t(lambda: ops.linear_backward([x], [w], y, [dy], ACT_RELU, p, 1, [1], True, math))
        print("debug=%s %s p=%.1f  fwd %.1f us  wgrad(+dz) %.1f us  wgrad+dgrad %.1f us" % (tag, math, p, f, wg, bw), flush=True)
        if p > 0:
            bits = [ops.dropout_bits(p, 1, 1, M * K, "cuda")]
            tb = t(lambda: ops.dropout_bits(p, 1, 1, M * K, "cuda"))
            y2 = ops.linear_forward([x], [w], [b], ACT_RELU, p, 1, [1], math, bits=bits)
            print("   bits path: max|y-y2| %.3e  bits kernel %.1f us" % ((y2[0] - y[0]).abs().max().item(), tb))
            f = t(lambda: ops.linear_forward([x], [w], [b], ACT_RELU, p, 1, [1], math, outs=y, bits=bits))
            wg = t(lambda: ops.linear_backward([x], [w], y, [dy], ACT_RELU, p, 1, [1], False, math, bits=bits))
            bw = t(lambda: ops.linear_backward([x], [w], y, [dy], ACT_RELU, p, 1, [1], True, math, bits=bits))
            print("debug=%s %s p=%.1f  fwd %.1f us  wgrad(+dz) %.1f us  wgrad+dgrad %.1f us  (keep-bits)" % (tag, math, p, f, wg, bw), flush=True)
