"""cuBLAS yardstick for the dominant GEMM shape (calibration only; never on the product path)."""
import torch
M, K, N = 9216, 2048, 310
x = torch.relu(torch.randn(M, K, device="cuda"))
w = torch.randn(N, K, device="cuda") / K ** 0.5
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
torch.backends.cuda.matmul.allow_tf32 = False
print("cuBLAS fp32      %.1f us" % t(lambda: x @ w.t()))
torch.backends.cuda.matmul.allow_tf32 = True
print("cuBLAS tf32      %.1f us" % t(lambda: x @ w.t()))
xb, wb = x.bfloat16(), w.bfloat16()
print("cuBLAS bf16      %.1f us" % t(lambda: xb @ wb.t()))
wp = torch.randn(320, K, device="cuda")
print("cuBLAS tf32 N=320 %.1f us" % t(lambda: x @ wp.t()))
dz = torch.randn(M, N, device="cuda")
print("cuBLAS tf32 wgrad %.1f us" % t(lambda: dz.t() @ x))
print("cuBLAS tf32 dgrad %.1f us" % t(lambda: dz @ w))
y = torch.empty(M, K, device="cuda")
print("copy 75MB         %.1f us" % t(lambda: y.copy_(x)))
