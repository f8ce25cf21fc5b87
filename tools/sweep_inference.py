"""Inference-only attention sweep (BASELINE.json configs[4]): regions 10-100 x batch 1-4096 x region dim 1024/2048,
per KERNEL: achieved GB/s against the measured copy bandwidth for the bandwidth-bound attention kernels
(attention logits + region softmax, attention pooling, CoR compound objects, ODA's factorised eval attention) and
TFLOP/s against the measured dense bf16 rate for the region-compression GEMM that feeds them (bf16 and bf16x3).
Kernel times come from the C ABI's per-kernel CUDA-event records (vqa_profile_begin/end: 'k:<name> hbm=<bytes>' /
'k:<gemm> M.. N.. K..'), inputs rotate over enough buffers to exceed the 126 MB L2.  GPU box only:

    python tools/sweep_inference.py > profiles/r2_inference_sweep.md
"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vqa_playground_pytorch_b200 import _lib, ops                    # noqa: E402
from vqa_playground_pytorch_b200._lib import ACT_RELU                # noqa: E402

H, F = 310, 510


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d["hbm_gbs"], d["bf16_tflops"], "measured (MEASURED_PEAKS.json, burst: kernels timed alone)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def profile(fn, reps):
    L = _lib.lib()
    for _ in range(2):
        fn(0)
    torch.cuda.synchronize()
    L.vqa_profile_begin()
    for i in range(reps):
        fn(i)
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 16)
    L.vqa_profile_end(buf, len(buf))
    out = {}
    for item in buf.value.decode().split(";"):
        if item.startswith("k:"):
            name, rest = item.rsplit("=", 1)
            tot, cnt = rest.split("/")
            out[name[2:]] = float(tot) / int(cnt)
    return out


def main():
    hbm, tf, src = peaks()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    print("# Inference-only attention sweep (round 2)\n")
    print("Per-kernel CUDA-event times through the C ABI; HBM kernels: algorithmic GB/s and %% of %.0f GB/s; GEMM: "
          "algorithmic TFLOP/s (one pass) and %% of %.0f TFLOP/s dense bf16; peaks %s.\n" % (hbm, tf, src))
    print("| regions | batch | D | kernel | us | GB/s or TFLOP/s | % of peak |")
    print("|---|---|---|---|---|---|---|")
    for D in (1024, 2048):
        for N in (10, 36, 64, 100):
            for B in (1, 16, 256, 4096):
                nbuf = max(2, min(8, int(3e8 // max(1, B * N * D * 4)) + 1))
                xs = [torch.randn(B, N, D, device=dev, generator=g).relu_() for _ in range(nbuf)]
                fuse = torch.randn(B, N, F, device=dev, generator=g)
                wc = torch.randn(4, F, 1, device=dev, generator=g) / F ** 0.5
                bc = torch.zeros(4, device=dev)
                g1, g2 = torch.rand(B, D, device=dev, generator=g), torch.rand(B, D, device=dev, generator=g)
                vl = torch.randn(B, N, H, device=dev, generator=g).relu_()
                ql = torch.randn(B, H, device=dev, generator=g).relu_()
                wo = torch.randn(4, N * H, 1, device=dev, generator=g) / (N * H) ** 0.5
                w = torch.randn(H, D, device=dev, generator=g) / D ** 0.5
                b = torch.zeros(H, device=dev)
                reps = 6 if B >= 256 else 12

                def attention(i):
                    x = xs[i % nbuf]
                    with torch.no_grad():
                        pooled, alpha = ops.RegionSoftmaxPoolFn.apply(x, fuse, wc, bc, 0.0, 0, 0)
                        ops.CorCompoundFn.apply(x, pooled, alpha, g1, g2)
                        ops.OdaPairAttnFn.apply(x, vl, ql, wo, bc, 0.0, 0, 0)

                rec = profile(attention, reps)
                for prec in ("bf16", "bf16x3"):
                    def gemm(i, prec=prec):
                        ops.linear_forward([xs[i % nbuf].view(B * N, D)], [w], [b], ACT_RELU, 0.0, 0, [0], prec)
                    for k, v in profile(gemm, reps).items():
                        if " M" in k:
                            rec[k] = v
                for name, ms in rec.items():
                    if " hbm=" in name:
                        kn, work = name.rsplit(" hbm=", 1)
                        rate = float(work) / (ms * 1e-3) / 1e9
                        print("| %d | %d | %d | %s | %.1f | %.0f GB/s | %.1f |" % (N, B, D, kn, ms * 1e3, rate, 100 * rate / hbm))
                    elif " M" in name and name.split()[0].startswith(("tc16", "tc_")):
                        dims = {t[0]: int(t[1:]) for t in name.split()[1:]}
                        fl = 2.0 * dims["M"] * dims["N"] * dims["K"] * dims["g"]
                        rate = fl / (ms * 1e-3) / 1e12
                        print("| %d | %d | %d | %s | %.1f | %.1f TFLOP/s | %.1f |" % (N, B, D, name.split()[0], ms * 1e3, rate,
                                                                                     100 * rate / tf))
                del xs
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
