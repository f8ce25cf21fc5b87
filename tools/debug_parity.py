"""Verbose parity report (per tensor) of the CUDA path against the oracle. GPU box only.
    python tools/debug_parity.py [model] [B] [N] [train_seed|-1] [precision]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import parity  # noqa: E402


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "CoR2"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 36
    seed = int(sys.argv[4]) if len(sys.argv) > 4 else -1
    prec = sys.argv[5] if len(sys.argv) > 5 else "fp32"
    seed = None if seed < 0 else seed
    C = 2000 if model == "CoR2" else 3000
    sd, (v, q, a), _ = parity.oracle_case(model, B, C, N=N, run=False)
    out = parity.run_cuda_model(model, sd, v, q, a, N=N, train_seed=seed, precision=prec)
    ref = parity.oracle_with_same_relu_pattern(model, sd, v, q, a, out, N=N, train_seed=seed, tie_tol=1e-3)
    print("relu ties replayed:", ref["relu_ties"])
    print("== %s B=%d N=%d seed=%s precision=%s" % (model, B, N, seed, prec))
    print("loss ref %.6f new %.6f" % (ref["loss"].item(), out["loss"].item()))
    print("logits err %.3e" % parity.rel_err(out["logits"], ref["logits"]))
    fa, fb = parity.flatten_alpha(out["alpha_dict"]), parity.flatten_alpha(ref["alpha_dict"])
    for k in fb:
        print("alpha %-10s err %.3e" % (k, parity.rel_err(fa[k], fb[k])))
    gmax = max(g.abs().max().item() for g in ref["grads"].values())
    floor = 1e-6 * gmax
    worst = 0.0
    for k, g in ref["grads"].items():
        e = parity.rel_err(out["grads"][k], g, floor)
        worst = max(worst, e)
        flag = "  <<<<<" if e > 1e-4 else ""
        print("grad %-48s |ref|max %.3e err %.3e%s" % (k, g.abs().max().item(), e, flag))
    print("WORST grad err %.3e" % worst)


if __name__ == "__main__":
    main()
