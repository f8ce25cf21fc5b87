"""SkipThoughts question encoder (blocks.SkipThoughts) fwd+bwd at batch 256 x 26 tokens, alone and in front of the CoR2
core (GPU box).  Eager launches (the encoder is not graph-captured yet); CUDA-event timing."""
import argparse, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256); ap.add_argument("--tokens", type=int, default=26)
    ap.add_argument("--vocab", type=int, default=14000); ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    from vqa_playground_pytorch_b200 import blocks, ops, _lib
    dev = torch.device("cuda", 0)
    B, T, V = args.batch, args.tokens, args.vocab
    torch.manual_seed(10)
    enc = blocks.SkipThoughts(["w%d" % i for i in range(V)], af="relu").to(dev).train()
    g = torch.Generator().manual_seed(1)
    idx = torch.randint(1, V, (B, T), generator=g)
    lens = torch.randint(4, T + 1, (B,), generator=g)
    for b in range(B):
        idx[b, lens[b]:] = 0
    idx = idx.to(dev)
    dx = torch.randn(B, 2400, generator=g).to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, n=args.steps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def enc_step():
        for p in enc.parameters():
            p.grad = None
        enc(idx).backward(dx)
    with torch.no_grad():
        ms_f = timed(lambda: enc(idx))
    ms = timed(enc_step)
    flop_f = 2.0 * B * T * (3 * 620 * 2400 + 3 * 2400 * 2400)
    print("encoder fwd %.3f ms, fwd+bwd %.3f ms/step at B=%d T=%d  (%.0f samples/s; fwd %.1f GFLOP -> %.1f TFLOP/s; "
          "fwd+bwd ~3x)" % (ms_f, ms, B, T, B / ms * 1e3, flop_f / 1e9, flop_f / ms_f / 1e9))
    cf = importlib.import_module("vqa_playground_pytorch_b200.config.CoR2")
    model = cf.Model(None, 2000, seq2vec=enc).to(dev).train()
    v = torch.randn(B, 36, 2048, generator=g).abs().to(dev)
    a = torch.softmax(torch.randn(B, 2000, generator=g), 1).to(dev)

    def full_step():
        for p in model.parameters():
            p.grad = None
        ops.kld_loss(model({"v": v, "q_idxes": idx}), a).backward()
    ms_full = timed(full_step)
    print("encoder + CoR2 core, eager fwd+loss+bwd: %.3f ms/step (%.0f samples/s)" % (ms_full, B / ms_full * 1e3))


if __name__ == "__main__":
    main()
