"""Per-tensor parity of the SkipThoughts encoder against its oracle at a few small shapes (GPU box)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import parity
from oracle import skipthoughts as st
from vqa_playground_pytorch_b200 import blocks
for (V, B, T, seed, gs) in ((24, 3, 5, 5, 2), (40, 5, 7, 17, 3)):
    g = torch.Generator().manual_seed(gs)
    idx = torch.randint(1, V, (B, T), generator=g); idx[1, 3:] = 0
    dx = torch.randn(B, 2400, generator=g)
    sd = st.synth_state_dict(V, seed=10)
    enc = blocks.SkipThoughts(["w%d" % i for i in range(V)], af="relu").to("cuda:0"); enc.load_state_dict(sd); enc.train(True); enc.fixed_seed = seed
    x = enc(idx.to("cuda:0")); x.backward(dx.to("cuda:0"))
    ref = st.step(sd, idx, dx, "relu", st.seq_masks(seed, B, 620, 2400, 0.25))
    gmax = max(t.abs().max().item() for t in ref["grads"].values())
    errs = {k.split("gru_cell.")[-1]: parity.rel_err(p.grad.detach().cpu(), ref["grads"][k], 1e-6 * gmax) for k, p in enc.named_parameters()}
    print("B=%d T=%d: x err %.2e; grads %s" % (B, T, parity.rel_err(x.detach().cpu(), ref["x"]), {k: "%.1e" % v for k, v in errs.items()}), flush=True)
