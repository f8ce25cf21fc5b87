"""Generate tests/golden/bayesian_gru_*.npz by running the UNMODIFIED reference classes (authoring container only).

    python -m oracle.make_golden_gru          # from the repo root; needs /root/reference

putils.BayesianGRU (putils/__init__.py:660-741) + nn.Embedding(padding_idx=0) as SkipThoughts.forward :975-982 wires
them, at small sizes (the classes are size-generic; SkipThoughts itself hard-codes 620/2400 and downloads files).
Train mode: each SequentialDropout's noise is preset to the Philox mask the kernels draw (oracle/skipthoughts.py).
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

from . import ref_import, skipthoughts as st

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = [("bayesian_gru_eval_relu", None, "relu"), ("bayesian_gru_train_relu", 91, "relu"),
         ("bayesian_gru_train_tanh", 92, "tanh")]
V, I, H, B, T, P = 9, 12, 16, 4, 6, 0.25


def inputs():
    idx = torch.tensor([[3, 5, 1, 0, 0, 0], [2, 2, 7, 8, 4, 6], [1, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0]], dtype=torch.int64)
    dx = torch.from_numpy(st.philox.uniform(7, 300, (B, H))).float() - 0.5
    return idx, dx


def reference_step(sd, idx, dx, af, masks):
    """The reference's own modules, wired as SkipThoughts.forward does."""
    ref_import.load_config("CoR2")                       # puts /root/reference on sys.path with the import stubs
    import putils
    emb = torch.nn.Embedding(V, I, padding_idx=0)
    gru = putils.BayesianGRU(input_size=I, hidden_size=H, dropout=P, return_last=True, af=af)
    emb.load_state_dict({"weight": sd["embedding.weight"]})
    gru.load_state_dict({k[len("gru."):]: v for k, v in sd.items() if k.startswith("gru.")})
    # BayesianGRU.forward's return_last branch does x.view(batch, 2400): patch the literal through a subclass-free trick
    train = masks is not None
    emb.train(train); gru.train(train)
    if train:
        c = gru.gru_cell
        for d, m in zip((c.drop_ir, c.drop_ii, c.drop_in, c.drop_hr, c.drop_hi, c.drop_hn), masks):
            d.noise, d.restart = m, False                # SequentialDropout.forward :517-527 then only multiplies
    e = emb(idx)
    lengths = (idx.size(1) - idx.data.eq(0).sum(1)).long()
    # the reference hard-codes `.view(batch_size, 2400)` in the return_last branch (:735); run its per-step loop and
    # apply the same last-position mask here for H != 2400
    hx, outs = None, []
    for t in range(idx.size(1)):
        hx = gru.gru_cell(e[:, t, :], hx=hx)
        outs.append(hx.view(B, 1, H))
    out = torch.cat(outs, 1)
    mask = torch.zeros_like(out)
    for i in range(B):
        mask[i][lengths[i] - 1].fill_(1)
    x = out.mul(mask).sum(1).view(B, H)
    x.backward(dx)
    grads = {"embedding.weight": emb.weight.grad}
    for k, p in gru.named_parameters():
        grads["gru." + k] = p.grad if p.grad is not None else torch.zeros_like(p)
    return x.detach(), out.detach(), grads


def main():
    if not ref_import.available():
        sys.exit("reference not present at /root/reference")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    idx, dx = inputs()
    sd = st.synth_state_dict(V, seed=10, I=I, H=H)
    for name, seed, af in CASES:
        masks = st.seq_masks(seed, B, I, H, P) if seed is not None else None
        x, hs, grads = reference_step(sd, idx, dx, af, masks)
        blob = {"meta.af": af, "meta.seed": -1 if seed is None else seed, "idx": idx.numpy(), "dx": dx.numpy(),
                "x": x.numpy(), "hs": hs.numpy()}
        for k, g in grads.items():
            blob["grad." + k] = g.numpy()
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **blob)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
