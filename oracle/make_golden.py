"""Generate tests/golden/*.npz by running the UNMODIFIED reference (authoring container only).

    python -m oracle.make_golden            # from the repo root; needs /root/reference

Each fixture holds the reference's own outputs for one (model, mode) case on the committed
synthetic inputs (oracle/reasoning_core.py: synth_state_dict / synth_inputs, Philox-derived,
so they can be regenerated bit-for-bit anywhere without torch's RNG):
  logits, loss, every alpha_dict entry, and for every parameter gradient (plus dv, dq):
  max|g|, ||g||_2, the full tensor when it has <= 4096 elements, else 512 strided samples.
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

from . import reasoning_core as rc
from . import ref_import

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (fixture name, model, batch, num_ans, train_seed)
CASES = [
    ("oda_eval_b4", "ODA", 4, 3000, None),
    ("oda_train_b4", "ODA", 4, 3000, 77),
    ("cor2_eval_b4", "CoR2", 4, 2000, None),
    ("cor2_train_b4", "CoR2", 4, 2000, 77),
    ("cor2_eval_b2_small_ans", "CoR2", 2, 16, None),
]
WEIGHT_SEED = 10
INPUT_SEED = 1234
FULL_LIMIT = 4096
NSAMPLES = 512


def summarize(t):
    f = t.detach().reshape(-1).to(torch.float64)
    out = {"absmax": np.float64(f.abs().max().item()), "l2": np.float64(f.norm().item()),
           "numel": np.int64(f.numel())}
    if f.numel() <= FULL_LIMIT:
        out["full"] = t.detach().numpy().astype(np.float32)
    else:
        stride = f.numel() // NSAMPLES
        out["stride"] = np.int64(stride)
        out["sample"] = t.detach().reshape(-1)[::stride][:NSAMPLES].numpy().astype(np.float32)
    return out


def flatten_alpha(alpha_dict):
    out = {}
    for k, val in alpha_dict.items():
        if isinstance(val, (tuple, list)):
            out[k] = torch.cat([t.detach() for t in val], dim=2)      # G x [B,N,1] -> [B,N,G]
        else:
            out[k] = val.detach()
    return out


def make_case(name, model, B, num_ans, train_seed):
    sd = rc.synth_state_dict(model, num_ans, seed=WEIGHT_SEED)
    v, q, a = rc.synth_inputs(B, 36, num_ans, seed=INPUT_SEED)
    ref = ref_import.reference_step(model, sd, v, q, a, train_seed=train_seed, want_input_grads=True)
    blob = {"meta.model": model, "meta.B": B, "meta.num_ans": num_ans, "meta.num_regions": 36,
            "meta.train_seed": -1 if train_seed is None else train_seed,
            "meta.weight_seed": WEIGHT_SEED, "meta.input_seed": INPUT_SEED,
            "logits": ref["logits"].numpy(), "loss": np.float64(ref["loss"].item())}
    for k, t in flatten_alpha(ref["alpha_dict"]).items():
        blob["alpha." + k] = t.numpy()
    grads = dict(ref["grads"])
    grads["__dv"] = ref["dv"]
    grads["__dq"] = ref["dq"]
    for k, g in grads.items():
        if k.startswith("seq2vec"):
            continue
        for f, val in summarize(g).items():
            blob[f"grad.{k}.{f}"] = val
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **blob)
    return path


def main():
    if not ref_import.available():
        sys.exit("reference not present at /root/reference; fixtures can only be made in the authoring container")
    torch.manual_seed(0)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for case in CASES:
        p = make_case(*case)
        print("wrote", p, os.path.getsize(p), "bytes")


if __name__ == "__main__":
    main()
