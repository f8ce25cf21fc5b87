"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import os

import numpy as np
import torch

from oracle import reasoning_core as rc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["oda_eval_b4", "oda_train_b4", "cor2_eval_b4", "cor2_train_b4", "cor2_eval_b2_small_ans"]

# Parity metric of SURVEY.md §8d: max|new-ref| / max(max|ref|, floor), floor = 1e-6 * largest gradient max-abs,
# so structurally-zero gradients (conv_att biases, fusion_vq biases) compare as ~0 instead of noise/noise.
FP32_TOL = 1e-4
# precision names that claim the fp32-parity bound (1e-4): CUDA-core fp32 and the error-compensated tensor-core modes
PARITY_MODES = ["fp32", "tf32x3", "bf16x3"]


def rel_err(new, ref, floor=0.0):
    new = torch.as_tensor(new).detach().double().cpu()
    ref = torch.as_tensor(ref).detach().double().cpu()
    assert new.shape == ref.shape, (new.shape, ref.shape)
    denom = max(ref.abs().max().item(), floor, 1e-30)
    return (new - ref).abs().max().item() / denom


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = {k[5:]: z[k].item() for k in z.files if k.startswith("meta.")}
    return z, meta


def golden_grad_names(z):
    return sorted({k[5:].rsplit(".", 1)[0] for k in z.files if k.startswith("grad.")})


def golden_grad_floor(z):
    return 1e-6 * max(float(z[f"grad.{n}.absmax"]) for n in golden_grad_names(z) if not n.startswith("__"))


def compare_grad_to_golden(z, name, g, floor):
    """g: full gradient tensor from the implementation under test. Returns the §8d error."""
    g = g.detach().float().cpu()
    absmax = float(z[f"grad.{name}.absmax"])
    denom = max(absmax, floor, 1e-30)
    if f"grad.{name}.full" in z.files:
        ref = torch.from_numpy(z[f"grad.{name}.full"])
        return (g.reshape(ref.shape) - ref).abs().max().item() / denom
    stride = int(z[f"grad.{name}.stride"])
    ref = torch.from_numpy(z[f"grad.{name}.sample"])
    got = g.reshape(-1)[::stride][:ref.numel()]
    e_samp = (got - ref).abs().max().item() / denom
    e_norm = abs(g.double().norm().item() - float(z[f"grad.{name}.l2"])) / max(float(z[f"grad.{name}.l2"]),
                                                                                  floor * np.sqrt(g.numel()), 1e-30)
    return max(e_samp, e_norm)


def flatten_alpha(alpha_dict):
    out = {}
    for k, val in alpha_dict.items():
        if isinstance(val, (tuple, list)):
            out[k] = torch.cat([t.detach() for t in val], dim=2)
        else:
            out[k] = val.detach()
    return out


def oracle_case(model, B, num_ans, N=36, train_seed=None, weight_seed=10, input_seed=1234, want_input_grads=False,
                gain=1.0, run=True, ties=None):
    sd = rc.synth_state_dict(model, num_ans, seed=weight_seed, num_regions=N, gain=gain)
    v, q, a = rc.synth_inputs(B, N, num_ans, seed=input_seed)
    if not run:
        return sd, (v, q, a), None
    drop = rc.no_drop if train_seed is None else rc.PhiloxDrop(train_seed)
    ref = rc.step(model, sd, v, q, a, drop=drop, num_regions=N, want_input_grads=want_input_grads, ties=ties)
    return sd, (v, q, a), ref


RELU_STASHES = {"CoR2": ["compress_v", "compress_v2", "compress_q", "compress_q_1", "compress_q_2", "linear_q",
                         "glimpses"],
                "ODA": ["compress_v", "compress_q", "linear_q", "glimpses"]}


def oracle_with_same_relu_pattern(model, sd, v, q, a, out, N=36, train_seed=None, tie_tol=1e-4):
    """Oracle fwd+bwd replaying the ReLU activation pattern the CUDA path produced (out['relu_masks']), after
    checking that the two patterns differ only at pre-activations within `tie_tol` of zero (relative to the
    layer's max |z|) — i.e. only at genuine rounding-level ties.  See oracle.reasoning_core.ReluTies."""
    ties = rc.ReluTies(masks=out["relu_masks"])
    drop = rc.no_drop if train_seed is None else rc.PhiloxDrop(train_seed)
    ref = rc.step(model, sd, v, q, a, drop=drop, num_regions=N, ties=ties)
    nflips = 0
    for name, mask in out["relu_masks"].items():
        z = ties.pre[name]
        if isinstance(z, dict):
            zz = torch.zeros(mask.shape)
            for (c0, c1), t in z.items():
                zz[:, c0:c1] = t
            z = zz
        z = z.reshape(mask.shape)
        flipped = (z > 0) != mask
        if flipped.any():
            worst = z[flipped].abs().max().item() / max(z.abs().max().item(), 1e-30)
            assert worst <= tie_tol, "ReLU pattern of %s differs at |z|/max|z| = %.3e (not a rounding tie)" % (name, worst)
            nflips += int(flipped.sum())
    ref["relu_ties"] = nflips
    return ref


def oracle_chunked(model, sd, v, q, a, chunk, N=36, train_seed=None, relu_masks=None, tie_tol=1e-4):
    """The oracle's fwd+bwd over a large batch in chunks of `chunk` samples: samples are independent through the
    forward and the loss is a SUM (train.py:541), so logits / alphas concatenate and gradients add.  Dropout masks
    are those of the full-batch tensors (PhiloxDrop.batch_offset).  With `relu_masks` (the activation patterns of the
    implementation under test, full batch) each chunk replays its rows, as oracle_with_same_relu_pattern does."""
    B = v.shape[0]
    logits, alphas, grads, loss, nflips = [], {}, None, 0.0, 0
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        ties = None
        if relu_masks is not None:
            sl = {}
            for name, m in relu_masks.items():
                per = m.shape[0] // B
                sl[name] = m[b0 * per:b1 * per]
            ties = rc.ReluTies(masks=sl)
        drop = rc.no_drop if train_seed is None else rc.PhiloxDrop(train_seed, batch_offset=b0)
        r = rc.step(model, sd, v[b0:b1], q[b0:b1], a[b0:b1], drop=drop, num_regions=N, ties=ties)
        if ties is not None:
            for name, mask in ties.masks.items():
                z = ties.pre[name]
                if isinstance(z, dict):
                    zz = torch.zeros(mask.shape)
                    for (c0, c1), t in z.items():
                        zz[:, c0:c1] = t
                    z = zz
                z = z.reshape(mask.shape)
                flipped = (z > 0) != mask
                if flipped.any():
                    worst = z[flipped].abs().max().item() / max(z.abs().max().item(), 1e-30)
                    assert worst <= tie_tol, "ReLU pattern of %s differs at |z|/max|z| = %.3e" % (name, worst)
                    nflips += int(flipped.sum())
        logits.append(r["logits"])
        loss += r["loss"].item()
        for k, val in flatten_alpha(r["alpha_dict"]).items():
            alphas.setdefault(k, []).append(val)
        if grads is None:
            grads = {k: g.clone() for k, g in r["grads"].items()}
        else:
            for k, g in r["grads"].items():
                grads[k] += g
    return {"logits": torch.cat(logits), "loss": torch.tensor(loss), "alpha_dict": {k: torch.cat(t) for k, t in alphas.items()},
            "grads": grads, "relu_ties": nflips}


def run_cuda_model(model, sd, v, q, a, N=36, train_seed=None, precision="tf32x3", device="cuda:0"):
    """fwd + KLD loss + bwd through the product's public API (config.<model>.Model)."""
    import importlib
    cf = importlib.import_module("vqa_playground_pytorch_b200.config." + model)
    from vqa_playground_pytorch_b200 import ops
    num_ans = a.shape[1]
    m = cf.Model(None, num_ans, num_regions=N, precision=precision)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    m = m.to(device)
    m.train(train_seed is not None)
    m.fixed_seed = train_seed
    sample = {"v": v.to(device), "q_idxes": q.to(device)}
    logits = m(sample)
    loss = ops.kld_loss(logits, a.to(device))
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().cpu() for n, p in m.named_parameters()}
    masks = {name: (ops.stash_tensor(name) > 0).cpu() for name in RELU_STASHES[model]}
    return {"relu_masks": masks, "logits": logits.detach().cpu(), "loss": loss.detach().cpu(), "alpha_dict":
            {k: (tuple(t.cpu() for t in val) if isinstance(val, tuple) else val.cpu()) for k, val in m.alpha_dict.items()},
            "grads": grads, "model": m}
