"""Import the UNMODIFIED reference models read-only from /root/reference (authoring container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box; callers must check
`available()` and skip.  Recipe: SURVEY.md §8c.
  * empty stubs for third-party modules that utils.py imports at module top but the hot path
    never touches (/root/reference/utils.py:20-38);
  * `SkipThoughts` replaced by a pass-through BEFORE the model is built, because its __init__
    downloads files (/root/reference/putils/__init__.py:902-911); `sample['q_idxes']` then
    carries the 2400-d float embedding directly;
  * train-mode parity: `cf.F.dropout` is swapped for the Philox mask of oracle/philox.py, with
    layer ids handed out in call order (oracle/reasoning_core.py: ODA_LAYERS / COR2_LAYERS).
"""
import contextlib
import importlib
import io
import os
import sys
import types

import torch

from . import philox

REF_ROOT = "/root/reference"
_STUBS = ["deepdish", "deepdish.io", "h5py", "nltk", "nltk.tokenize", "nltk.parse", "nltk.parse.stanford",
          "yagmail", "munch", "configobj", "passlib", "passlib.hash", "redis", "lda", "tables"]


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "config", "CoR2.py"))


class PassThrough(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, q):
        return q


_cache = {}


def load_config(name):
    """name in {'ODA','CoR2'} -> the reference's config module (imported once)."""
    if name in _cache:
        return _cache[name]
    if not available():
        raise RuntimeError("reference sources not present at %s" % REF_ROOT)
    sys.dont_write_bytecode = True
    for m in _STUBS:
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["configobj"].ConfigObj = object
    sys.modules["passlib.hash"].sha512_crypt = object
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        putils = importlib.import_module("putils")
        putils.SkipThoughts = PassThrough
        cf = importlib.import_module("config." + name)
    cf.SkipThoughts = PassThrough
    _cache[name] = cf
    return cf


def build_model(name, num_ans, state_dict=None):
    cf = load_config(name)
    model = cf.Model(["PAD", "UNK"], num_ans)
    if state_dict is not None:
        missing = model.load_state_dict(state_dict, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
    return cf, model


@contextlib.contextmanager
def philox_dropout(cf, seed):
    """Replace the reference's F.dropout by the deterministic Philox mask, ids in call order."""
    real = cf.F.dropout
    counter = {"n": 0}

    def fake(x, p=0.5, training=True, inplace=False):
        layer = counter["n"]
        counter["n"] += 1
        if not training:
            return x
        m = torch.from_numpy(philox.dropout_mask(seed, layer, tuple(x.shape), p)).to(x.dtype)
        return x * m * (1.0 / (1.0 - p))

    cf.F.dropout = fake
    try:
        yield counter
    finally:
        cf.F.dropout = real


def reference_step(name, state_dict, v, q, a, train_seed=None, want_input_grads=False):
    """fwd + KLD loss + bwd through the reference's own Model (train.py:63-78, :536-544).
    train_seed=None -> eval mode (no dropout); else train mode with the Philox mask."""
    num_ans = a.shape[1]
    cf, model = build_model(name, num_ans, state_dict)
    model.train(train_seed is not None)
    v = v.detach().clone().requires_grad_(want_input_grads)
    q = q.detach().clone().requires_grad_(want_input_grads)
    ctx = philox_dropout(cf, train_seed) if train_seed is not None else contextlib.nullcontext()
    with ctx:
        logits = model({"v": v, "q_idxes": q})
    loss = torch.nn.KLDivLoss(reduction="sum")(torch.nn.functional.log_softmax(logits, dim=1), a)
    loss.backward()
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}
    out = {"logits": logits.detach(), "loss": loss.detach(), "alpha_dict": model.alpha_dict, "grads": grads}
    if want_input_grads:
        out["dv"], out["dq"] = v.grad, q.grad
    return out
