"""CPU restatement of the reference's question encoder — TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench CPU legs).

Follows /root/reference/putils/__init__.py: SkipThoughts.forward :975-982 (embedding -> BayesianGRU ->
hidden state at the last non-PAD token), BayesianGRU.forward :689-741 (return_last=True branch :727-738),
BayesianGRUCell.forward :622-637, SequentialDropout.forward :517-527 (one mask per sequence and call site,
`noise.bernoulli_(1-p).div_(1-p)`, shared by all time steps).  Pinned against the live reference classes in
tests/test_skipthoughts_oracle.py (the reference's dropout noise is replaced by the same Philox masks).
"""
import torch

from . import philox

INPUT_MASK_LAYER, HIDDEN_MASK_LAYER = 64, 67     # = ops.GRU_INPUT_MASK_LAYER / GRU_HIDDEN_MASK_LAYER
NAMES = ("ir", "ii", "in", "hr", "hi", "hn")


def param_shapes(vocab, I=620, H=2400):
    sh = {"embedding.weight": (vocab, I)}
    for g in ("ir", "ii", "in"):
        sh["gru.gru_cell.weight_%s.weight" % g] = (H, I)
        sh["gru.gru_cell.weight_%s.bias" % g] = (H,)
    for g in ("hr", "hi", "hn"):
        sh["gru.gru_cell.weight_%s.weight" % g] = (H, H)
    return sh


def synth_state_dict(vocab, seed=10, I=620, H=2400):
    """Philox-derived weights (reproducible anywhere), scaled like nn.Linear's default init; PAD row zero."""
    sd = {}
    for k, (name, shape) in enumerate(param_shapes(vocab, I, H).items()):
        fan_in = shape[1] if len(shape) == 2 else I
        w = torch.from_numpy(philox.uniform(seed, 200 + k, shape)).to(torch.float32) * 2.0 - 1.0
        sd[name] = w * (0.5 if name == "embedding.weight" else fan_in ** -0.5)
    sd["embedding.weight"][0].zero_()
    return sd


def seq_masks(seed, B, I, H, p):
    """The six sequence-tied multipliers (0 or 1/(1-p)), in the order of NAMES; None when p == 0."""
    if not p:
        return None
    out = []
    for m in range(3):
        out.append(torch.from_numpy(philox.dropout_mask(seed, INPUT_MASK_LAYER + m, (B, I), p)).float() / (1.0 - p))
    for m in range(3):
        out.append(torch.from_numpy(philox.dropout_mask(seed, HIDDEN_MASK_LAYER + m, (B, H), p)).float() / (1.0 - p))
    return out


def encode(sd, idx, af="relu", masks=None):
    """idx [B,T] int64 (0 = PAD) -> (x [B,H], all hidden states [B,T,H])."""
    B, T = idx.shape
    g = lambda n: sd["gru.gru_cell.weight_%s.weight" % n]
    b = lambda n: sd["gru.gru_cell.weight_%s.bias" % n]
    act = torch.relu if af == "relu" else torch.tanh
    emb = torch.nn.functional.embedding(idx, sd["embedding.weight"], padding_idx=0)
    m = masks if masks is not None else [1.0] * 6
    h = torch.zeros(B, g("hr").shape[0], dtype=emb.dtype)
    hs = []
    for t in range(T):
        x = emb[:, t, :]
        r = torch.sigmoid((x * m[0]) @ g("ir").t() + b("ir") + (h * m[3]) @ g("hr").t())
        i = torch.sigmoid((x * m[1]) @ g("ii").t() + b("ii") + (h * m[4]) @ g("hi").t())
        n = act((x * m[2]) @ g("in").t() + b("in") + r * ((h * m[5]) @ g("hn").t()))
        h = (1 - i) * n + i * h
        hs.append(h)
    hs = torch.stack(hs, 1)
    lengths = T - (idx == 0).sum(1)
    pos = (lengths - 1) % T                      # mask[i][lengths[i] - 1]: -1 wraps to the last position
    return hs[torch.arange(B), pos], hs


def step(sd, idx, dx, af="relu", masks=None):
    """forward + backward with the output gradient dx; returns x and every parameter gradient."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    x, hs = encode(leaves, idx, af, masks)
    x.backward(dx)
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return {"x": x.detach(), "hs": hs.detach(), "grads": grads}
