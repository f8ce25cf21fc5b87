"""CPU restatement of the reference's ODA / CoR2 reasoning core (torch, CPU, fp32 or fp64).

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` leg may import this file; the product path
(`vqa-playground-pytorch_b200/`) never does and fails loudly without its CUDA library.

Parity pinning: the reference ships no tests or golden vectors, so this
restatement is pinned against the reference ITSELF: `oracle/make_golden.py` imports the
unmodified `/root/reference/config/{ODA,CoR2}.py` (stubs for absent third-party modules,
pass-through question encoder), runs it on the committed synthetic inputs and writes
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks this file against those vectors
and, when `/root/reference` is present, against the live reference as well.

The functions follow the reference's MATERIALISED form (the N x N pairwise tensor, the
N x N x D compound tensor); they are deliberately not the factorised algebra the CUDA
kernels use, so that agreement between the two is evidence and not a tautology.
Gradients come from torch autograd over these functions.

`drop(x, p, layer_id)` is a pluggable dropout: identity in eval mode, the shared Philox
mask (oracle/philox.py) in deterministic-train mode.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import philox

H = 310        # compressed dim   (config/CoR2.py:168, config/ODA.py:185)
F_DIM = 510    # Mutan hidden dim (config/CoR2.py:173)
G = 4          # glimpses         (config/CoR2.py:174)
A_DIM = 620    # attended dim     (config/CoR2.py:174)
Q_DIM = 2400   # SkipThoughts dim
D_DIM = 2048   # region feature dim

# Dropout call sites, numbered in the order the reference's forward() reaches them.
ODA_LAYERS = ["compress_v", "compress_q", "att.conv_att",
              "att.list_linear_v_fusion.0", "att.list_linear_v_fusion.1",
              "att.list_linear_v_fusion.2", "att.list_linear_v_fusion.3",
              "linear_q", "linear_classif"]
COR2_LAYERS = ["compress_q", "compress_v", "att1.conv_att",
               "att1.list_linear_v_fusion.0", "att1.list_linear_v_fusion.1",
               "att1.list_linear_v_fusion.2", "att1.list_linear_v_fusion.3",
               "compress_q_1", "expand_q_1", "compress_q_2", "expand_q_2",
               "compress_v2", "att2.conv_att",
               "att2.list_linear_v_fusion.0", "att2.list_linear_v_fusion.1",
               "att2.list_linear_v_fusion.2", "att2.list_linear_v_fusion.3",
               "linear_q", "linear_classif"]
ODA_LAYER_ID = {n: i for i, n in enumerate(ODA_LAYERS)}
COR2_LAYER_ID = {n: i for i, n in enumerate(COR2_LAYERS)}


# ----------------------------------------------------------------------------- dropout
def no_drop(x, p, layer_id):
    return x


def torch_drop(x, p, layer_id):
    """Stock F.dropout (torch's generator), as the reference runs it; used for CPU timing only."""
    return F.dropout(x, p=p, training=True)


class PhiloxDrop:
    """Deterministic train-mode dropout: x * keep / (1-p), keep from oracle/philox.py."""

    def __init__(self, seed, batch_offset=0):
        self.seed = int(seed)
        self.batch_offset = int(batch_offset)     # x holds samples [batch_offset, batch_offset + x.shape[0]) of the batch

    def __call__(self, x, p, layer_id):
        start = self.batch_offset * int(x[0].numel()) if self.batch_offset else 0
        m = philox.dropout_mask(self.seed, layer_id, tuple(x.shape), p, start=start)
        return x * torch.from_numpy(m).to(x.dtype) * (1.0 / (1.0 - p))


# ----------------------------------------------------------------------------- blocks
class ReluTies:
    """ReLU is not differentiable at 0, so two correct fp32 implementations whose pre-activations differ by a
    rounding error can disagree on relu'(z) where z ~ 0, and ONE such tie moves a weight-gradient row by
    ~1/sqrt(rows).  To compare gradients meaningfully the oracle can replay a given activation pattern
    (`masks[name]`, the pattern the implementation under test produced) and records its own pre-activations
    (`pre[name]`) so the test can assert that the patterns differ only where |z| is at rounding level."""

    def __init__(self, masks=None):
        self.masks = masks or {}
        self.pre = {}

    def relu(self, z, name, cols=None):
        if cols is None:
            self.pre[name] = z.detach()
        else:
            self.pre.setdefault(name, {})[cols] = z.detach()
        if name in self.masks:
            m = self.masks[name]
            if cols is not None:
                m = m[:, cols[0]:cols[1]]
            return z * m.reshape(z.shape).to(z.dtype)
        return torch.relu(z)


_NO_TIES = None


def _activate(x, af, ties, name, cols=None, dim=None):
    if af == "softmax":
        return F.softmax(x, dim=dim)
    if af == "relu" and ties is not None and name is not None:
        return ties.relu(x, name, cols)
    if af:
        return getattr(torch, af)(x)
    return x


def my_conv1d(x, w, b, p, af, drop, layer_id, dim=None, ties=None, name=None):
    """config/CoR2.py:72-88 == config/ODA.py:89-105. w is [Cout, Cin, 1]."""
    if x.dim() != 3:
        raise ValueError("input_dim (%s) should equal to 3" % x.dim())
    if p:
        x = drop(x, p, layer_id)
    x = F.conv1d(x.transpose(1, 2), w, b).transpose(1, 2)
    return _activate(x, af, ties, name, dim=dim)


def my_linear(x, w, b, p, af, drop, layer_id, ties=None, name=None, cols=None):
    """config/CoR2.py:106-119 == config/ODA.py:123-136."""
    if x.size(-1) != w.size(1):
        raise ValueError("last dimension of input(%s) should equal to in_features(%s)" % (x.size(-1), w.size(1)))
    if p:
        x = drop(x, p, layer_id)
    x = F.linear(x, w, b)
    return _activate(x, af, ties, name, cols)


def mutan_fusion(sd, prefix, x1, x2, R):
    """putils/__init__.py:232-238 with bmul (:98-104): per-sample broadcasting product."""
    total = 0
    for r in range(R):
        h1 = F.linear(x1, sd[f"{prefix}.list_linear1.{r}.linear.weight"], sd[f"{prefix}.list_linear1.{r}.linear.bias"])
        h2 = F.linear(x2, sd[f"{prefix}.list_linear2.{r}.linear.weight"], sd[f"{prefix}.list_linear2.{r}.linear.bias"])
        if h1.dim() == 3 and h2.dim() == 2:        # bmul: h1[b] * h2[b] broadcasts [N,F] * [F]
            h2 = h2.unsqueeze(1)
        total = total + h1 * h2
    return total


def my_att(sd, prefix, inputs, fuse, drop, layer_ids, ties=None, col0=0):
    """config/CoR2.py:137-154 == config/ODA.py:154-171. Returns (x_v [B,620], x_att [B,N,G], tmp [B,G,D])."""
    x_att = my_conv1d(fuse, sd[f"{prefix}.conv_att.conv.weight"], sd[f"{prefix}.conv_att.conv.bias"],
                      0.5, "softmax", drop, layer_ids[f"{prefix}.conv_att"], dim=1)
    tmp = torch.bmm(x_att.transpose(1, 2), inputs)                       # bmatmul, putils/__init__.py:89-95
    list_v = []
    for g in range(x_att.size(2)):
        name = f"{prefix}.list_linear_v_fusion.{g}"
        w_ = sd[f"{name}.linear.weight"]
        list_v.append(my_linear(tmp[:, g, :], w_, sd[f"{name}.linear.bias"], 0.5, "relu", drop, layer_ids[name],
                                ties, "glimpses", (col0 + g * w_.size(0), col0 + (g + 1) * w_.size(0))))
    return torch.cat(list_v, 1), x_att, tmp


# ----------------------------------------------------------------------------- models
def oda_forward(sd, v, q, drop=no_drop, num_regions=36, ties=None):
    """config/ODA.py:200-240 (q is the 2400-d question embedding, i.e. seq2vec's output).
    Returns (logits, alpha_dict) with alpha_dict as the reference sets it (:228-230)."""
    L = ODA_LAYER_ID
    N = num_regions
    v = v.contiguous().view(-1, N, D_DIM)
    b = v.size(0)
    vl = my_conv1d(v, sd["compress_v.conv.weight"], sd["compress_v.conv.bias"], 0.5, "relu", drop, L["compress_v"],
                   ties=ties, name="compress_v")
    ql = my_linear(q, sd["compress_q.linear.weight"], sd["compress_q.linear.bias"], 0.5, "relu", drop, L["compress_q"],
                   ties, "compress_q")
    # :216-222  vq[b,i,j*H+k] = (vl[b,i,k]-vl[b,j,k]) * ql[b,k]
    vq = ((vl.unsqueeze(2) - vl.unsqueeze(1)) * ql.view(b, 1, 1, -1)).reshape(b, N, N * vl.size(-1))
    v_final, x_att, _ = my_att(sd, "att", v, vq, drop, L, ties)
    alpha_dict = {"alphas": x_att[:, :, 0:1]}
    q_final = my_linear(q, sd["linear_q.linear.weight"], sd["linear_q.linear.bias"], 0.5, "relu", drop, L["linear_q"],
                        ties, "linear_q")
    x = mutan_fusion(sd, "fusion_final", v_final, q_final, 5)
    x = my_linear(x, sd["linear_classif.linear.weight"], sd["linear_classif.linear.bias"], 0.5, None, drop,
                  L["linear_classif"])
    return x, alpha_dict


def decare_cat(sd, block1, block2, guidance, drop, ties=None):
    """config/CoR2.py:191-199: [B,m,m,d] = block1[b,i,:]*g1[b,:] + block2[b,j,:]*g2[b,:]."""
    L = COR2_LAYER_ID
    b, m, d = block1.size()
    f1 = block1.view(-1, m, 1, d).expand(b, m, m, d)
    f2 = block2.view(-1, 1, m, d).expand(b, m, m, d)
    g1 = my_linear(my_linear(guidance, sd["compress_q_1.linear.weight"], sd["compress_q_1.linear.bias"], 0.5, "relu",
                             drop, L["compress_q_1"], ties, "compress_q_1"),
                   sd["expand_q_1.linear.weight"], sd["expand_q_1.linear.bias"], 0.5, "sigmoid", drop, L["expand_q_1"])
    g2 = my_linear(my_linear(guidance, sd["compress_q_2.linear.weight"], sd["compress_q_2.linear.bias"], 0.5, "relu",
                             drop, L["compress_q_2"], ties, "compress_q_2"),
                   sd["expand_q_2.linear.weight"], sd["expand_q_2.linear.bias"], 0.5, "sigmoid", drop, L["expand_q_2"])
    return f1 * g1.view(b, 1, 1, d) + f2 * g2.view(b, 1, 1, d)


def cor2_forward(sd, v, q, drop=no_drop, num_regions=36, ties=None):
    """config/CoR2.py:201-237. Returns (logits, alpha_dict)."""
    L = COR2_LAYER_ID
    N = num_regions
    v = v.contiguous().view(-1, N, D_DIM)
    b = v.size(0)
    ql = my_linear(q, sd["compress_q.linear.weight"], sd["compress_q.linear.bias"], 0.5, "relu", drop, L["compress_q"],
                   ties, "compress_q")
    vl = my_conv1d(v, sd["compress_v.conv.weight"], sd["compress_v.conv.bias"], 0.5, "relu", drop, L["compress_v"],
                   ties=ties, name="compress_v")
    v1_att, alpha1, _ = my_att(sd, "att1", v, mutan_fusion(sd, "fusion_vq1", vl, ql, 2), drop, L, ties, 0)
    v2_cat = decare_cat(sd, v, v, q, drop, ties)
    v2 = (alpha1[:, :, 0].contiguous().view(b, N, 1, 1) * v2_cat).sum(1)          # :216
    v2l = my_conv1d(v2, sd["compress_v2.conv.weight"], sd["compress_v2.conv.bias"], 0.5, "relu", drop, L["compress_v2"],
                    ties=ties, name="compress_v2")
    v2_att, alpha2, _ = my_att(sd, "att2", v2, mutan_fusion(sd, "fusion_vq2", v2l, ql, 2), drop, L, ties, A_DIM)
    alpha_dict = {"alpha1": torch.split(alpha1, 1, dim=2), "alpha2": torch.split(alpha2, 1, dim=2),
                  "feature": v2[:, [0, 1], :]}
    v_f = torch.cat([v1_att, v2_att], dim=1)
    q_final = my_linear(q, sd["linear_q.linear.weight"], sd["linear_q.linear.bias"], 0.5, "relu", drop, L["linear_q"],
                        ties, "linear_q")
    x = mutan_fusion(sd, "fusion_final", v_f, q_final, 2)
    x = my_linear(x, sd["linear_classif.linear.weight"], sd["linear_classif.linear.bias"], 0.5, None, drop,
                  L["linear_classif"])
    return x, alpha_dict


def cor_layers(steps):
    """Dropout call sites of a CoR chain with `steps` attention steps, in forward-call order (steps = 2: COR2_LAYERS)."""
    out = COR2_LAYERS[:17]
    for s in range(3, steps + 1):
        out += ["compress_q_%d" % (2 * s - 3), "expand_q_%d" % (2 * s - 3), "compress_q_%d" % (2 * s - 2),
                "expand_q_%d" % (2 * s - 2), "compress_v%d" % s, "att%d.conv_att" % s]
        out += ["att%d.list_linear_v_fusion.%d" % (s, g) for g in range(G)]
    return out + COR2_LAYERS[17:]


def cor_forward(sd, v, q, drop=no_drop, num_regions=36, steps=3):
    """The chain of reasoning with `steps` attention steps.  UNPINNED for steps > 2: the reference ships only the
    two-step config/CoR2.py (the 3-step model is known from the `alpha3` / `v3_feature` keys of
    visu.py:2491-2495 and CoR_Visulization.py:107-109).  Built from the reference's OWN blocks composed once more:
    every further step s calls decare_cat(previous objects, v, q) (config/CoR2.py:191-199) with its own gates, takes
    the alpha_{s-1}[0]-weighted sum over i (:215-216), compresses, fuses with the question and attends (:218-219)."""
    L = {n: i for i, n in enumerate(cor_layers(steps))}
    N = num_regions
    v = v.contiguous().view(-1, N, D_DIM)
    b = v.size(0)
    ql = my_linear(q, sd["compress_q.linear.weight"], sd["compress_q.linear.bias"], 0.5, "relu", drop, L["compress_q"])
    x, prev, feats, alphas, objects = v, v, [], [], []
    for s in range(1, steps + 1):
        if s > 1:
            k1, k2 = 2 * s - 3, 2 * s - 2
            f1 = prev.view(-1, N, 1, D_DIM).expand(b, N, N, D_DIM)             # block1: the previous step's objects
            f2 = v.view(-1, 1, N, D_DIM).expand(b, N, N, D_DIM)                # block2: v
            g = []
            for k in (k1, k2):
                h = my_linear(q, sd["compress_q_%d.linear.weight" % k], sd["compress_q_%d.linear.bias" % k], 0.5, "relu",
                              drop, L["compress_q_%d" % k])
                g.append(my_linear(h, sd["expand_q_%d.linear.weight" % k], sd["expand_q_%d.linear.bias" % k], 0.5,
                                   "sigmoid", drop, L["expand_q_%d" % k]))
            cat = f1 * g[0].view(b, 1, 1, D_DIM) + f2 * g[1].view(b, 1, 1, D_DIM)
            x = (alphas[-1][:, :, 0].contiguous().view(b, N, 1, 1) * cat).sum(1)
            objects.append(x)
        name = "compress_v" if s == 1 else "compress_v%d" % s
        xl = my_conv1d(x, sd[name + ".conv.weight"], sd[name + ".conv.bias"], 0.5, "relu", drop, L[name])
        att, alpha, _ = my_att(sd, "att%d" % s, x, mutan_fusion(sd, "fusion_vq%d" % s, xl, ql, 2), drop, L)
        feats.append(att)
        alphas.append(alpha)
        prev = x
    alpha_dict = {"alpha%d" % (s + 1): torch.split(a, 1, dim=2) for s, a in enumerate(alphas)}
    alpha_dict["feature"] = objects[0][:, [0, 1], :]
    for s, o in enumerate(objects[1:], start=3):
        alpha_dict["v%d_feature" % s] = o[:, [0, 1], :]
    q_final = my_linear(q, sd["linear_q.linear.weight"], sd["linear_q.linear.bias"], 0.5, "relu", drop, L["linear_q"])
    xf = mutan_fusion(sd, "fusion_final", torch.cat(feats, dim=1), q_final, 2)
    return my_linear(xf, sd["linear_classif.linear.weight"], sd["linear_classif.linear.bias"], 0.5, None, drop,
                     L["linear_classif"]), alpha_dict


def kld_loss(logits, target):
    """train.py:536-544: KLDivLoss(size_average=False)(log_softmax(x), a) = sum a*(log a - log p)."""
    return F.kl_div(F.log_softmax(logits, dim=1), target, reduction="sum")


# ----------------------------------------------------------------------------- synthetic data / weights
def param_shapes(model, num_ans, num_regions=36):
    """state_dict keys and shapes of the reference Model, excluding seq2vec.* (SURVEY.md §8b)."""
    N = num_regions

    def lin(name, o, i):
        return [(f"{name}.linear.weight", (o, i)), (f"{name}.linear.bias", (o,))]

    def conv(name, o, i):
        return [(f"{name}.conv.weight", (o, i, 1)), (f"{name}.conv.bias", (o,))]

    def att(name, fuse_dim):
        out = conv(f"{name}.conv_att", G, fuse_dim)
        for g in range(G):
            out += lin(f"{name}.list_linear_v_fusion.{g}", A_DIM // G, D_DIM)
        return out

    def mutan(name, d1, d2, R):
        out = []
        for r in range(R):
            out += lin(f"{name}.list_linear1.{r}", F_DIM, d1)
        for r in range(R):
            out += lin(f"{name}.list_linear2.{r}", F_DIM, d2)
        return out

    if model == "ODA":      # registration order of config/ODA.py:183-198
        s = conv("compress_v", H, D_DIM) + lin("compress_q", H, Q_DIM) + att("att", N * H)
        s += lin("linear_q", H, Q_DIM) + mutan("fusion_final", A_DIM, H, 5) + lin("linear_classif", num_ans, F_DIM)
    elif model == "CoR2":   # registration order of config/CoR2.py:166-189
        s = conv("compress_v", H, D_DIM) + conv("compress_v2", H, D_DIM) + lin("compress_q", H, Q_DIM)
        s += mutan("fusion_vq1", H, H, 2) + att("att1", F_DIM) + mutan("fusion_vq2", H, H, 2) + att("att2", F_DIM)
        s += lin("linear_q", H, Q_DIM) + mutan("fusion_final", 2 * A_DIM, H, 2) + lin("linear_classif", num_ans, F_DIM)
        s += lin("compress_q_1", H, Q_DIM) + lin("expand_q_1", D_DIM, H)
        s += lin("compress_q_2", H, Q_DIM) + lin("expand_q_2", D_DIM, H)
    else:
        raise ValueError(model)
    return s


def synth_state_dict(model, num_ans, seed=10, num_regions=36, dtype=torch.float32, gain=1.0):
    """Version-independent weights: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like nn.Linear/Conv1d default
    init, drawn from the Philox stream (layer = 1000 + tensor index) so fixtures do not depend on
    torch's generator."""
    sd = {}
    for t, (name, shape) in enumerate(param_shapes(model, num_ans, num_regions)):
        fan_in = shape[1] if len(shape) > 1 else dict(param_shapes(model, num_ans, num_regions))[
            name.replace("bias", "weight")][1]
        bound = gain / math.sqrt(fan_in)
        u = philox.uniform(seed, 1000 + t, shape)
        sd[name] = torch.from_numpy(((u * 2.0 - 1.0) * bound).astype(np.float32)).to(dtype)
    return sd


def synth_inputs(B, num_regions, num_ans, seed=1234, dtype=torch.float32):
    """SURVEY.md §8d synthetic batch: v = relu(n), q = 0.1*relu(n), soft target with mass (.6,.3,.1)."""
    v = np.maximum(philox.pseudo_normal(seed, 1, (B, num_regions, D_DIM)), 0.0)
    q = 0.1 * np.maximum(philox.pseudo_normal(seed, 2, (B, Q_DIM)), 0.0)
    a = np.zeros((B, num_ans), dtype=np.float32)
    cls = (philox.uniform(seed, 3, (B, 3)) * num_ans).astype(np.int64) % num_ans
    for k, mass in enumerate((0.6, 0.3, 0.1)):
        np.add.at(a, (np.arange(B), cls[:, k]), mass)
    return (torch.from_numpy(v.astype(np.float32)).to(dtype), torch.from_numpy(q.astype(np.float32)).to(dtype),
            torch.from_numpy(a).to(dtype))


FORWARD = {"ODA": oda_forward, "CoR2": cor2_forward}
LAYERS = {"ODA": ODA_LAYERS, "CoR2": COR2_LAYERS}


def step(model, sd, v, q, a, drop=no_drop, num_regions=36, want_input_grads=False, ties=None):
    """One fwd+bwd (train.py:63-78 without the optimizer). Returns dict(logits, loss, alpha_dict, grads)."""
    sd = {k: t.detach().clone().requires_grad_(True) for k, t in sd.items()}
    v = v.detach().clone().requires_grad_(want_input_grads)
    q = q.detach().clone().requires_grad_(want_input_grads)
    logits, alpha = FORWARD[model](sd, v, q, drop, num_regions, ties)
    loss = kld_loss(logits, a)
    loss.backward()
    grads = {k: (t.grad if t.grad is not None else torch.zeros_like(t)) for k, t in sd.items()}
    out = {"logits": logits.detach(), "loss": loss.detach(), "alpha_dict": alpha, "grads": grads}
    if want_input_grads:
        out["dv"], out["dq"] = v.grad, q.grad
    return out
