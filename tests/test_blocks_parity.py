"""The host-side mirror of the reference's building blocks (blocks.MyLinear / MyConv1d / MutanFusion / MyATT, the
classes a user composes by hand) and the standalone C-ABI entry points behind them (vqa_linear_*, vqa_mutan_*,
vqa_region_softmax_pool_*, vqa_cor_compound_*, vqa_oda_pair_attn_*), each forward AND backward against a plain torch
fp32 restatement of the reference lines cited, with the oracle's Philox masks in train mode.
Tolerance: 1e-4 relative to the tensor's max (the fp32-parity bound)."""
import pytest
import torch

from parity import rel_err

pytestmark = pytest.mark.gpu

MODES = ["fp32", "tf32x3"]


def _mask(seed, layer, shape, p=0.5):
    from oracle import philox
    return torch.from_numpy(philox.dropout_mask(seed, layer, shape, p)).cuda() / (1.0 - p)


def _grads(outs, leaves, douts):
    return torch.autograd.grad(outs, leaves, douts, allow_unused=True)


@pytest.mark.parametrize("math", MODES)
@pytest.mark.parametrize("train", [False, True])
def test_my_linear_and_conv1d_forward_backward(cuda, math, train):
    """config/CoR2.py:72-88 (MyConv1d k=1) and :106-119 (MyLinear): dropout on the INPUT, then linear, then act."""
    from vqa_playground_pytorch_b200 import blocks, ops
    torch.manual_seed(0)
    lin = blocks.MyLinear(310, 2048, p=0.5, af="sigmoid").cuda()          # expand_q_1's shape: K = 310 rows (not 16-byte)
    conv = blocks.MyConv1d(2048, 310, 1, 1, p=0.5, af="relu").cuda()
    for m, lid in ((lin, 3), (conv, 5)):
        m.math, m.layer_id = math, lid
        m.train(train)
    x1 = torch.randn(24, 310, device="cuda").relu_().requires_grad_()
    x2 = torch.randn(6, 36, 2048, device="cuda").relu_().requires_grad_()
    ops.manual_seed(77)
    s1 = (77 << 32) | 1
    s2 = (77 << 32) | 2
    y1 = lin(x1)
    y2 = conv(x2)
    d1, d2 = torch.randn_like(y1), torch.randn_like(y2)
    g = _grads([y1, y2], [x1, x2, lin.linear.weight, lin.linear.bias, conv.conv.weight, conv.conv.bias], [d1, d2])
    m1 = _mask(s1, 3, (24, 310)) if train else 1.0
    m2 = _mask(s2, 5, (6, 36, 2048)) if train else 1.0
    r1 = torch.sigmoid(torch.nn.functional.linear(x1 * m1, lin.linear.weight, lin.linear.bias))
    z2 = torch.nn.functional.conv1d((x2 * m2).transpose(1, 2), conv.conv.weight, conv.conv.bias).transpose(1, 2)
    r2 = z2 * (y2 > 0)                                  # same activation pattern (rounding-level ties)
    assert rel_err(y1, r1) < 1e-4 and rel_err(y2, z2.relu()) < 1e-4
    flips = ((z2 > 0) != (y2 > 0))
    assert not flips.any() or z2[flips].abs().max() <= 1e-4 * z2.abs().max()
    gr = _grads([r1, r2], [x1, x2, lin.linear.weight, lin.linear.bias, conv.conv.weight, conv.conv.bias], [d1, d2])
    for got, want in zip(g, gr):
        assert rel_err(got, want) < 1e-4
    with pytest.raises(ValueError):
        lin(torch.randn(4, 311, device="cuda"))
    with pytest.raises(ValueError):
        conv(torch.randn(4, 2048, device="cuda"))


@pytest.mark.parametrize("math", MODES)
def test_mutan_fusion_module(cuda, math):
    """putils/__init__.py:232-238 with the per-sample bmul broadcast (:98-104): x1 [B,N,310], x2 [B,310]."""
    from vqa_playground_pytorch_b200 import blocks
    torch.manual_seed(1)
    m = blocks.MutanFusion(310, 310, 510, 2).cuda()
    m.math = math
    x1 = torch.randn(5, 36, 310, device="cuda", requires_grad=True)
    x2 = torch.randn(5, 310, device="cuda", requires_grad=True)
    y = m(x1, x2)
    dy = torch.randn_like(y)
    leaves = [x1, x2] + list(m.parameters())
    g = _grads([y], leaves, [dy])
    ref = sum(m.list_linear1[r].linear(x1) * m.list_linear2[r].linear(x2).unsqueeze(1) for r in range(2))
    assert rel_err(y, ref) < 1e-4
    for got, want in zip(g, _grads([ref], leaves, [dy])):
        assert rel_err(got, want) < 1e-4
    with pytest.raises(ValueError):
        m(torch.randn(5, 36, 311, device="cuda"), x2)


@pytest.mark.parametrize("math", MODES)
@pytest.mark.parametrize("train", [False, True])
def test_my_att_module_with_downstream_use_of_alpha(cuda, math, train):
    """MyATT.forward (config/CoR2.py:137-154) composed by hand, with the attention weights USED downstream the way the
    reference uses alpha1[0] (config/CoR2.py:216): the gradient of alpha depends on the region index, so this checks
    the full [B,N,G] incoming-gradient path of vqa_region_softmax_pool_bwd (`dalpha_ext`)."""
    from vqa_playground_pytorch_b200 import blocks, ops
    torch.manual_seed(2)
    B, N = 6, 36
    att = blocks.MyATT(fuse_dim=510, glimpses=4, inputs_dim=2048, att_dim=620, af="relu").cuda()
    att.conv_att.layer_id = 2
    for g_, lin in enumerate(att.list_linear_v_fusion):
        lin.layer_id, lin.math = 3 + g_, math
    att.train(train)
    x = torch.randn(B, N, 2048, device="cuda").relu_().requires_grad_()
    fuse = torch.randn(B, N, 510, device="cuda", requires_grad=True)
    wdown = torch.randn(B, N, 4, device="cuda")
    ops.manual_seed(5)
    xv, alphas = att(x, fuse)
    alpha = torch.cat(alphas, 2)
    out = xv.sum() * 0.1 + (alpha * wdown).sum()
    leaves = [x, fuse] + list(att.parameters())
    g = torch.autograd.grad(out, leaves)
    seeds = [(5 << 32) | (i + 1) for i in range(5)]
    f = fuse * _mask(seeds[0], 2, (B, N, 510)) if train else fuse
    z = torch.nn.functional.conv1d(f.transpose(1, 2), att.conv_att.conv.weight, att.conv_att.conv.bias).transpose(1, 2)
    a_ref = torch.softmax(z, dim=1)
    tmp = torch.bmm(a_ref.transpose(1, 2), x)
    outs = []
    for g_, lin in enumerate(att.list_linear_v_fusion):
        t = tmp[:, g_, :] * _mask(seeds[1 + g_], 3 + g_, (B, 2048)) if train else tmp[:, g_, :]
        zz = lin.linear(t)
        outs.append(zz * (xv[:, g_ * 155:(g_ + 1) * 155] > 0))
    xv_ref = torch.cat(outs, 1)
    assert rel_err(alpha, a_ref) < 1e-4 and rel_err(xv, xv_ref) < 1e-4
    ref = xv_ref.sum() * 0.1 + (a_ref * wdown).sum()
    gr = torch.autograd.grad(ref, leaves)
    gmax = max(t.abs().max().item() for t in gr)
    for (name, _), got, want in zip([("x", 0), ("fuse", 0)] + list(att.named_parameters()), g, gr):
        if name == "conv_att.conv.bias":        # analytically zero (sum_i dz = 0)
            assert got.abs().max().item() <= 1e-5 * gmax
            continue
        assert rel_err(got, want, 1e-6 * gmax) < 1e-4, name


def test_cor_compound_standalone(cuda):
    """vqa_cor_compound_fwd/_bwd against the MATERIALISED reference form: decare_cat's [B,N,N,D] tensor and the
    alpha1[0]-weighted sum over i (config/CoR2.py:191-199, :215-216)."""
    from vqa_playground_pytorch_b200 import ops
    torch.manual_seed(3)
    B, N, D = 5, 36, 2048
    x = torch.randn(B, N, D, device="cuda").relu_()
    alpha = torch.softmax(torch.randn(B, N, 4, device="cuda"), 1).requires_grad_()
    g1 = torch.rand(B, D, device="cuda", requires_grad=True)
    g2 = torch.rand(B, D, device="cuda", requires_grad=True)
    pooled = torch.randn(B, 4, D, device="cuda")
    pooled[:, 0, :] = torch.einsum("bi,bid->bd", alpha[:, :, 0].detach(), x)
    pooled.requires_grad_()
    v2 = ops.CorCompoundFn.apply(x, pooled, alpha, g1, g2)
    dv2 = torch.randn_like(v2)
    g = torch.autograd.grad(v2, [pooled, alpha, g1, g2], dv2)
    # reference form: v2_cat[b,i,j,:] = x[b,i,:]*g1 + x[b,j,:]*g2; v2[b,j,:] = sum_i alpha[b,i,0] v2_cat[b,i,j,:]
    a0 = alpha[:, :, 0]
    vt = pooled[:, 0, :]                                   # = sum_i alpha_i x_i in the model; an input here
    ref = vt.unsqueeze(1) * g1.unsqueeze(1) + a0.sum(1).view(B, 1, 1) * x * g2.unsqueeze(1)
    cat = x.view(B, N, 1, D) * g1.view(B, 1, 1, D) + x.view(B, 1, N, D) * g2.view(B, 1, 1, D)
    mat = (a0.detach().view(B, N, 1, 1) * cat).sum(1)
    assert rel_err(v2, mat) < 1e-5 and rel_err(v2, ref) < 1e-5
    for got, want in zip(g, torch.autograd.grad(ref, [pooled, alpha, g1, g2], dv2)):
        assert rel_err(got, want) < 1e-4


@pytest.mark.parametrize("N", [10, 36])
@pytest.mark.parametrize("train", [False, True])
def test_oda_pair_attn_standalone(cuda, N, train):
    """vqa_oda_pair_attn_fwd/_bwd against the reference's materialised pair tensor + conv_att + softmax + bmatmul
    (config/ODA.py:216-226)."""
    from vqa_playground_pytorch_b200 import ops
    torch.manual_seed(4)
    B, H, D, seed, layer = 4, 310, 2048, 991, 2
    x = torch.randn(B, N, D, device="cuda").relu_()
    vl = torch.randn(B, N, H, device="cuda").relu_().requires_grad_()
    ql = torch.randn(B, H, device="cuda").relu_().requires_grad_()
    w = (torch.randn(4, N * H, 1, device="cuda") / (N * H) ** 0.5).requires_grad_()
    bc = torch.randn(4, device="cuda", requires_grad=True)
    pooled, alpha = ops.OdaPairAttnFn.apply(x, vl, ql, w, bc, 0.5 if train else 0.0, seed, layer)
    dp = torch.randn_like(pooled)
    g = torch.autograd.grad(pooled, [vl, ql, w, bc], dp)
    vq = ((vl.unsqueeze(2) - vl.unsqueeze(1)) * ql.view(B, 1, 1, H)).reshape(B, N, N * H)
    if train:
        vq = vq * _mask(seed, layer, (B, N, N * H))
    z = torch.nn.functional.conv1d(vq.transpose(1, 2), w, bc).transpose(1, 2)
    a_ref = torch.softmax(z, dim=1)
    p_ref = torch.bmm(a_ref.transpose(1, 2), x)
    assert rel_err(alpha, a_ref) < 1e-4 and rel_err(pooled, p_ref) < 1e-4
    gr = torch.autograd.grad(p_ref, [vl, ql, w, bc], dp)
    gmax = max(t.abs().max().item() for t in gr)
    for name, got, want in zip(("vl", "ql", "w", "bc"), g, gr):
        if name == "bc":
            assert got.abs().max().item() <= 1e-5 * gmax
            continue
        assert rel_err(got, want, 1e-6 * gmax) < 1e-4, name


@pytest.mark.parametrize("math", ["tf32x3", "tf32"])
def test_unsupported_tensor_core_requests_fail_loudly(cuda, math):
    """A tensor-core math mode is never silently replaced by the CUDA-core GEMM: an operand TMA cannot address (and that
    the op cannot repack) or an unbuilt mode is a ValueError with a message."""
    from vqa_playground_pytorch_b200 import ops
    from vqa_playground_pytorch_b200._lib import ACT_NONE
    x = torch.randn(64, 2048 + 1, device="cuda")[:, 1:]            # rows 4-byte aligned only, stride 2049
    w = torch.randn(32, 2048, device="cuda")
    b = torch.zeros(32, device="cuda")
    y = ops.linear_forward([x], [w], [b], ACT_NONE, 0.0, 0, [0], math)[0]       # repacked, still on tensor cores
    assert rel_err(y, x @ w.t()) < (1e-4 if math == "tf32x3" else 2e-2)
    with pytest.raises(ValueError, match="not built|tensor-core"):
        ops.linear_forward([x], [w], [b], ACT_NONE, 0.0, 0, [0], 77)
