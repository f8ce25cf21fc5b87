"""Op-level parity of C-ABI features the model plans rely on, each against a plain torch fp32 restatement with the
oracle's Philox masks (oracle/philox.py).  Tolerance: 1e-4 relative to the tensor's max (fp32-parity bound)."""
import numpy as np
import pytest
import torch

from parity import rel_err

pytestmark = pytest.mark.gpu


def _mask(seed, layer, shape, p=0.5):
    from oracle import philox
    return torch.from_numpy(philox.dropout_mask(seed, layer, shape, p)).cuda() / (1.0 - p)


def test_keep_bits_batch_matches_the_oracle_masks(cuda):
    from oracle import philox
    from vqa_playground_pytorch_b200 import ops
    sites = [(3, 1000), (7, 16), (11, 36 * 310 + 5), (0, 4096)]
    outs = ops.dropout_bits_batch(0.5, 0xABCDEF12345, sites, "cuda")
    single = ops.dropout_bits(0.5, 0xABCDEF12345, 11, sites[2][1], "cuda")
    for (layer, n), out in zip(sites, outs):
        bits = np.unpackbits(out.cpu().numpy(), bitorder="little")[:n]
        want = philox.mask_bytes(0xABCDEF12345, layer, n) >= philox.threshold(0.5)
        assert np.array_equal(bits.astype(bool), want), layer
    assert torch.equal(single, outs[2])
    # a threshold that is not a power of two goes through the per-byte compare as well
    out = ops.dropout_bits_batch(0.3, 99, [(5, 333)], "cuda")[0]
    bits = np.unpackbits(out.cpu().numpy(), bitorder="little")[:333]
    assert np.array_equal(bits.astype(bool), philox.mask_bytes(99, 5, 333) >= philox.threshold(0.3))


@pytest.mark.parametrize("K", [310, 2048])
def test_grouped_linear_with_cached_masks(cuda, K):
    """Rows of K=310 floats are not a multiple of 4 long: a quad's mask bits run across byte boundaries."""
    from vqa_playground_pytorch_b200 import ops
    from vqa_playground_pytorch_b200._lib import ACT_RELU
    g = torch.Generator(device="cuda").manual_seed(K)
    M, N, G_, seed, layers = 256, 155, 2, 4242, [3, 4]
    xs = [torch.randn(M, K, device="cuda", generator=g).relu_() for _ in range(G_)]
    ws = [torch.randn(N, K, device="cuda", generator=g) / K ** 0.5 for _ in range(G_)]
    bs = [0.1 * torch.randn(N, device="cuda", generator=g) for _ in range(G_)]
    dys = [torch.randn(M, N, device="cuda", generator=g) for _ in range(G_)]
    bits = ops.dropout_bits_batch(0.5, seed, [(l, M * K) for l in layers], "cuda")
    ys = ops.linear_forward(xs, ws, bs, ACT_RELU, 0.5, seed, layers, "tf32x3", bits=bits)
    ys_philox = ops.linear_forward(xs, ws, bs, ACT_RELU, 0.5, seed, layers, "tf32x3")
    dws, dbs, dxs = ops.linear_backward(xs, ws, ys, dys, ACT_RELU, 0.5, seed, layers, True, "tf32x3", bits=bits)
    for i in range(G_):
        xd = xs[i] * _mask(seed, layers[i], (M, K))
        z = xd @ ws[i].t() + bs[i]
        assert rel_err(ys[i], z.relu()) < 1e-4
        assert rel_err(ys[i], ys_philox[i]) < 1e-5           # cached and regenerated masks are the same mask
        dz = dys[i] * (ys[i] > 0)
        assert rel_err(dws[i], dz.t() @ xd) < 1e-4
        assert rel_err(dbs[i], dz.sum(0)) < 1e-4
        assert rel_err(dxs[i], (dz @ ws[i]) * _mask(seed, layers[i], (M, K))) < 1e-4


@pytest.mark.parametrize("math", ["tf32x3", "fp32"])
def test_dgrad_with_pooling_addend(cuda, math):
    """dX = mask (.) (dZ W) + sum_g alpha[m,g] dpooled[b,g,:]  (compress_v2 + att2 pooling share their input)."""
    from vqa_playground_pytorch_b200 import ops
    from vqa_playground_pytorch_b200._lib import ACT_RELU
    g = torch.Generator(device="cuda").manual_seed(5)
    B, R, K, N, seed = 12, 36, 512, 310, 77
    M = B * R
    x = torch.randn(M, K, device="cuda", generator=g).relu_()
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.zeros(N, device="cuda")
    dy = torch.randn(M, N, device="cuda", generator=g)
    alpha = torch.rand(M, 4, device="cuda", generator=g)
    dpooled = torch.randn(B, 4, K, device="cuda", generator=g)
    bits = ops.dropout_bits_batch(0.5, seed, [(9, M * K)], "cuda") if math != "fp32" else None
    y = ops.linear_forward([x], [w], [b], ACT_RELU, 0.5, seed, [9], math, bits=bits)
    _, _, dxs = ops.linear_backward([x], [w], y, [dy], ACT_RELU, 0.5, seed, [9], True, math, bits=bits,
                                    pool=(alpha, dpooled, R))
    dz = dy * (y[0] > 0)
    want = (dz @ w) * _mask(seed, 9, (M, K)) + torch.einsum("brg,bgk->brk", alpha.view(B, R, 4), dpooled).reshape(M, K)
    assert rel_err(dxs[0], want) < 1e-4


def test_mutan_backward_fused_gradient_operands(cuda):
    """MutanFn (fused dH kernel + wgrad/dgrad GEMMs) against autograd of the restated fusion, rows_per = 36."""
    from vqa_playground_pytorch_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    B, R_, K1, K2, F, ranks = 8, 36, 310, 310, 510, 2
    x1 = torch.randn(B * R_, K1, device="cuda", generator=g, requires_grad=True)
    x2 = torch.randn(B, K2, device="cuda", generator=g, requires_grad=True)
    W1 = [(torch.randn(F, K1, device="cuda", generator=g) / K1 ** 0.5).requires_grad_() for _ in range(ranks)]
    W2 = [(torch.randn(F, K2, device="cuda", generator=g) / K2 ** 0.5).requires_grad_() for _ in range(ranks)]
    b1 = [(0.1 * torch.randn(F, device="cuda", generator=g)).requires_grad_() for _ in range(ranks)]
    b2 = [(0.1 * torch.randn(F, device="cuda", generator=g)).requires_grad_() for _ in range(ranks)]
    dy = torch.randn(B * R_, F, device="cuda", generator=g)
    leaves = [x1, x2] + W1 + b1 + W2 + b2

    ref = sum((x1 @ W1[r].t() + b1[r]) * (x2 @ W2[r].t() + b2[r]).repeat_interleave(R_, 0) for r in range(ranks))
    ref_grads = torch.autograd.grad(ref, leaves, dy)
    wb = [t for r in range(ranks) for t in (W1[r], b1[r])] + [t for r in range(ranks) for t in (W2[r], b2[r])]
    y = ops.MutanFn.apply(x1, x2, "tf32x3", ranks, *wb)
    assert rel_err(y, ref) < 1e-4
    grads = torch.autograd.grad(y, leaves, dy)
    for got, want in zip(grads, ref_grads):
        assert rel_err(got, want) < 1e-4


def test_eval_tail_argmax_and_answer_mapping(cuda):
    """ops.argmax_rows / engine.predict_answers against the reference's `output.data.cpu().max(1)` and its
    MultipleChoice scan (train.py:146-169), ties included (first maximum wins on both sides)."""
    from oracle import reasoning_core as rc
    from vqa_playground_pytorch_b200 import ops
    from vqa_playground_pytorch_b200.config import ODA
    from vqa_playground_pytorch_b200.engine import predict_answers
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(37, 3000, device="cuda", generator=g)
    x[5, 100] = x[5, 2000] = 9.0                      # a tie: index 100 wins
    x[6] = 0.0                                        # all equal: index 0
    pred, best = ops.argmax_rows(x, return_best=True)
    want_v, want_i = x.cpu().max(1)
    assert torch.equal(pred.cpu(), want_i) and torch.equal(best.cpu(), want_v)
    mc = torch.randint(0, 3000, (37, 18), device="cuda", generator=g)
    mc[:, 15:] = -1
    mc[3] = -1                                        # no candidate at all
    pm = ops.argmax_rows(x, mc).cpu()
    xc, mcc = x.cpu(), mc.cpu()
    for j in range(37):                               # the reference's scan (train.py:156-164)
        cand = [e for e in mcc[j].tolist() if e != -1]
        p, prob = -1, 0.0
        for k in range(3000):
            if k in cand and (p == -1 or prob < xc[j, k]):
                p, prob = k, xc[j, k]
        assert pm[j].item() == p, j

    class Vocab:
        def idx2word(self, i):
            return "ans%d" % i
    C = 50
    m = ODA.Model(None, C)
    m.load_state_dict(rc.synth_state_dict("ODA", C, seed=3))
    m = m.cuda().train()
    v, q, _ = (t.cuda() for t in rc.synth_inputs(6, 36, C, seed=1))
    sample = {"v": v, "q_idxes": q, "q_id": torch.arange(100, 106)}
    items = predict_answers(m, sample, Vocab())
    assert m.training                                  # mode restored
    m.eval()
    with torch.no_grad():
        ref = m(sample).cpu().max(1)[1]
    assert items == [{"question_id": 100 + j, "answer": "ans%d" % int(ref[j])} for j in range(6)]


def test_bf16_feature_shards_through_the_prefetcher(cuda):
    """engine.pack_feature_shard + HostPrefetcher: features shipped as bf16 and widened on the device give exactly the
    step the fp32 copies of the same (bf16-representable) values give — plain and widened into a graph's static buffers."""
    from oracle import reasoning_core as rc
    from vqa_playground_pytorch_b200 import ops
    from vqa_playground_pytorch_b200.config import ODA
    from vqa_playground_pytorch_b200.engine import GraphedStep, HostPrefetcher, pack_feature_shard
    C = 40
    m = ODA.Model(None, C)
    m.load_state_dict(rc.synth_state_dict("ODA", C, seed=3))
    m = m.cuda().eval()
    batches = []
    for i in range(3):
        v, q, a = rc.synth_inputs(4, 36, C, seed=20 + i)
        batches.append({"v": v.to(torch.bfloat16).float(), "q_idxes": q, "a": a})
    shard = pack_feature_shard(batches)
    assert shard[0]["v"].dtype == torch.bfloat16 and shard[0]["v"].is_pinned()
    want = [ops.kld_loss(m({k: t.cuda() for k, t in b.items()}), b["a"].cuda()).item() for b in batches]
    pf = HostPrefetcher(shard, "cuda:0")
    got = [ops.kld_loss(m(s), s["a"]).item() for s in pf]
    assert pf.bytes_per_batch == sum(t.numel() * t.element_size() for t in shard[0].values())
    assert got == want
    m = ODA.Model(None, C)                  # a fresh model: autograd caches per-parameter streams at first use
    m.load_state_dict(rc.synth_state_dict("ODA", C, seed=3))
    m = m.cuda().eval()
    step = GraphedStep(m, {k: t.cuda() for k, t in batches[0].items()}, warmup=1)
    got2 = [step(s).item() for s in HostPrefetcher(shard, "cuda:0", widen_into=step.static)]
    assert all(abs(x - y) <= 1e-5 * abs(y) for x, y in zip(got2, want))
