"""Thin torch.autograd.Function wrappers over the C ABI (include/vqacore.h).

PyTorch is used only for device memory, the current stream and the autograd graph; every
forward and backward below is one C call that enqueues hand-written sm_100a kernels on
torch's current stream.  Nothing here computes with ATen, and there is no CPU fallback:
CPU tensors raise ValueError.
"""
import ctypes as C
import itertools

import torch

from . import _lib
from ._lib import GLIMPSES, MAXG, fp

H_DIM, F_DIM, A_DIM, Q_DIM, D_DIM = 310, 510, 620, 2400, 2048

_seed_counter = itertools.count(1)
_base_seed = 0x5EED


def manual_seed(seed):
    """Base key of the Philox dropout stream (each train-mode forward takes the next counter)."""
    global _base_seed, _seed_counter
    _base_seed = int(seed) & 0xFFFFFFFF
    _seed_counter = itertools.count(1)


def next_seed():
    return (_base_seed << 32) | (next(_seed_counter) & 0xFFFFFFFF)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t, name, dims=None):
    if not isinstance(t, torch.Tensor):
        raise ValueError("%s: expected a tensor" % name)
    if not t.is_cuda:
        raise ValueError("%s: must be a CUDA tensor (libvqacore has no CPU path)" % name)
    if t.dtype != torch.float32:
        raise ValueError("%s: must be float32, got %s" % (name, t.dtype))
    if dims is not None and t.dim() != dims:
        raise ValueError("%s: expected %d dims, got %d" % (name, dims, t.dim()))
    return t.contiguous()


def _p(t):
    return None if t is None else t.data_ptr()


def _workspace(nbytes, device):
    return torch.empty((int(nbytes),), device=device, dtype=torch.uint8) if nbytes else None


def _math(math):
    return _lib.MATH_BY_NAME[math] if isinstance(math, str) else int(math)


def _repack_bytes(tensors, K):
    """Extra scratch for operands TMA cannot address as given (base not 16-byte aligned or row stride not a multiple
    of 4 floats) when K itself is a multiple of 4 — the case vqa_linear_*_workspace_bytes() cannot see from the shapes
    (include/vqacore.h).  The tensor-core ops copy such an operand into padded scratch instead of leaving the tensor
    cores."""
    if K % 4 != 0:
        return 0                      # already counted by the workspace query
    bad = [t for t in tensors if t.data_ptr() % 16 or t.stride(0) % 4]
    return sum(t.shape[0] * ((K + 3) // 4 * 4) * 4 + 256 for t in tensors) if bad else 0


# =========================================================================== grouped linear
def linear_forward(xs, ws, bs, act, p, seed, layers, math=0, outs=None, bits=None):
    """Y_g = act(dropout(X_g) W_g^T + b_g). xs/ws/bs: lists (one entry per group), X_g [M,K] (row stride
    may exceed K). Returns list of Y_g [M,N] (or writes into `outs`, which may be strided views)."""
    g = len(xs)
    M, K = xs[0].shape
    N = ws[0].shape[0]
    pr = _lib.LinearFwd()
    pr.groups, pr.M, pr.K, pr.N, pr.act, pr.math, pr.p, pr.seed = g, M, K, N, act, _math(math), float(p), int(seed)
    if outs is None:
        outs = [torch.empty((M, N), device=xs[0].device, dtype=torch.float32) for _ in range(g)]
    for i in range(g):
        assert xs[i].stride(1) == 1 and outs[i].stride(1) == 1
        pr.X[i], pr.ldx[i] = xs[i].data_ptr(), xs[i].stride(0)
        pr.W[i], pr.b[i] = ws[i].data_ptr(), _p(bs[i])
        pr.Y[i], pr.ldy[i] = outs[i].data_ptr(), outs[i].stride(0)
        pr.layer[i], pr.drop_index_base[i] = int(layers[i]), 0
        pr.drop_bits[i] = _p(bits[i]) if bits is not None else None
    nbytes = _lib.lib().vqa_linear_fwd_workspace_bytes(pr.math, g, M, K, N) + (_repack_bytes(xs, K) if pr.math else 0)
    ws = _workspace(nbytes, xs[0].device)
    pr.workspace, pr.workspace_bytes = _p(ws), (ws.numel() if ws is not None else 0)
    _lib.check(_lib.lib().vqa_linear_fwd(C.byref(pr), _stream()), "vqa_linear_fwd")
    return outs


def linear_backward(xs, ws, ys, dys, act, p, seed, layers, need_dx, math=0, dws=None, dbs=None, dxs=None,
                    accumulate_w=False, accumulate_x=False, bits=None, pool=None):
    """Backward of the grouped linear (vqa_linear_bwd).  pool = (alpha [M,4], dpooled [M/regions,4,K], regions) adds
    the gradient of an attention pooling over X_0 into the dX_0 store."""
    g = len(xs)
    M, K = xs[0].shape
    N = ws[0].shape[0]
    dev = xs[0].device
    pr = _lib.LinearBwd()
    pr.groups, pr.M, pr.K, pr.N, pr.act, pr.math, pr.p, pr.seed = g, M, K, N, act, _math(math), float(p), int(seed)
    pr.accumulate_w, pr.accumulate_x = int(accumulate_w), int(accumulate_x)
    if dws is None:
        dws = [torch.empty_like(w) for w in ws]
    if dbs is None:
        dbs = [torch.empty((N,), device=dev, dtype=torch.float32) for _ in range(g)]
    if dxs is None:
        dxs = [torch.empty((M, K), device=dev, dtype=torch.float32) if need_dx else None for _ in range(g)]
    for i in range(g):
        pr.X[i], pr.ldx[i] = xs[i].data_ptr(), xs[i].stride(0)
        pr.W[i] = ws[i].data_ptr()
        pr.Y[i], pr.ldy[i] = _p(ys[i]), (ys[i].stride(0) if ys[i] is not None else N)
        pr.dY[i], pr.lddy[i] = dys[i].data_ptr(), dys[i].stride(0)
        pr.dW[i], pr.db[i] = _p(dws[i]), _p(dbs[i])
        pr.dX[i], pr.lddx[i] = _p(dxs[i]), (dxs[i].stride(0) if dxs[i] is not None else K)
        pr.layer[i], pr.drop_index_base[i] = int(layers[i]), 0
        pr.drop_bits[i] = _p(bits[i]) if bits is not None else None
    if pool is not None:
        pr.pool_alpha, pr.pool_dpooled, pr.pool_regions = pool[0].data_ptr(), pool[1].data_ptr(), int(pool[2])
    nbytes = _lib.lib().vqa_linear_bwd_workspace_bytes(pr.math, g, M, K, N) + (_repack_bytes(xs, K) if pr.math else 0)
    ws = _workspace(nbytes, dev)
    pr.workspace, pr.workspace_bytes = _p(ws), (ws.numel() if ws is not None else 0)
    _lib.check(_lib.lib().vqa_linear_bwd(C.byref(pr), _stream()), "vqa_linear_bwd")
    return dws, dbs, dxs


def seed_advance(seed_dev):
    """*seed_dev += 1 on the current stream (captured at the head of a graphed step)."""
    _lib.check(_lib.lib().vqa_seed_advance(seed_dev.data_ptr(), _stream()), "vqa_seed_advance")


def dropout_bits(p, seed, layer, n, device):
    """Packed keep-bits of n elements (vqa_dropout_bits); 4 spare bytes at the end, which the kernels that read a
    quad's bits across a byte boundary may touch."""
    out = torch.zeros(((n + 15) // 16 * 2 + 4,), device=device, dtype=torch.uint8)
    _lib.check(_lib.lib().vqa_dropout_bits(float(p), int(seed), None, int(layer), int(n), out.data_ptr(), _stream()),
               "vqa_dropout_bits")
    return out


def dropout_bits_batch(p, seed, sites, device):
    """Keep-bits of several dropout sites in one launch (vqa_dropout_bits_batch). sites: [(layer, n), ...];
    returns one uint8 tensor per site (with a few spare bytes, as the kernels that read quads across byte
    boundaries expect)."""
    segs = (_lib.BitsSegment * len(sites))()
    outs = []
    for i, (layer, n) in enumerate(sites):
        out = torch.zeros(((n + 15) // 16 * 2 + 4,), device=device, dtype=torch.uint8)
        segs[i].layer, segs[i].n, segs[i].out = int(layer), int(n), out.data_ptr()
        outs.append(out)
    _lib.check(_lib.lib().vqa_dropout_bits_batch(float(p), int(seed), None, segs, len(sites), _stream()),
               "vqa_dropout_bits_batch")
    return outs


class LinearFn(torch.autograd.Function):
    """Single-group linear with fused input dropout / bias / activation (MyLinear, MyConv1d k=1)."""

    @staticmethod
    def forward(ctx, x, w, b, act, p, seed, layer, math):
        x2 = _chk(x, "x").reshape(-1, x.shape[-1])
        w2 = _chk(w, "weight").reshape(w.shape[0], -1)
        y = linear_forward([x2], [w2], [b], act, p, seed, [layer], math)[0]
        ctx.save_for_backward(x2, w2, y)
        ctx.meta = (act, p, seed, layer, math, x.shape, w.shape, b is not None)
        return y.reshape(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w2, y = ctx.saved_tensors
        act, p, seed, layer, math, xshape, wshape, has_b = ctx.meta
        dy2 = dy.contiguous().reshape(-1, dy.shape[-1])
        dws, dbs, dxs = linear_backward([x2], [w2], [y], [dy2], act, p, seed, [layer], ctx.needs_input_grad[0], math)
        dx = dxs[0].reshape(xshape) if dxs[0] is not None else None
        return dx, dws[0].reshape(wshape), (dbs[0] if has_b else None), None, None, None, None, None


# =========================================================================== Mutan fusion
def _ptr_table(tensors, n=MAXG):
    arr = (fp * n)()
    for i, t in enumerate(tensors):
        arr[i] = _p(t)
    return arr


class MutanFn(torch.autograd.Function):
    """sum_r (x1 W1_r^T + b1_r) (.) (x2 W2_r^T + b2_r), x2 broadcast over the region axis of x1."""

    @staticmethod
    def forward(ctx, x1, x2, math, R, *wb):
        # wb = W1_0, b1_0, ..., W1_{R-1}, b1_{R-1}, W2_0, b2_0, ...
        x1c, x2c = _chk(x1, "inputs1"), _chk(x2, "inputs2")
        a = x1c.reshape(-1, x1c.shape[-1])
        c = x2c.reshape(-1, x2c.shape[-1])
        M, K1 = a.shape
        Mh, K2 = c.shape
        if M % Mh != 0:
            raise ValueError("MutanFusion: inputs1 rows (%d) not a multiple of inputs2 rows (%d)" % (M, Mh))
        W1, b1 = wb[0:2 * R:2], wb[1:2 * R:2]
        W2, b2 = wb[2 * R::2], wb[2 * R + 1::2]
        Fd = W1[0].shape[0]
        dev = a.device
        H1 = torch.empty((R, M, Fd), device=dev, dtype=torch.float32)
        H2 = torch.empty((R, Mh, Fd), device=dev, dtype=torch.float32)
        y = torch.empty((M, Fd), device=dev, dtype=torch.float32)
        pr = _lib.MutanFwd()
        pr.R, pr.M, pr.K1, pr.K2, pr.F, pr.rows_per_h2, pr.math = R, M, K1, K2, Fd, M // Mh, _math(math)
        pr.X1, pr.ldx1, pr.X2, pr.ldx2 = a.data_ptr(), K1, c.data_ptr(), K2
        for r in range(R):
            pr.W1[r], pr.b1[r], pr.W2[r], pr.b2[r] = W1[r].data_ptr(), _p(b1[r]), W2[r].data_ptr(), _p(b2[r])
        pr.H1, pr.H2, pr.Y, pr.ldy = H1.data_ptr(), H2.data_ptr(), y.data_ptr(), Fd
        ws = _workspace(_lib.lib().vqa_mutan_workspace_bytes(pr.math, R, M, M // Mh, K1, K2, Fd, 0), dev)
        pr.workspace, pr.workspace_bytes = _p(ws), (ws.numel() if ws is not None else 0)
        _lib.check(_lib.lib().vqa_mutan_fwd(C.byref(pr), _stream()), "vqa_mutan_fwd")
        ctx.save_for_backward(a, c, H1, H2, *wb)
        ctx.meta = (math, R, x1.shape, x2.shape)
        return y.reshape(*x1.shape[:-1], Fd)

    @staticmethod
    def backward(ctx, dy):
        a, c, H1, H2, *wb = ctx.saved_tensors
        math, R, x1shape, x2shape = ctx.meta
        W1, W2 = wb[0:2 * R:2], wb[2 * R::2]
        M, K1 = a.shape
        Mh, K2 = c.shape
        Fd = W1[0].shape[0]
        dev = a.device
        dy2 = dy.contiguous().reshape(M, Fd)
        grads = [torch.empty_like(t) for t in wb]
        dH2 = torch.empty_like(H2)
        dx1 = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        dx2 = torch.empty_like(c) if ctx.needs_input_grad[1] else None
        pr = _lib.MutanBwd()
        pr.R, pr.M, pr.K1, pr.K2, pr.F, pr.rows_per_h2, pr.math = R, M, K1, K2, Fd, M // Mh, _math(math)
        pr.X1, pr.ldx1, pr.X2, pr.ldx2 = a.data_ptr(), K1, c.data_ptr(), K2
        for r in range(R):
            pr.W1[r], pr.W2[r] = W1[r].data_ptr(), W2[r].data_ptr()
            pr.dW1[r], pr.db1[r] = grads[2 * r].data_ptr(), grads[2 * r + 1].data_ptr()
            pr.dW2[r], pr.db2[r] = grads[2 * R + 2 * r].data_ptr(), grads[2 * R + 2 * r + 1].data_ptr()
        pr.H1, pr.H2, pr.dY, pr.lddy, pr.dH2 = H1.data_ptr(), H2.data_ptr(), dy2.data_ptr(), Fd, dH2.data_ptr()
        pr.dX1, pr.lddx1, pr.dX2, pr.lddx2 = _p(dx1), K1, _p(dx2), K2
        ws = _workspace(_lib.lib().vqa_mutan_workspace_bytes(pr.math, R, M, M // Mh, K1, K2, Fd, 1), dev)
        pr.workspace, pr.workspace_bytes = _p(ws), (ws.numel() if ws is not None else 0)
        _lib.check(_lib.lib().vqa_mutan_bwd(C.byref(pr), _stream()), "vqa_mutan_bwd")
        return (dx1.reshape(x1shape) if dx1 is not None else None,
                dx2.reshape(x2shape) if dx2 is not None else None, None, None, *grads)


# =========================================================================== region softmax + pooling
class RegionSoftmaxPoolFn(torch.autograd.Function):
    """alpha = softmax_regions(conv_att(dropout(fuse))); pooled = alpha^T x.  Returns (pooled, alpha).
    Both outputs are differentiable: an incoming gradient of alpha [B,N,G] (the reference uses alpha1[0] downstream,
    config/CoR2.py:216) is added to <dpooled, x> before the softmax backward (vqa_region_softmax_pool_bwd,
    `dalpha_ext`)."""

    @staticmethod
    def forward(ctx, x, fuse, wc, bc, p, seed, layer):
        xc, fc = _chk(x, "inputs", 3), _chk(fuse, "fuse", 3)
        B, N, Dd = xc.shape
        Ff = fc.shape[2]
        wc2 = _chk(wc, "conv_att.weight").reshape(GLIMPSES, Ff)
        alpha = torch.empty((B, N, GLIMPSES), device=xc.device, dtype=torch.float32)
        pooled = torch.empty((B, GLIMPSES, Dd), device=xc.device, dtype=torch.float32)
        pr = _lib.PoolFwd()
        pr.B, pr.N, pr.Ff, pr.D = B, N, Ff, Dd
        pr.drop.p, pr.drop.layer, pr.drop.seed = float(p), int(layer), int(seed)
        pr.fuse, pr.Wc, pr.bc, pr.x = fc.data_ptr(), wc2.data_ptr(), bc.data_ptr(), xc.data_ptr()
        pr.alpha, pr.pooled = alpha.data_ptr(), pooled.data_ptr()
        _lib.check(_lib.lib().vqa_region_softmax_pool_fwd(C.byref(pr), _stream()), "vqa_region_softmax_pool_fwd")
        ctx.save_for_backward(xc, fc, wc2, alpha)
        ctx.meta = (p, seed, layer, wc.shape)
        return pooled, alpha

    @staticmethod
    def backward(ctx, dpooled, dalpha_in):
        xc, fc, wc2, alpha = ctx.saved_tensors
        p, seed, layer, wshape = ctx.meta
        B, N, Dd = xc.shape
        Ff = fc.shape[2]
        dev = xc.device
        if dpooled is None:
            dpooled = torch.zeros((B, GLIMPSES, Dd), device=dev, dtype=torch.float32)
        dpooled = dpooled.contiguous()
        ext = dalpha_in.contiguous() if dalpha_in is not None else None
        dalpha = torch.empty_like(alpha)
        dz = torch.empty_like(alpha)
        dwc = torch.empty_like(wc2)
        dbc = torch.empty((GLIMPSES,), device=dev, dtype=torch.float32)
        dfuse = torch.empty_like(fc) if ctx.needs_input_grad[1] else None
        dx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        pr = _lib.PoolBwd()
        pr.B, pr.N, pr.Ff, pr.D = B, N, Ff, Dd
        pr.drop.p, pr.drop.layer, pr.drop.seed = float(p), int(layer), int(seed)
        pr.accumulate_w, pr.accumulate_x = 0, 0
        pr.fuse, pr.Wc, pr.x, pr.alpha = fc.data_ptr(), wc2.data_ptr(), xc.data_ptr(), alpha.data_ptr()
        pr.dpooled, pr.dalpha0_ext, pr.dalpha_ext = dpooled.data_ptr(), None, _p(ext)
        pr.dalpha, pr.dz, pr.dWc, pr.dbc = dalpha.data_ptr(), dz.data_ptr(), dwc.data_ptr(), dbc.data_ptr()
        pr.dfuse, pr.dx = _p(dfuse), _p(dx)
        _lib.check(_lib.lib().vqa_region_softmax_pool_bwd(C.byref(pr), _stream()), "vqa_region_softmax_pool_bwd")
        return dx, dfuse, dwc.reshape(wshape), dbc, None, None, None


# =========================================================================== CoR2 compound objects
class CorCompoundFn(torch.autograd.Function):
    """v2[b,j,:] = pooled[b,0,:]*g1[b,:] + (sum_i alpha[b,i,0]) * x[b,j,:]*g2[b,:]  — decare_cat + the alpha1[0]-weighted
    sum of config/CoR2.py:191-199, :215-216 in collapsed form (vqa_cor_compound_fwd / _bwd).
    Differentiable in pooled (glimpse 0), alpha (glimpse 0), g1, g2; `x` is the graph input v there and gets no
    gradient (requesting one raises)."""

    @staticmethod
    def forward(ctx, x, pooled, alpha, g1, g2):
        xc, pc, ac = _chk(x, "x", 3), _chk(pooled, "pooled", 3), _chk(alpha, "alpha", 3)
        g1c, g2c = _chk(g1, "g1", 2), _chk(g2, "g2", 2)
        B, N, Dd = xc.shape
        v2 = torch.empty_like(xc)
        pr = _lib.CompoundFwd()
        pr.B, pr.N, pr.D = B, N, Dd
        pr.x, pr.pooled, pr.alpha, pr.g1, pr.g2, pr.v2 = (xc.data_ptr(), pc.data_ptr(), ac.data_ptr(), g1c.data_ptr(),
                                                          g2c.data_ptr(), v2.data_ptr())
        _lib.check(_lib.lib().vqa_cor_compound_fwd(C.byref(pr), _stream()), "vqa_cor_compound_fwd")
        ctx.save_for_backward(xc, pc, ac, g1c, g2c)
        return v2

    @staticmethod
    def backward(ctx, dv2):
        xc, pc, ac, g1c, g2c = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("CorCompoundFn: no gradient for x (it is the graph input v in CoR2; "
                                      "config/CoR2.py:215 passes v, v)")
        B, N, Dd = xc.shape
        dv2 = dv2.contiguous()
        dg1, dg2 = torch.empty_like(g1c), torch.empty_like(g2c)
        dpooled = torch.zeros_like(pc)                 # the kernel ACCUMULATES into the glimpse-0 slice
        ext = torch.empty((B,), device=xc.device, dtype=torch.float32)
        pr = _lib.CompoundBwd()
        pr.B, pr.N, pr.D = B, N, Dd
        pr.x, pr.pooled, pr.alpha, pr.g1, pr.g2 = xc.data_ptr(), pc.data_ptr(), ac.data_ptr(), g1c.data_ptr(), g2c.data_ptr()
        pr.dv2, pr.dg1, pr.dg2, pr.dpooled, pr.dalpha0_ext = (dv2.data_ptr(), dg1.data_ptr(), dg2.data_ptr(),
                                                              dpooled.data_ptr(), ext.data_ptr())
        _lib.check(_lib.lib().vqa_cor_compound_bwd(C.byref(pr), _stream()), "vqa_cor_compound_bwd")
        dalpha = torch.zeros_like(ac)
        dalpha[:, :, 0] = ext.unsqueeze(1)              # d s / d alpha[b,i,0] = 1 for every region i
        return None, dpooled, dalpha, dg1, dg2


# =========================================================================== ODA object-difference attention
class OdaPairAttnFn(torch.autograd.Function):
    """ODA's pairwise-difference attention (config/ODA.py:216-226): logits over the never-materialised
    [B,N,N*H] tensor, region softmax, pooling.  Returns (pooled [B,G,D], alpha [B,N,G]); gradients for vl, ql, the
    conv_att weight [G,N*H(,1)] and bias.  x (= v) is a graph input; alpha is a side output here."""

    @staticmethod
    def forward(ctx, x, vl, ql, w, bc, p, seed, layer):
        xc, vlc, qlc = _chk(x, "inputs", 3), _chk(vl, "vl", 3), _chk(ql, "ql", 2)
        B, N, Dd = xc.shape
        Hd = vlc.shape[2]
        w2 = _chk(w, "conv_att.weight").reshape(GLIMPSES, N * Hd)
        dev = xc.device
        alpha = torch.empty((B, N, GLIMPSES), device=dev, dtype=torch.float32)
        pooled = torch.empty((B, GLIMPSES, Dd), device=dev, dtype=torch.float32)
        wsum = torch.empty((GLIMPSES, Hd), device=dev, dtype=torch.float32)
        pr = _lib.OdaFwd()
        pr.B, pr.N, pr.H, pr.D, pr.train = B, N, Hd, Dd, int(p > 0.0)
        pr.drop.p, pr.drop.layer, pr.drop.seed = float(p), int(layer), int(seed)
        pr.vl, pr.ql, pr.W, pr.bc, pr.x = vlc.data_ptr(), qlc.data_ptr(), w2.data_ptr(), bc.data_ptr(), xc.data_ptr()
        pr.wsum, pr.alpha, pr.pooled = wsum.data_ptr(), alpha.data_ptr(), pooled.data_ptr()
        # train mode: keep-bit cache of the dropped [B,N,N*H] tensor + partial logits, filled here, read by backward
        nws = int(_lib.lib().vqa_oda_pair_attn_workspace_bytes(B, N, Hd)) if p > 0.0 else 0
        ws = torch.empty((max(nws, 1),), device=dev, dtype=torch.uint8)
        pr.workspace, pr.workspace_bytes, pr.keep_bits_ready = ws.data_ptr(), nws, 0
        _lib.check(_lib.lib().vqa_oda_pair_attn_fwd(C.byref(pr), _stream()), "vqa_oda_pair_attn_fwd")
        ctx.save_for_backward(xc, vlc, qlc, w2, alpha, wsum, ws)
        ctx.meta = (p, seed, layer, w.shape)
        ctx.mark_non_differentiable(alpha)
        return pooled, alpha

    @staticmethod
    def backward(ctx, dpooled, _dalpha):
        xc, vlc, qlc, w2, alpha, wsum, ws = ctx.saved_tensors
        p, seed, layer, wshape = ctx.meta
        B, N, Dd = xc.shape
        Hd = vlc.shape[2]
        dev = xc.device
        dpooled = dpooled.contiguous()
        dalpha, dz = torch.empty_like(alpha), torch.empty_like(alpha)
        dwsum = torch.empty_like(wsum)
        dW, dbc = torch.empty_like(w2), torch.empty((GLIMPSES,), device=dev, dtype=torch.float32)
        dvl, dql = torch.empty_like(vlc), torch.empty_like(qlc)
        pr = _lib.OdaBwd()
        pr.B, pr.N, pr.H, pr.D, pr.train = B, N, Hd, Dd, int(p > 0.0)
        pr.drop.p, pr.drop.layer, pr.drop.seed = float(p), int(layer), int(seed)
        pr.accumulate_w = 0
        pr.vl, pr.ql, pr.W, pr.x, pr.alpha, pr.wsum = (vlc.data_ptr(), qlc.data_ptr(), w2.data_ptr(), xc.data_ptr(),
                                                       alpha.data_ptr(), wsum.data_ptr())
        pr.dpooled, pr.dalpha, pr.dz, pr.dwsum = dpooled.data_ptr(), dalpha.data_ptr(), dz.data_ptr(), dwsum.data_ptr()
        pr.dW, pr.dbc, pr.dvl, pr.dql = dW.data_ptr(), dbc.data_ptr(), dvl.data_ptr(), dql.data_ptr()
        pr.workspace, pr.workspace_bytes = ws.data_ptr(), (ws.numel() if p > 0.0 else 0)
        _lib.check(_lib.lib().vqa_oda_pair_attn_bwd(C.byref(pr), _stream()), "vqa_oda_pair_attn_bwd")
        return None, dvl, dql, dW.reshape(wshape), dbc, None, None, None


# =========================================================================== SkipThoughts question encoder
GRU_INPUT_MASK_LAYER, GRU_HIDDEN_MASK_LAYER = 64, 67     # Philox layer ids of drop_ir/ii/in and drop_hr/hi/hn


def seq_dropout_masks(p, seed, layer0, B, dim, nmasks, device):
    """[nmasks, B, dim] sequence-tied dropout multipliers (0 or 1/(1-p)) — SequentialDropout, putils/__init__.py:503-539."""
    out = torch.empty((nmasks, B, dim), device=device, dtype=torch.float32)
    _lib.check(_lib.lib().vqa_seq_dropout_masks(float(p), int(seed), None, int(layer0), B, dim, nmasks, out.data_ptr(),
                                                _stream()), "vqa_seq_dropout_masks")
    return out


def _gru_act(af):
    if af not in ("relu", "tanh"):
        raise ValueError("BayesianGRU: af must be 'relu' or 'tanh', got %r" % (af,))
    return 1 if af == "relu" else 3


class BayesianGruFn(torch.autograd.Function):
    """SkipThoughts.forward (putils/__init__.py:975-982): embedding -> BayesianGRU over T tokens -> hidden state at the
    last non-PAD token.  One autograd node; every kernel inside is libvqacore's: the three input projections of all
    time steps are ONE grouped tensor-core linear over [T*B, 620], each step is one grouped linear for the three
    recurrent projections plus one fused gate kernel, the backward walks the steps in reverse with dgrad-only linears
    and finishes with ONE wgrad over all steps per weight."""

    @staticmethod
    def forward(ctx, idx, emb_w, w_ir, b_ir, w_ii, b_ii, w_in, b_in, w_hr, w_hi, w_hn, p, seed, af, math):
        if not (isinstance(idx, torch.Tensor) and idx.is_cuda and idx.dtype == torch.int64 and idx.dim() == 2):
            raise ValueError("q_idxes: expected a [B, T] int64 CUDA tensor of token ids")
        idx = idx.contiguous()
        B, T = idx.shape
        emb_w = _chk(emb_w, "embedding.weight", 2)
        I, H = emb_w.shape[1], w_hr.shape[0]
        wi = [_chk(w, "weight_i*", 2) for w in (w_ir, w_ii, w_in)]
        bi = [_chk(b, "bias_i*", 1) for b in (b_ir, b_ii, b_in)]
        wh = [_chk(w, "weight_h*", 2) for w in (w_hr, w_hi, w_hn)]
        dev, L, act = idx.device, _lib.lib(), _gru_act(af)
        train = p > 0.0
        imask = seq_dropout_masks(p, seed, GRU_INPUT_MASK_LAYER, B, I, 3, dev) if train else None
        hmask = seq_dropout_masks(p, seed, GRU_HIDDEN_MASK_LAYER, B, H, 3, dev) if train else None
        X = torch.empty((3, T * B, I), device=dev, dtype=torch.float32)
        _lib.check(L.vqa_gru_embed_fwd(B, T, I, idx.data_ptr(), emb_w.data_ptr(), _p(imask), X.data_ptr(), _stream()),
                   "vqa_gru_embed_fwd")
        gi = linear_forward([X[0], X[1], X[2]], wi, bi, 0, 0.0, 0, [0, 0, 0], math)        # 3 x [T*B, H]
        hs = torch.empty((T, B, H), device=dev, dtype=torch.float32)
        r, i, n, ghn = (torch.empty_like(hs) for _ in range(4))
        hm = torch.empty((3, T, B, H), device=dev, dtype=torch.float32)
        gh_r, gh_i = torch.empty((B, H), device=dev), torch.empty((B, H), device=dev)
        for t in range(T):
            pr = _lib.GruGateFwd()
            pr.B, pr.H, pr.act = B, H, act
            if t > 0:
                linear_forward([hm[0, t - 1], hm[1, t - 1], hm[2, t - 1]], wh, [None] * 3, 0, 0.0, 0, [0, 0, 0], math,
                               outs=[gh_r, gh_i, ghn[t]])
                pr.gh[0], pr.gh[1], pr.gh[2] = gh_r.data_ptr(), gh_i.data_ptr(), ghn[t].data_ptr()
                pr.h_prev = hs[t - 1].data_ptr()
            for g in range(3):
                pr.gi[g] = gi[g][t * B:(t + 1) * B].data_ptr()
                pr.hmask[g] = hmask[g].data_ptr() if train else None
                pr.hm[g] = hm[g, t].data_ptr()
            pr.h, pr.r, pr.i, pr.n = hs[t].data_ptr(), r[t].data_ptr(), i[t].data_ptr(), n[t].data_ptr()
            _lib.check(L.vqa_gru_gate_fwd(C.byref(pr), _stream()), "vqa_gru_gate_fwd")
        last_pos = torch.empty((B,), device=dev, dtype=torch.int64)
        _lib.check(L.vqa_gru_last_pos(B, T, idx.data_ptr(), last_pos.data_ptr(), _stream()), "vqa_gru_last_pos")
        out = torch.empty((B, H), device=dev, dtype=torch.float32)
        _lib.check(L.vqa_gru_select_last(B, H, hs.data_ptr(), last_pos.data_ptr(), out.data_ptr(), _stream()),
                   "vqa_gru_select_last")
        ctx.save_for_backward(idx, X, hs, r, i, n, ghn, hm, last_pos, *wi, *wh)
        ctx.masks = (imask, hmask)
        ctx.meta = (act, math, emb_w.shape, ctx.needs_input_grad[1])
        ctx.all_hiddens = hs
        return out

    @staticmethod
    def backward(ctx, dx):
        idx, X, hs, r, i, n, ghn, hm, last_pos, w_ir, w_ii, w_in, w_hr, w_hi, w_hn = ctx.saved_tensors
        imask, hmask = ctx.masks
        act, math, emb_shape, need_emb = ctx.meta
        B, T = idx.shape
        I, H = X.shape[2], hs.shape[2]
        dev, L = idx.device, _lib.lib()
        dx = dx.contiguous()
        wi, wh = [w_ir, w_ii, w_in], [w_hr, w_hi, w_hn]
        dA = torch.empty((3, T, B, H), device=dev, dtype=torch.float32)        # d(pre-activations) = d gi_* (= d gh_r, d gh_i)
        dghn = torch.zeros((T, B, H), device=dev, dtype=torch.float32)         # d gh_n
        dh_part = [torch.empty((B, H), device=dev), torch.empty((B, H), device=dev)]
        dhm, have_part = None, False
        for t in range(T - 1, -1, -1):
            pr = _lib.GruGateBwd()
            pr.B, pr.H, pr.act, pr.t = B, H, act, t
            pr.dh_partial = dh_part[(t + 1) & 1].data_ptr() if have_part else None
            for g in range(3):
                pr.dhm[g] = dhm[g].data_ptr() if dhm is not None else None
                pr.hmask[g] = hmask[g].data_ptr() if hmask is not None else None
                pr.da[g] = dA[g, t].data_ptr()
            pr.dx_last, pr.last_pos = dx.data_ptr(), last_pos.data_ptr()
            pr.r, pr.i, pr.n = r[t].data_ptr(), i[t].data_ptr(), n[t].data_ptr()
            pr.gh_n = ghn[t].data_ptr() if t > 0 else None
            pr.h_prev = hs[t - 1].data_ptr() if t > 0 else None
            pr.dgh_n, pr.dh_partial_out = dghn[t].data_ptr(), dh_part[t & 1].data_ptr()
            _lib.check(L.vqa_gru_gate_bwd(C.byref(pr), _stream()), "vqa_gru_gate_bwd")
            have_part = True
            if t > 0:       # gradients of the masked copies of h_{t-1}: dgrad only, the wgrad of all steps follows below
                _, _, dhm = linear_backward([hm[0, t - 1], hm[1, t - 1], hm[2, t - 1]], wh, [None] * 3,
                                            [dA[0, t], dA[1, t], dghn[t]], 0, 0.0, 0, [0, 0, 0], True, math,
                                            dws=[None] * 3, dbs=[None] * 3)
        dwh = [torch.zeros_like(w) for w in wh]
        if T > 1:
            linear_backward([hm[g, :T - 1].reshape(-1, H) for g in range(3)], wh, [None] * 3,
                            [dA[0, 1:].reshape(-1, H), dA[1, 1:].reshape(-1, H), dghn[1:].reshape(-1, H)], 0, 0.0, 0,
                            [0, 0, 0], False, math, dws=dwh, dbs=[None] * 3)
        dX = torch.empty((3, T * B, I), device=dev, dtype=torch.float32) if need_emb else None
        dwi, dbi, _ = linear_backward([X[0], X[1], X[2]], wi, [None] * 3, [dA[g].reshape(-1, H) for g in range(3)], 0,
                                      0.0, 0, [0, 0, 0], need_emb, math,
                                      dxs=[dX[0], dX[1], dX[2]] if need_emb else [None] * 3)
        demb = None
        if need_emb:
            demb = torch.zeros(emb_shape, device=dev, dtype=torch.float32)
            _lib.check(L.vqa_gru_embed_bwd(B, T, I, idx.data_ptr(), _p(imask), dX.data_ptr(), demb.data_ptr(), _stream()),
                       "vqa_gru_embed_bwd")
        return (None, demb, dwi[0], dbi[0], dwi[1], dbi[1], dwi[2], dbi[2], dwh[0], dwh[1], dwh[2], None, None, None, None)


# =========================================================================== loss
class KldLogSoftmaxFn(torch.autograd.Function):
    """KLDivLoss(size_average=False)(log_softmax(x,1), a) — train.py:536-544."""

    @staticmethod
    def forward(ctx, logits, target):
        x, a = _chk(logits, "logits", 2), _chk(target, "target", 2)
        B, Cc = x.shape
        rows = torch.empty((B,), device=x.device, dtype=torch.float32)
        dlogits = torch.empty_like(x)
        pr = _lib.KldParams()
        pr.B, pr.C, pr.grad_scale = B, Cc, 1.0
        pr.logits, pr.target, pr.loss_rows, pr.dlogits = x.data_ptr(), a.data_ptr(), rows.data_ptr(), dlogits.data_ptr()
        _lib.check(_lib.lib().vqa_kld_logsoftmax_fwd_bwd(C.byref(pr), _stream()), "vqa_kld_logsoftmax_fwd_bwd")
        ctx.save_for_backward(dlogits)
        return rows

    @staticmethod
    def backward(ctx, drows):
        (dlogits,) = ctx.saved_tensors
        # drows is all-ones for loss = rows.sum(); general case scales per row
        return dlogits * drows.unsqueeze(1), None


def kld_loss_rows(logits, target):
    return KldLogSoftmaxFn.apply(logits, target)


class KldLossFn(torch.autograd.Function):
    """Scalar KLDivLoss(size_average=False)(log_softmax(x,1), a) (train.py:536-544): per-row loss + gradient in one
    kernel, a fixed-order row sum, and the backward scaling by the incoming gradient read on the device."""

    @staticmethod
    def forward(ctx, logits, target):
        x, a = _chk(logits, "logits", 2), _chk(target, "target", 2)
        B, Cc = x.shape
        rows = torch.empty((B,), device=x.device, dtype=torch.float32)
        dlogits = torch.empty_like(x)
        loss = torch.empty((), device=x.device, dtype=torch.float32)
        pr = _lib.KldParams()
        pr.B, pr.C, pr.grad_scale = B, Cc, 1.0
        pr.logits, pr.target, pr.loss_rows, pr.dlogits = x.data_ptr(), a.data_ptr(), rows.data_ptr(), dlogits.data_ptr()
        L = _lib.lib()
        _lib.check(L.vqa_kld_logsoftmax_fwd_bwd(C.byref(pr), _stream()), "vqa_kld_logsoftmax_fwd_bwd")
        _lib.check(L.vqa_sum_rows(B, rows.data_ptr(), loss.data_ptr(), _stream()), "vqa_sum_rows")
        ctx.save_for_backward(dlogits)
        ctx.rows = rows
        return loss

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        out = torch.empty_like(dlogits)
        gg = g.contiguous().to(torch.float32)
        _lib.check(_lib.lib().vqa_scale_by_device_scalar(dlogits.numel(), dlogits.data_ptr(), gg.data_ptr(),
                                                         out.data_ptr(), _stream()), "vqa_scale_by_device_scalar")
        return out, None


def kld_loss(logits, target):
    return KldLossFn.apply(logits, target)


def cast_bf16_to_f32(src, out=None):
    """fp32 copy of a bf16 CUDA tensor on the current stream (vqa_cast_bf16_f32; exact)."""
    if not (isinstance(src, torch.Tensor) and src.is_cuda and src.dtype == torch.bfloat16):
        raise ValueError("cast_bf16_to_f32: expected a bfloat16 CUDA tensor")
    src = src.contiguous()
    if out is None:
        out = torch.empty(src.shape, device=src.device, dtype=torch.float32)
    _lib.check(_lib.lib().vqa_cast_bf16_f32(src.numel(), src.data_ptr(), out.data_ptr(), _stream()), "vqa_cast_bf16_f32")
    return out


def argmax_rows(logits, mc_idx=None, return_best=False):
    """pred[b] = first argmax of logits[b, :] (or of the candidate columns mc_idx[b, :], -1 = padding) on the GPU
    (vqa_argmax_rows) — the `output.data.cpu().max(1)` of the reference's eval loop (train.py:146-164) without moving
    the [B, num_ans] logits to the host."""
    x = _chk(logits.detach(), "logits", 2)
    B, Cc = x.shape
    pred = torch.empty((B,), device=x.device, dtype=torch.int64)
    best = torch.empty((B,), device=x.device, dtype=torch.float32) if return_best else None
    mc, n_mc = None, 0
    if mc_idx is not None:
        mc = mc_idx.to(device=x.device, dtype=torch.int64).contiguous()
        if mc.dim() != 2 or mc.shape[0] != B:
            raise ValueError("a_mc_idx must be [B, n_candidates]")
        n_mc = mc.shape[1]
    _lib.check(_lib.lib().vqa_argmax_rows(B, Cc, x.data_ptr(), _p(mc), n_mc, pred.data_ptr(), _p(best), _stream()),
               "vqa_argmax_rows")
    return (pred, best) if return_best else pred


# =========================================================================== whole-model plans
_MODEL = {
    "CoR2": ("vqa_cor2_workspace_bytes", "vqa_cor2_fwd", "vqa_cor2_bwd"),
    "ODA": ("vqa_oda_workspace_bytes", "vqa_oda_fwd", "vqa_oda_bwd"),
}


def stash_tensor(name):
    """A ReLU output of the most recent forward plan, as a [rows, cols] view into its workspace (tests only)."""
    model, B, N, Cc, ws = ModelCoreFn.last_workspace
    off, rows, cols, ld = C.c_size_t(), C.c_int64(), C.c_int64(), C.c_int64()
    _lib.check(_lib.lib().vqa_stash_info(0 if model == "CoR2" else 1, name.encode(), B, N, Cc, C.byref(off),
                                         C.byref(rows), C.byref(cols), C.byref(ld)), "vqa_stash_info")
    flat = ws[off.value:off.value + rows.value * ld.value * 4].view(torch.float32)
    return flat.view(rows.value, ld.value)[:, :cols.value]


def _fill_model_params(pr, B, N, Cc, train, math, seed, v, q, ptab, logits, alpha1, alpha2, v2, ws, seed_dev=None):
    pr.B, pr.N, pr.C, pr.train, pr.math, pr.seed = B, N, Cc, int(train), _math(math), int(seed)
    pr.seed_dev = _p(seed_dev)
    pr.v, pr.q, pr.params = v.data_ptr(), q.data_ptr(), ptab
    pr.logits, pr.alpha1, pr.alpha2, pr.v2 = logits.data_ptr(), alpha1.data_ptr(), _p(alpha2), _p(v2)
    pr.workspace, pr.workspace_bytes = ws.data_ptr(), ws.numel()


class ModelCoreFn(torch.autograd.Function):
    """Model.forward of config/CoR2.py:201-237 or config/ODA.py:200-240 as ONE C call (and one more
    for the whole backward).  Returns (logits, alpha1, alpha2, v2); the last three are
    non-differentiable side outputs feeding `alpha_dict`."""
    last_workspace = None

    @staticmethod
    def forward(ctx, model, v, q, train, math, seed, seed_dev, num_regions, num_ans, grad_sink, *params):
        wsb, fwd, _ = _MODEL[model]
        L = _lib.lib()
        vc = _chk(v, "sample['v']").reshape(-1, num_regions, D_DIM)
        qc = _chk(q, "question embedding", 2)
        B, N = vc.shape[0], num_regions
        if qc.shape[0] != B or qc.shape[1] != Q_DIM:
            raise ValueError("question embedding must be [%d,%d], got %s" % (B, Q_DIM, tuple(qc.shape)))
        Cc = int(num_ans)
        dev = vc.device
        params = [_chk(t, "parameter") for t in params]
        ws = torch.empty((int(getattr(L, wsb)(B, N, Cc)),), device=dev, dtype=torch.uint8)
        logits = torch.empty((B, Cc), device=dev, dtype=torch.float32)
        alpha1 = torch.empty((B, N, GLIMPSES), device=dev, dtype=torch.float32)
        cor2 = model == "CoR2"
        alpha2 = torch.empty((B, N, GLIMPSES), device=dev, dtype=torch.float32) if cor2 else None
        v2 = torch.empty((B, N, D_DIM), device=dev, dtype=torch.float32) if cor2 else None
        ptab = _ptr_table(params, len(params))
        pr = _lib.ModelFwd()
        _fill_model_params(pr, B, N, Cc, train, math, seed, vc, qc, ptab, logits, alpha1, alpha2, v2, ws, seed_dev)
        _lib.check(getattr(L, fwd)(C.byref(pr), _stream()), fwd)
        ctx.model, ctx.meta, ctx.grad_sink = model, (B, N, Cc, train, math, seed, seed_dev), grad_sink
        ModelCoreFn.last_workspace = (model, B, N, Cc, ws)      # test introspection (vqa_stash_info)
        ctx.keep = (vc, qc, params, ws, logits, alpha1, alpha2, v2)
        # alpha / v2 are side outputs (alpha_dict, visualisation): without this autograd zero-fills a [B,N,2048]
        # gradient for v2 on every backward (75 MB, 14 us at B=256)
        ctx.set_materialize_grads(False)
        if cor2:
            ctx.mark_non_differentiable(alpha1, alpha2, v2)
            return logits, alpha1, alpha2, v2
        ctx.mark_non_differentiable(alpha1)
        return logits, alpha1, None, None

    @staticmethod
    def backward(ctx, dlogits, *_unused):
        _, _, bwd = _MODEL[ctx.model]
        L = _lib.lib()
        B, N, Cc, train, math, seed, seed_dev = ctx.meta
        vc, qc, params, ws, logits, alpha1, alpha2, v2 = ctx.keep
        if dlogits is None:
            dlogits = torch.zeros_like(logits)
        dlogits = dlogits.contiguous()
        sink = ctx.grad_sink
        if sink is not None:
            # data-parallel engine: gradients land directly in its flat buffer (no autograd copies)
            grads, accumulate, ret, flat = sink.slices, int(sink.accumulate), [None] * len(params), sink.flat
        else:
            flat = torch.empty((sum(t.numel() for t in params),), device=dlogits.device, dtype=torch.float32)
            grads, off = [], 0
            for t in params:
                grads.append(flat[off:off + t.numel()].view(t.shape))
                off += t.numel()
            accumulate, ret = 0, grads
        ptab = _ptr_table(params, len(params))
        gtab = _ptr_table(grads, len(grads))
        pr = _lib.ModelBwd()
        _fill_model_params(pr.fwd, B, N, Cc, train, math, seed, vc, qc, ptab, logits, alpha1, alpha2, v2, ws, seed_dev)
        pr.dlogits, pr.grads, pr.accumulate = dlogits.data_ptr(), gtab, accumulate
        pr.grads_flat, pr.grads_flat_bytes = flat.data_ptr(), flat.numel() * 4
        for k, ev in enumerate(getattr(sink, "group_events", None) or ()):
            pr.group_events[k] = ev.cuda_event      # recorded by the plan as gradient group k completes
        dq = torch.empty_like(qc) if ctx.needs_input_grad[2] else None     # a trainable question encoder in front
        pr.dq = _p(dq)
        _lib.check(getattr(L, bwd)(C.byref(pr), _stream()), bwd)
        if sink is not None:
            sink.after_backward()
        ctx.keep = None
        return (None, None, dq, None, None, None, None, None, None, None, *ret)
