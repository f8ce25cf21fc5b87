"""Counter-based random bits shared bit-for-bit by the oracle and the CUDA kernels.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): nothing in the product path imports this
file.  The CUDA twin is `csrc/philox.cuh`; both implement Philox4x32-10 (Salmon et al.,
SC'11) with the standard constants, so that a dropout mask generated here on the CPU is
the mask the kernels regenerate in registers.

Why it exists: the reference draws its dropout masks from torch's global generator
(`F.dropout`, /root/reference/config/CoR2.py:78, :116; config/ODA.py:95, :129), which cannot
be reproduced inside a fused kernel.  Train-mode parity is therefore checked by
monkey-patching the reference's `F.dropout` with `dropout_mask` below.

Mask definition (the contract, restated in include/vqacore.h): one Philox call covers a group of 16
consecutive indices, each element owning one byte of the 128-bit output (little-endian over the 4 words):
    out   = Philox4x32-10(key = (seed_lo, seed_hi), ctr = (g_lo, g_hi, layer, 0)),  g = idx >> 4
    byte  = (out[(idx >> 2) & 3] >> (8 * (idx & 3))) & 0xFF
    keep(seed, layer, idx) = byte >= floor(p * 256)        (p = 0.5, the reference's only rate, is exact)
`idx` is the row-major linear index of the element in the LOGICAL tensor the reference
applies dropout to (e.g. (b*N + i)*D + c for compress_v's input, ((b*N+i)*N+j)*H+k for
ODA's pairwise tensor, config/ODA.py:222).
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)
_SH32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. Counter words are uint64 arrays holding 32-bit values,
    keys are python ints. Returns four uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.broadcast_to(np.asarray(c1, dtype=np.uint64), c0.shape)
    c2 = np.broadcast_to(np.asarray(c2, dtype=np.uint64), c0.shape)
    c3 = np.broadcast_to(np.asarray(c3, dtype=np.uint64), c0.shape)
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _SH32, p0 & _MASK32
        hi1, lo1 = p1 >> _SH32, p1 & _MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def words(seed, layer, n, start=0, stream=0):
    """uint32 word for each linear index in [start, start+n)."""
    idx = np.arange(start, start + n, dtype=np.uint64)
    q = idx >> np.uint64(2)
    r = philox4x32_10(q & _MASK32, q >> _SH32, np.uint64(layer & 0xFFFFFFFF), np.uint64(stream),
                      int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    sel = (idx & np.uint64(3)).astype(np.int64)
    out = np.where(sel == 0, r[0], np.where(sel == 1, r[1], np.where(sel == 2, r[2], r[3])))
    return out.astype(np.uint32)


def threshold(p):
    return min(int(np.floor(float(p) * 256.0)), 255)


def mask_bytes(seed, layer, n, start=0):
    """uint8 byte for each linear index in [start, start+n): 16 indices per Philox call."""
    idx = np.arange(start, start + n, dtype=np.uint64)
    g = idx >> np.uint64(4)
    ug, inv = np.unique(g, return_inverse=True)
    r = philox4x32_10(ug & _MASK32, ug >> _SH32, np.uint64(layer & 0xFFFFFFFF), np.uint64(0),
                      int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    table = np.stack(r, axis=1)                                   # [groups, 4] uint32
    word = table[inv, ((idx >> np.uint64(2)) & np.uint64(3)).astype(np.int64)]
    shift = (np.uint32(8) * (idx & np.uint64(3)).astype(np.uint32))
    return ((word >> shift) & np.uint32(0xFF)).astype(np.uint8)


def dropout_mask(seed, layer, shape, p, start=0):
    """float32 {0,1} keep-mask of `shape` (row-major linear index = start + element index; `start` lets a batch be
    processed in chunks with the masks of the full-batch tensor)."""
    n = int(np.prod(shape))
    keep = mask_bytes(seed, layer, n, start=start) >= np.uint8(threshold(p))
    return keep.astype(np.float32).reshape(shape)


def uniform(seed, layer, shape, stream=1):
    """float32 uniform in [0,1): top 24 bits of the word (stream 1 keeps it disjoint from masks)."""
    n = int(np.prod(shape))
    w = words(seed, layer, n, stream=stream)
    return ((w >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).reshape(shape)


def pseudo_normal(seed, layer, shape):
    """Irwin-Hall(4) shifted/scaled to zero mean, unit variance. Version-independent synthetic data."""
    n = int(np.prod(shape))
    acc = np.zeros(n, dtype=np.float32)
    for s in range(4):
        acc += uniform(seed, layer, (n,), stream=2 + s)
    return ((acc - np.float32(2.0)) * np.float32(np.sqrt(3.0))).reshape(shape)
