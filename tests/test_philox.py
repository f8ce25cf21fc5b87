"""Known-answer tests for the Philox4x32-10 the oracle and the kernels share (Random123 KAT vectors)."""
import numpy as np

from oracle import philox


def _run(ctr, key):
    r = philox.philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], key[0], key[1])
    return [int(x[0]) for x in r]


def test_known_answers():
    assert _run((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _run((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _run((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_mask_is_index_addressed():
    full = philox.words(77, 3, 64)
    part = philox.words(77, 3, 16, start=20)
    assert np.array_equal(full[20:36], part)
    m = philox.dropout_mask(77, 3, (1000, 100), 0.5)
    assert 0.48 < m.mean() < 0.52
    assert set(np.unique(m)) == {0.0, 1.0}
    assert not np.array_equal(philox.words(77, 3, 64), philox.words(77, 4, 64))
    assert not np.array_equal(philox.words(77, 3, 64), philox.words(78, 3, 64))


def test_threshold():
    assert philox.threshold(0.5) == 128
    assert philox.threshold(0.0) == 0


def test_mask_bytes_layout():
    """Element idx owns byte (idx & 15) of the Philox output of group idx >> 4 (little-endian over the 4 words)."""
    b = philox.mask_bytes(99, 7, 40, start=8)
    for j, idx in enumerate(range(8, 48)):
        g = idx >> 4
        r = philox.philox4x32_10(np.array([g], dtype=np.uint64), 0, 7, 0, 99, 0)
        w = int(r[(idx >> 2) & 3][0])
        assert int(b[j]) == (w >> (8 * (idx & 3))) & 0xFF
    m1 = philox.dropout_mask(5, 1, (3, 37), 0.5)
    m2 = philox.dropout_mask(5, 1, (111,), 0.5)
    assert np.array_equal(m1.reshape(-1), m2)
