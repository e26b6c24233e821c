"""CPU: the numpy restatement of the dropout stream (oracle/philox_oracle.py) against the Random123 known-answer vectors of
Philox4x32-10 (the round function; the masks run it for 7 rounds), plus the mask layout the kernels share with it (16 columns per call, 15-bit lanes, threshold / scale)."""
import numpy as np
import pytest

from oracle import philox_oracle as P


def test_philox4x32_10_known_answers():
    for ctr, key, want in P.KAT:
        got = P.philox4x32_10(*ctr, *key)
        assert [int(x) for x in got] == list(want)


def test_seven_round_variant_is_a_prefix_of_the_same_round_function():
    # 7 rounds followed by 3 more rounds with the key schedule continued = the 10-round known answers
    W0, W1 = 0x9E3779B9, 0xBB67AE85
    for ctr, key, want in P.KAT:
        mid = P.philox4x32(*ctr, *key, rounds=7)
        k0, k1 = (key[0] + 7 * W0) & 0xFFFFFFFF, (key[1] + 7 * W1) & 0xFFFFFFFF
        got = P.philox4x32(*mid, k0, k1, rounds=3)
        assert [int(x) for x in got] == list(want)


def test_mask_layout_threshold_and_scale():
    assert P.threshold(0.1) == 3277 and P.threshold(0.0) == 0 and P.threshold(0.96875) == 0x7C00
    with pytest.raises(ValueError):
        P.threshold(0.97)
    assert abs(P.scale_of(P.threshold(0.1)) - 1 / (1 - 3277 / 32768)) < 1e-12
    m = P.keep_mask(64, 37, 7, 123456789, 5, 0.3)
    assert m.shape == (64, 37)
    # element (r, c): call with counter (c >> 4, r, site, step); e = c & 15 -> high byte = byte (e & 3) of word e >> 2, low byte = the
    # same byte position of the neighbouring word (e >> 2) ^ 1; 15 bits
    assert P.ROUNDS == 7
    w = [int(x) for x in P.philox4x32(np.uint32(2), np.uint32(9), np.uint32(7), np.uint32(5), 123456789 & 0xFFFFFFFF, 123456789 >> 32)]
    want = []
    for e in range(5):
        i, j = e >> 2, e & 3
        u15 = ((((w[i] >> (8 * j)) & 0xFF) << 8) | ((w[i ^ 1] >> (8 * j)) & 0xFF)) & 0x7FFF
        want.append(u15 >= P.threshold(0.3))
    assert want == m[9, 32:37].tolist()
    big = P.keep_mask(512, 512, 1, 1, 1, 0.1)
    assert abs(big.mean() - (1 - 3277 / 32768)) < 4 * (0.1 * 0.9 / big.size) ** 0.5
    assert P.keep_mask(4, 9, 1, 1, 1, 0.0).all()
    # the 15-bit lanes of a block are uniform: every lane position has the expected keep rate, neighbours are uncorrelated
    for p in (0.1, 0.5):
        mm = P.keep_mask(4096, 256, 3, 99, 4, p).reshape(4096, 16, 16)
        pe = P.threshold(p) / 32768
        assert np.abs(mm.mean(axis=(0, 1)) - (1 - pe)).max() < 5 * (pe * (1 - pe) / (4096 * 16)) ** 0.5
        a, b = mm[:, :, 0::2].ravel(), mm[:, :, 1::2].ravel()
        assert abs(np.corrcoef(a, b)[0, 1]) < 0.01
        a, b = mm[:, :, 0:4].ravel(), mm[:, :, 4:8].ravel()  # columns that share bytes (word i and word i ^ 1)
        assert abs(np.corrcoef(a, b)[0, 1]) < 0.02
    # another step / site / seed decorrelates
    for other in (P.keep_mask(512, 512, 1, 1, 2, 0.1), P.keep_mask(512, 512, 2, 1, 1, 0.1), P.keep_mask(512, 512, 1, 2, 1, 0.1)):
        assert abs((big == other).mean() - (0.9 * 0.9 + 0.1 * 0.1)) < 0.01
