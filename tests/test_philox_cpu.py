"""CPU: the numpy restatement of the dropout stream (oracle/philox_oracle.py) against the Random123 known-answer vectors of
Philox4x32-10 (the round function; the masks run it for 7 rounds), plus the mask layout the kernels share with it (8 columns per call, 16-bit lanes, threshold / scale)."""
import numpy as np

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
    assert P.threshold(0.1) == 6554 and P.threshold(0.0) == 0 and P.threshold(0.99999) == 65535
    assert abs(P.scale_of(P.threshold(0.1)) - 1 / (1 - 6554 / 65536)) < 1e-12
    m = P.keep_mask(64, 21, 7, 123456789, 5, 0.3)
    assert m.shape == (64, 21)
    # element (r, c) = lane (c & 7) of the call with counter (c >> 3, r, site, step)
    assert P.ROUNDS == 7
    w = P.philox4x32(np.uint32(2), np.uint32(9), np.uint32(7), np.uint32(5), 123456789 & 0xFFFFFFFF, 123456789 >> 32)
    lanes = []
    for x in w:
        lanes += [int(x) & 0xFFFF, int(x) >> 16]
    assert [bool(v >= P.threshold(0.3)) for v in lanes[:5]] == m[9, 16:21].tolist()
    big = P.keep_mask(512, 512, 1, 1, 1, 0.1)
    assert abs(big.mean() - (1 - 6554 / 65536)) < 4 * (0.1 * 0.9 / big.size) ** 0.5
    assert P.keep_mask(4, 9, 1, 1, 1, 0.0).all()
    # another step / site / seed decorrelates
    for other in (P.keep_mask(512, 512, 1, 1, 2, 0.1), P.keep_mask(512, 512, 2, 1, 1, 0.1), P.keep_mask(512, 512, 1, 2, 1, 0.1)):
        assert abs((big == other).mean() - (0.9 * 0.9 + 0.1 * 0.1)) < 0.01
