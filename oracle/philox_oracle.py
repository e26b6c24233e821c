"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the dropout masks of liteasr_b200 (csrc/philox.cuh).

The reference draws its dropout masks from torch's global generator (nn.Dropout / F.dropout, nets/feed_forward.py:19,
nets/conformer_layer.py:42-63, nets/attention.py:55, nets/positional_encoding.py:55,75, nets/ctc.py:29); fused kernels cannot
replay that stream, so the product uses its OWN counter-based stream: Philox4x32 (Salmon et al., SC'11 --
"Parallel random numbers: as easy as 1, 2, 3"; the Random123 constants) with ROUNDS = 7 rounds (the paper's minimum
Crush-resistant count, Random123's philox4x32_R<7>; csrc/philox.cuh says why not 10), keyed by the seed and indexed by
(step, site, row, 8-column group).  This file restates that published algorithm and the product's indexing so that the float64
oracle can apply EXACTLY the masks the kernels apply; the round function is pinned by the Random123 known-answer vectors
below, which exist for 10 rounds (same function, `rounds=10`).

    keep(row, col) of a logical (rows, n) tensor:  w = philox4x32(ctr = (col >> 4, row, site, step), key = (seed_lo, seed_hi))
                                                   e = col & 15, i = e >> 2, j = e & 3
                                                   u15 = ((byte j of w[i]) << 8 | (byte j of w[i ^ 1])) & 0x7fff
                                                   keep iff u15 >= thr,  thr = round(p * 32768) <= 0x7c00;  scale = 32768 / (32768 - thr)
(one call serves 16 columns; 15-bit lanes so that the kernels can compare two of them with one half2 instruction)
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


ROUNDS = 7  # csrc/philox.cuh LASR_PHILOX_ROUNDS


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=ROUNDS):
    """Vectorised over numpy uint32 arrays (broadcastable).  Returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(rounds):
            p0 = c0.astype(np.uint64) * M0
            p1 = c2.astype(np.uint64) * M1
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    return philox4x32(c0, c1, c2, c3, k0, k1, rounds=10)


# Random123 known-answer vectors (kat_vectors: "philox4x32 10")
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


THR_MAX = 0x7C00


def threshold(p: float) -> int:
    """Drop threshold on a 15-bit lane; p_eff = thr / 32768 (|p_eff - p| <= 2^-16); p <= 0.96875."""
    thr = int(max(0, round(float(p) * 32768.0)))
    if thr > THR_MAX:
        raise ValueError("dropout rate above 0.96875")
    return thr


def scale_of(thr: int) -> float:
    return 32768.0 / (32768.0 - thr)


def lanes15(w):
    """The 16 fifteen-bit lanes of Philox blocks w = (w0, w1, w2, w3) (uint32 arrays of one shape) -> uint32 array (..., 16)."""
    out = np.empty(np.shape(w[0]) + (16,), dtype=np.uint32)
    for e in range(16):
        i, j = e >> 2, e & 3
        hi = (w[i] >> np.uint32(8 * j)) & np.uint32(0xFF)
        lo = (w[i ^ 1] >> np.uint32(8 * j)) & np.uint32(0xFF)
        out[..., e] = ((hi << np.uint32(8)) | lo) & np.uint32(0x7FFF)
    return out


def keep_mask(rows: int, n: int, site: int, seed: int, step: int, p: float) -> np.ndarray:
    """bool (rows, n): True = kept."""
    thr = threshold(p)
    if thr == 0:
        return np.ones((rows, n), dtype=bool)
    groups = (n + 15) // 16
    g = np.arange(groups, dtype=np.uint32)[None, :]
    r = np.arange(rows, dtype=np.uint32)[:, None]
    w = philox4x32(g, r, np.uint32(site), np.uint32(step & 0xFFFFFFFF), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    w = [np.broadcast_to(x, (rows, groups)) for x in w]
    return (lanes15(w).reshape(rows, groups * 16)[:, :n] >= thr)


def apply(x, site: int, seed: int, step: int, p: float):
    """x: torch tensor (..., n) viewed as (rows, n) row-major -> x * keep * scale (same dtype)."""
    import torch
    thr = threshold(p)
    if thr == 0:
        return x
    n = x.shape[-1]
    rows = x.numel() // n
    m = torch.from_numpy(keep_mask(rows, n, site, seed, step, p)).view(x.shape)
    return x * (m.to(x.dtype) * scale_of(thr))
