"""Deterministic synthetic inputs and weights.

There is no dataset or checkpoint access, so every test / bench uses:
  * ``synth_batch``: the collator output contract of the reference
    (/root/reference/liteasr/dataset/asr_dataset.py:115-126): ``xs`` f32 (B,Tmax,F) zero padded,
    ``xlens`` i64 sorted descending, ``ys`` i64 (B,Lmax) padded with -1, ``ylens`` i64;
  * ``synth_state_dict``: a full ``U2`` state_dict with the reference key names/shapes.
Both are generated on the CPU with a seeded ``torch.Generator`` so they are identical in the
dev container (where the golden fixtures are made) and on the GPU box.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from ..schema import U2Dims, u2_schema


def pred_len(xlens: torch.Tensor) -> torch.Tensor:
    """models/u2.py:319-321."""
    return ((xlens - 1) // 2 - 1) // 2


def synth_batch(batch: int, tmax: int, lmax: int, vocab: int, seed: int = 42, feat: int = 80,
                min_frac: float = 0.6) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    lo = max(7, int(min_frac * tmax))
    xlens = torch.randint(lo, tmax + 1, (batch,), generator=g)
    xlens[0] = tmax
    xlens, _ = torch.sort(xlens, descending=True)
    xs = torch.randn(batch, tmax, feat, generator=g)
    xs = xs * (torch.arange(tmax).view(1, -1, 1) < xlens.view(-1, 1, 1))
    tp = pred_len(xlens)
    ylens = torch.randint(max(1, lmax // 2), lmax + 1, (batch,), generator=g)
    ylens[0] = lmax
    ylens = torch.minimum(ylens, torch.clamp(tp // 2, min=1))
    real_lmax = int(ylens.max())
    ys = torch.randint(1, vocab - 1, (batch, real_lmax), generator=g)  # 0 = blank, V-1 = sos/eos
    ys = ys.masked_fill(torch.arange(real_lmax).view(1, -1) >= ylens.view(-1, 1), -1)
    return xs.contiguous(), xlens, ys.contiguous(), ylens


def sinusoid_table(n: int, d: int) -> torch.Tensor:
    """nets/positional_encoding.py:29-38 (float32, (1,n,d))."""
    pe = torch.zeros(n, d)
    pos = torch.arange(0, n).unsqueeze(1).float()
    div = torch.exp(torch.arange(0, d, 2).float() * -(math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0)


def synth_state_dict(dims: U2Dims, seed: int = 42) -> Dict[str, torch.Tensor]:
    """Random but well-conditioned weights: dense ~ U(+-1/sqrt(fan_in)); norm gains ~ 1 +- 0.1;
    biases small non-zero (so every gradient path is exercised); BN running stats non-trivial."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    for name, shape, kind in u2_schema(dims):
        if kind == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            sd[name] = uni(shape, 1.0 / math.sqrt(fan_in))
        elif kind == "emb":
            sd[name] = torch.randn(shape, generator=g) * 0.5
        elif kind == "pb":
            sd[name] = uni(shape, 0.3)
        elif kind in ("b", "lnb", "bnb"):
            sd[name] = uni(shape, 0.1)
        elif kind in ("lnw", "bnw"):
            sd[name] = 1.0 + uni(shape, 0.1)
        elif kind == "rm":
            sd[name] = uni(shape, 0.2)
        elif kind == "rv":
            sd[name] = 1.0 + uni(shape, 0.3)
        elif kind == "nbt":
            sd[name] = torch.tensor(3, dtype=torch.long)
        elif kind == "pe":
            sd[name] = sinusoid_table(shape[1], shape[2])
        else:  # pragma: no cover
            raise ValueError(kind)
    return sd
