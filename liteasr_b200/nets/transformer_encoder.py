"""Conformer encoder with relative-position attention (nets/transformer_encoder.py:28-127 of the reference)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from .. import functions as F
from .layers import (Conv2DLayer, Convolution, LayerNorm, PositionwiseFeedForward, RelativeEncoderLayer,
                     RelativeMultiHeadAttention, RelativePositionalEncoding, Swish, check_rates)


class TransformerEncoder(nn.Module):
    """Same constructor as the reference (transformer_encoder.py:29-43).  Implemented configuration = the north-star one:
    ``use_rel=True, arch="conformer", activation="swish"`` (the other encoder variants are out of scope)."""

    def __init__(self, use_rel: bool, i_dim: int, h_dim: int, ff_dim: int, n_head: int, n_layer: int, dropout_rate: float,
                 pos_dropout_rate: float, attn_dropout_rate: float, ff_dropout_rate: float, activation: str, arch: str) -> None:
        super().__init__()
        if not use_rel or arch != "conformer" or activation != "swish":
            raise NotImplementedError("liteasr_b200.TransformerEncoder implements use_rel=True, arch='conformer', activation='swish'")
        check_rates(self, dropout_rate, pos_dropout_rate, attn_dropout_rate, ff_dropout_rate)
        self.i_dim, self.h_dim, self.n_head = i_dim, h_dim, n_head
        self.embed = Conv2DLayer(i_dim, h_dim, dropout_rate)
        self.pe = RelativePositionalEncoding(h_dim, dropout_rate=pos_dropout_rate)
        act = Swish()
        self.enc_layers = nn.ModuleList([
            RelativeEncoderLayer(
                size=h_dim,
                self_attn=RelativeMultiHeadAttention(n_head, h_dim, attn_dropout_rate),
                feed_forward=PositionwiseFeedForward(h_dim, ff_dim, dropout_rate=ff_dropout_rate, activation=act),
                feed_forward_macaron=PositionwiseFeedForward(h_dim, ff_dim, dropout_rate=ff_dropout_rate, activation=act),
                conv=Convolution(h_dim, 15, activation=act),
                dropout_rate=dropout_rate,
            ) for _ in range(n_layer)
        ])
        self.after_norm = LayerNorm(h_dim)

    def forward_lens(self, x: Tensor, xlens: Optional[Tensor]) -> Tensor:
        """x (B,T,i_dim) fp32, xlens (B,) int64 valid lengths (None = no key mask).  -> (B,T',h_dim) fp32."""
        st, _, pfx = F.bind(self, x.device)
        t2 = ((x.size(1) - 3) // 2 + 1 - 3) // 2 + 1
        self.pe.ensure(t2, x.device)
        return F.EncoderFn.apply(self, x, xlens, st.anchor, *F.net_params(self, st, pfx))

    def forward(self, x: Tensor, mask: Optional[Tensor] = None) -> Tensor:
        """Reference signature: ``mask`` bool (B,T), True = padding (utils/mask.py:8-27), or None."""
        xlens = None
        if mask is not None:
            assert mask.size() == x.size()[:2]  # transformer_encoder.py:114
            xlens = (~mask).sum(dim=1)  # padding masks are prefix masks: recover the lengths
        return self.forward_lens(x, xlens)
