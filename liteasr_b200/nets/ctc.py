"""CTC head (nets/ctc.py:7-30 of the reference)."""
from __future__ import annotations

import torch.nn as nn

from .. import functions as F
from .layers import check_rates


class CTC(nn.Module):
    def __init__(self, i_dim: int, o_dim: int, dropout_rate: float):
        super().__init__()
        # quirk Q3: the reference applies F.dropout(p) even in eval(); p = 0 (the U2Config default) makes it the identity
        check_rates(self, dropout_rate)
        self.ctc_lo = nn.Linear(i_dim, o_dim)
        self.dropout_rate = dropout_rate

    def forward(self, xs):
        """(B,T',d) fp32 -> logits (B,T',V) (operand dtype: bf16 in 'bf16' mode, fp32 in 'fp32' mode)."""
        st, _, pfx = F.bind(self, xs.device)
        return F.CTCHeadFn.apply(self, xs, st.anchor, *F.net_params(self, st, pfx))

    def log_softmax(self, x):
        from .. import decoding
        return decoding.log_softmax(self.forward(x))
