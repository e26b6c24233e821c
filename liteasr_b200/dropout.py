"""Dropout sites of the U2 hot path and the per-store RNG state.

The reference places ``nn.Dropout`` / ``F.dropout`` at (paths relative to /root/reference/liteasr):
  encoder   pos-enc on ``x * sqrt(d)`` and on ``pos_emb``             nets/positional_encoding.py:68-75
            per Conformer layer: FFN inner (after the activation)       nets/feed_forward.py:19
                                 four sub-layer outputs                  nets/conformer_layer.py:42,54,63,125
                                 attention probabilities                 nets/attention.py:55
  decoder   pos-enc on ``embed * sqrt(d) + pe``                        nets/positional_encoding.py:49-56
            per layer: self / source attention probabilities + outputs, FFN inner + output   nets/transformer_layer.py:29-61,161-177
  CTC head  on its input, ALWAYS (``F.dropout`` without ``training=``, quirk Q3)             nets/ctc.py:29
with rates from ``U2Config`` (models/u2.py:39-66; ``config/model/my_U2.yaml``: 0.1 everywhere except the three attention
rates, 0.0).  Masks here come from the library's own Philox stream (csrc/philox.cuh, include/lasr.h "Dropout"): every site has
an id, every forward pass of a net takes a snapshot of the store's ``{seed, step}`` pair after advancing ``step``, and forward
and backward kernels of that pass index the stream by (step, site, row, column) -- no mask tensor exists.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops

NET_ENC, NET_DEC, NET_CTC = 1, 2, 3
# encoder
ENC_POS_X, ENC_POS_EMB, ENC_FFM_INNER, ENC_FFM_OUT, ENC_ATT_PROB, ENC_ATT_OUT, ENC_CONV_OUT, ENC_FF_INNER, ENC_FF_OUT = range(1, 10)
# decoder
DEC_POS, DEC_SELF_PROB, DEC_SELF_OUT, DEC_SRC_PROB, DEC_SRC_OUT, DEC_FF_INNER, DEC_FF_OUT = range(1, 8)
# CTC head
CTC_IN = 1
GLOBAL_LAYER = 0xFFFF  # sites that exist once per net (positional encodings, the CTC input)


def site_id(net: int, layer: int, kind: int) -> int:
    return ((net & 0xFF) << 24) | ((layer & 0xFFFF) << 8) | (kind & 0xFF)


class RngState:
    """Device ``{seed, step}`` (two int64) of one ParamStore.  ``begin_pass`` advances ``step`` on the device and returns a
    private snapshot for the forward AND backward kernels of that pass (a later forward of another net, or of the next
    micro-step, must not change the masks an outstanding backward regenerates).  Everything is stream-ordered and CUDA-graph
    capturable: a captured step draws fresh masks on every replay."""

    def __init__(self, device: torch.device, seed: Optional[int] = None):
        if seed is None:
            seed = torch.initial_seed()  # torch.manual_seed(...) therefore also seeds the dropout stream
        self.state = torch.tensor([seed & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=device)

    def seed(self, seed: int, step: int = 0) -> None:
        self.state.copy_(torch.tensor([seed & 0x7FFFFFFFFFFFFFFF, step], dtype=torch.int64))

    def begin_pass(self) -> torch.Tensor:
        ops.rng_advance(self.state)
        return self.state.clone()

    def host(self):
        s = self.state.cpu().tolist()
        return int(s[0]), int(s[1])


def drop(rng: Optional[torch.Tensor], net: int, layer: int, kind: int, p) -> Optional[ops.Drop]:
    """``ops.Drop`` for one site, or None when dropout is off there (no rng snapshot = eval mode, or p rounds to 0)."""
    if rng is None or p is None or float(p) <= 0.0:
        return None
    d = ops.Drop(rng, site_id(net, layer, kind), float(p))
    return d if d.thr else None


def has_dropout(module) -> bool:
    """True if any sub-module of ``module`` carries a non-zero ``dropout_rate`` (cached on the module)."""
    cached = getattr(module, "_lasr_has_dropout", None)
    if cached is None:
        cached = any(float(getattr(m, "dropout_rate", 0.0) or 0.0) > 0.0 for m in module.modules()
                     if type(m).__name__ != "Conv2DLayer")  # the front end ignores its rate (quirk Q8)
        module._lasr_has_dropout = cached
    return cached


def check_rates(module, *rates) -> None:
    for r in rates:
        if not (isinstance(r, (int, float)) and 0.0 <= float(r) <= ops.DROP_P_MAX):
            raise ValueError(f"{type(module).__name__}: dropout rate {r!r} must be a number in [0, {ops.DROP_P_MAX}]")
