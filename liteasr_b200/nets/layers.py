"""Parameter containers mirroring the reference's building blocks.

These classes own the ``nn.Parameter``s under the reference's attribute names (so checkpoints interchange, SURVEY.md
section 8b) but carry no arithmetic of their own: the enclosing ``TransformerEncoder`` / ``TransformerDecoder`` / ``CTC``
run the whole stack through ``liteasr_b200.engine`` (hand-written kernels).  torch's stock layer classes are used purely
as initialisers/containers, which also reproduces the reference's default initialisation distributions.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

_STANDALONE = ("{} is a parameter container in liteasr_b200; run it through TransformerEncoder / TransformerDecoder / CTC "
               "(the fused sm_100a pipeline), not stand-alone")


class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise NotImplementedError(_STANDALONE.format(type(self).__name__))


class LayerNorm(nn.LayerNorm):
    """nets/layer_norm.py:8-29 (eps = 1e-12)."""

    def __init__(self, nout: int, dim: int = -1):
        super().__init__(nout, eps=1e-12)
        self.dim = dim

    def forward(self, x):  # pragma: no cover - guard
        raise NotImplementedError(_STANDALONE.format("LayerNorm"))


class Swish(_Container):
    """nets/swish.py:7-16 (x * sigmoid(x)); fused into the fc1 GEMM epilogue / BatchNorm kernel."""


class PositionwiseFeedForward(_Container):
    """nets/feed_forward.py:4-19."""

    def __init__(self, i_dim: int, h_units: int, dropout_rate: float, activation: nn.Module = None):
        super().__init__()
        self.fc1 = nn.Linear(i_dim, h_units)
        self.fc2 = nn.Linear(h_units, i_dim)
        self.dropout_rate = dropout_rate
        self.activation = activation if activation is not None else nn.ReLU()


class MultiHeadAttention(_Container):
    """nets/attention.py:8-71."""

    def __init__(self, n_head: int, i_dim: int, dropout_rate: float):
        super().__init__()
        assert i_dim % n_head == 0
        self.d_k = i_dim // n_head
        self.scaling = self.d_k ** -0.5
        self.h = n_head
        self.linear_q = nn.Linear(i_dim, i_dim)
        self.linear_k = nn.Linear(i_dim, i_dim)
        self.linear_v = nn.Linear(i_dim, i_dim)
        self.linear_o = nn.Linear(i_dim, i_dim)
        self.dropout_rate = dropout_rate


class RelativeMultiHeadAttention(MultiHeadAttention):
    """nets/attention.py:74-154 (legacy rel_shift, absolute-position pos_emb of length T')."""

    def __init__(self, n_head: int, i_dim: int, dropout_rate: float):
        super().__init__(n_head, i_dim, dropout_rate)
        self.linear_pos = nn.Linear(i_dim, i_dim, bias=False)
        self.pos_bias_u = nn.Parameter(torch.Tensor(self.h, self.d_k))
        self.pos_bias_v = nn.Parameter(torch.Tensor(self.h, self.d_k))
        nn.init.xavier_uniform_(self.pos_bias_u)
        nn.init.xavier_uniform_(self.pos_bias_v)


class Convolution(_Container):
    """nets/conformer_convolution.py:4-57."""

    def __init__(self, channels: int, kernel_size: int, bias: bool = True, activation: nn.Module = None):
        super().__init__()
        assert (kernel_size - 1) % 2 == 0  # 'SAME' padding needs an odd kernel
        if kernel_size != 15 or not bias:
            raise NotImplementedError("liteasr_b200 implements the reference's Conformer setting: kernel 15 with bias")
        self.pointwise_conv1 = nn.Conv1d(channels, 2 * channels, kernel_size=1, stride=1, padding=0, bias=bias)
        self.depthwise_conv = nn.Conv1d(channels, channels, kernel_size, stride=1, padding=(kernel_size - 1) // 2,
                                        groups=channels, bias=bias)
        self.pointwise_conv2 = nn.Conv1d(channels, channels, kernel_size=1, stride=1, padding=0, bias=bias)
        self.norm = nn.BatchNorm1d(num_features=channels)
        self.activation = activation if activation is not None else nn.ReLU()


class Conv2DLayer(_Container):
    """nets/subsampling.py:9-48 (the ``dropout_rate`` argument is unused there too, quirk Q8)."""

    def __init__(self, i_dim: int, o_dim: int, dropout_rate: float):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(1, o_dim, 3, 2), nn.ReLU(), nn.Conv2d(o_dim, o_dim, 3, 2), nn.ReLU())
        f_dim = (i_dim - 3) // 2 + 1
        f_dim = (f_dim - 3) // 2 + 1
        self.out = nn.Linear(o_dim * f_dim, o_dim)


class PositionalEncoding(_Container):
    """nets/positional_encoding.py:9-56: persistent (1, max_len, d) sinusoid buffer ``pe`` (quirk Q12)."""

    def __init__(self, h_dim: int, dropout_rate: float, max_len: int = 5000):
        super().__init__()
        if h_dim % 2 != 0:
            raise ValueError("Cannot use sin/cos positional encoding with odd dim (got dim={:d})".format(h_dim))
        self.h_dim = h_dim
        self.scale = math.sqrt(h_dim)
        self.dropout_rate = dropout_rate
        self.max_len = max_len
        self.register_buffer("pe", self.init_pe())

    def init_pe(self) -> torch.Tensor:
        pe = torch.zeros(self.max_len, self.h_dim)
        position = torch.arange(0, self.max_len).unsqueeze(1).float()
        div_term = torch.exp(torch.arange(0, self.h_dim, 2).float() * -(math.log(10000.0) / self.h_dim))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        return pe.unsqueeze(0)

    def ensure(self, length: int, device) -> None:
        """extend_pe (positional_encoding.py:40-53) + device placement."""
        if self.pe.size(1) < length:
            self.max_len = length
            self.pe = self.init_pe()
        if self.pe.device != device or self.pe.dtype != torch.float32:
            self.pe = self.pe.to(device=device, dtype=torch.float32)


class RelativePositionalEncoding(PositionalEncoding):
    """nets/positional_encoding.py:59-75."""


class RelativeEncoderLayer(_Container):
    """Conformer block, nets/conformer_layer.py:84-147 (+ base ctor nets/transformer_layer.py:10-27)."""

    def __init__(self, size, self_attn, feed_forward, feed_forward_macaron, conv, dropout_rate, normalize_before=True,
                 concat_after=False):
        super().__init__()
        if not normalize_before or concat_after:
            raise NotImplementedError("only the reference defaults (pre-norm, no concat) are implemented")
        self.self_attn = self_attn
        self.feed_forward = feed_forward
        self.self_attn_norm = LayerNorm(size)
        self.feed_forward_norm = LayerNorm(size)
        self.dropout_rate = dropout_rate
        self.size = size
        self.normalize_before = normalize_before
        self.feed_forward_macaron = feed_forward_macaron
        self.conv = conv
        self.feed_forward_macaron_norm = LayerNorm(size)
        self.conv_norm = LayerNorm(size)
        self.final_norm = LayerNorm(size)
        self.feed_forward_scale = 0.5


class DecoderLayer(_Container):
    """nets/transformer_layer.py:139-221."""

    def __init__(self, size, self_attn, src_attn, feed_forward, dropout_rate, normalize_before=True, concat_after=False):
        super().__init__()
        if not normalize_before or concat_after:
            raise NotImplementedError("only the reference defaults (pre-norm, no concat) are implemented")
        self.self_attn = self_attn
        self.feed_forward = feed_forward
        self.self_attn_norm = LayerNorm(size)
        self.feed_forward_norm = LayerNorm(size)
        self.dropout_rate = dropout_rate
        self.size = size
        self.normalize_before = normalize_before
        self.src_attn = src_attn
        self.src_attn_norm = LayerNorm(size)


from ..dropout import check_rates  # noqa: E402,F401  (re-exported: the nets validate their rates with it)
