"""state_dict schema of the reference ``U2`` model (the checkpoint-compatibility surface).

Names, shapes and dtypes follow what ``liteasr.models.u2.U2(cfg).state_dict()`` yields in the
reference (probed): /root/reference/liteasr/models/u2.py:72-114,
nets/transformer_encoder.py:28-105, nets/transformer_decoder.py:13-56, nets/ctc.py:15-23.
``U2Dims`` carries only the fields of ``U2Config`` (models/u2.py:35-67) that change shapes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

PE_MAX_LEN = 5000  # nets/positional_encoding.py:14
DW_KERNEL = 15  # nets/transformer_encoder.py:98


@dataclass
class U2Dims:
    input_dim: int = 80
    vocab_size: int = 500
    enc_dim: int = 256
    enc_ff_dim: int = 2048
    enc_attn_heads: int = 4
    enc_layers: int = 12
    dec_dim: int = 256
    dec_ff_dim: int = 2048
    dec_attn_heads: int = 4
    dec_layers: int = 6

    @property
    def freq_out(self) -> int:
        f = (self.input_dim - 3) // 2 + 1  # nets/subsampling.py:38-39
        return (f - 3) // 2 + 1

    @classmethod
    def from_cfg(cls, cfg) -> "U2Dims":
        return cls(**{k: int(getattr(cfg, k)) for k in cls.__dataclass_fields__})


# kind: "w" dense weight, "b" bias, "lnw"/"lnb" LayerNorm affine, "bnw"/"bnb" BatchNorm affine,
#       "rm"/"rv"/"nbt" BatchNorm buffers, "pe" sinusoid buffer, "emb" embedding, "pb" pos_bias_{u,v}
Entry = Tuple[str, Tuple[int, ...], str]


def _ln(p: str, d: int) -> List[Entry]:
    return [(p + ".weight", (d,), "lnw"), (p + ".bias", (d,), "lnb")]


def _lin(p: str, o: int, i: int, bias: bool = True) -> List[Entry]:
    e: List[Entry] = [(p + ".weight", (o, i), "w")]
    if bias:
        e.append((p + ".bias", (o,), "b"))
    return e


def _mha(p: str, d: int) -> List[Entry]:
    e: List[Entry] = []
    for n in ("linear_q", "linear_k", "linear_v", "linear_o"):
        e += _lin(f"{p}.{n}", d, d)
    return e


def u2_schema(c: U2Dims) -> List[Entry]:
    d, f, h = c.enc_dim, c.enc_ff_dim, c.enc_attn_heads
    e: List[Entry] = [
        ("encoder.embed.conv.0.weight", (d, 1, 3, 3), "w"),
        ("encoder.embed.conv.0.bias", (d,), "b"),
        ("encoder.embed.conv.2.weight", (d, d, 3, 3), "w"),
        ("encoder.embed.conv.2.bias", (d,), "b"),
    ]
    e += _lin("encoder.embed.out", d, d * c.freq_out)
    e.append(("encoder.pe.pe", (1, PE_MAX_LEN, d), "pe"))
    for i in range(c.enc_layers):
        p = f"encoder.enc_layers.{i}"
        e += [(p + ".self_attn.pos_bias_u", (h, d // h), "pb"), (p + ".self_attn.pos_bias_v", (h, d // h), "pb")]
        e += _mha(p + ".self_attn", d)
        e += _lin(p + ".self_attn.linear_pos", d, d, bias=False)
        e += _lin(p + ".feed_forward.fc1", f, d) + _lin(p + ".feed_forward.fc2", d, f)
        e += _ln(p + ".self_attn_norm", d) + _ln(p + ".feed_forward_norm", d)
        e += _lin(p + ".feed_forward_macaron.fc1", f, d) + _lin(p + ".feed_forward_macaron.fc2", d, f)
        e += [
            (p + ".conv.pointwise_conv1.weight", (2 * d, d, 1), "w"),
            (p + ".conv.pointwise_conv1.bias", (2 * d,), "b"),
            (p + ".conv.depthwise_conv.weight", (d, 1, DW_KERNEL), "w"),
            (p + ".conv.depthwise_conv.bias", (d,), "b"),
            (p + ".conv.pointwise_conv2.weight", (d, d, 1), "w"),
            (p + ".conv.pointwise_conv2.bias", (d,), "b"),
            (p + ".conv.norm.weight", (d,), "bnw"),
            (p + ".conv.norm.bias", (d,), "bnb"),
            (p + ".conv.norm.running_mean", (d,), "rm"),
            (p + ".conv.norm.running_var", (d,), "rv"),
            (p + ".conv.norm.num_batches_tracked", (), "nbt"),
        ]
        e += _ln(p + ".feed_forward_macaron_norm", d) + _ln(p + ".conv_norm", d) + _ln(p + ".final_norm", d)
    e += _ln("encoder.after_norm", d)

    dd, df = c.dec_dim, c.dec_ff_dim
    e.append(("decoder.embed.weight", (c.vocab_size, dd), "emb"))
    e.append(("decoder.pe.pe", (1, PE_MAX_LEN, dd), "pe"))
    for i in range(c.dec_layers):
        p = f"decoder.dec_layers.{i}"
        e += _mha(p + ".self_attn", dd)
        e += _lin(p + ".feed_forward.fc1", df, dd) + _lin(p + ".feed_forward.fc2", dd, df)
        e += _ln(p + ".self_attn_norm", dd) + _ln(p + ".feed_forward_norm", dd)
        e += _mha(p + ".src_attn", dd)
        e += _ln(p + ".src_attn_norm", dd)
    e += _ln("decoder.after_norm", dd)
    e += _lin("decoder.linear_out", c.vocab_size, dd)
    e += _lin("ctc.ctc_lo", c.vocab_size, d)
    return e


BUFFER_KINDS = ("rm", "rv", "nbt", "pe")


def is_buffer(kind: str) -> bool:
    return kind in BUFFER_KINDS
