"""U2 (shared Conformer encoder + CTC head + attention decoder) -- the reference's ``liteasr/models/u2.py`` on sm_100a.

Same config fields/defaults (models/u2.py:35-67), same attributes (ignore=-1, blank=0, sos=eos=V-1, :111-114), same
``forward`` / ``get_pred_len`` / ``get_target`` / ``get_target_len`` contract and the same state_dict schema.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from typing import Optional, Tuple

import torch
from torch import Tensor

from .. import functions as F
from ..config import II, MISSING, LiteasrDataclass, resolve_interpolations
from ..nets.ctc import CTC
from ..nets.transformer_decoder import TransformerDecoder
from ..nets.transformer_encoder import TransformerEncoder
from . import LiteasrModel, register_model


class EncoderArch(Enum):
    Transformer = "transformer"
    Conformer = "conformer"


class DecoderArch(Enum):
    Transformer = "transformer"


@dataclass
class U2Config(LiteasrDataclass):
    name: Optional[str] = field(default="U2")
    dropout_rate: float = field(default=0.0)
    # shared encoder
    enc_arch: EncoderArch = field(default=EncoderArch.Conformer)
    use_rel: bool = field(default=True)
    input_dim: int = field(default=MISSING)
    enc_dim: int = field(default=256)
    enc_ff_dim: int = field(default=2048)
    enc_attn_heads: int = field(default=4)
    enc_dropout_rate: float = II("model.dropout_rate")
    enc_pos_dropout_rate: float = II("model.enc_dropout_rate")
    enc_attn_dropout_rate: float = II("model.enc_dropout_rate")
    enc_ff_dropout_rate: float = II("model.enc_dropout_rate")
    enc_layers: int = field(default=12)
    activation: str = field(default="swish")
    # attention decoder
    dec_arch: DecoderArch = field(default=DecoderArch.Transformer)
    vocab_size: int = field(default=MISSING)
    dec_dim: int = field(default=256)
    dec_ff_dim: int = field(default=2048)
    dec_attn_heads: int = field(default=4)
    dec_dropout_rate: float = II("model.dropout_rate")
    dec_pos_dropout_rate: float = II("model.dec_dropout_rate")
    dec_self_attn_dropout_rate: float = II("model.dec_dropout_rate")
    dec_src_attn_dropout_rate: float = II("model.dec_dropout_rate")
    dec_ff_dropout_rate: float = II("model.dec_dropout_rate")
    dec_layers: int = field(default=6)
    # liteasr_b200 extension (not in the reference): GEMM operand precision, "bf16" (tcgen05) or "fp32" (SIMT parity mode)
    precision: str = field(default="bf16")


def _arch_value(a, enum_cls=None):
    """Enum member, its value ("conformer") or its name ("Conformer", what a YAML file holds before OmegaConf converts it)."""
    if isinstance(a, Enum):
        return a.value
    s = str(a)
    if enum_cls is not None:
        for e in enum_cls:
            if s == e.value or s == e.name:
                return e.value
    return s


@register_model("U2", dataclass=U2Config)
class U2(LiteasrModel):
    def __init__(self, cfg: U2Config, task=None):
        super().__init__()
        import dataclasses
        cfg = resolve_interpolations(cfg, "model", fields=[f.name for f in dataclasses.fields(U2Config)])
        assert _arch_value(cfg.enc_arch, EncoderArch) in [e.value for e in EncoderArch]
        self.encoder = TransformerEncoder(
            use_rel=cfg.use_rel, i_dim=cfg.input_dim, h_dim=cfg.enc_dim, ff_dim=cfg.enc_ff_dim, n_head=cfg.enc_attn_heads,
            n_layer=cfg.enc_layers, dropout_rate=cfg.enc_dropout_rate, pos_dropout_rate=cfg.enc_pos_dropout_rate,
            attn_dropout_rate=cfg.enc_attn_dropout_rate, ff_dropout_rate=cfg.enc_ff_dropout_rate, activation=cfg.activation,
            arch=_arch_value(cfg.enc_arch, EncoderArch))
        assert _arch_value(cfg.dec_arch, DecoderArch) in [e.value for e in DecoderArch]
        self.decoder = TransformerDecoder(
            i_dim=cfg.vocab_size, h_dim=cfg.dec_dim, ff_dim=cfg.dec_ff_dim, n_head=cfg.dec_attn_heads, n_layer=cfg.dec_layers,
            dropout_rate=cfg.dec_dropout_rate, pos_dropout_rate=cfg.dec_pos_dropout_rate,
            self_attn_dropout_rate=cfg.dec_self_attn_dropout_rate, src_attn_dropout_rate=cfg.dec_src_attn_dropout_rate,
            ff_dropout_rate=cfg.dec_ff_dropout_rate, arch=_arch_value(cfg.dec_arch, DecoderArch))
        self.ctc = CTC(i_dim=cfg.enc_dim, o_dim=cfg.vocab_size, dropout_rate=cfg.dropout_rate)
        if cfg.enc_dim != cfg.dec_dim:
            raise NotImplementedError("enc_dim != dec_dim is not implemented (the reference configs use equal dims)")
        self.ignore = -1
        self.blank = 0
        self.sos = cfg.vocab_size - 1
        self.eos = cfg.vocab_size - 1
        self.vocab_size = cfg.vocab_size
        F.set_precision(self, getattr(cfg, "precision", "bf16"))
        self.last_losses = None

    # ------------------------------------------------------------------ reference contract
    def forward(self, xs, xlens, ys, ylens) -> Tuple[Tensor, Tensor]:
        """-> (h_attn (B,Lmax+1,V), h_ctc (B,T',V)) raw logits (models/u2.py:116-159)."""
        F.bind(self, xs.device)
        h_enc = self.encoder.forward_lens(xs, xlens)
        h_attn = self.decoder.forward_lens(self.decoder_tokens(ys), ylens, h_enc, xlens)
        h_ctc = self.ctc(h_enc)
        return h_attn, h_ctc

    def decoder_tokens(self, ys: Tensor) -> Tensor:
        """ys_in = [sos | ys with ignore -> eos]  (models/u2.py:346-353)."""
        ys_ = ys.masked_fill(ys == self.ignore, self.eos)
        sos = torch.full((ys.size(0), 1), self.sos, dtype=ys.dtype, device=ys.device)
        return torch.cat([sos, ys_], dim=1)

    def get_pred_len(self, xlens) -> Tensor:
        return ((xlens - 1) // 2 - 1) // 2  # models/u2.py:319-321

    def get_target(self, ys, ylens) -> Tuple[Tensor, Tensor]:
        """models/u2.py:323-333."""
        ignore = torch.full((ys.size(0), 1), self.ignore, dtype=ys.dtype, device=ys.device)
        tgt_attn = torch.cat([ys, ignore], dim=1)
        tgt_attn[torch.arange(len(ylens), device=ys.device), ylens] = self.eos
        return tgt_attn, ys

    def get_target_len(self, ylens) -> Tensor:
        return ylens

    def inference(self, x):
        from .. import decoding
        return decoding.attention_rescore(self, x)

    def ctc_prefix_beam_search(self, x):
        from .. import decoding
        return decoding.ctc_prefix_beam_search(self, x)[0][0][0]

    def greedy_ctc(self, xs, xlens=None):
        from .. import decoding
        return decoding.greedy_ctc(self, xs, xlens)

    @classmethod
    def build_model(cls, cfg: U2Config, task=None):
        cfg.input_dim = task.feat_dim
        cfg.vocab_size = task.vocab_size
        return cls(cfg, task)
