"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference hot path.

Imports ``liteasr.models.u2`` / ``liteasr.criterions.hybrid_ctc_attn`` straight from
``/root/reference`` (read-only, exists only in the dev container, never on the GPU box)
so that ``oracle/make_golden.py`` can pin ``oracle/u2_oracle.py`` against the real thing.

The reference cannot be imported as shipped:
  * ``liteasr/__init__.py:3-9`` eagerly imports every sub-package (needs hydra/soundfile);
  * ``liteasr/config/__init__.py:55,93-98`` uses dataclass instances as defaults, which
    Python >= 3.11 rejects.
Three shims fix that without touching a single reference file:
  1. a bare ``liteasr`` package module whose ``__path__`` points at the reference tree;
  2. stub ``omegaconf`` / ``hydra.core.config_store`` modules;
  3. a stub ``liteasr.config`` exporting ``LiteasrDataclass`` and friends.
Nothing in the product (``liteasr_b200``) imports this file.
"""
from __future__ import annotations

import dataclasses
import os
import sys
import types
from typing import Optional

REFERENCE_ROOT = os.environ.get("LITEASR_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "liteasr", "nets"))


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def install() -> None:
    """Install the shims (idempotent)."""
    if "liteasr" in sys.modules and getattr(sys.modules["liteasr"], "_lasr_shim", False):
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    # (1) bare package: skips liteasr/__init__.py
    pkg = types.ModuleType("liteasr")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "liteasr")]
    pkg._lasr_shim = True
    sys.modules["liteasr"] = pkg

    # (2) omegaconf / hydra stubs
    class _OmegaConf:
        @staticmethod
        def merge(*a, **k):
            raise NotImplementedError("omegaconf is stubbed")

        @staticmethod
        def set_struct(*a, **k):
            pass

    if "omegaconf" not in sys.modules:
        _stub("omegaconf", II=lambda s: "${%s}" % s, MISSING="???", OmegaConf=_OmegaConf)
        _stub("omegaconf.listconfig", ListConfig=list)

    class _ConfigStore:
        _inst = None

        @classmethod
        def instance(cls):
            if cls._inst is None:
                cls._inst = cls()
            return cls._inst

        def store(self, **kw):
            pass

    if "hydra" not in sys.modules:
        _stub("hydra")
        _stub("hydra.core")
        _stub("hydra.core.config_store", ConfigStore=_ConfigStore)

    # (3) liteasr.config stub
    @dataclasses.dataclass
    class LiteasrDataclass:
        name: Optional[str] = None

    def _empty(n):
        return dataclasses.dataclass(type(n, (LiteasrDataclass,), {}))

    _stub(
        "liteasr.config",
        LiteasrDataclass=LiteasrDataclass,
        DatasetConfig=_empty("DatasetConfig"),
        PostProcessConfig=_empty("PostProcessConfig"),
        DistributedConfig=_empty("DistributedConfig"),
        LiteasrConfig=_empty("LiteasrConfig"),
        InferenceConfig=_empty("InferenceConfig"),
        _SpecAugmentConfig=_empty("_SpecAugmentConfig"),
        CommonConfig=_empty("CommonConfig"),
        OptimizationConfig=_empty("OptimizationConfig"),
    )


_DROPOUT_FIELDS = (
    "dropout_rate",
    "enc_dropout_rate",
    "enc_pos_dropout_rate",
    "enc_attn_dropout_rate",
    "enc_ff_dropout_rate",
    "dec_dropout_rate",
    "dec_pos_dropout_rate",
    "dec_self_attn_dropout_rate",
    "dec_src_attn_dropout_rate",
    "dec_ff_dropout_rate",
)


def build_reference(cfg: dict, smoothing: float, ctc_weight: float, dropout: float = 0.0):
    """Build the reference ``U2`` + ``HybridCTCLoss`` from a plain dict.

    ``cfg`` keys: input_dim, vocab_size, enc_dim, enc_ff_dim, enc_attn_heads, enc_layers,
    dec_dim, dec_ff_dim, dec_attn_heads, dec_layers.  All ``II(...)``-interpolated dropout
    fields (models/u2.py:49-52,62-66) are overwritten with floats.
    """
    install()
    from liteasr.criterions.hybrid_ctc_attn import HybridCTCLoss, HybridCTCLossConfig
    from liteasr.models.u2 import U2, U2Config

    mcfg = U2Config(**cfg)
    for f in _DROPOUT_FIELDS:
        setattr(mcfg, f, float(dropout))
    model = U2(mcfg)
    ccfg = HybridCTCLossConfig(
        vocab_size=cfg["vocab_size"], smoothing=smoothing, ctc_weight=ctc_weight
    )
    crit = HybridCTCLoss(ccfg)
    return model, crit
