"""Optimizers with the reference's names/config fields (optims/adam.py:12-38, optims/noam.py:10-58), fused on the GPU.

``FusedAdam`` / ``FusedNoam`` keep ``step()`` / ``zero_grad()`` / ``param_groups``-style access but run ONE fused pass over the
flat parameter store: global-norm clip + non-finite skip + (Noam) learning rate + Adam (``lasr_clip_adam_step``).  The
step counter and the skip decision live on the device, so -- unlike ``trainer.py:157`` -- no host sync is needed."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import torch

from .. import ops
from ..config import LiteasrDataclass
from ..store import ParamStore

OPTIMIZER_REGISTRY = {}
OPTIMIZER_DATACLASS_REGISTRY = {}


def register_optimzer(name, dataclass=None):  # [sic] the reference's spelling (optims/__init__.py:73)
    def deco(cls):
        OPTIMIZER_REGISTRY[name] = cls
        if dataclass is not None:
            OPTIMIZER_DATACLASS_REGISTRY[name] = dataclass
        return cls
    return deco


@dataclass
class AdamConfig(LiteasrDataclass):
    name: Optional[str] = field(default="adam")
    lr: float = field(default=1e-3)
    beta1: float = field(default=0.9)
    beta2: float = field(default=0.999)
    eps: float = field(default=1e-8)
    weight_decay: float = field(default=0.0)
    amsgrad: bool = field(default=False)


@dataclass
class NoamConfig(AdamConfig):
    """optims/noam.py:10-17: an AdamConfig with beta2 = 0.98, eps = 1e-9 and the schedule fields."""
    name: Optional[str] = field(default="noam")
    beta2: float = field(default=0.98)
    eps: float = field(default=1e-9)
    model_dim: int = field(default=256)
    factor: float = field(default=1.0)
    warmup: int = field(default=25000)


def _reject_amsgrad(cfg) -> None:
    if getattr(cfg, "amsgrad", False):
        raise NotImplementedError("amsgrad=True is not implemented by the fused optimizer step (the reference default is False)")


class _FusedFlatOptimizer:
    """Difference from the reference, on purpose: ``trainer.py:157`` skips the update only when the gradient norm is NaN; an
    *infinite* norm there makes ``clip_grad_norm_`` scale by 0 and step on NaN/0 gradients.  The fused step skips on any
    non-finite norm (NaN or inf)."""

    def __init__(self, store: ParamStore, *, beta1, beta2, eps, weight_decay, lr=0.0, noam_factor=0.0, model_dim=256.0,
                 warmup=25000.0):
        self.store = store
        dev = store.device
        self.exp_avg = torch.zeros_like(store.flat)
        self.exp_avg_sq = torch.zeros_like(store.flat)
        self.state = torch.zeros(8, dtype=torch.float32, device=dev)  # [step, grad_norm, lr, skipped, grad_scale, ...]
        self.ws = torch.zeros(1024, dtype=torch.float32, device=dev)
        self.hp = dict(beta1=beta1, beta2=beta2, eps=eps, weight_decay=weight_decay, lr=lr, noam_factor=noam_factor,
                       model_dim=float(model_dim), warmup=float(warmup))

    def step(self, clip_grad_norm: float = 5.0, grad_mult: float = 1.0) -> None:
        st = self.store
        ops.clip_adam_step(st.flat, st.gflat, self.exp_avg, self.exp_avg_sq, self.state, self.ws, grad_mult=grad_mult,
                           max_norm=clip_grad_norm, **self.hp)

    def zero_grad(self) -> None:
        self.store.zero_grads()

    # host-side introspection (syncs; for logging only)
    def rate(self) -> float:
        return float(self.state[2])

    def num_updates(self) -> int:
        return int(self.state[0])

    def last_grad_norm(self) -> float:
        return float(self.state[1])


@register_optimzer("adam", dataclass=AdamConfig)
class FusedAdam(_FusedFlatOptimizer):
    def __init__(self, store: ParamStore, cfg: AdamConfig = None):
        cfg = cfg or AdamConfig()
        _reject_amsgrad(cfg)
        super().__init__(store, beta1=cfg.beta1, beta2=cfg.beta2, eps=cfg.eps, weight_decay=cfg.weight_decay, lr=cfg.lr)


@register_optimzer("noam", dataclass=NoamConfig)
class FusedNoam(_FusedFlatOptimizer):
    """Adam (NoamConfig defaults: betas (0.9, 0.98), eps 1e-9) with lr = factor * d^-0.5 * min(step^-0.5, step * warmup^-1.5)
    (optims/noam.py:10-46; the ``lr`` field is overwritten by the schedule before every step, as there)."""

    def __init__(self, store: ParamStore, cfg: NoamConfig = None):
        cfg = cfg or NoamConfig()
        _reject_amsgrad(cfg)
        if not cfg.factor > 0:
            raise ValueError("noam: factor must be > 0")
        super().__init__(store, beta1=cfg.beta1, beta2=cfg.beta2, eps=cfg.eps, weight_decay=cfg.weight_decay, noam_factor=cfg.factor,
                         model_dim=cfg.model_dim, warmup=cfg.warmup)
