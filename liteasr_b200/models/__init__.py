"""Model registry with the reference's protocol (models/__init__.py:12-86): ``@register_model(name, dataclass=...)``,
``build_model(cfg, task)``, ``LiteasrModel`` base class.  Importing this package registers ``U2``."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..config import LiteasrDataclass, store_in_hydra

MODEL_REGISTRY = {}
MODEL_DATACLASS_REGISTRY = {}


class LiteasrModel(nn.Module):
    """models/__init__.py:21-50."""

    def __init__(self):
        super().__init__()

    @classmethod
    def build_model(cls, cfg, task):
        raise NotImplementedError

    def inference(self, x):
        raise NotImplementedError

    def save(self, model_path):
        torch.save(self.state_dict(), model_path)

    def get_pred_len(self, xlens):
        raise NotImplementedError

    def get_target(self, ys, ylens):
        raise NotImplementedError

    def get_target_len(self, ylens):
        raise NotImplementedError


def register_model(name, dataclass=None):
    def register_model_cls(cls):
        MODEL_REGISTRY[name] = cls  # re-registering silently overwrites, like the reference (:74)
        if dataclass is not None:
            assert issubclass(dataclass, LiteasrDataclass)
            MODEL_DATACLASS_REGISTRY[name] = dataclass
            store_in_hydra("model", name, dataclass)
        return cls

    return register_model_cls


def build_model(cfg, task) -> LiteasrModel:
    """cfg: a (duck-typed) config with ``name``; merged over the registered dataclass defaults (:53-69)."""
    name = getattr(cfg, "name", None)
    cls = MODEL_REGISTRY[name]
    dc = MODEL_DATACLASS_REGISTRY[name]()
    for k in vars(dc):
        if hasattr(cfg, k) and getattr(cfg, k) is not None:
            setattr(dc, k, getattr(cfg, k))
    return cls.build_model(dc, task)


from . import u2  # noqa: E402,F401  (registers "U2")
