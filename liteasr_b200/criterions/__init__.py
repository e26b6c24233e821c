"""Criterion registry with the reference's protocol (criterions/__init__.py:11-56).  Importing registers ``hybrid_ctc``."""
from __future__ import annotations

from ..config import LiteasrDataclass, store_in_hydra

CRITERION_REGISTRY = {}
CRITERION_DATACLASS_REGISTRY = {}
CRITERION_CLASS_NAMES = set()


class LiteasrLoss(object):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg

    @classmethod
    def build_criterion(cls, cfg, task):
        raise NotImplementedError


def register_criterion(name, dataclass=None):
    def register_criterion_cls(cls):
        CRITERION_REGISTRY[name] = cls
        CRITERION_CLASS_NAMES.add(cls.__name__)
        if dataclass is not None:
            assert issubclass(dataclass, LiteasrDataclass)
            CRITERION_DATACLASS_REGISTRY[name] = dataclass
            store_in_hydra("criterion", name, dataclass)
        return cls

    return register_criterion_cls


def build_criterion(cfg, task) -> LiteasrLoss:
    name = getattr(cfg, "name", None)
    cls = CRITERION_REGISTRY[name]
    dc = CRITERION_DATACLASS_REGISTRY[name]()
    for k in vars(dc):
        if hasattr(cfg, k) and getattr(cfg, k) is not None:
            setattr(dc, k, getattr(cfg, k))
    return cls.build_criterion(dc, task)


from . import hybrid_ctc_attn  # noqa: E402,F401  (registers "hybrid_ctc")
