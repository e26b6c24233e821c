"""Post-processing transforms (utils/transform/__init__.py:1-46 of the reference): a name -> class registry and the
``PostProcess`` workflow wrapper.  The B200 variants run on CUDA batches after the host-to-device copy."""
from __future__ import annotations

TRANS_REGISTRY = {}


def register_transformation(name):
    def register_transformation_cls(cls):
        TRANS_REGISTRY[name] = cls
        return cls

    return register_transformation_cls


from . import spec_augment  # noqa: E402,F401  (registers "spec_aug")


class PostProcess(object):
    """cfg.workflow = list of registered names, cfg.<name> = that transform's config (utils/transform/__init__.py:36-46)."""

    def __init__(self, cfg):
        self.workflow = [TRANS_REGISTRY[name](getattr(cfg, name)) for name in cfg.workflow]

    def __call__(self, x):
        for transformation in self.workflow:
            x = transformation(x)
        return x
