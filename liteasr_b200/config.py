"""Plain-dataclass stand-ins for the pieces of ``liteasr.config`` the hot path touches.

hydra / omegaconf are not required (and not installed in the build image): registries fall back to plain dataclasses, and
when hydra *is* importable the decorators also store the nodes in its ConfigStore exactly like the reference
(models/__init__.py:72-86, criterions/__init__.py:41-56)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

try:  # pragma: no cover - optional dependency
    from omegaconf import II, MISSING  # type: ignore
    HAVE_OMEGACONF = True
except Exception:  # noqa: BLE001
    HAVE_OMEGACONF = False
    MISSING = "???"

    def II(path: str) -> str:
        return "${%s}" % path


def _reference_base():
    """``liteasr.config.LiteasrDataclass`` when LiteASR itself is importable: the reference registries assert
    ``issubclass(dataclass, LiteasrDataclass)`` (models/__init__.py:77, criterions/__init__.py:46, optims/__init__.py), so the
    config dataclasses of this package must derive from THAT class to be registered there."""
    try:
        import sys
        mod = sys.modules.get("liteasr.config")
        if mod is None:
            import importlib
            mod = importlib.import_module("liteasr.config")
        base = getattr(mod, "LiteasrDataclass", None)
        import dataclasses
        if isinstance(base, type) and dataclasses.is_dataclass(base):
            return base
    except Exception:  # noqa: BLE001  (liteasr absent, or not importable: hydra missing / Python >= 3.11 default-factory error)
        pass
    return None


_REF_BASE = _reference_base()

if _REF_BASE is not None:
    LiteasrDataclass = _REF_BASE
else:
    @dataclass
    class LiteasrDataclass:  # stand-in with the reference's single field (config/__init__.py:14-16)
        name: Optional[str] = None


def resolve_interpolations(cfg, root_name: str = "model", fields=None):
    """Resolve the ``${model.xxx}`` defaults (what OmegaConf does on access).  ``cfg`` may be a plain dataclass instance, any
    attribute bag, or an OmegaConf ``DictConfig`` (already resolved on access: nothing to do).  With ``fields`` given the values
    are read with ``getattr`` and written back with ``setattr`` (works for objects that do not keep them in ``__dict__``)."""
    prefix = "${%s." % root_name
    names = list(fields) if fields is not None else list(vars(cfg).keys())
    for _ in range(4):
        changed = False
        for k in names:
            v = getattr(cfg, k, None)
            if isinstance(v, str) and v.startswith(prefix) and v.endswith("}"):
                ref = v[len(prefix):-1]
                tgt = getattr(cfg, ref)
                if not (isinstance(tgt, str) and tgt.startswith("${")):
                    setattr(cfg, k, tgt)
                    changed = True
        if not changed:
            break
    return cfg


def store_in_hydra(group: str, name: str, dataclass_type) -> None:
    try:  # pragma: no cover - optional dependency
        from hydra.core.config_store import ConfigStore  # type: ignore
        node = dataclass_type()
        node._name = name
        ConfigStore.instance().store(name=name, group=group, node=node)
    except Exception:  # noqa: BLE001
        pass
