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


@dataclass
class LiteasrDataclass:
    name: Optional[str] = None


def resolve_interpolations(cfg, root_name: str = "model"):
    """Resolve the ``${model.xxx}`` defaults of a plain dataclass instance (what OmegaConf would do)."""
    prefix = "${%s." % root_name
    for _ in range(4):
        changed = False
        for k, v in list(vars(cfg).items()):
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
