"""Checkpoint interchange and averaging (utils/checkpoint.py:15-73 of the reference).

Checkpoints are model-only ``state_dict`` files named ``model.ep.<name>.pt`` (``LiteasrModel.save``, models/__init__.py:31-32;
trainer.py writes one per epoch).  Because the product keeps the reference's state_dict schema (``liteasr_b200/schema.py``), files
written by either side load into the other with ``strict=True``.  ``load_ckpt`` mirrors the reference's selection rules:

* ``model_avg`` false: the single file ``model.ep.<ckpt_name>.pt``;
* ``model_avg`` true, ``avg_policy`` None: the ``avg_num`` files ending at ``ckpt_name`` in modification-time order;
* ``avg_policy`` = path of a training log: the ``avg_num`` files with the lowest ``valid loss`` among those up to ``ckpt_name``
  (one ``valid loss: <x>`` line per epoch, matched to the files in order).

Averaging sums every entry and divides by ``avg_num`` (true division for floating tensors, floor division for integer ones such
as BatchNorm's ``num_batches_tracked``), exactly like the reference -- host-side, no kernels involved.
"""
from __future__ import annotations

import glob
import logging
import os
import re
from typing import Dict, List, Optional, Sequence

import torch

logger = logging.getLogger(__name__)

_VALID_LOSS = re.compile(r".*valid loss: ([\d\.]+)")


def _read(path: str) -> Dict[str, torch.Tensor]:
    return torch.load(path, map_location="cpu")


def ckpt_file(ckpt_path: str, ckpt_name) -> str:
    return f"{ckpt_path}/model.ep.{ckpt_name}.pt"


def average_state_dicts(paths: Sequence[str], avg_num: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Entry-wise sum of the state dicts in ``paths`` divided by ``avg_num`` (default: their count)."""
    if not paths:
        raise ValueError("no checkpoints to average")
    n = len(paths) if avg_num is None else int(avg_num)
    total = _read(paths[0])
    for p in paths[1:]:
        other = _read(p)
        for key in total:
            total[key] += other[key]
    for key, value in total.items():
        if value is None:
            continue
        if value.is_floating_point():
            value /= n
        else:
            value //= n
    return total


def valid_losses(log_path: str) -> List[float]:
    """One entry per ``valid loss: <x>`` line of a training log (trainer.py:205-209 writes one per epoch)."""
    out = []
    with open(log_path, "r") as f:
        for line in f:
            m = _VALID_LOSS.match(line.strip())
            if m:
                out.append(float(m.group(1)))
    return out


def select_checkpoints(ckpt_path: str, ckpt_name, avg_num: int, avg_policy: Optional[str] = None) -> List[str]:
    """The files ``load_ckpt`` averages (see the module docstring)."""
    files = sorted(glob.glob(f"{ckpt_path}/*"), key=os.path.getmtime)
    last = files.index(ckpt_file(ckpt_path, ckpt_name))
    if last - avg_num + 1 < 0:
        raise AssertionError(f"only {last + 1} checkpoints up to {ckpt_name}, cannot average {avg_num}")
    if avg_policy is None:
        return files[last - avg_num + 1: last + 1]
    losses = valid_losses(avg_policy)
    ranked = sorted(zip(files[: last + 1], losses[: last + 1]), key=lambda fl: fl[1])
    return [f for f, _ in ranked[:avg_num]]


def load_ckpt(cfg) -> Dict[str, torch.Tensor]:
    """``cfg``: anything with the reference's ``InferenceConfig`` fields ckpt_path, ckpt_name, model_avg, avg_num, avg_policy."""
    if not getattr(cfg, "model_avg", False):
        logger.info("loading checkpoint: %s/%s", cfg.ckpt_path, cfg.ckpt_name)
        return _read(ckpt_file(cfg.ckpt_path, cfg.ckpt_name))
    picked = select_checkpoints(cfg.ckpt_path, cfg.ckpt_name, cfg.avg_num, getattr(cfg, "avg_policy", None))
    logger.info("loading average checkpoint from:\n\t%s", "\n\t".join(picked))
    return average_state_dicts(picked, cfg.avg_num)
