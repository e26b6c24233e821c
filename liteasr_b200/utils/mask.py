"""Mask helpers with the reference's semantics (utils/mask.py:8-90).  Host-side plumbing only: the kernels take lengths."""
import torch
from torch import Tensor


def padding_mask(size: Tensor) -> Tensor:
    """True marks padding; width = max(size)  (utils/mask.py:8-27)."""
    base = torch.arange(int(size.max()), device=size.device).unsqueeze(0)
    return base >= size.unsqueeze(1)


def triangle_mask(row: int, col: int = 0, stage: int = 1, diagonal: int = 1) -> Tensor:
    """utils/mask.py:30-90."""
    col = row if col == 0 else col
    r = torch.arange(row).unsqueeze(1)
    c = torch.arange(col).unsqueeze(0)
    return (c // stage) > ((r // stage) + (diagonal - 1))
