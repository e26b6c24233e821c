"""SpecAugment on the GPU (utils/transform/spec_augment.py:19-125 of the reference).

Same config fields (config/__init__.py:43-51: time_warp, freq_mask, freq_mask_times, time_mask, time_mask_times, inplace,
replace_with_zero) and the same random decisions: ``draw`` consumes Python's ``random`` and ``numpy.random`` in the reference's
call order, so a run seeded like the reference masks / warps the same frames.  The arithmetic runs in one kernel launch for a
whole padded batch (``lasr_spec_augment``): PIL's BICUBIC time warp reproduced bit for bit, then the frequency and time masks.
"""
from __future__ import annotations

import ctypes as C
import random
from typing import Dict, List, Sequence

import numpy as np
import torch

from ... import _lib
from . import register_transformation

MAX_MASKS = 8


@register_transformation("spec_aug")
class SpecAugment(object):
    def __init__(self, cfg):
        self.cfg = cfg
        if cfg.freq_mask_times > MAX_MASKS or cfg.time_mask_times > MAX_MASKS:
            raise NotImplementedError(f"at most {MAX_MASKS} masks of each kind")

    # ------------------------------------------------------------------ host: the reference's random decisions
    def draw(self, t: int, f: int) -> Dict:
        cfg = self.cfg
        p = dict(t=t, center=-1, warped=-1, freq=[], time=[])
        window = cfg.time_warp
        if t - window > window:                                     # spec_augment.py:31-36
            center = random.randrange(window, t - window)
            p["center"], p["warped"] = center, random.randrange(center - window, center + window) + 1
        for width, end in np.random.randint(0, cfg.freq_mask, size=(cfg.freq_mask_times, 2)):   # :65-78
            f_zero = random.randrange(0, f - width)
            if width:
                p["freq"].append((int(f_zero), int(end + f_zero)))
        for width, end in np.random.randint(0, cfg.time_mask, size=(cfg.time_mask_times, 2)):   # :96-113
            if t - width <= 0:
                continue
            t_zero = random.randrange(0, t - width)
            if width:
                p["time"].append((int(t_zero), int(end + t_zero)))
        return p

    @staticmethod
    def pack(params: Sequence[Dict]) -> torch.Tensor:
        npar = int(_lib.lib().lasr_spec_augment_npar())
        out = np.zeros((len(params), npar), dtype=np.int32)
        for i, p in enumerate(params):
            out[i, :5] = (p["t"], p["center"], p["warped"], len(p["freq"]), len(p["time"]))
            for j, (lo, hi) in enumerate(p["freq"]):
                out[i, 5 + 2 * j: 7 + 2 * j] = (lo, hi)
            for j, (lo, hi) in enumerate(p["time"]):
                out[i, 5 + 2 * MAX_MASKS + 2 * j: 7 + 2 * MAX_MASKS + 2 * j] = (lo, hi)
        return torch.from_numpy(out)

    # ------------------------------------------------------------------ device
    def apply_batch(self, xs: torch.Tensor, params: Sequence[Dict]) -> torch.Tensor:
        """xs (B, Tmax, F) fp32 CUDA, utterance b valid on rows [0, params[b]['t']).  Returns a new tensor (padding rows are
        copied through untouched: zero in the reference's collated batches)."""
        if not xs.is_cuda:
            raise RuntimeError("liteasr_b200 SpecAugment runs on CUDA tensors only (no CPU fallback)")
        xs = xs.contiguous().float()
        B, Tmax, F = xs.shape
        assert len(params) == B and all(p["t"] <= Tmax for p in params)
        out = xs.clone()
        pk = self.pack(params).to(xs.device, non_blocking=True)
        stream = C.c_void_p(torch.cuda.current_stream(xs.device).cuda_stream)
        _lib.check(_lib.lib().lasr_spec_augment(C.c_void_p(xs.data_ptr()), C.c_void_p(out.data_ptr()), C.c_int64(Tmax * F), C.c_int(F),
                                                C.c_void_p(pk.data_ptr()), C.c_int(B), C.c_int(int(bool(self.cfg.replace_with_zero))),
                                                stream), "spec_augment")
        return out

    def augment_batch(self, xs: torch.Tensor, xlens) -> torch.Tensor:
        """Draw (utterance by utterance, in batch order) and apply."""
        lens: List[int] = [int(v) for v in xlens]
        return self.apply_batch(xs, [self.draw(t, xs.shape[2]) for t in lens])

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """One (time, freq) utterance, like the reference's transform."""
        assert x.dim() == 2
        return self.apply_batch(x.unsqueeze(0), [self.draw(x.shape[0], x.shape[1])])[0]
