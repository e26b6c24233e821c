"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's SpecAugment (utils/transform/spec_augment.py:19-125).

Two halves, like the product:
  * ``draw``  -- the random decisions, taken from Python's ``random`` and ``numpy.random`` in EXACTLY the reference's call order
                 (time_warp :34-36, freq_mask :65-70, time_mask :96-104), so that equal seeds give equal decisions;
  * ``apply`` -- the deterministic arithmetic on a (time, freq) float32 array given those decisions.

Third-party arithmetic: the time warp resizes the two halves with ``PIL.Image.resize(..., BICUBIC)`` on mode-"F" images (Pillow
12.2.0 in this container; absent from /root/reference).  ``resize_rows_bicubic`` restates Pillow's published resampling
(src/libImaging/Resample.c: ``precompute_coeffs`` + ``ImagingResampleVertical_32bpc``): filter support 2.0 scaled by
max(in/out, 1), Keys cubic with a = -0.5, window bounds ``int(center -/+ support + 0.5)`` clipped to the image, coefficients
normalised by their sum, accumulation in double, result cast to float32.  Only the vertical pass runs because the width is
unchanged.  ``tests/test_specaug_cpu.py`` pins this file against outputs of the unmodified reference class
(tests/golden/specaug.json, produced by oracle/make_golden.py).
"""
from __future__ import annotations

import random
from typing import Dict

import numpy as np


def _cubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    if x < 2.0:
        return (((x - 5.0) * x + 8.0) * x - 4.0) * a
    return 0.0


def resize_rows_bicubic(img: np.ndarray, out_rows: int) -> np.ndarray:
    """(in_rows, cols) float32 -> (out_rows, cols) float32, Pillow BICUBIC along the row axis."""
    in_rows = img.shape[0]
    scale = in_rows / out_rows
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ss = 1.0 / filterscale
    out = np.empty((out_rows, img.shape[1]), dtype=np.float32)
    src = img.astype(np.float64)
    for yy in range(out_rows):
        center = (yy + 0.5) * scale
        ymin = max(int(center - support + 0.5), 0)
        ymax = min(int(center + support + 0.5), in_rows) - ymin
        w = np.array([_cubic((k + ymin - center + 0.5) * ss) for k in range(ymax)], dtype=np.float64)
        w /= w.sum()
        out[yy] = (src[ymin:ymin + ymax] * w[:, None]).sum(0).astype(np.float32)
    return out


def draw(t: int, f: int, cfg) -> Dict:
    """The reference's random decisions for one (t, f) utterance, in its RNG call order."""
    p = dict(center=-1, warped=-1, freq=[], time=[])
    window = cfg.time_warp
    if t - window > window:
        center = random.randrange(window, t - window)
        p["center"], p["warped"] = center, random.randrange(center - window, center + window) + 1
    fs = np.random.randint(0, cfg.freq_mask, size=(cfg.freq_mask_times, 2))
    for width, end in fs:
        f_zero = random.randrange(0, f - width)
        if width == 0:
            continue
        p["freq"].append((int(f_zero), int(end + f_zero)))   # [f_zero, f_zero + second draw): the reference's quirk (:71)
    ts = np.random.randint(0, cfg.time_mask, size=(cfg.time_mask_times, 2))
    for width, end in ts:
        if t - width <= 0:
            continue
        t_zero = random.randrange(0, t - width)
        if width == 0:
            continue
        p["time"].append((int(t_zero), int(end + t_zero)))
    return p


def apply(x: np.ndarray, p: Dict, replace_with_zero: bool = False) -> np.ndarray:
    x = np.array(x, dtype=np.float32, copy=True)
    t = x.shape[0]
    if p["center"] >= 0:
        c, w = p["center"], p["warped"]
        x = np.concatenate((resize_rows_bicubic(x[:c], w), resize_rows_bicubic(x[c:], t - w)), 0)
    for lo, hi in p["freq"]:
        x[:, lo:hi] = 0 if replace_with_zero else x.mean()
    for lo, hi in p["time"]:
        x[lo:hi] = 0 if replace_with_zero else x.mean()
    return x


def spec_augment(x: np.ndarray, cfg) -> np.ndarray:
    return apply(x, draw(x.shape[0], x.shape[1], cfg), cfg.replace_with_zero)
