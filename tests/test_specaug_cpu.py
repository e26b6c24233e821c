"""CPU: the SpecAugment oracle (oracle/specaug_oracle.py, incl. its restatement of Pillow's BICUBIC resize) against outputs of
the unmodified reference class under the same seeds (tests/golden/specaug.json, made by oracle/make_golden.py)."""
import json
import os
import random
from types import SimpleNamespace

import numpy as np
import torch

from oracle import specaug_oracle as S

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "specaug.json")


def test_oracle_reproduces_reference_outputs_bit_exactly():
    g = json.load(open(GOLDEN))
    assert len(g["cases"]) >= 6
    for rec in g["cases"]:
        cfg = SimpleNamespace(**rec["cfg"])
        x = (torch.randn(rec["T"], rec["F"], generator=torch.Generator().manual_seed(rec["seed"])) * 3 + 1).numpy()
        random.seed(rec["seed"])
        np.random.seed(rec["seed"])
        y = S.spec_augment(x, cfg)
        flat = y.reshape(-1)
        assert flat[:16].tolist() == rec["head"]
        assert flat[:: max(1, flat.size // 61)].tolist() == rec["sample"]
        assert float(np.float64(y).sum()) == rec["sum"] and float(np.abs(np.float64(y)).sum()) == rec["abs_sum"]
        if "full" in rec:
            assert np.array_equal(y, np.array(rec["full"], dtype=np.float32))


def test_bicubic_identity_and_constant():
    x = np.random.default_rng(0).normal(size=(37, 5)).astype(np.float32)
    assert np.array_equal(S.resize_rows_bicubic(x, 37), x)                      # same size: centre tap only
    c = np.full((20, 3), 2.5, dtype=np.float32)
    assert np.allclose(S.resize_rows_bicubic(c, 33), 2.5) and np.allclose(S.resize_rows_bicubic(c, 7), 2.5)
