"""CPU: checkpoint interchange + averaging (utils/checkpoint.py:15-73 of the reference)."""
import os
import time
from types import SimpleNamespace

import pytest
import torch

from liteasr_b200.schema import U2Dims
from liteasr_b200.utils import checkpoint as C
from liteasr_b200.utils.synthetic import synth_state_dict


def _write(tmp_path, n):
    dims = U2Dims(80, 30, 32, 64, 2, 1, 32, 64, 2, 1)
    sds = []
    for i in range(n):
        sd = synth_state_dict(dims, seed=100 + i)
        for k, v in sd.items():
            if k.endswith("num_batches_tracked"):
                v.fill_(10 * (i + 1) + i)  # 10, 21, 32, ... : floor division is visible
        p = tmp_path / f"model.ep.{i + 1}.pt"
        torch.save(sd, p)
        os.utime(p, (time.time() + i, time.time() + i))  # strictly increasing modification times
        sds.append(sd)
    return sds


def test_single_checkpoint_and_schema_round_trip(tmp_path):
    sds = _write(tmp_path, 2)
    got = C.load_ckpt(SimpleNamespace(ckpt_path=str(tmp_path), ckpt_name=2, model_avg=False))
    assert got.keys() == sds[1].keys()
    assert all(torch.equal(got[k], sds[1][k]) for k in got)


def test_average_last_n(tmp_path):
    sds = _write(tmp_path, 4)
    cfg = SimpleNamespace(ckpt_path=str(tmp_path), ckpt_name=4, model_avg=True, avg_num=3, avg_policy=None)
    got = C.load_ckpt(cfg)
    for k in got:
        stack = torch.stack([sd[k] for sd in sds[1:4]])
        if got[k].is_floating_point():
            assert torch.allclose(got[k], stack.sum(0) / 3, rtol=1e-6, atol=1e-7), k
        else:
            assert torch.equal(got[k], stack.sum(0) // 3), k
    with pytest.raises(AssertionError):
        C.select_checkpoints(str(tmp_path), 2, 3)


def test_average_by_validation_loss(tmp_path):
    sds = _write(tmp_path, 4)
    log = tmp_path.parent / "train.log"
    log.write_text("\n".join(["epoch 1 valid loss: 3.5", "noise", "epoch 2 valid loss: 1.25", "epoch 3 valid loss: 2.0", "epoch 4 valid loss: 0.5"]))
    picked = C.select_checkpoints(str(tmp_path), 3, 2, str(log))  # only epochs <= 3 are eligible
    assert [os.path.basename(p) for p in picked] == ["model.ep.2.pt", "model.ep.3.pt"]
    got = C.load_ckpt(SimpleNamespace(ckpt_path=str(tmp_path), ckpt_name=3, model_avg=True, avg_num=2, avg_policy=str(log)))
    k = "encoder.after_norm.weight"
    assert torch.allclose(got[k], (sds[1][k] + sds[2][k]) / 2)
    assert C.valid_losses(str(log)) == [3.5, 1.25, 2.0, 0.5]
