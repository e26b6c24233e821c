"""TEST INFRASTRUCTURE ONLY -- pin the oracle's dropout PLACEMENT and SCALING against the UNMODIFIED reference.

    python -m oracle.make_dropout_golden          (dev container only: needs /root/reference)

The reference draws dropout masks from torch's global generator inside ``torch.nn.functional.dropout`` (reached by every
``nn.Dropout`` and by ``F.dropout`` in nets/ctc.py:29).  Here that ONE function is replaced, for the duration of a reference
run, by a deterministic stand-in: the k-th call with p > 0 multiplies its input by ``mask_k / (1 - p)`` where ``mask_k`` comes
from ``torch.Generator().manual_seed(BASE + k)``, and logs ``(shape, p)``.  Nothing of the reference is edited.  The oracle
(``oracle/u2_oracle.py``) is then run with ``SequenceDropper`` -- the same k-th-call masks -- and must reproduce the
reference's loss and every parameter gradient: that holds only if the oracle applies dropout at the same sites, in the same
order, with the same rates, on tensors of the same logical shape.  The results go to tests/golden/u2_dropout_placement.json and
tests/test_oracle_golden.py::test_dropout_placement_matches_the_unmodified_reference re-checks the oracle against them on any
machine.  Ten DISTINCT rates are used so that a rate wired to the wrong site cannot cancel out.
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from liteasr_b200.schema import U2Dims  # noqa: E402
from liteasr_b200.utils.synthetic import synth_batch, synth_state_dict  # noqa: E402

BASE = 9000
DIMS = U2Dims(80, 41, 64, 96, 2, 2, 64, 96, 2, 2)
BATCH, TMAX, LMAX, CTC_W, SMOOTH, SEED = 3, 75, 5, 0.3, 0.1, 17
RATES = dict(dropout_rate=0.11, enc_dropout_rate=0.12, enc_pos_dropout_rate=0.13, enc_attn_dropout_rate=0.14, enc_ff_dropout_rate=0.15,
             dec_dropout_rate=0.16, dec_pos_dropout_rate=0.17, dec_self_attn_dropout_rate=0.18, dec_src_attn_dropout_rate=0.19,
             dec_ff_dropout_rate=0.21)


def kth_mask(k: int, shape, p: float, dtype) -> torch.Tensor:
    g = torch.Generator().manual_seed(BASE + k)
    return (torch.rand(tuple(shape), generator=g, dtype=torch.float64) >= p).to(dtype) / (1.0 - p)


class SequenceDropper:
    """Drop-in for ``oracle.u2_oracle.Dropper``: masks by call order instead of by Philox site."""

    def __init__(self, rates, training: bool):
        self.r, self.training, self.k, self.log = rates, training, 0, []

    def __call__(self, x, net, layer, kind, p, always=False):
        if p <= 0.0 or not (self.training or always):
            return x
        m = kth_mask(self.k, x.shape, p, x.dtype)
        self.log.append([list(x.shape), p])
        self.k += 1
        return x * m


def run_reference(training: bool):
    from oracle import ref_shims
    ref_shims.install()
    import torch.nn.functional as F
    from liteasr.criterions.hybrid_ctc_attn import HybridCTCLoss, HybridCTCLossConfig
    from liteasr.models.u2 import U2, U2Config
    cfg = U2Config(**DIMS.__dict__)
    for k, v in RATES.items():
        setattr(cfg, k, v)
    model = U2(cfg)
    crit = HybridCTCLoss(HybridCTCLossConfig(vocab_size=DIMS.vocab_size, smoothing=SMOOTH, ctc_weight=CTC_W))
    model.load_state_dict(synth_state_dict(DIMS, seed=SEED), strict=True)
    model = model.double()
    model.train(training)
    xs, xlens, ys, ylens = synth_batch(BATCH, TMAX, LMAX, DIMS.vocab_size, seed=SEED)
    state = {"k": 0, "log": []}
    orig = F.dropout

    def fake(input, p=0.5, training=True, inplace=False):
        if p <= 0.0 or not training:
            return input
        m = kth_mask(state["k"], input.shape, p, input.dtype)
        state["log"].append([list(input.shape), p])
        state["k"] += 1
        return input * m

    F.dropout = fake
    try:
        if training:
            loss = crit(model, xs.double(), xlens, ys, ylens)
            loss.backward()
        else:
            with torch.no_grad():
                loss = crit(model, xs.double(), xlens, ys, ylens)
    finally:
        F.dropout = orig
    rec = dict(loss=float(loss), calls=state["log"])
    if training:
        rec["grad_l2"] = {k: float(p.grad.norm()) for k, p in model.named_parameters()}
        rec["grad_sum"] = {k: float(p.grad.sum()) for k, p in model.named_parameters()}
    return rec


def run_oracle(training: bool):
    from oracle import u2_oracle as O
    sd = synth_state_dict(DIMS, seed=SEED)
    sd64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k and ".pe.pe" not in k
                else (v.double() if v.is_floating_point() else v)) for k, v in sd.items()}
    xs, xlens, ys, ylens = synth_batch(BATCH, TMAX, LMAX, DIMS.vocab_size, seed=SEED)
    dp = SequenceDropper(O.DropRates(**RATES), training)
    out = O.hybrid_loss(sd64, O.U2Shape(**DIMS.__dict__), xs.double(), xlens, ys, ylens, CTC_W, SMOOTH, training, {}, dp)
    rec = dict(loss=float(out["loss"]), calls=dp.log)
    if training:
        out["loss"].backward()
        rec["grad_l2"] = {k: float(v.grad.norm()) for k, v in sd64.items() if getattr(v, "grad", None) is not None}
        rec["grad_sum"] = {k: float(v.grad.sum()) for k, v in sd64.items() if getattr(v, "grad", None) is not None}
    return rec


def main():
    gold = dict(dims=DIMS.__dict__, batch=BATCH, tmax=TMAX, lmax=LMAX, ctc_weight=CTC_W, smoothing=SMOOTH, seed=SEED, rates=RATES,
                base=BASE, train=run_reference(True), eval=run_reference(False))
    path = os.path.join(ROOT, "tests", "golden", "u2_dropout_placement.json")
    with open(path, "w") as f:
        json.dump(gold, f)
    o = run_oracle(True)
    print("reference loss", gold["train"]["loss"], "oracle", o["loss"], "calls", len(gold["train"]["calls"]), len(o["calls"]))
    print("eval: reference", gold["eval"]["loss"], "oracle", run_oracle(False)["loss"], "calls", gold["eval"]["calls"])
    print("wrote", path)


if __name__ == "__main__":
    main()
