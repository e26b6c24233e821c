"""CPU: pin oracle/u2_oracle.py (+ ctc_oracle.c) against fixtures generated from the UNMODIFIED
reference (oracle/make_golden.py -> tests/golden/*.json)."""
import json
import math
import os

import numpy as np
import pytest
import torch

from liteasr_b200.schema import U2Dims
from liteasr_b200.utils.synthetic import synth_batch, synth_state_dict
from oracle import u2_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def check_summary(t, rec, rtol, atol):
    f = t.detach().double().reshape(-1)
    assert list(t.shape) == rec["shape"]
    got = torch.stack([f[i] for i in rec["idx"]])
    want = torch.tensor(rec["val"], dtype=torch.float64)
    assert torch.allclose(got, want, rtol=rtol, atol=atol), (got - want).abs().max()
    assert math.isclose(float(f.norm()), rec["l2"], rel_tol=max(rtol, 1e-9), abs_tol=atol)


@pytest.mark.parametrize("case", ["tiny", "tiny_odd", "c1"])
def test_u2_oracle_matches_reference_f64(case):
    g = load(f"u2_{case}.json")
    dims = U2Dims(**g["dims"])
    xs, xlens, ys, ylens = synth_batch(g["batch"], g["tmax"], g["lmax"], dims.vocab_size, seed=g["seed"])
    assert xlens.tolist() == g["xlens"] and ylens.tolist() == g["ylens"]
    sd32 = synth_state_dict(dims, seed=g["seed"])
    sd = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k and ".pe.pe" not in k
              else (v.double() if v.is_floating_point() else v)) for k, v in sd32.items()}
    cfg = O.U2Shape(**g["dims"])
    bn = {}
    out = O.hybrid_loss(sd, cfg, xs.double(), xlens, ys, ylens, g["ctc_weight"], g["smoothing"], True, bn)
    ref = g["f64"]
    assert math.isclose(float(out["loss"]), ref["loss"], rel_tol=1e-10)
    assert math.isclose(float(out["loss_ctc"]), ref["loss_ctc"], rel_tol=1e-9)
    assert math.isclose(float(out["loss_attn"]), ref["loss_attn"], rel_tol=1e-9)
    check_summary(out["h_attn"], ref["h_attn"], 1e-9, 1e-10)
    check_summary(out["h_ctc"], ref["h_ctc"], 1e-9, 1e-10)
    out["loss"].backward()
    for k, rec in ref["grads"].items():
        assert sd[k].grad is not None, k
        check_summary(sd[k].grad, rec, 1e-7, 1e-9)
    for k, rec in ref["bn"].items():
        check_summary(bn[k], rec, 1e-9, 1e-12)
    for k, v in ref["nbt"].items():
        assert int(bn[k]) == v
    # eval-mode greedy CTC after the running-stat update (token ids bit-exact)
    sd_eval = {k: v.detach() for k, v in sd.items()}
    sd_eval.update(bn)
    toks, _ = O.greedy_ctc(sd_eval, cfg, xs.double(), xlens)
    assert toks == ref["greedy"]


def test_u2_oracle_fp32_loss_close():
    g = load("u2_c1.json")
    dims = U2Dims(**g["dims"])
    xs, xlens, ys, ylens = synth_batch(g["batch"], g["tmax"], g["lmax"], dims.vocab_size, seed=g["seed"])
    sd = synth_state_dict(dims, seed=g["seed"])
    with torch.no_grad():
        out = O.hybrid_loss(sd, O.U2Shape(**g["dims"]), xs, xlens, ys, ylens, g["ctc_weight"], g["smoothing"])
    assert math.isclose(float(out["loss"]), g["f32"]["loss"], rel_tol=2e-6)


def test_ctc_restatements_match_torch_ctcloss():
    g = load("ctc_golden.json")
    for c in g["cases"]:
        logits = torch.tensor(c["logits"], dtype=torch.float64)
        lp = logits.log_softmax(-1).numpy()
        tg = np.array(c["targets"], dtype=np.int64)
        il = np.array(c["in_len"], dtype=np.int64)
        tl = np.array(c["tgt_len"], dtype=np.int64)
        for fn in (O.ctc_alpha_beta_numpy, O.ctc_alpha_beta_c):
            nll, dlp = fn(lp, np.clip(tg, 0, None), il, tl)
            for b, want in enumerate(c["nll"]):
                if want == "inf":
                    assert np.isinf(nll[b])
                else:
                    assert math.isclose(nll[b], want, rel_tol=1e-12, abs_tol=1e-12)
            if c["grad_logits"] is not None:
                # compose with log-softmax backward: dlogits = dlp - softmax * sum_c dlp
                fin = np.array(c["finite"])
                d = np.where(np.isnan(dlp), 0.0, dlp) * fin[None, :, None]
                sm = np.exp(lp)
                dl = d - sm * d.sum(-1, keepdims=True)
                assert np.allclose(dl, np.array(c["grad_logits"]), rtol=1e-10, atol=1e-12)


def test_rel_shift_closed_form_matches_pad_view_trick():
    # nets/attention.py:99-118 restated literally here as the check
    for t in (1, 2, 5, 7, 31):
        x = torch.randn(2, 3, t, t, dtype=torch.float64)
        zp = torch.zeros(2, 3, t, 1, dtype=torch.float64)
        xp = torch.cat([zp, x], -1).view(2, 3, t + 1, t)[:, :, 1:].reshape(2, 3, t, t)
        assert torch.equal(O.rel_shift(x), xp)


def test_mask_docstring_examples():
    # utils/mask.py:17-20 and :47-52
    m = O.pad_mask(torch.tensor([5, 3, 1]))
    assert m.int().tolist() == [[0, 0, 0, 0, 0], [0, 0, 0, 1, 1], [0, 1, 1, 1, 1]]
    c = O.causal_mask(5).int().tolist()
    assert c == [[0, 1, 1, 1, 1], [0, 0, 1, 1, 1], [0, 0, 0, 1, 1], [0, 0, 0, 0, 1], [0, 0, 0, 0, 0]]


@pytest.mark.parametrize("case,nutt", [("tiny", 3), ("tiny_odd", 2), ("c1", 1)])
def test_oracle_inference_matches_reference(case, nutt):
    """Batch-1 CTC prefix beam search + attention rescoring (models/u2.py:221-317) and greedy CTC: the oracle restatement
    reproduces the unmodified reference's n-best lists (prefixes bit-exact, scores to 1e-9) and its final hypothesis."""
    g = load(f"u2_{case}.json")
    dims = U2Dims(**g["dims"])
    xs, xlens, _, _ = synth_batch(g["batch"], g["tmax"], g["lmax"], dims.vocab_size, seed=g["seed"])
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in synth_state_dict(dims, seed=g["seed"]).items()}
    cfg = O.U2Shape(**g["dims"])
    with torch.no_grad():
        for rec in g["f64"]["inference"][:nutt]:
            i = rec["utt"]
            x = xs[i:i + 1, : rec["frames"]].double()
            out = O.attention_rescore(sd, cfg, x)
            assert [p for p, _ in out["hyps"]] == [p for p, _ in rec["hyps"]]
            assert np.allclose([s for _, s in out["hyps"]], [s for _, s in rec["hyps"]], rtol=1e-9, atol=1e-9)
            assert out["best"] == rec["best"]
            greedy, _ = O.greedy_ctc(sd, cfg, x, None)
            assert greedy[0] == rec["greedy"]


def test_log_add_and_prefix_search_small_known_answer():
    """Hand-checkable lattice: 2 frames, 3 classes (blank 0).  p(prefix) sums every alignment collapsing to it."""
    lp = torch.log(torch.tensor([[0.6, 0.3, 0.1], [0.5, 0.4, 0.1]], dtype=torch.float64))
    hyps = dict((p, s) for p, s in O.prefix_beam_search_logp(lp, 3))
    # () : blank,blank = .30 ; (1,): 1b + b1 + 11 = .15 + .24 + .12 = .51 ; (2,): .05 + .06 + .01 = .12 ; (1,2): .03 ; (2,1): .04
    assert math.isclose(math.exp(hyps[()]), 0.30, rel_tol=1e-12)
    assert math.isclose(math.exp(hyps[(1,)]), 0.51, rel_tol=1e-12)
    assert math.isclose(math.exp(hyps[(2,)]), 0.12, rel_tol=1e-12)
    assert math.isclose(O.log_add([math.log(0.25), math.log(0.75)]), 0.0, abs_tol=1e-15)
    assert O.log_add([-float("inf"), -float("inf")]) == -float("inf")


def test_dropout_placement_matches_the_unmodified_reference():
    """tests/golden/u2_dropout_placement.json was produced by the UNMODIFIED reference with ``torch.nn.functional.dropout``
    swapped for a call-order-deterministic stand-in (oracle/make_dropout_golden.py).  The oracle with the same call-order masks
    must give the same loss, the same 30 dropout calls (shape and rate, in order) and the same gradients: sites, order, rates,
    1/(1-p) scaling and the always-on CTC-head site (quirk Q3, eval record) are thereby pinned."""
    import json
    from oracle import make_dropout_golden as G
    g = json.load(open(os.path.join(GOLDEN, "u2_dropout_placement.json")))
    assert g["rates"] == G.RATES and g["dims"] == G.DIMS.__dict__ and g["base"] == G.BASE
    tr = G.run_oracle(True)
    assert tr["calls"] == g["train"]["calls"] and len(tr["calls"]) == 30
    assert math.isclose(tr["loss"], g["train"]["loss"], rel_tol=1e-12)
    for k, v in g["train"]["grad_l2"].items():
        assert math.isclose(tr["grad_l2"][k], v, rel_tol=1e-9, abs_tol=1e-12), k
        assert math.isclose(tr["grad_sum"][k], g["train"]["grad_sum"][k], rel_tol=1e-7, abs_tol=1e-10), k
    ev = G.run_oracle(False)
    assert ev["calls"] == g["eval"]["calls"] == [[[3, 18, 64], 0.11]]   # eval(): only F.dropout of nets/ctc.py:29 is live
    assert math.isclose(ev["loss"], g["eval"]["loss"], rel_tol=1e-12)
