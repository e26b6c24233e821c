"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.json from the UNMODIFIED reference.

Run in the dev container (the only place /root/reference exists):
    python -m oracle.make_golden
It (1) builds the reference ``U2`` + ``HybridCTCLoss`` through ``oracle/ref_shims.py``,
(2) loads the deterministic synthetic weights / batch from ``liteasr_b200.utils.synthetic``,
(3) runs forward+backward in float64 (and float32 for the record) and (4) stores losses,
output samples, per-parameter gradient norms/samples, BatchNorm running-stat updates and
eval-mode greedy-CTC token ids.  It also stores torch.nn.CTCLoss known answers
(the third-party arithmetic behind criterions/hybrid_ctc_attn.py:32) for the CTC restatement.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from liteasr_b200.schema import U2Dims  # noqa: E402
from liteasr_b200.utils.synthetic import synth_batch, synth_state_dict  # noqa: E402
from oracle import ref_shims  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

MODEL_CASES = {
    # name: (dims, batch, tmax, lmax, ctc_weight, smoothing, seed)
    "tiny": (U2Dims(80, 50, 128, 256, 2, 2, 128, 256, 2, 2), 3, 67, 6, 0.3, 0.1, 7),
    "tiny_odd": (U2Dims(80, 37, 64, 96, 1, 1, 64, 96, 1, 1), 2, 43, 4, 0.5, 0.0, 11),
    "c1": (U2Dims(80, 500, 256, 2048, 4, 4, 256, 2048, 4, 6), 8, 500, 30, 0.3, 0.1, 42),
}


def sample_idx(n: int, k: int = 16):
    if n <= k:
        return list(range(n))
    return [int(i) for i in np.linspace(0, n - 1, k).astype(np.int64)]


def summarize(t: torch.Tensor, k: int = 16):
    f = t.detach().double().reshape(-1)
    idx = sample_idx(f.numel(), k)
    return dict(shape=list(t.shape), sum=float(f.sum()), abssum=float(f.abs().sum()), l2=float(f.norm()),
                idx=idx, val=[float(f[i]) for i in idx])


def greedy(ids_row, n):
    out, prev = [], -1
    for t in range(n):
        c = int(ids_row[t])
        if c != prev and c != 0:
            out.append(c)
        prev = c
    return out


def run_model_case(name):
    dims, b, tmax, lmax, w, eps, seed = MODEL_CASES[name]
    xs, xlens, ys, ylens = synth_batch(b, tmax, lmax, dims.vocab_size, seed=seed)
    sd = synth_state_dict(dims, seed=seed)
    out = dict(case=name, dims=dims.__dict__, batch=b, tmax=tmax, lmax=lmax, ctc_weight=w, smoothing=eps, seed=seed,
               xlens=xlens.tolist(), ylens=ylens.tolist())
    g64 = {}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        model, crit = ref_shims.build_reference(dims.__dict__, eps, w)
        model.load_state_dict(sd, strict=True)
        model = model.to(dtype)
        model.train()
        loss = crit(model, xs.to(dtype), xlens, ys, ylens)
        loss.backward()
        rec = dict(loss=float(loss))
        if tag == "f32":
            # the reference's OWN fp32 deviation from its fp64 run, per parameter (max-abs): fp32 ReLU/argmax flips and
            # summation order put a floor under any fp32 implementation's distance to the fp64 truth
            rec["grad_maxabs_err_vs_f64"] = {k: float((p.grad.double() - g64[k]).abs().max()) for k, p in model.named_parameters()}
        if tag == "f64":
            g64 = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
            # recompute pieces for the record
            with torch.no_grad():
                model2, _ = ref_shims.build_reference(dims.__dict__, eps, w)
                model2.load_state_dict(sd, strict=True)
                model2 = model2.to(dtype).train()
                h_attn, h_ctc = model2(xs.to(dtype), xlens, ys, ylens)
                h_enc = None
            rec["h_attn"] = summarize(h_attn)
            rec["h_ctc"] = summarize(h_ctc)
            rec["grads"] = {k: summarize(p.grad, 6) for k, p in model.named_parameters()}
            rec["bn"] = {k: summarize(v, 6) for k, v in model.state_dict().items() if ".conv.norm.running" in k}
            rec["nbt"] = {k: int(v) for k, v in model.state_dict().items() if k.endswith("num_batches_tracked")}
            # split losses
            import liteasr.criterions.hybrid_ctc_attn as H  # noqa
            lp = h_ctc.transpose(0, 1).log_softmax(-1)
            lctc = torch.nn.functional.ctc_loss(lp, ys, model.get_pred_len(xlens), ylens, reduction="sum") / b
            rec["loss_ctc"] = float(lctc)
            rec["loss_attn"] = float((float(loss) - w * float(lctc)) / (1 - w))
        # eval-mode greedy CTC (encoder with mask, running BN stats)
        model.eval()
        with torch.no_grad():
            from liteasr.utils.mask import padding_mask
            h = model.encoder(xs.to(dtype), mask=padding_mask(xlens))
            lp = model.ctc.log_softmax(h)
            ids = lp.argmax(-1)
            plen = model.get_pred_len(xlens)
            rec["greedy"] = [greedy(ids[i], int(plen[i])) for i in range(b)]
            rec["eval_h_enc"] = summarize(h)
            # batch-1 inference exactly as `liteasr-infer` runs it (infer.py:97-120 -> models/u2.py:160-161,221-317): maskless
            # encoder on ONE utterance, CTC prefix beam search (beam 10), attention rescoring.  The reference crashes in
            # attention_rescore at HEAD (quirk Q13: Python lists reach `ylens + 1` / padding_mask); `_preprocess` is wrapped so the
            # two length lists become tensors -- nothing else is touched.
            # A FRESH model (pristine BatchNorm running statistics: the training forward above updated `model`'s).
            minf, _ = ref_shims.build_reference(dims.__dict__, eps, w)
            minf.load_state_dict(sd, strict=True)
            minf = minf.to(dtype).eval()
            orig_pre = minf._preprocess

            def pre(xs, xlens, ys, ylens, _orig=orig_pre):
                return _orig(xs, torch.as_tensor(xlens), ys, torch.as_tensor(ylens))

            minf._preprocess = pre
            inf = []
            for i in range(min(b, 3)):
                xi = xs[i:i + 1, : int(xlens[i])].to(dtype)
                hyps, _ = minf._ctc_prefix_beam_search(xi)
                best = minf.attention_rescore(xi)
                ids1 = minf.ctc.log_softmax(minf.encoder(xi)).argmax(-1)[0]
                inf.append(dict(utt=i, frames=int(xlens[i]), hyps=[[list(p), float(sc)] for p, sc in hyps], best=list(best),
                                greedy=greedy(ids1, ids1.numel())))
            rec["inference"] = inf
        out[tag] = rec
    return out


def ctc_cases():
    """Known answers from torch.nn.CTCLoss(reduction='none') + log_softmax backward."""
    g = torch.Generator().manual_seed(1234)
    cases = []
    specs = [
        # T, B, V, targets(list of lists), in_len
        (5, 1, 4, [[1, 2]], [5]),
        (6, 2, 5, [[1, 1, 2], [3]], [6, 4]),           # repeated label needs a blank between
        (4, 2, 3, [[1, 1], [2, 2]], [3, 4]),            # first utt: T=3 < 2L-ish -> exactly feasible (1 _ 1)
        (3, 1, 3, [[1, 1]], [2]),                       # infeasible -> inf
        (7, 3, 6, [[], [5, 4, 3], [2]], [7, 7, 1]),     # empty target, len-1 input
        (12, 2, 9, [[8, 7, 7, 1], [1, 2, 3, 4, 5]], [12, 10]),
    ]
    for T, B, V, tg, il in specs:
        logits = torch.randn(T, B, V, generator=g, dtype=torch.float64)
        logits.requires_grad_(True)
        lmax = max(1, max(len(t) for t in tg))
        tgt = torch.full((B, lmax), -1, dtype=torch.long)
        for i, t in enumerate(tg):
            tgt[i, : len(t)] = torch.tensor(t, dtype=torch.long)
        tl = torch.tensor([len(t) for t in tg])
        ilen = torch.tensor(il)
        lp = logits.log_softmax(-1)
        nll = torch.nn.functional.ctc_loss(lp, tgt.clamp(min=0), ilen, tl, reduction="none", zero_infinity=False)
        finite = torch.isfinite(nll)
        (nll * finite).sum().backward() if finite.any() else None
        cases.append(dict(T=T, B=B, V=V, logits=logits.detach().tolist(), targets=tgt.tolist(), in_len=il,
                          tgt_len=tl.tolist(), nll=[float(x) if np.isfinite(float(x)) else "inf" for x in nll],
                          grad_logits=(logits.grad.tolist() if logits.grad is not None else None),
                          finite=finite.tolist()))
    return cases


SPECAUG_CASES = [
    # seed, T, F, time_warp, freq_mask, freq_mask_times, time_mask, time_mask_times, replace_with_zero, store full output
    (0, 40, 12, 5, 6, 2, 10, 2, False, True),
    (1, 200, 80, 80, 27, 1, 100, 1, False, False),   # reference defaults (config/__init__.py:43-51)
    (2, 170, 80, 80, 27, 2, 100, 2, False, False),
    (3, 300, 40, 30, 10, 2, 40, 3, True, False),
    (4, 100, 80, 80, 27, 1, 100, 1, False, False),   # too short for the warp (t - window <= window)
    (5, 161, 23, 80, 8, 2, 50, 2, False, True),      # shortest warpable length: extreme scale factors
    (6, 500, 80, 80, 27, 2, 100, 2, False, False),
]


def specaug_cases():
    """Outputs of the UNMODIFIED reference SpecAugment (utils/transform/spec_augment.py) under seeded `random` / `numpy.random`
    on seeded inputs (torch.randn(T, F, generator=seed) * 3 + 1)."""
    import importlib
    import random
    from types import SimpleNamespace
    ref_shims.install()
    mod = importlib.import_module("liteasr.utils.transform.spec_augment")
    import PIL
    out = []
    for seed, T, F, tw, fm, fmt, tm, tmt, rz, full in SPECAUG_CASES:
        cfg = SimpleNamespace(time_warp=tw, freq_mask=fm, freq_mask_times=fmt, time_mask=tm, time_mask_times=tmt, inplace=True,
                              replace_with_zero=rz)
        x = torch.randn(T, F, generator=torch.Generator().manual_seed(seed)) * 3 + 1
        random.seed(seed)
        np.random.seed(seed)
        y = mod.SpecAugment(cfg)(x.clone())
        rec = dict(seed=seed, T=T, F=F, cfg=vars(cfg), sum=float(y.double().sum()), abs_sum=float(y.double().abs().sum()),
                   head=y.flatten()[:16].tolist(), sample=y.flatten()[:: max(1, y.numel() // 61)].tolist())
        if full:
            rec["full"] = y.tolist()
        out.append(rec)
    return dict(pillow=PIL.__version__, numpy=np.__version__, cases=out)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name in MODEL_CASES:
        rec = run_model_case(name)
        with open(os.path.join(GOLDEN_DIR, f"u2_{name}.json"), "w") as f:
            json.dump(rec, f)
        print(name, "loss f64", rec["f64"]["loss"], "f32", rec["f32"]["loss"])
    with open(os.path.join(GOLDEN_DIR, "ctc_golden.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, cases=ctc_cases()), f)
    with open(os.path.join(GOLDEN_DIR, "specaug.json"), "w") as f:
        json.dump(specaug_cases(), f)
    print("golden written to", GOLDEN_DIR)


if __name__ == "__main__":
    main()
