"""Developer tool (GPU): stage-by-stage parity of the CUDA path against the CPU oracle on a golden case."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from liteasr_b200 import functions as F  # noqa: E402
from liteasr_b200.criterions.hybrid_ctc_attn import HybridCTCLoss, HybridCTCLossConfig  # noqa: E402
from liteasr_b200.models.u2 import U2, U2Config  # noqa: E402
from liteasr_b200.schema import U2Dims  # noqa: E402
from liteasr_b200.utils.synthetic import synth_batch, synth_state_dict  # noqa: E402
from oracle import u2_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30)), float((a - b).abs().max())


def main(case="tiny", precision="fp32"):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", f"u2_{case}.json")))
    dims = U2Dims(**g["dims"])
    xs, xlens, ys, ylens = synth_batch(g["batch"], g["tmax"], g["lmax"], dims.vocab_size, seed=g["seed"])
    sd = synth_state_dict(dims, seed=g["seed"])
    cfg = U2Config(**{**g["dims"], "precision": precision})
    model = U2(cfg)
    model.load_state_dict(sd)
    model = model.cuda().train()
    crit = HybridCTCLoss(HybridCTCLossConfig(vocab_size=dims.vocab_size, smoothing=g["smoothing"], ctc_weight=g["ctc_weight"]))
    dxs, dxl, dys, dyl = xs.cuda(), xlens.cuda(), ys.cuda(), ylens.cuda()

    # oracle (float64)
    sd64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k and ".pe.pe" not in k
                else (v.double() if v.is_floating_point() else v)) for k, v in sd.items()}
    ocfg = O.U2Shape(**g["dims"])
    bn = {}
    out = O.hybrid_loss(sd64, ocfg, xs.double(), xlens, ys, ylens, g["ctc_weight"], g["smoothing"], True, bn)
    out["loss"].backward()

    # staged forward (no grad)
    with torch.no_grad():
        h_attn, h_ctc = model(dxs, dxl, dys, dyl)
    print("h_ctc  rel/max", rel(h_ctc.float(), out["h_ctc"]))
    print("h_attn rel/max", rel(h_attn.float(), out["h_attn"]))
    # reset BN stats changed by that forward
    model.load_state_dict(sd)
    loss = crit(model, dxs, dxl, dys, dyl)
    loss.backward()
    torch.cuda.synchronize()
    print("loss", float(loss), "oracle", float(out["loss"]), "golden", g["f64"]["loss"], "parts", model.last_losses.tolist())
    rows = []
    for n, p in model.named_parameters():
        r, m = rel(p.grad, sd64[n].grad)
        rows.append((r, m, n))
    rows.sort(reverse=True)
    for r, m, n in rows[:25]:
        print(f"{r:10.3e} {m:10.3e} {n}")
    print("median rel", sorted(r for r, _, _ in rows)[len(rows) // 2])
    for k, v in bn.items():
        if "running" in k:
            r, m = rel(model.state_dict()[k], v)
            if r > 1e-5:
                print("BN", k, r, m)


if __name__ == "__main__":
    main(*(sys.argv[1:]))
