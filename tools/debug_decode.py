"""Developer tool (GPU): eval-mode encoder / CTC-logit error of the fp32 mode against the float64 oracle, and the oracle's
top-2 margin at every frame whose argmax differs."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from liteasr_b200 import decoding
from liteasr_b200.models.u2 import U2, U2Config
from liteasr_b200.schema import U2Dims
from liteasr_b200.utils.synthetic import synth_batch, synth_state_dict
from oracle import u2_oracle as O

case = sys.argv[1] if len(sys.argv) > 1 else "c1"
g = json.load(open(os.path.join(ROOT, "tests", "golden", f"u2_{case}.json")))
dims = U2Dims(**g["dims"])
xs, xlens, ys, ylens = synth_batch(g["batch"], g["tmax"], g["lmax"], dims.vocab_size, seed=g["seed"])
sd = synth_state_dict(dims, seed=g["seed"])
model = U2(U2Config(**{**g["dims"], "precision": "fp32"}))
model.load_state_dict(sd)
model = model.cuda().eval()
cfg = O.U2Shape(**g["dims"])
for dt in (torch.float64, torch.float32):
    sdd = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in sd.items()}
    with torch.no_grad():
        mask = O.pad_mask(xlens, xs.size(1))
        h = O.encoder(sdd, cfg, xs.to(dt), mask, training=False)
        lg = O.dense(sdd, "ctc.ctc_lo", h)
    if dt == torch.float64:
        h64, lg64 = h, lg
    else:
        print("oracle f32 vs f64: h", float((h.double() - h64).abs().max()), "logits", float((lg.double() - lg64).abs().max()))
with torch.no_grad():
    hg, lgg, Tp = decoding._encode(model, xs.cuda(), xlens.cuda())
lgg = lgg[:, : dims.vocab_size].float().cpu().view(xs.size(0), Tp, -1).double()
print("gpu fp32 vs f64: h", float((hg.cpu().double() - h64).abs().max()), "logits", float((lgg - lg64).abs().max()))
ids64 = lg64.argmax(-1)
idsg = lgg.argmax(-1)
plen = O.subsampled_len(xlens)
for b in range(xs.size(0)):
    for t in range(int(plen[b])):
        if ids64[b, t] != idsg[b, t]:
            top = lg64[b, t].topk(2).values
            print(f"flip b={b} t={t}: oracle {int(ids64[b,t])} gpu {int(idsg[b,t])} oracle top-2 margin {float(top[0]-top[1]):.3e}")
# per-layer error growth
