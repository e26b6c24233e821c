"""Developer tool (GPU): which prior use of a model breaks the CUDA-graph capture of TrainStep."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from liteasr_b200.criterions.hybrid_ctc_attn import HybridCTCLoss, HybridCTCLossConfig
from liteasr_b200.models.u2 import U2, U2Config
from liteasr_b200.schema import U2Dims
from liteasr_b200.trainer import TrainStep
from liteasr_b200.utils.synthetic import synth_batch, synth_state_dict

g = json.load(open(os.path.join(ROOT, "tests", "golden", "u2_tiny.json")))
dims = U2Dims(**g["dims"])
dev = torch.device("cuda:0")
batch = tuple(t.to(dev) for t in synth_batch(g["batch"], g["tmax"], g["lmax"], dims.vocab_size, seed=g["seed"]))
crit = HybridCTCLoss(HybridCTCLossConfig(vocab_size=dims.vocab_size, smoothing=g["smoothing"], ctc_weight=g["ctc_weight"]))
for variant in sys.argv[1:]:
    if variant.startswith("two"):
        sd = synth_state_dict(dims, seed=g["seed"])
        for precision in ("bf16", "fp32"):
            model = U2(U2Config(**g["dims"], precision=precision))
            if variant != "two_noload":
                model.load_state_dict(sd)
            model = model.to(dev).train()
            loss = crit(model, *batch)
            loss.backward()
            torch.cuda.synchronize()
            if variant == "two_grad":
                _ = dict(model.named_parameters())["encoder.enc_layers.0.feed_forward.fc1.weight"].grad.double().cpu()
        try:
            step = TrainStep(model, crit, device=dev)
            print(variant, "ok", float(step(*batch)))
        except Exception as e:  # noqa: BLE001
            print(variant, "FAILED", str(e).splitlines()[0])
            if os.environ.get("LASR_DEBUG_TRACE"):
                import traceback
                traceback.print_exc()
        continue
    model = U2(U2Config(**g["dims"], precision="fp32")).to(dev).train()
    if variant in ("fwd", "bwd", "bwd_none"):
        loss = crit(model, *batch)
        if variant != "fwd":
            loss.backward()
        if variant == "bwd_none":
            model.zero_grad(set_to_none=True)
        del loss
        torch.cuda.synchronize()
    try:
        step = TrainStep(model, crit, device=dev)
        l0 = float(step(*batch))
        print(variant, "ok", l0)
    except Exception as e:  # noqa: BLE001
        print(variant, "FAILED", str(e).splitlines()[0])
        break
