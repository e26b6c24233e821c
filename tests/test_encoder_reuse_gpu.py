"""GPU: the accelerated encoder as a stand-alone module inside ANOTHER model (the reference's Transducer and
Paraformer build ``TransformerEncoder`` with exactly these keyword arguments, models/transducer.py:87-100 and
models/paraformer.py:70-83, and put their own torch decoders behind it).  A host model owns ``self.encoder`` plus a plain
torch head; forward through the reference signature ``encoder(xs, mask)`` and autograd through the head + the hand-written
encoder backward must match the float64 oracle encoder with the same head.  Tolerances: fp32 rel-L2 <= 1e-5 on the output,
<= 1e-4 on gradients; bf16 rel-L2 <= 2e-2 / 5e-2."""
import json
import os

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


class HostModel(nn.Module):
    """Stands in for Transducer / Paraformer: reference constructor call of the encoder + a torch head (joint / predictor)."""

    def __init__(self, d, precision):
        super().__init__()
        from liteasr_b200 import functions as F
        from liteasr_b200.nets.transformer_encoder import TransformerEncoder  # the reference import path (models/transducer.py:21)
        self.encoder = TransformerEncoder(use_rel=True, i_dim=d["input_dim"], h_dim=d["enc_dim"], ff_dim=d["enc_ff_dim"],
                                          n_head=d["enc_attn_heads"], n_layer=d["enc_layers"], dropout_rate=0.0, pos_dropout_rate=0.0,
                                          attn_dropout_rate=0.0, ff_dropout_rate=0.0, activation="swish", arch="conformer")
        self.head = nn.Linear(d["enc_dim"], 7)
        F.set_precision(self, precision)


@pytest.mark.parametrize("precision,out_tol,grad_tol", [("fp32", 1e-5, 1e-4), ("bf16", 2e-2, 5e-2)])
def test_encoder_inside_another_model_matches_oracle(precision, out_tol, grad_tol):
    from liteasr_b200.schema import U2Dims
    from liteasr_b200.utils.mask import padding_mask
    from liteasr_b200.utils.synthetic import synth_batch, synth_state_dict
    from oracle import u2_oracle as O
    g = json.load(open(os.path.join(GOLDEN, "u2_tiny.json")))
    dims = U2Dims(**g["dims"])
    xs, xlens, _, _ = synth_batch(g["batch"], g["tmax"], g["lmax"], dims.vocab_size, seed=g["seed"])
    sd = synth_state_dict(dims, seed=g["seed"])
    enc_sd = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}

    torch.manual_seed(0)
    model = HostModel(g["dims"], precision)
    model.encoder.load_state_dict(enc_sd, strict=True)
    model = model.cuda().train()
    mask = padding_mask(xlens).cuda()
    assert mask.dtype == torch.bool and bool(mask[-1, -1])  # True = padding (utils/mask.py:8-27)
    h = model.encoder(xs.cuda(), mask)  # reference signature: bool (B,T) mask, True = padding
    out = model.head(h)
    w = torch.linspace(-1.0, 1.0, out.numel(), device="cuda").view_as(out)
    (out * w).sum().backward()

    # float64 oracle encoder + the same head
    sd64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k and ".pe.pe" not in k
                else (v.double() if v.is_floating_point() else v)) for k, v in sd.items()}
    xs_mask = torch.arange(xs.size(1))[None, :] >= xlens[:, None]
    h_ref = O.encoder(sd64, O.U2Shape(**g["dims"]), xs.double(), xs_mask, True, {})
    hw, hb = model.head.weight.detach().double().cpu(), model.head.bias.detach().double().cpu()
    out_ref = h_ref @ hw.t() + hb
    (out_ref * w.double().cpu()).sum().backward()

    rel = float((h.detach().double().cpu() - h_ref.detach()).norm() / h_ref.detach().norm())
    assert rel <= out_tol, rel
    checked = 0
    for n, p in model.encoder.named_parameters():
        gr = sd64["encoder." + n].grad
        if gr is None or float(gr.abs().max()) < 1e-9:
            continue
        r = float((p.grad.double().cpu() - gr).norm() / gr.norm())
        if n.endswith("feed_forward.fc1.weight") or n.endswith("linear_q.weight") or n.endswith("embed.conv.0.weight") or n.endswith("after_norm.weight"):
            assert r <= grad_tol, (n, r)
            checked += 1
    assert checked >= 4
    assert model.head.weight.grad is not None and torch.isfinite(model.head.weight.grad).all()
