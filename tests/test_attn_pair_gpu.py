"""GPU: the paired attention-backward contractions (csrc/attn_pair.cu, include/lasr.h ``lasr_attn_bwd_pair``) against torch fp32
matmuls of the same bf16 operands -- the arithmetic autograd performs for matrix_ac / matrix_bd of nets/attention.py:120-154:
dl = X . r, dr = X^T . l per (utterance, head), dr optionally summed over the batch (the shared positional projection)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _case(B, H, Tq, Tk, ld, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    d = 64 * H
    x = torch.full((B, H, Tq, ld), float("nan"), dtype=torch.bfloat16)
    x[..., :Tk] = (torch.randn(B, H, Tq, Tk, generator=g) * 0.5).bfloat16()   # padding columns hold NaN: they must not be read
    wide = torch.randn(B * Tk, 3 * d, generator=g).bfloat16()                 # r = a column slice of a wider matrix (like k in qkv)
    l = torch.randn(B * Tq, d, generator=g).bfloat16()
    pos = torch.randn(Tk, d, generator=g).bfloat16()
    return [t.to(DEV) for t in (x, wide, l, pos)]


def _ref(x, r, l, B, H, Tq, Tk, batched):
    xf = x[..., :Tk].float()
    rf = (r.float().view(B, Tk, H, 64) if batched else r.float().view(1, Tk, H, 64).expand(B, Tk, H, 64)).permute(0, 2, 1, 3)
    lf = l.float().view(B, Tq, H, 64).permute(0, 2, 1, 3)
    dl = torch.matmul(xf, rf).permute(0, 2, 1, 3).reshape(B * Tq, H * 64)
    dr = torch.matmul(xf.transpose(-1, -2), lf).permute(0, 2, 1, 3)           # (B, Tk, H, 64)
    return dl, dr


SHAPES = [(3, 4, 299, 299, 320), (2, 2, 127, 127, 128), (2, 4, 320, 320, 320), (2, 1, 64, 64, 64), (3, 2, 1, 1, 8), (2, 2, 200, 200, 200),
          (2, 4, 41, 299, 304), (5, 4, 130, 257, 264)]


@pytest.mark.parametrize("B,H,Tq,Tk,ld", SHAPES)
def test_pair_per_utterance(B, H, Tq, Tk, ld):
    from liteasr_b200 import ops
    assert ops.attn_bwd_pair_supported(Tk, 64) and not ops.attn_bwd_pair_supported(321, 64) and not ops.attn_bwd_pair_supported(Tk, 32)
    d = 64 * H
    x, wide, l, _ = _case(B, H, Tq, Tk, ld)
    r = wide[:, d:2 * d]
    dl = torch.full((B * Tq, d), float("nan"), device=DEV, dtype=torch.bfloat16)
    drw = torch.full((B * Tk, 3 * d), float("nan"), device=DEV, dtype=torch.bfloat16)
    cs = torch.zeros(d, device=DEV)
    ops.attn_bwd_pair(x, r, l, dl, drw[:, d:2 * d], B, H, Tq, Tk, 64, colsum=cs)
    torch.cuda.synchronize()
    want_dl, want_dr = _ref(x, r, l, B, H, Tq, Tk, True)
    dr = drw[:, d:2 * d].float().view(B, Tk, H, 64)
    assert torch.isnan(drw[:, :d]).all() and torch.isnan(drw[:, 2 * d:]).all()       # neighbours of the slice are untouched
    tol_l = 1e-2 * float(want_dl.abs().max()) + 1e-3
    tol_r = 1e-2 * float(want_dr.abs().max()) + 1e-3
    assert float((dl.float() - want_dl).abs().max()) <= tol_l
    assert float((dr - want_dr).abs().max()) <= tol_r
    want_cs = want_dr.sum(dim=(0, 1)).reshape(-1)
    assert float((cs - want_cs).abs().max()) <= 2e-3 * float(want_cs.abs().max()) + 1e-2


@pytest.mark.parametrize("B,H,Tq,Tk,ld", [(3, 4, 299, 299, 320), (130, 4, 64, 64, 64), (7, 2, 150, 150, 152), (300, 1, 5, 5, 8)])
def test_pair_batch_reduced_shared_r(B, H, Tq, Tk, ld):
    """X = dbd: r = the positional projection shared by every utterance, dr accumulates (+=) over the batch in fp32; a CTA that owns
    several utterances of one head flushes once (B = 130, 300: more units than SMs, head boundaries inside a CTA's range)."""
    from liteasr_b200 import ops
    d = 64 * H
    x, _, l, pos = _case(B, H, Tq, Tk, ld, seed=1)
    dl = torch.empty((B * Tq, d), device=DEV, dtype=torch.bfloat16)
    base = torch.randn(Tk, d, device=DEV)
    dr = base.clone()
    ops.attn_bwd_pair(x, pos, l, dl, dr, B, H, Tq, Tk, 64, r_batched=False, reduce_b=True)
    torch.cuda.synchronize()
    want_dl, want_dr = _ref(x, pos, l, B, H, Tq, Tk, False)
    want = base + want_dr.sum(0).reshape(Tk, d)
    assert float((dl.float() - want_dl).abs().max()) <= 1e-2 * float(want_dl.abs().max()) + 1e-3
    assert float((dr - want).abs().max()) <= 1e-4 * float(want.abs().max()) + 1e-3


def test_engine_paired_and_unpaired_backward_agree(monkeypatch):
    """The training step with LASR_FUSED_ATTN_BWD=0 (two batched GEMMs per gradient tensor) and with the paired kernel."""
    import json
    import os
    from liteasr_b200.criterions.hybrid_ctc_attn import HybridCTCLoss, HybridCTCLossConfig
    from liteasr_b200.models.u2 import U2, U2Config
    from liteasr_b200.utils.synthetic import synth_batch, synth_state_dict
    from liteasr_b200.schema import U2Dims
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "u2_tiny.json")))
    kw = dict(g["dims"])  # enc_dim 128, 2 heads: dk = 64
    dims = U2Dims(**kw)
    sd = synth_state_dict(dims, seed=3)
    batch = [t.to(DEV) for t in synth_batch(3, 400, 12, dims.vocab_size, seed=3)]
    crit = HybridCTCLoss(HybridCTCLossConfig(vocab_size=dims.vocab_size, smoothing=0.1, ctc_weight=0.3))
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("LASR_FUSED_ATTN_BWD", flag)
        model = U2(U2Config(**kw, precision="bf16"))
        model.load_state_dict(sd)
        model = model.to(DEV).train()
        loss = crit(model, *batch)
        loss.backward()
        torch.cuda.synchronize()
        out[flag] = (float(loss.detach()), {n: p.grad.float().clone() for n, p in model.named_parameters()})
    assert out["0"][0] == out["1"][0]
    gmax = max(float(v.abs().max()) for v in out["0"][1].values())
    for n, a in out["0"][1].items():
        err = float((a - out["1"][1][n]).abs().max())
        assert err <= 2e-3 * gmax, (n, err, gmax)
