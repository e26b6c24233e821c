"""GPU: the fused rel-pos attention forward (csrc/attn_fused.cu, ``lasr_rel_attn_fwd``) through the C ABI against

  * a plain torch fp32 restatement of the reference ops (nets/attention.py:46-59 apply_attention, :99-118 legacy rel_shift,
    :137-154 matrix_ac / matrix_bd / scaling), fed the same bf16-rounded operands, and
  * the unfused kernel sequence it replaces (two lasr_gemm -> lasr_attn_softmax_fwd -> lasr_gemm).

Tolerances (stated): probabilities are bf16 (relative 2^-8 rounding) and the shifted bd term passes through fp16 (2^-11 of a
logit of magnitude <~ 16 -> <= 1e-2 absolute on a logit): |p - p_ref| <= 2.5e-2 * p_ref + 3e-4; O rel-L2 <= 1e-2.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_shift_legacy(x):
    """nets/attention.py:99-118 (zero column in front, re-view, drop the first row)."""
    B, H, T1, T2 = x.shape
    zero_pad = torch.zeros((B, H, T1, 1), device=x.device, dtype=x.dtype)
    x_padded = torch.cat([zero_pad, x], dim=-1).view(B, H, T2 + 1, T1)
    return x_padded[:, :, 1:].view_as(x)


def torch_reference(qu, qv, k, v, pos, klen, B, H, T, dk):
    f = lambda t: t.float().view(B, T, H, dk).transpose(1, 2)  # noqa: E731
    p = pos.float().view(1, T, H, dk).transpose(1, 2)
    ac = torch.matmul(f(qu), f(k).transpose(-2, -1))
    bd = rel_shift_legacy(torch.matmul(f(qv), p.transpose(-2, -1)))
    scores = (ac + bd) * dk ** -0.5
    if klen is not None:
        mask = torch.arange(T, device=qu.device)[None, :] >= klen[:, None]  # (B, T) True = padded key
        scores = scores.masked_fill(mask[:, None, None, :], -1e38)
    probs = torch.softmax(scores, dim=-1)
    o = torch.matmul(probs, f(v)).transpose(1, 2).contiguous().view(B * T, H * dk)
    return probs, o


def make_inputs(B, H, T, dk, seed, amp):
    g = torch.Generator(device="cuda").manual_seed(seed)
    d = H * dk
    qkv = (torch.randn(B * T, 3 * d, generator=g, device="cuda") * amp).bfloat16()
    q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    u = torch.randn(d, generator=g, device="cuda") * 0.5
    vb = torch.randn(d, generator=g, device="cuda") * 0.5
    qu = (q.float() + u).bfloat16()
    qv = (q.float() + vb).bfloat16()
    pos = (torch.randn(T, d, generator=g, device="cuda") * amp).bfloat16()
    return qu, qv, k, v, pos


def run_fused(qu, qv, k, v, pos, lens, mask_mode, B, H, T, dk):
    from liteasr_b200 import ops
    ld = (T + 7) // 8 * 8
    probs = torch.full((B, H, T, ld), float("nan"), device="cuda", dtype=torch.bfloat16)
    o = torch.full((B * T, H * dk), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.rel_attn_fwd(qu, qv, k, v, pos, probs, o, lens, mask_mode, dk ** -0.5, B, H, T, dk)
    torch.cuda.synchronize()
    return probs, o


def run_unfused(qu, qv, k, v, pos, lens, mask_mode, B, H, T, dk):
    from liteasr_b200 import ops
    ld = (T + 7) // 8 * 8
    d = H * dk
    ac = torch.empty((B, H, T, ld), device="cuda")
    bd = torch.empty((B, H, T, ld), device="cuda")
    ops.gemm(qu, k, ac, T, T, dk, lda=qu.stride(0), ldb=k.stride(0), ldc=ld, batch=(B, H), sa=(T * qu.stride(0), dk),
             sb=(T * k.stride(0), dk), sc=(H * T * ld, T * ld), n_store=ld)
    ops.gemm(qv, pos, bd, T, T, dk, lda=qv.stride(0), ldb=pos.stride(0), ldc=ld, batch=(B, H), sa=(T * qv.stride(0), dk),
             sb=(0, dk), sc=(H * T * ld, T * ld), n_store=ld)
    probs = torch.empty((B, H, T, ld), device="cuda", dtype=torch.bfloat16)
    ops.attn_softmax_fwd(ac, bd, probs, lens, mask_mode, 0, dk ** -0.5, T)
    o = torch.empty((B * T, d), device="cuda", dtype=torch.bfloat16)
    ops.gemm(probs, v, o, T, dk, T, lda=ld, ldb=v.stride(0), ldc=d, tb=True, batch=(B, H), sa=(H * T * ld, T * ld),
             sb=(T * v.stride(0), dk), sc=(T * d, dk))
    torch.cuda.synchronize()
    return probs, o


SHAPES = [(2, 4, 50), (1, 2, 127), (2, 2, 128), (2, 1, 129), (3, 4, 299), (1, 1, 320), (2, 4, 255), (2, 3, 1), (2, 2, 7), (1, 4, 161)]


@pytest.mark.parametrize("B,H,T", SHAPES)
@pytest.mark.parametrize("masked", [False, True])
def test_rel_attn_fwd_vs_torch_and_unfused(B, H, T, masked):
    from liteasr_b200 import ops
    dk = 64
    assert ops.rel_attn_fwd_supported(T, dk)
    qu, qv, k, v, pos = make_inputs(B, H, T, dk, seed=1000 * T + B, amp=1.0)
    lens = klen = None
    mode = 0
    if masked:  # encoder mask: klen = #{j : 4j < xlens[b]} (mask_mode 3); includes a fully valid and a short utterance
        xl = torch.tensor([4 * T - 1] + [max(1, (4 * T * (b + 1)) // (2 * B + 1)) for b in range(1, B)], device="cuda", dtype=torch.int64)
        lens, mode = xl, 3
        klen = torch.clamp((xl + 3) // 4, max=T)
    probs, o = run_fused(qu, qv, k, v, pos, lens, mode, B, H, T, dk)
    pr, orf = torch_reference(qu, qv, k, v, pos, klen, B, H, T, dk)
    pf = probs.float()
    assert torch.isfinite(pf).all() and torch.isfinite(o.float()).all()
    assert (pf[..., T:] == 0).all(), "padding columns [T, ld) must be exact zeros"
    err = (pf[..., :T] - pr).abs()
    assert (err <= 2.5e-2 * pr + 3e-4).all(), f"probabilities: max abs err {err.max().item():.3e}"
    rel = ((o.float() - orf).norm() / orf.norm()).item()
    assert rel <= 1e-2, f"O rel-L2 {rel:.3e}"
    # the kernel sequence it replaces: same operands, same bf16 probabilities up to the fp16 staging of the shifted term
    pu, ou = run_unfused(qu, qv, k, v, pos, lens, mode, B, H, T, dk)
    assert (pf - pu.float()).abs().max().item() <= 2e-2
    rel_u = ((o.float() - ou.float()).norm() / ou.float().norm()).item()
    assert rel_u <= 1e-2, f"O vs unfused rel-L2 {rel_u:.3e}"


def test_rel_attn_fwd_all_keys_masked_row_is_uniform():
    """klen = 0 (quirk Q4: masked_fill(-1e38) then softmax, no post-softmax zeroing) -> uniform 1/T over the T keys."""
    B, H, T, dk = 2, 2, 40, 64
    qu, qv, k, v, pos = make_inputs(B, H, T, dk, seed=5, amp=1.0)
    lens = torch.tensor([160, 0], device="cuda", dtype=torch.int64)
    probs, _ = run_fused(qu, qv, k, v, pos, lens, 3, B, H, T, dk)
    assert torch.allclose(probs[1, :, :, :T].float(), torch.full((H, T, T), 1.0 / T, device="cuda"), rtol=1e-2, atol=0)


def test_rel_attn_fwd_large_logits():
    """Logits of magnitude ~100 (saturated softmax): still finite and close to the fp32 reference."""
    B, H, T, dk = 1, 2, 200, 64
    qu, qv, k, v, pos = make_inputs(B, H, T, dk, seed=9, amp=3.0)
    probs, o = run_fused(qu, qv, k, v, pos, None, 0, B, H, T, dk)
    pr, orf = torch_reference(qu, qv, k, v, pos, None, B, H, T, dk)
    assert torch.isfinite(probs.float()).all()
    assert (probs.float()[..., :T] - pr).abs().max().item() <= 8e-2  # fp16 staging: 2^-11 of a logit ~100 = 0.05 on the logit
    rel = ((o.float() - orf).norm() / orf.norm()).item()
    assert rel <= 4e-2, rel


def test_rel_attn_fwd_rejects_unsupported_shapes():
    from liteasr_b200 import ops
    assert not ops.rel_attn_fwd_supported(321, 64)
    assert not ops.rel_attn_fwd_supported(100, 32)
    B, H, T, dk = 1, 2, 400, 64
    qu, qv, k, v, pos = make_inputs(B, H, T, dk, seed=1, amp=1.0)
    probs = torch.empty((B, H, T, T), device="cuda", dtype=torch.bfloat16)
    o = torch.empty((B * T, H * dk), device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.rel_attn_fwd(qu, qv, k, v, pos, probs, o, None, 0, dk ** -0.5, B, H, T, dk)


def test_engine_fused_and_unfused_paths_agree(monkeypatch):
    """One C1-shaped training step (bf16) with the fused forward against LASR_FUSED_ATTN=0: loss and a gradient agree."""
    import json
    import os
    from liteasr_b200.criterions.hybrid_ctc_attn import HybridCTCLoss, HybridCTCLossConfig
    from liteasr_b200.models.u2 import U2, U2Config
    from liteasr_b200.schema import U2Dims
    from liteasr_b200.utils.synthetic import synth_batch, synth_state_dict
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = json.load(open(os.path.join(root, "tests", "golden", "u2_tiny.json")))
    dims = U2Dims(**g["dims"])
    if dims.enc_dim // dims.enc_attn_heads != 64:
        pytest.skip("fused kernel needs dk = 64")
    xs, xlens, ys, ylens = synth_batch(g["batch"], g["tmax"], g["lmax"], dims.vocab_size, seed=g["seed"])
    sd = synth_state_dict(dims, seed=g["seed"])
    crit = HybridCTCLoss(HybridCTCLossConfig(vocab_size=dims.vocab_size, smoothing=g["smoothing"], ctc_weight=g["ctc_weight"]))
    batch = tuple(t.cuda() for t in (xs, xlens, ys, ylens))
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("LASR_FUSED_ATTN", flag)
        model = U2(U2Config(**g["dims"], precision="bf16"))
        model.load_state_dict(sd)
        model = model.cuda().train()
        loss = crit(model, *batch)
        loss.backward()
        torch.cuda.synchronize()
        n = "encoder.enc_layers.0.self_attn.linear_q.weight"
        out[flag] = (float(loss), dict(model.named_parameters())[n].grad.float().clone())
    assert abs(out["1"][0] - out["0"][0]) <= 2e-3 * abs(out["0"][0])
    rel = ((out["1"][1] - out["0"][1]).norm() / out["0"][1].norm()).item()
    assert rel <= 5e-2, rel
