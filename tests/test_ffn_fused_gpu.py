"""GPU: the fused feed-forward backward kernel (csrc/ffn_fused.cu: dh = alpha (dy @ W2) * g, db1 += colsum(dh), dln = dh @ W1)
against a torch fp32 reference of the same three expressions (nets/feed_forward.py:18-19 backward) and against the two
separate lasr_gemm calls it replaces."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _case(m, d, f, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed + m + d + f)
    dy = (torch.randn(m, d, generator=g, device=DEV) * 0.5).bfloat16()
    w2 = (torch.randn(d, f, generator=g, device=DEV) * 0.05).bfloat16()
    w1 = (torch.randn(f, d, generator=g, device=DEV) * 0.05).bfloat16()
    h = torch.randn(m, f, generator=g, device=DEV) * 2
    sg = torch.sigmoid(h)
    gd = (sg * (1 + h * (1 - sg)))
    gd = torch.where(torch.rand(m, f, generator=g, device=DEV) < 0.1, torch.zeros_like(gd), gd).bfloat16()  # inner-dropout zeros
    return dy, w2, w1, gd


@pytest.mark.parametrize("m,d,f", [(300, 256, 2048), (129, 128, 384), (5, 64, 128), (1000, 256, 128), (37674, 256, 2048)])
def test_ffn_bwd_fused_matches_reference(m, d, f):
    from liteasr_b200 import ops
    assert ops.ffn_bwd_supported(d, f)
    dy, w2, w1, gd = _case(m, d, f)
    alpha = 0.5 * 1.1112
    dh = torch.empty(m, f, device=DEV, dtype=torch.bfloat16)
    dln = torch.empty(m, d, device=DEV, dtype=torch.bfloat16)
    cs = torch.zeros(f, device=DEV)
    ops.ffn_bwd(dy, gd, w2, w1, dh, dln, colsum=cs, alpha=alpha)
    torch.cuda.synchronize()
    dh_ref = alpha * (dy.float() @ w2.float()) * gd.float()
    # dh: one bf16 rounding of an fp32-accumulated product
    assert torch.allclose(dh.float(), dh_ref, rtol=1e-2, atol=1e-2 * float(dh_ref.abs().max()))
    assert (dh.float()[gd.float() == 0] == 0).all()
    # the second contraction consumes the ROUNDED dh (what the unfused path reads back from HBM as well)
    dln_ref = dh.float() @ w1.float()
    assert torch.allclose(dln.float(), dln_ref, rtol=1e-2, atol=1e-2 * float(dln_ref.abs().max()))
    cs_ref = dh_ref.sum(0)   # column sums of the fp32 values (taken before the bf16 rounding of the stored copy)
    assert torch.allclose(cs, cs_ref, rtol=1e-3, atol=2e-3 * float(cs_ref.abs().max()) + 1e-4)


def test_ffn_bwd_fused_equals_the_two_gemms_it_replaces():
    from liteasr_b200 import ops
    m, d, f = 777, 256, 2048
    dy, w2, w1, gd = _case(m, d, f, seed=5)
    dh = torch.empty(m, f, device=DEV, dtype=torch.bfloat16)
    dln = torch.empty(m, d, device=DEV, dtype=torch.bfloat16)
    cs = torch.zeros(f, device=DEV)
    ops.ffn_bwd(dy, gd, w2, w1, dh, dln, colsum=cs, alpha=0.5)
    dh2 = torch.empty_like(dh)
    dln2 = torch.empty_like(dln)
    cs2 = torch.zeros_like(cs)
    ops.gemm(dy, w2, dh2, m, f, d, lda=d, ldb=f, ldc=f, tb=True, alpha=0.5, dact=gd, act=ops.ACT_MUL, colsum=cs2)
    ops.gemm(dh2, w1, dln2, m, d, f, lda=f, ldb=d, ldc=d, tb=True)
    assert torch.equal(dh, dh2)   # same MMA order per element (K = d in one accumulation chain), same rounding
    assert torch.allclose(dln.float(), dln2.float(), rtol=1e-2, atol=1e-2 * float(dln2.float().abs().max()))
    assert torch.allclose(cs, cs2, rtol=1e-3, atol=2e-3 * float(cs2.abs().max()))


def test_ffn_bwd_rejects_unsupported_shapes():
    from liteasr_b200 import ops
    assert not ops.ffn_bwd_supported(512, 2048) and not ops.ffn_bwd_supported(144, 320) and not ops.ffn_bwd_supported(256, 192) and not ops.ffn_bwd_supported(256, 4096)
    x = torch.zeros(8, 512, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.ffn_bwd(x, torch.zeros(8, 64, device=DEV, dtype=torch.bfloat16), torch.zeros(512, 64, device=DEV, dtype=torch.bfloat16),
                    torch.zeros(64, 512, device=DEV, dtype=torch.bfloat16), torch.zeros(8, 64, device=DEV, dtype=torch.bfloat16),
                    torch.zeros(8, 512, device=DEV, dtype=torch.bfloat16))


# ------------------------------------------------------------------------------------------------------------------------------
# fused forward (lasr_ffn_fwd): a = drop_in(swish(ln W1^T + b1)), g = swish'(.), out = res + drop_out(alpha (a W2^T + b2))
# ------------------------------------------------------------------------------------------------------------------------------
def _fwd_case(m, d, f, seed=0):
    gen = torch.Generator(device=DEV).manual_seed(seed + 3 * m + d + f)
    ln = torch.randn(m, d, generator=gen, device=DEV).bfloat16()
    w1 = (torch.randn(f, d, generator=gen, device=DEV) * d ** -0.5).bfloat16()
    w2 = (torch.randn(d, f, generator=gen, device=DEV) * f ** -0.5).bfloat16()
    b1 = torch.randn(f, generator=gen, device=DEV) * 0.1
    b2 = torch.randn(d, generator=gen, device=DEV) * 0.1
    res = torch.randn(m, d, generator=gen, device=DEV)
    return ln, w1, b1, w2, b2, res


def _run_fwd(case, alpha, di=None, do=None):
    from liteasr_b200 import ops
    ln, w1, b1, w2, b2, res = case
    m, d = ln.shape
    f = w1.shape[0]
    a = torch.full((m, f), float("nan"), device=DEV, dtype=torch.bfloat16)
    g = torch.full((m, f), float("nan"), device=DEV, dtype=torch.bfloat16)
    out = torch.full((m, d), float("nan"), device=DEV)
    ops.ffn_fwd(ln, w1, b1, w2, b2, res, a, g, out, alpha=alpha, drop_in=di, drop_out=do)
    torch.cuda.synchronize()
    return a, g, out


@pytest.mark.parametrize("m,d,f", [(300, 256, 2048), (129, 128, 384), (5, 64, 128), (1000, 192, 256), (37674, 256, 2048)])
def test_ffn_fwd_fused_matches_reference(m, d, f):
    from liteasr_b200 import ops
    assert ops.ffn_fwd_supported(d, f)
    case = _fwd_case(m, d, f)
    ln, w1, b1, w2, b2, res = case
    a, g, out = _run_fwd(case, 0.5)
    h = ln.float() @ w1.float().t() + b1
    sg = torch.sigmoid(h)
    a_ref, g_ref = h * sg, sg * (1 + h * (1 - sg))
    # tanh.approx-based swish / swish' (rel. error ~2^-11) + one bf16 rounding
    assert torch.allclose(a.float(), a_ref, rtol=1.2e-2, atol=2e-3)
    assert torch.allclose(g.float(), g_ref, rtol=1.2e-2, atol=4e-3)
    out_ref = res + 0.5 * (a.float() @ w2.float().t() + b2)   # the second contraction consumes the ROUNDED a
    assert torch.allclose(out, out_ref, rtol=1e-4, atol=2e-3)


def test_ffn_fwd_fused_equals_the_two_gemms_it_replaces_with_dropout():
    """Same masks (site ids, RNG snapshot), same arithmetic as lasr_gemm(act = Swish, aux_deriv, drop) + lasr_gemm(res, drop)."""
    from liteasr_b200 import ops
    m, d, f = 777, 256, 2048
    case = _fwd_case(m, d, f, seed=9)
    ln, w1, b1, w2, b2, res = case
    st = torch.tensor([4321, 17], dtype=torch.int64, device=DEV)
    for p_in, p_out in ((0.0, 0.0), (0.1, 0.1), (0.3, 0.0), (0.0, 0.25)):
        di = ops.Drop(st, 0x01000103, p_in) if p_in else None
        do = ops.Drop(st, 0x01000104, p_out) if p_out else None
        a, g, out = _run_fwd(case, 0.5, di, do)
        a2 = torch.empty_like(a)
        g2 = torch.empty_like(g)
        out2 = torch.empty_like(out)
        ops.gemm(ln, w1, a2, m, f, d, lda=d, ldb=d, ldc=f, bias=b1, aux=g2, act=ops.ACT_SWISH, drop=di, drop_mark_aux=True, aux_deriv=True)
        ops.gemm(a2, w2, out2, m, d, f, lda=f, ldb=f, ldc=d, bias=b2, res=res, ldres=d, alpha=0.5, drop=do)
        torch.cuda.synchronize()
        assert torch.equal(a, a2) and torch.equal(g, g2), (p_in, p_out)   # same accumulation chain, same epilogue arithmetic, same masks
        assert torch.allclose(out, out2, rtol=1e-5, atol=1e-4), (p_in, p_out, float((out - out2).abs().max()))
        if p_in:
            dropped = float((a.float() == 0).float().mean())
            assert abs(dropped - di.thr / 32768.0) < 0.01 and bool(((a.float() == 0) == (g.float() == 0)).all())


def test_engine_ffn_fused_forward_and_unfused_agree(monkeypatch):
    import json
    import os
    from liteasr_b200.criterions.hybrid_ctc_attn import HybridCTCLoss, HybridCTCLossConfig
    from liteasr_b200.models.u2 import U2, U2Config
    from liteasr_b200.schema import U2Dims
    from liteasr_b200.utils.synthetic import synth_batch, synth_state_dict
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "u2_tiny.json")))
    kw = dict(g["dims"])
    dims = U2Dims(**kw)
    sd = synth_state_dict(dims, seed=3)
    batch = [t.to(DEV) for t in synth_batch(3, 400, 12, dims.vocab_size, seed=3)]
    crit = HybridCTCLoss(HybridCTCLossConfig(vocab_size=dims.vocab_size, smoothing=0.1, ctc_weight=0.3))
    out = {}
    for flag in ("0", "1"):  # "1" = the fused forward kernel (off by default)
        monkeypatch.setenv("LASR_FUSED_FFN_FWD", flag)
        torch.manual_seed(11)
        model = U2(U2Config(**kw, precision="bf16", dropout_rate=0.1, enc_attn_dropout_rate=0.0, dec_self_attn_dropout_rate=0.0,
                            dec_src_attn_dropout_rate=0.0))  # config/model/my_U2.yaml
        model.load_state_dict(sd)
        model = model.to(DEV).train()
        loss = crit(model, *batch)
        loss.backward()
        torch.cuda.synchronize()
        out[flag] = (float(loss.detach()), {n: p.grad.float().clone() for n, p in model.named_parameters()})
    assert abs(out["0"][0] - out["1"][0]) <= 1e-4 * abs(out["0"][0])
    gmax = max(float(v.abs().max()) for v in out["0"][1].values())
    for n, a in out["0"][1].items():
        err = float((a - out["1"][1][n]).abs().max())
        assert err <= 2e-3 * gmax, (n, err, gmax)
