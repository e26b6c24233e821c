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
