"""GPU: the B-stationary tcgen05 GEMM variant (K <= 256, the weight tile resident in shared memory, 16-column epilogue staging)
against fp32 torch on shapes that select it (many M tiles per CTA, even split), for every fused epilogue."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _select_b_stationary(monkeypatch):
    monkeypatch.setenv("LASR_GEMM_BS", "1")  # the variant is off by default


def _mk(rows, cols, g, ld=None):
    ld = ld or (cols + 7) // 8 * 8
    return (torch.randn(rows, ld, generator=g, device="cuda") * 0.5).bfloat16()[:, :cols]


def _swish(v):
    return v * torch.sigmoid(v)


def _dswish(v):
    s = torch.sigmoid(v)
    return s * (1 + v * (1 - s))


@pytest.mark.parametrize("m,n,k", [(37674, 256, 256), (37674, 2048, 256), (37674, 512, 200), (56000, 256, 64), (37674, 4233, 256)])
@pytest.mark.parametrize("tb", [False, True])
def test_bs_plain_bias_bf16_and_f32(m, n, k, tb):
    from liteasr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    x = _mk(m, k, g)
    w = _mk(k, n, g) if tb else _mk(n, k, g)
    bias = torch.randn(n, generator=g, device="cuda")
    ref = x.float() @ (w.float() if tb else w.float().t())
    ldn = (n + 7) // 8 * 8
    for cdt in (torch.bfloat16, torch.float32):
        use_bias = not tb
        c = torch.full((m, ldn), float("nan"), device="cuda", dtype=cdt)[:, :n]
        ops.gemm(x, w, c, m, n, k, lda=x.stride(0), ldb=w.stride(0), ldc=c.stride(0), tb=tb, bias=bias if use_bias else None, alpha=0.5)
        want = 0.5 * (ref + (bias if use_bias else 0.0))
        assert torch.isfinite(c.float()).all()
        err = (c.float() - want).abs().max().item()
        assert err <= (3e-2 if cdt == torch.bfloat16 else 2e-3) * max(1.0, want.abs().max().item()), (cdt, err)


def test_bs_swish_aux_relu_res():
    from liteasr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    m, n, k = 37674, 2048, 256
    x, w = _mk(m, k, g), _mk(n, k, g)
    bias = torch.randn(n, generator=g, device="cuda")
    pre = x.float() @ w.float().t() + bias
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    aux = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    ops.linear(x, w, out, bias=bias, aux=aux, act=ops.ACT_SWISH)
    assert (aux.float() - pre).abs().max().item() <= 3e-2 * pre.abs().max().item()
    assert (out.float() - _swish(pre)).abs().max().item() <= 3e-2 * pre.abs().max().item()
    ops.linear(x, w, out, bias=bias, act=ops.ACT_RELU, alpha=2.0)
    assert (out.float() - 2.0 * torch.relu(pre)).abs().max().item() <= 3e-2 * 2 * pre.abs().max().item()
    # fp32 C with residual (N = 256: one N tile shared by every CTA)
    n2 = 256
    w2 = _mk(n2, k, g)
    res = torch.randn(m, n2, generator=g, device="cuda")
    o32 = torch.empty(m, n2, device="cuda")
    ops.linear(x, w2, o32, bias=bias[:n2].contiguous(), res=res, alpha=0.5)
    want = 0.5 * (x.float() @ w2.float().t() + bias[:n2]) + res
    assert (o32 - want).abs().max().item() <= 2e-2 * want.abs().max().item()


@pytest.mark.parametrize("act", ["swish", "relu"])
def test_bs_fused_activation_backward_and_colsum(act):
    from liteasr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(13)
    m, n, k = 37674, 2048, 256
    dy = _mk(m, k, g)
    w = _mk(k, n, g)  # (N_out = k, K_in = n) weight, MN-major B
    saved = _mk(m, n, g)
    dx = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    cs = torch.zeros(n, device="cuda")
    ops.gemm(dy, w, dx, m, n, k, lda=dy.stride(0), ldb=w.stride(0), ldc=n, tb=True, alpha=0.5, dact=saved,
             act=ops.ACT_SWISH if act == "swish" else ops.ACT_RELU, colsum=cs)
    d = _dswish(saved.float()) if act == "swish" else (saved.float() > 0).float()
    want = 0.5 * (dy.float() @ w.float()) * d
    assert (dx.float() - want).abs().max().item() <= 3e-2 * max(1.0, want.abs().max().item())
    ref_cs = want.sum(0)
    assert (cs - ref_cs).abs().max().item() <= 2e-2 * max(1.0, ref_cs.abs().max().item()) + 0.5
