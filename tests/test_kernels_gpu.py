"""GPU: the bandwidth kernels (LayerNorm, conv-module middle, attention softmax) through the C ABI vs plain torch fp32
references of the same op (the torch ops the reference calls: nets/layer_norm.py:15, nets/conformer_convolution.py:49-53,
nets/attention.py:46-59,99-118)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(a, b, tol):
    a, b = a.float(), b.float()
    return (a - b).abs().max().item() <= tol * max(1.0, b.abs().max().item())


@pytest.mark.parametrize("rows,d", [(37, 64), (1000, 128), (9568, 256), (700, 512), (129, 1024)])
@pytest.mark.parametrize("dy_dtype", [torch.float32, torch.bfloat16])
def test_layernorm_fwd_bwd_fused_outputs(rows, d, dy_dtype):
    from liteasr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + d)
    x = torch.randn(rows, d, generator=g, device="cuda") * 2 + 0.5
    gamma = torch.randn(d, generator=g, device="cuda")
    beta = torch.randn(d, generator=g, device="cuda")
    y = torch.empty(rows, d, device="cuda")
    mean = torch.empty(rows, device="cuda")
    rstd = torch.empty(rows, device="cuda")
    ops.layernorm_fwd(x, gamma, beta, y, mean, rstd, 1e-12)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-12)
    assert _close(y, yr, 1e-5)
    ylo = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    ops.layernorm_fwd(x, gamma, beta, ylo, None, None, 1e-12)
    assert _close(ylo, yr, 1e-2)
    dy = torch.randn(rows, d, generator=g, device="cuda").to(dy_dtype)
    yr.backward(dy.float())
    for accumulate in (False, True):
        base = torch.randn(rows, d, generator=g, device="cuda")
        dx = base.clone()
        dg, db = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
        cs = torch.ones(d, device="cuda")
        lo = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
        ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, dg, db, accumulate, dx_lo=lo, colsum=cs, colsum_scale=0.5)
        ref = xr.grad + (base if accumulate else 0)
        assert _close(dx, ref, 1e-4)
        assert _close(lo, ref, 1e-2)
        assert _close(cs - 1.0, 0.5 * ref.sum(0), 1e-4)
        assert _close(dg, gr.grad, 1e-4) and _close(db, br.grad, 1e-4)
        # the plain call (no fused outputs) gives the same dx
        dx2 = base.clone()
        ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx2, torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda"), accumulate)
        assert torch.equal(dx, dx2)


def _conv_module_ref(y2, w, bias, gamma, beta, eps=1e-5):
    """GLU -> depthwise conv k=15 pad 7 -> BatchNorm (batch stats over all B*T frames) -> Swish, on (B,T,2d) input."""
    B, T, d2 = y2.shape
    d = d2 // 2
    u = torch.nn.functional.glu(y2.transpose(1, 2), dim=1)                      # (B,d,T)
    z = torch.nn.functional.conv1d(u, w.view(d, 1, -1), bias, padding=w.shape[1] // 2, groups=d)
    zn = torch.nn.functional.batch_norm(z, None, None, gamma, beta, True, 0.1, eps)
    a = zn * torch.sigmoid(zn)
    return z.transpose(1, 2), a.transpose(1, 2)


@pytest.mark.parametrize("B,T,d", [(3, 50, 64), (5, 299, 256), (2, 31, 128), (4, 97, 512), (2, 1, 64), (3, 64, 64), (2, 33, 192), (1, 7, 144)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_module_middle_fwd_bwd(B, T, d, dtype):
    from liteasr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B * T + d)
    y2 = torch.randn(B, T, 2 * d, generator=g, device="cuda").to(dtype)
    w = torch.randn(d, 15, generator=g, device="cuda") * 0.3
    bias = torch.randn(d, generator=g, device="cuda") * 0.1
    gamma = torch.rand(d, generator=g, device="cuda") + 0.5
    beta = torch.randn(d, generator=g, device="cuda") * 0.1
    rows = B * T
    y2f = y2.float().clone().requires_grad_(True)
    wr, br, gr, ber = (t.clone().requires_grad_(True) for t in (w, bias, gamma, beta))
    z_ref, a_ref = _conv_module_ref(y2f, wr, br, gr, ber)
    # forward
    z = torch.empty(rows, d, device="cuda")
    nblk = B * ((T + 31) // 32)
    partial = torch.empty(nblk, 2, d, device="cuda")
    ops.glu_dwconv_fwd(y2.view(rows, 2 * d), w, bias, z, partial, B, T, d)
    # bf16 operands: the streaming kernel's GLU uses sigmoid(x) = 0.5 tanh.approx(x / 2) + 0.5 (rel. error ~2^-11, below the 2^-9
    # resolution of the operands it multiplies)
    bf = dtype == torch.bfloat16
    assert _close(z, z_ref.reshape(rows, d), 2e-3 if bf else 1e-5)
    mean, rstd = torch.empty(d, device="cuda"), torch.empty(d, device="cuda")
    rm, rv = torch.zeros(d, device="cuda"), torch.ones(d, device="cuda")
    nbt = torch.zeros((), dtype=torch.int64, device="cuda")
    ops.bn_finalize(partial, nblk, d, rows, mean, rstd, rm, rv, nbt, True)
    zr = z_ref.detach().reshape(rows, d)
    assert _close(mean, zr.mean(0), 5e-4 if bf else 1e-5) and _close(rstd, (zr.var(0, unbiased=False) + 1e-5).rsqrt(), 1e-3 if bf else 1e-4)
    assert _close(rm, 0.1 * zr.mean(0), 5e-4 if bf else 1e-5) and _close(rv, 0.9 + 0.1 * zr.var(0, unbiased=True), 1e-3 if bf else 1e-4) and int(nbt) == 1
    zs = z.view(B, T, d).double()   # the statistics are exact for the z the kernel wrote
    assert _close(mean, zs.mean((0, 1)).float(), 1e-5) and _close(rstd, (zs.var((0, 1), unbiased=False) + 1e-5).rsqrt().float(), 1e-4)
    a = torch.empty(rows, d, device="cuda", dtype=dtype)
    ops.bn_swish_fwd(z, mean, rstd, gamma, beta, a)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert _close(a, a_ref.reshape(rows, d), tol)
    # backward
    da = torch.randn(rows, d, generator=g, device="cuda").to(dtype)
    a_ref.backward(da.float().view(B, T, d))
    part2 = torch.empty((rows + 31) // 32, 2, d, device="cuda")
    sums = torch.empty(2, d, device="cuda")
    dgam, dbet = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    ops.bn_swish_bwd_stats(da, z, mean, rstd, gamma, beta, part2, sums, dgam, dbet)
    assert _close(dgam, gr.grad, 2e-4 if dtype == torch.float32 else 2e-2) and _close(dbet, ber.grad, 2e-4 if dtype == torch.float32 else 2e-2)
    dy2 = torch.empty(rows, 2 * d, device="cuda", dtype=dtype)
    dw, dbias = torch.zeros(d, 15, device="cuda"), torch.zeros(d, device="cuda")
    cs = torch.zeros(2 * d, device="cuda")
    ops.dwconv_glu_bwd(da, z, y2.view(rows, 2 * d), mean, rstd, gamma, beta, sums, w, dy2, dw, dbias, B, T, d, colsum=cs)
    # two-stage (deterministic) reduction of the same gradients
    dw2, dbias2, cs2 = torch.zeros(d, 15, device="cuda"), torch.zeros(d, device="cuda"), torch.zeros(2 * d, device="cuda")
    wpart = torch.empty(nblk, 18, d, device="cuda")
    ops.dwconv_glu_bwd(da, z, y2.view(rows, 2 * d), mean, rstd, gamma, beta, sums, w, torch.empty_like(dy2), dw2, dbias2, B, T, d,
                       colsum=cs2, wpartial=wpart)
    assert _close(dw2, dw, 1e-4) and _close(dbias2, dbias, 1e-4) and _close(cs2, cs, 1e-4)
    tolg = 2e-4 if dtype == torch.float32 else 3e-2
    assert _close(dy2, y2f.grad.reshape(rows, 2 * d), tolg)
    assert _close(dw, wr.grad, tolg) and _close(dbias, br.grad, tolg)
    assert _close(cs, dy2.float().sum(0), 1e-3 if dtype == torch.float32 else 3e-2)


def _rel_shift(x):
    """nets/attention.py:99-118 (legacy, no zero_triu)."""
    b, h, t1, t2 = x.shape
    zp = torch.zeros(b, h, t1, 1, device=x.device, dtype=x.dtype)
    xp = torch.cat([zp, x], dim=-1).view(b, h, t2 + 1, t1)
    return xp[:, :, 1:].view_as(x)


@pytest.mark.parametrize("B,H,T", [(2, 2, 7), (3, 4, 124), (2, 4, 299), (1, 8, 399)])
@pytest.mark.parametrize("pdt", [torch.float32, torch.bfloat16])
def test_attn_softmax_relshift_fwd_bwd(B, H, T, pdt):
    from liteasr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B * H * T)
    ld = (T + 7) // 8 * 8
    ac = torch.zeros(B, H, T, ld, device="cuda")
    bd = torch.zeros(B, H, T, ld, device="cuda")
    ac[..., :T] = torch.randn(B, H, T, T, generator=g, device="cuda") * 3
    bd[..., :T] = torch.randn(B, H, T, T, generator=g, device="cuda") * 3
    lens = torch.randint(max(1, 4 * T - 11), 4 * T + 1, (B,), generator=g, device="cuda")   # raw frame counts: klen = ceil(l/4)
    lens[0] = 4 * T
    scale = 0.125
    acr = ac[..., :T].clone().requires_grad_(True)
    bdr = bd[..., :T].clone().requires_grad_(True)
    klen = (lens + 3) // 4
    mask = torch.arange(T, device="cuda")[None, :] >= klen[:, None]
    sc = ((acr + _rel_shift(bdr)) * scale).masked_fill(mask[:, None, None, :], -1e38)
    pr = torch.softmax(sc, -1)
    probs = torch.full((B, H, T, ld), 9.0, device="cuda", dtype=pdt)
    ops.attn_softmax_fwd(ac, bd, probs, lens, 3, 0, scale, T)
    tol = 1e-5 if pdt == torch.float32 else 1e-2
    assert _close(probs[..., :T], pr, tol)
    assert (probs[..., T:] == 0).all()
    dp = torch.zeros(B, H, T, ld, device="cuda")
    dp[..., :T] = torch.randn(B, H, T, T, generator=g, device="cuda")
    pr.backward(dp[..., :T])
    # feed the kernel the exact probabilities it would have saved
    dsc = torch.empty(B, H, T, ld, device="cuda", dtype=pdt)
    dbd = torch.empty(B, H, T, ld, device="cuda", dtype=pdt)
    ops.attn_softmax_bwd(probs, dp, dsc, dbd, scale, T)
    tolb = 1e-4 if pdt == torch.float32 else 3e-2
    assert _close(dsc[..., :T], acr.grad, tolb)
    assert _close(dbd[..., :T], bdr.grad, tolb)
    assert (dsc[..., T:] == 0).all()


def test_attn_softmax_plain_causal_and_cross():
    from liteasr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    B, H, Tq, Tk = 3, 2, 9, 9
    ld = 16
    ac = torch.zeros(B, H, Tq, ld, device="cuda")
    ac[..., :Tk] = torch.randn(B, H, Tq, Tk, generator=g, device="cuda")
    ylens = torch.tensor([8, 5, 3], device="cuda")
    probs = torch.empty(B, H, Tq, ld, device="cuda")
    ops.attn_softmax_fwd(ac, None, probs, ylens, 2, 1, 0.5, Tk)   # decoder self-attention: keys j <= i and j < ylen+1
    j = torch.arange(Tk, device="cuda")
    mask = (j[None, None, :] > j[None, :, None]) | (j[None, None, :] >= (ylens + 1)[:, None, None])
    ref = torch.softmax((ac[..., :Tk] * 0.5).masked_fill(mask[:, None], -1e38), -1)
    assert _close(probs[..., :Tk], ref, 1e-5)


@pytest.mark.parametrize("rows,d", [(37674, 256), (1000, 512), (77, 64), (5166, 144)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pos_bias_bwd(rows, d, dtype):
    """dq = dqu + dqv, du += colsum(dqu), dv += colsum(dqv), dq-bias += colsum(dq): the backward of q + pos_bias_u / q + pos_bias_v
    (nets/attention.py:135-139); bf16 with d % 8 == 0 takes the wide kernel, everything else the generic one."""
    from liteasr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + d)
    wide = torch.randn(rows, 2 * d + 8, generator=g, device="cuda").to(dtype)   # strided views of a wider buffer
    dqu, dqv = wide[:, :d], wide[:, d:2 * d]
    out = torch.full((rows, 3 * d), float("nan"), device="cuda", dtype=dtype)
    du0, dv0, db0 = (torch.randn(d, generator=g, device="cuda") for _ in range(3))
    du, dv, db = du0.clone(), dv0.clone(), db0.clone()
    ops.pos_bias_bwd(dqu, dqv, out[:, :d], du, dv, db)
    torch.cuda.synchronize()
    want = (dqu.float() + dqv.float()).to(dtype)
    assert torch.equal(out[:, :d], want) and torch.isnan(out[:, d:]).all()
    su, sv = dqu.double().sum(0), dqv.double().sum(0)
    scale = max(1.0, float(su.abs().max()), float(sv.abs().max()))
    assert float((du.double() - du0.double() - su).abs().max()) <= 1e-4 * scale
    assert float((dv.double() - dv0.double() - sv).abs().max()) <= 1e-4 * scale
    assert float((db.double() - db0.double() - su - sv).abs().max()) <= 2e-4 * scale
