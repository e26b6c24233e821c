"""GPU: the parity-plane sub-sampling front end (conv1 planes + implicit-GEMM conv2 forward / input gradient / weight gradient,
include/lasr.h) against torch.nn.functional.conv2d + autograd on the same bf16-rounded operands (nets/subsampling.py:32-35)."""
import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu


def unplane(h1p, T1, F1, U, V):
    """(B,4,U*V,d) -> (B,T1,F1,d)"""
    B, _, _, d = h1p.shape
    p = h1p.view(B, 2, 2, U, V, d)
    out = torch.zeros(B, 2 * U, 2 * V, d, dtype=h1p.dtype, device=h1p.device)
    for pt in range(2):
        for pf in range(2):
            out[:, pt::2, pf::2] = p[:, pt, pf]
    return out[:, :T1, :F1]


def to_planes(h1, U, V):
    B, T1, F1, d = h1.shape
    full = torch.zeros(B, 2 * U, 2 * V, d, dtype=h1.dtype, device=h1.device)
    full[:, :T1, :F1] = h1
    p = torch.stack([torch.stack([full[:, pt::2, pf::2] for pf in range(2)], 1) for pt in range(2)], 1)  # (B,2,2,U,V,d)
    return p.reshape(B, 4, U * V, d).contiguous()


@pytest.mark.parametrize("B,T,F,d", [(2, 67, 80, 128), (3, 100, 83, 64), (1, 50, 20, 256), (2, 131, 80, 256), (2, 67, 80, 512), (5, 203, 80, 256)])
def test_plane_front_end_matches_torch(B, T, F, d):
    from liteasr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B * T + d)
    x = torch.randn(B, T, F, generator=g, device="cuda")
    w1 = torch.randn(d, 9, generator=g, device="cuda") * 0.3
    b1 = torch.randn(d, generator=g, device="cuda") * 0.1
    w2 = (torch.randn(d, d, 3, 3, generator=g, device="cuda") * (1.0 / (3 * d ** 0.5))).bfloat16()
    b2 = torch.randn(d, generator=g, device="cuda") * 0.1
    T1, F1, U, V, T2, F2 = ops.plane_dims(T, F)
    assert F2 == V - 1
    # ---- conv1 -> planes
    h1p = torch.full((B, 4, U * V, d), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.conv1_fwd_planes(x, w1, b1, h1p)
    ref1 = TF.relu(TF.conv2d(x.unsqueeze(1), w1.view(d, 1, 3, 3), b1, stride=2)).permute(0, 2, 3, 1)  # (B,T1,F1,d)
    got1 = unplane(h1p, T1, F1, U, V)
    assert torch.isfinite(h1p.float()).all()
    assert torch.allclose(got1.float(), ref1, atol=2e-2, rtol=1e-2)
    assert torch.equal(to_planes(got1, U, V), h1p)  # every slot without a (t1, f1) is exactly zero
    # ---- conv2 forward (implicit GEMM) vs conv2d on the same bf16 h1
    w2k = w2.permute(0, 2, 3, 1).reshape(d, 9 * d).contiguous()  # (co, kh, kw, ci)
    h2p = torch.full((B * T2, V * d), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.conv2_fwd(h1p, w2k, b2, h2p, B, T, F)
    h1f = got1.float().permute(0, 3, 1, 2).requires_grad_(True)  # (B,d,T1,F1)
    w2f = w2.float().requires_grad_(True)
    ref2 = TF.relu(TF.conv2d(h1f, w2f, b2, stride=2))  # (B,d,T2,F2)
    got2 = h2p.view(B, T2, V, d)[:, :, :F2].float()
    assert torch.isfinite(h2p.float()).all()
    err = (got2 - ref2.permute(0, 2, 3, 1)).abs().max().item()
    assert err <= 2e-2 * max(1.0, ref2.abs().max().item()), err
    # ---- backward: dY (ReLU-masked upstream gradient) in the padded layout, padding slot zero
    dy = (torch.randn(B, T2, F2, d, generator=g, device="cuda") * (ref2.permute(0, 2, 3, 1) > 0)).bfloat16()
    dyp = torch.zeros(B, T2, V, d, dtype=torch.bfloat16, device="cuda")
    dyp[:, :, :F2] = dy
    pre = TF.conv2d(h1f, w2f, None, stride=2)
    pre.backward(dy.float().permute(0, 3, 1, 2))
    # weight gradient
    gw = torch.zeros(d, 9 * d, dtype=torch.float32, device="cuda")
    ops.conv2_wgrad(dyp.view(B * T2, V * d), h1p, gw, B, T, F)
    ref_gw = w2f.grad.permute(0, 2, 3, 1).reshape(d, 9 * d)
    assert (gw - ref_gw).abs().max().item() <= 2e-3 * ref_gw.abs().max().item() + 1e-4
    # input gradient with conv1's ReLU mask
    dh1p = torch.full((B, 4, U * V, d), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.conv2_dgrad(dyp.view(B * T2, V * d), w2k, h1p, dh1p, B, T, F)
    ref_dh1 = (h1f.grad * (h1f > 0)).permute(0, 2, 3, 1)
    got_dh1 = unplane(dh1p, T1, F1, U, V).float()
    assert torch.isfinite(dh1p.float()).all()
    assert (got_dh1 - ref_dh1).abs().max().item() <= 2e-2 * max(1.0, ref_dh1.abs().max().item())
    assert torch.equal(to_planes(unplane(dh1p, T1, F1, U, V), U, V), dh1p)
    # conv1 weight / bias gradient from the planes
    dw1 = torch.zeros(d, 9, device="cuda")
    db1 = torch.zeros(d, device="cuda")
    ops.conv1_bwd_planes(x, dh1p, dw1, db1)
    xr = x.unsqueeze(1).clone()
    w1r = w1.view(d, 1, 3, 3).clone().requires_grad_(True)
    b1r = b1.clone().requires_grad_(True)
    TF.conv2d(xr, w1r, b1r, stride=2).backward(got_dh1.permute(0, 3, 1, 2))
    assert torch.allclose(dw1, w1r.grad.view(d, 9), rtol=2e-3, atol=2e-3 * w1r.grad.abs().max().item())
    assert torch.allclose(db1, b1r.grad, rtol=2e-3, atol=2e-3 * b1r.grad.abs().max().item())
