"""CPU: the index arithmetic and operand packings the tcgen05 kernels rely on, restated in torch and checked against the
reference formulations (no GPU, no kernel call -- these pin the DESIGN, the GPU tests pin the kernels).

  * csrc/attn_fused.cu: the legacy ``rel_shift`` (nets/attention.py:99-118) as a flat re-view per 127-row tile: bd row r0 + r
    written at ``r (T+1) + c + (r0 + 1 - T)`` (zero at one position earlier), shifted row r' read at ``r' T + j``.
  * csrc/conv1_fwd_tc.cu / conv1_wgrad_tc.cu: Conv2d(1 -> d, 3x3, stride 2) as a K = 32 GEMM on bf16 head + tail splits.
"""
import pytest
import torch


def rel_shift_legacy(x):
    """nets/attention.py:99-118."""
    B, H, T1, T2 = x.shape
    zero_pad = torch.zeros((B, H, T1, 1), dtype=x.dtype)
    x_padded = torch.cat([zero_pad, x], dim=-1).view(B, H, T2 + 1, T1)
    return x_padded[:, :, 1:].view_as(x)


@pytest.mark.parametrize("T", [1, 2, 7, 50, 127, 128, 129, 254, 255, 299, 320])
def test_flat_review_per_tile_equals_legacy_rel_shift(T):
    TOUT, TM, PAD = 127, 128, 320
    g = torch.Generator().manual_seed(T)
    bd = torch.randn(T, T, generator=g, dtype=torch.float64)
    ref = rel_shift_legacy(bd.view(1, 1, T, T))[0, 0]
    out = torch.full((T, T), float("nan"), dtype=torch.float64)
    for tile in range((T + TOUT - 1) // TOUT):
        r0 = tile * TOUT
        flat = torch.full((PAD + 129 * 320 + 128,), float("nan"), dtype=torch.float64)  # the kernel's buffer incl. front padding
        for r in range(TM):                      # thread = bd row r0 + r (rows >= T are TMA zero fill)
            off = r * (T + 1) + r0 - T           # position of the padded zero of this row
            assert off + PAD >= 0
            flat[PAD + off] = 0.0
            row = bd[r0 + r] if r0 + r < T else torch.zeros(T, dtype=torch.float64)
            flat[PAD + off + 1: PAD + off + 1 + T] = row  # columns >= T are never written
        for rp in range(TOUT):                   # thread = attention row r0 + rp
            i = r0 + rp
            if i < T:
                out[i] = flat[PAD + rp * T: PAD + rp * T + T]
    assert torch.equal(out, ref)


def _split(v):
    h = v.to(torch.bfloat16).to(torch.float32)
    return h, (v - h).to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("d", [128, 256])
def test_conv1_split_bf16_gemm_packing_reproduces_fp32_conv(d):
    """A row = [xh(9) | xl(9) | xh(9) | 1 | 1 | 0 0 0], B row = [wh(9) | wh(9) | wl(9) | bh | bl | 0 0 0] with bf16 entries and
    fp32 accumulation == bias + conv to ~2^-16 (the xl.wl term is dropped)."""
    g = torch.Generator().manual_seed(d)
    B, T, F = 2, 21, 17
    x = torch.randn(B, T, F, generator=g) * 3
    w = torch.randn(d, 9, generator=g) * 0.3
    b = torch.randn(d, generator=g) * 0.1
    ref = torch.nn.functional.conv2d(x.unsqueeze(1).double(), w.view(d, 1, 3, 3).double(), b.double(), stride=2)  # (B,d,T1,F1)
    T1, F1 = ref.shape[2], ref.shape[3]
    patches = torch.stack([x[:, kh:kh + 2 * T1 - 1:2, kw:kw + 2 * F1 - 1:2] for kh in range(3) for kw in range(3)], -1)  # (B,T1,F1,9)
    xh, xl = _split(patches)
    wh, wl = _split(w)
    bh, bl = _split(b)
    ones = torch.ones_like(xh[..., :1])
    A = torch.cat([xh, xl, xh, ones, ones, torch.zeros_like(xh[..., :3])], -1)                       # (..., 32)
    Bm = torch.cat([wh, wh, wl, bh[:, None], bl[:, None], torch.zeros(d, 3)], -1)                      # (d, 32)
    assert A.shape[-1] == 32 and Bm.shape[-1] == 32
    assert torch.equal(A, A.to(torch.bfloat16).float()) and torch.equal(Bm, Bm.to(torch.bfloat16).float())  # exactly bf16
    got = (A.double() @ Bm.double().t()).permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item()
    assert err <= 2.0 ** -14 * ref.abs().max().item(), err
    # weight-gradient packing: columns [xh(9) | 1 | ... | xl(9)] against the gradient rows give dW and dbias
    dy = torch.randn(B, T1, F1, d, generator=g).to(torch.bfloat16).float()
    cols = torch.cat([xh, ones, torch.zeros_like(xh[..., :6]), xl, torch.zeros_like(xh[..., :7])], -1)  # (..., 32): n<9 | 9 | 16+n
    D = torch.einsum("btfc,btfn->cn", dy.double(), cols.double())
    dw = D[:, :9] + D[:, 16:25]
    ref_dw = torch.einsum("btfc,btfn->cn", dy.double(), patches.double())
    assert (dw - ref_dw).abs().max().item() <= 2.0 ** -14 * ref_dw.abs().max().item()
    assert torch.allclose(D[:, 9], dy.double().sum((0, 1, 2)))
