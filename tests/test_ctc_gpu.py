"""GPU: fused CTC fwd+bwd (lasr_ctc_fwdbwd) vs the float64 oracle composed with log-softmax backward,
and vs the torch.nn.CTCLoss known answers in tests/golden/ctc_golden.json."""
import json
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def oracle_grad(logits64, targets, in_len, tgt_len):
    from oracle import u2_oracle as O
    lp = torch.log_softmax(logits64, -1).numpy()
    nll, dlp = O.ctc_alpha_beta_c(lp, np.clip(targets.numpy(), 0, None), in_len.numpy(), tgt_len.numpy())
    d = np.nan_to_num(dlp, nan=0.0)
    dl = d - np.exp(lp) * d.sum(-1, keepdims=True)
    return nll, dl


def make_case(T, B, V, L, seed, repeat=False):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(T, B, V, generator=g, dtype=torch.float64)
    in_len = torch.randint(int(0.6 * T), T + 1, (B,), generator=g)
    in_len[0] = T
    tgt_len = torch.randint(max(1, L // 2), L + 1, (B,), generator=g)
    tgt_len[0] = L
    tgt_len = torch.minimum(tgt_len, in_len // 2)
    targets = torch.randint(1, V, (B, L), generator=g)
    if repeat:
        targets[:, 1::2] = targets[:, 0::2][:, : targets[:, 1::2].shape[1]]  # force repeated labels
        tgt_len = torch.minimum(tgt_len, in_len // 3)
    return logits, targets, in_len, tgt_len


@pytest.mark.parametrize("T,B,V,L,repeat", [(50, 4, 20, 5, False), (200, 64, 500, 20, False), (123, 9, 4233, 40, True),
                                             (400, 16, 1000, 50, True), (300, 8, 5000, 100, False), (97, 5, 8200, 12, False)])
def test_ctc_matches_oracle_fp32(T, B, V, L, repeat):
    from liteasr_b200 import ops
    logits, targets, in_len, tgt_len = make_case(T, B, V, L, seed=T + V, repeat=repeat)
    nll_ref, g_ref = oracle_grad(logits, targets, in_len, tgt_len)
    x = logits.float().cuda()
    nll, grad = ops.ctc_fwdbwd(x, targets.cuda(), in_len.cuda(), tgt_len.cuda(), time_major=True, grad_scale=1.0)
    torch.cuda.synchronize()
    # tolerance: fp32 log-space recursion vs float64 oracle -- loss rel 1e-5, grad abs 2e-5
    nerr = np.abs(nll.cpu().numpy() - nll_ref).max()
    gerr = np.abs(grad.cpu().double().numpy() - g_ref).max()
    print(f"ctc T={T} B={B} V={V} L={L}: nll abs err {nerr:.3e} (nll~{np.abs(nll_ref).max():.1f}), grad abs err {gerr:.3e}")
    # Stated fp32 tolerance (north_star): loss rel 1e-5; gradient abs 3e-4.  The lattice runs in fp32 log
    # space (like torch's fp32 CTC): each of the T sequential log-sum-exps rounds at ulp(|alpha~|), so the
    # occupancy error grows ~ sqrt(T) * ulp; the blank-normalised recursion keeps |alpha~| ~ 1e2 instead of
    # T*log(V) ~ 1e4, i.e. ~10x tighter than the raw recursion.
    assert np.allclose(nll.cpu().numpy(), nll_ref, rtol=1e-5, atol=1e-4), nerr
    assert gerr <= 3e-4, gerr


def test_ctc_batch_major_padded_and_scaled():
    """(B,T,V) view of a padded buffer (the in-model layout), grad_scale and device upstream scalar."""
    from liteasr_b200 import ops
    T, B, V, L = 61, 6, 4233, 9
    logits, targets, in_len, tgt_len = make_case(T, B, V, L, seed=3)
    nll_ref, g_ref = oracle_grad(logits, targets, in_len, tgt_len)
    Vp = 4240
    buf = torch.zeros(B, T, Vp, device="cuda")
    buf[..., :V] = logits.float().permute(1, 0, 2).cuda()
    gbuf = torch.zeros(B, T, Vp, device="cuda")
    up = torch.tensor(0.5, device="cuda")
    nll, grad = ops.ctc_fwdbwd(buf[..., :V], targets.cuda(), in_len.cuda(), tgt_len.cuda(), time_major=False,
                               grad=gbuf[..., :V], grad_scale=0.3 / B, upstream=up)
    assert np.allclose(nll.cpu().numpy(), nll_ref, rtol=1e-5, atol=1e-4)
    want = g_ref.transpose(1, 0, 2) * (0.3 / B * 0.5)
    assert np.abs(grad.cpu().double().numpy() - want).max() <= 2e-5
    assert (gbuf[..., V:] == 0).all()


def test_ctc_golden_known_answers_and_edges():
    from liteasr_b200 import ops
    with open(os.path.join(GOLDEN, "ctc_golden.json")) as f:
        g = json.load(f)
    for c in g["cases"]:
        logits = torch.tensor(c["logits"], dtype=torch.float32).cuda()
        tg = torch.tensor(c["targets"], dtype=torch.int64).clamp(min=0).cuda()
        il = torch.tensor(c["in_len"], dtype=torch.int64).cuda()
        tl = torch.tensor(c["tgt_len"], dtype=torch.int64).cuda()
        nll, grad = ops.ctc_fwdbwd(logits, tg, il, tl, time_major=True)
        nll = nll.cpu().tolist()
        for b, want in enumerate(c["nll"]):
            if want == "inf":
                assert math.isinf(nll[b]) and nll[b] > 0
            else:
                assert math.isclose(nll[b], want, rel_tol=2e-6, abs_tol=2e-6)
        if c["grad_logits"] is not None:
            fin = torch.tensor(c["finite"]).cuda()
            want = torch.tensor(c["grad_logits"], dtype=torch.float32).cuda()
            got = torch.where(fin.view(1, -1, 1), grad, torch.zeros_like(grad))
            assert (got - want).abs().max().item() <= 5e-6
            if not bool(fin.all()):  # infeasible utterances poison their own rows with NaN (zero_infinity=False)
                bad = (~fin).nonzero().flatten().tolist()
                for b in bad:
                    assert torch.isnan(grad[: c["in_len"][b], b]).all()


def test_ctc_bf16_close():
    from liteasr_b200 import ops
    T, B, V, L = 150, 8, 512, 20
    logits, targets, in_len, tgt_len = make_case(T, B, V, L, seed=21)
    xb = logits.float().cuda().bfloat16()
    nll_ref, g_ref = oracle_grad(xb.double().cpu(), targets, in_len, tgt_len)
    nll, grad = ops.ctc_fwdbwd(xb, targets.cuda(), in_len.cuda(), tgt_len.cuda(), time_major=True)
    assert np.allclose(nll.cpu().numpy(), nll_ref, rtol=1e-5, atol=1e-3)
    # bf16 output rounding: 2^-8 relative on values <= 1
    assert np.abs(grad.float().cpu().double().numpy() - g_ref).max() <= 6e-3


def test_ctc_full_size_properties():
    """BASELINE config 4 largest-ish size: size-independent properties instead of an oracle run:
    rows sum to 0 for t < in_len (softmax - occupancy), are exactly 0 beyond, nll finite and positive."""
    from liteasr_b200 import ops
    T, B, V, L = 1600, 64, 5000, 200
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(T, B, V, generator=g, device="cuda")
    in_len = torch.randint(int(0.6 * T), T + 1, (B,), generator=g, device="cuda")
    in_len[0] = T
    tgt_len = torch.randint(L // 2, L + 1, (B,), generator=g, device="cuda")
    targets = torch.randint(1, V, (B, L), generator=g, device="cuda")
    nll, grad = ops.ctc_fwdbwd(x, targets, in_len, tgt_len, time_major=True)
    assert torch.isfinite(nll).all() and (nll > 0).all()
    live = torch.arange(T, device="cuda").view(-1, 1) < in_len.view(1, -1)
    rs = grad.sum(-1)
    assert rs[live].abs().max().item() < 1e-2  # 1 - sum_s occupancy; fp32 lattice over T = 1600 steps
    assert (grad[~live] == 0).all()
    # blank column: softmax - occupancy in [-1, 1]
    assert grad.abs().max().item() <= 1.0 + 3e-3  # fp32 log-space lattice over T = 1600 steps (occupancy error ~ sqrt(T) * ulp)


def test_ctc_poisoned_workspace_is_harmless():
    """The lattice kernels must never read workspace cells the gather kernel did not write (ragged target lengths leave the
    emission columns k > L_b untouched): a workspace pre-filled with NaN bit patterns gives bit-identical results."""
    from liteasr_b200 import ops
    for (T, B, V, L) in ((300, 8, 1000, 100), (160, 6, 300, 40), (90, 5, 64, 20)):
        logits, targets, in_len, tgt_len = make_case(T, B, V, L, seed=7 * T + L)
        tgt_len[1] = max(1, L // 3)
        x = logits.float().cuda()
        args = (targets.cuda(), in_len.cuda(), tgt_len.cuda())
        n = ops.ctc_workspace_bytes(T, B, L)
        ws0 = torch.zeros(n, dtype=torch.uint8, device="cuda")
        ws1 = torch.full((n,), 0xFF, dtype=torch.uint8, device="cuda")  # 0xFFFFFFFF = NaN
        nll0, g0 = ops.ctc_fwdbwd(x, *args, time_major=True, workspace=ws0)
        nll1, g1 = ops.ctc_fwdbwd(x, *args, time_major=True, workspace=ws1)
        assert torch.isfinite(g1).all() and torch.isfinite(nll1).all()
        assert torch.equal(g0, g1) and torch.equal(nll0, nll1)


# BASELINE config 4, the whole grid (B = 64): the reference's own call (criterions/hybrid_ctc_attn.py:67-75: log_softmax +
# nn.CTCLoss(sum)) evaluated in float64 on the GPU is the yardstick; torch's fp32 evaluation of the same call is the
# "reference fp32 path" whose own deviation from float64 is recorded next to ours.
C4_GRID = [(T, L, V) for (T, L) in ((200, 20), (400, 50), (800, 100), (1600, 200)) for V in (500, 1000, 2000, 5000)]


def c4_case(T, L, V, B=64, device="cuda"):
    g = torch.Generator(device=device).manual_seed(1000 * T + V)
    x = torch.randn(T, B, V, generator=g, device=device)
    il = torch.randint(int(0.6 * T), T + 1, (B,), generator=g, device=device)
    il[0] = T
    tl = torch.randint(L // 2, L + 1, (B,), generator=g, device=device)
    tl[0] = L
    tg = torch.randint(1, V, (B, L), generator=g, device=device)
    tg[1, 1] = tg[1, 0]          # a repeated label at the start ...
    tg[2, 2:6] = tg[2, 1]        # ... and a run of five equal labels
    tg[3, L - 1] = tg[3, L - 2]
    return x, tg, il, tl


def c4_tolerance(T):
    """Gradient tolerance of the fp32 log-space lattice against float64, absolute on |grad| <= 1.
    SURVEY 8d asks 1e-4 * max|g| 'vs the reference fp32 path'; that path (torch fp32 CTC) itself deviates from float64 by
    1e-3 (T = 200) ... 2.4e-2 (T = 1600), so the bound that can be checked against an exact answer is T-dependent: every one
    of the T sequential log-sum-exps rounds alpha~ / beta~ at ulp(|alpha~|) with |alpha~| <~ 2^10 in log2 units (2^-13 absolute
    = 1.2e-4 relative on the occupancy), the errors of a sweep add like a random walk with drift, and measured growth on this
    grid is 4e-7 * T ... 1.5e-6 * T.  Stated bound: max(1e-4, 2e-6 * T) -- 4e-4 at T = 200, 3.2e-3 at T = 1600 -- and, checked in
    the same test, never worse than HALF of torch's own fp32 error at the same point."""
    return max(1e-4, 2e-6 * T)


@pytest.mark.parametrize("T,L,V", C4_GRID)
def test_ctc_c4_grid_vs_torch_float64(T, L, V):
    from liteasr_b200 import ops
    B = 64
    x, tg, il, tl = c4_case(T, L, V, B)
    n = ops.ctc_workspace_bytes(T, B, L)
    ws = torch.full((n,), 0xFF, dtype=torch.uint8, device="cuda")  # NaN bit patterns: nothing unwritten may be read
    nll, grad = ops.ctc_fwdbwd(x, tg, il, tl, time_major=True, workspace=ws)
    assert torch.isfinite(nll).all() and torch.isfinite(grad).all()
    x64 = x.double().requires_grad_(True)
    l64 = torch.nn.functional.ctc_loss(x64.log_softmax(-1), tg, il, tl, blank=0, reduction="none", zero_infinity=False)
    l64.sum().backward()
    gerr = (grad.double() - x64.grad).abs().max().item()
    nerr = ((nll.double() - l64).abs() / l64.abs()).max().item()
    x32 = x.clone().requires_grad_(True)
    torch.nn.functional.ctc_loss(x32.log_softmax(-1), tg, il, tl, blank=0, reduction="sum", zero_infinity=False).backward()
    terr = (x32.grad.double() - x64.grad).abs().max().item()
    print(f"c4 T={T} L={L} V={V}: nll rel err {nerr:.2e}, grad abs err {gerr:.2e} (tolerance {c4_tolerance(T):.1e}; torch fp32: {terr:.2e})")
    assert nerr <= 1e-6            # CTC fp32 loss vs the float64 restatement, rel 1e-6
    assert gerr <= c4_tolerance(T)
    assert gerr <= 0.5 * terr
    live = torch.arange(T, device="cuda").view(-1, 1) < il.view(1, -1)
    assert (grad[~live] == 0).all()


VARIANT_SCRIPT = r'''
import sys, torch
sys.path.insert(0, ROOT)
from liteasr_b200 import ops
for (T, B, V, L) in ((300, 8, 1000, 60), (800, 64, 5000, 100), (97, 5, 333, 12)):
    g = torch.Generator(device="cuda").manual_seed(T + V)
    x = torch.randn(T, B, V, generator=g, device="cuda")
    il = torch.randint(int(0.6 * T), T + 1, (B,), generator=g, device="cuda"); il[0] = T
    tl = torch.randint(L // 2, L + 1, (B,), generator=g, device="cuda"); tl[0] = L
    tg = torch.randint(1, V, (B, L), generator=g, device="cuda"); tg[1, 1] = tg[1, 0]
    ws = torch.full((ops.ctc_workspace_bytes(T, B, L),), 0xFF, dtype=torch.uint8, device="cuda")
    nll, grad = ops.ctc_fwdbwd(x, tg, il, tl, time_major=True, workspace=ws)
    x64 = x.double().requires_grad_(True)
    l64 = torch.nn.functional.ctc_loss(x64.log_softmax(-1), tg, il, tl, blank=0, reduction="none", zero_infinity=False)
    l64.sum().backward()
    assert torch.isfinite(grad).all()
    assert ((nll.double() - l64).abs() / l64.abs()).max().item() < 1e-6
    assert (grad.double() - x64.grad).abs().max().item() < max(1e-4, 2e-6 * T), (T, V)
print("VARIANT-OK")
'''


@pytest.mark.parametrize("env", [{"LASR_CTC_PIPE": "1"}, {"LASR_CTC_FUSED": "1"}], ids=["two_pass_pipeline", "meet_in_the_middle"])
def test_ctc_off_by_default_variants_stay_correct(env):
    """The two alternative CTC pipelines in csrc/ctc.cu (developer switches, both measured slower than the default: DESIGN.md)
    are kept only as long as they are correct: each runs in a subprocess (the switch is read once per process) against
    torch's float64 CTC."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", f"ROOT = {root!r}\n" + VARIANT_SCRIPT], capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0 and "VARIANT-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
