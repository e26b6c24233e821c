"""GPU: dropout.  The reference's masks come from torch's global generator and cannot be replayed by fused
kernels, so parity is stated as: the reference arithmetic (the float64 oracle, whose dropout PLACEMENT is pinned against the
unmodified reference in tests/test_oracle_golden.py::test_dropout_placement...) under the SAME masks the kernels draw -- the
product's Philox4x32-7 stream, restated in numpy (oracle/philox_oracle.py, pinned by the Random123 known-answer vectors)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _state(seed=1234, step=7):
    return torch.tensor([seed, step], dtype=torch.int64, device=DEV)


@pytest.mark.parametrize("rows,cols", [(5, 8), (33, 257), (1000, 64), (7, 3), (129, 4233)])
@pytest.mark.parametrize("p", [0.1, 0.5])
def test_dropout_kernel_bit_exact_vs_numpy_philox(rows, cols, p):
    from liteasr_b200 import ops
    from oracle import philox_oracle as P
    st = _state()
    d = ops.Drop(st, 0x01000203, p)
    assert d.thr == P.threshold(p) and math.isclose(d.scale, P.scale_of(d.thr))
    x = torch.randn(rows, cols, device=DEV)
    y = ops.dropout(x, torch.empty_like(x), d)
    keep = torch.from_numpy(P.keep_mask(rows, cols, d.site, 1234, 7, p)).to(DEV)
    want = torch.where(keep, x * d.scale, torch.zeros_like(x))
    assert torch.equal(y, want)
    # fp32 -> bf16 (the CTC-head input cast) and a strided view of a padded buffer, in place
    yb = ops.dropout(x, torch.empty(rows, cols, device=DEV, dtype=torch.bfloat16), d)
    assert torch.equal(yb, want.bfloat16())
    ld = (cols + 7) // 8 * 8 + 8
    buf = torch.zeros(rows, ld, device=DEV)
    buf[:, :cols] = x
    ops.dropout(buf[:, :cols], buf[:, :cols], d)
    assert torch.equal(buf[:, :cols], want) and (buf[:, cols:] == 0).all()
    if rows * cols > 50000:  # keep rate = 1 - thr / 32768 within 4 sigma
        pe = d.thr / 32768.0
        assert abs(float(keep.float().mean()) - (1 - pe)) < 4 * math.sqrt(pe * (1 - pe) / (rows * cols))


def test_gpu_round_function_matches_random123_known_answers():
    """The CUDA round function (csrc/philox.cuh) against the Random123 kat_vectors at 10 rounds, and the 7-round variant the
    masks use against the numpy restatement."""
    from liteasr_b200 import ops
    from oracle import philox_oracle as P

    def i32(vals):
        return torch.tensor([v - (1 << 32) if v >= (1 << 31) else v for v in vals], dtype=torch.int32, device=DEV)

    for ctr, key, want in P.KAT:
        got10 = ops.philox_raw(i32(list(ctr) + list(key)), 10).cpu().numpy().view(np.uint32)
        assert [int(v) for v in got10] == list(want)
        got7 = ops.philox_raw(i32(list(ctr) + list(key)), 7).cpu().numpy().view(np.uint32)
        assert [int(v) for v in got7] == [int(v) for v in P.philox4x32(*ctr, *key)]


def test_rates_above_the_fp16_lane_limit_are_rejected():
    """15-bit lanes are compared as fp16 bit patterns: thr = round(p * 32768) must not exceed 0x7c00 (p <= 0.96875)."""
    from liteasr_b200 import ops
    st = _state()
    assert ops.Drop(st, 1, 0.96875).thr == 0x7C00
    with pytest.raises(ValueError):
        ops.Drop(st, 1, 0.97)
    x = torch.ones(64, 64, device=DEV)
    y = ops.dropout(x, torch.empty_like(x), ops.Drop(st, 1, 0.96875))
    assert 0 < int((y > 0).sum()) < 64 * 64 // 8 and float(y.max()) == 32.0   # 1 / (1 - 0.96875)


def test_rng_advance_and_site_separation():
    from liteasr_b200 import ops
    from liteasr_b200.dropout import RngState
    rs = RngState(torch.device("cuda:0"), seed=99)
    assert rs.host() == (99, 0)
    snap1 = rs.begin_pass()
    snap2 = rs.begin_pass()
    assert rs.host() == (99, 2) and snap1.tolist() == [99, 1] and snap2.tolist() == [99, 2]
    x = torch.ones(256, 256, device=DEV)
    a = ops.dropout(x, torch.empty_like(x), ops.Drop(snap1, 5, 0.3))
    a2 = ops.dropout(x, torch.empty_like(x), ops.Drop(snap1, 5, 0.3))
    b = ops.dropout(x, torch.empty_like(x), ops.Drop(snap2, 5, 0.3))
    c = ops.dropout(x, torch.empty_like(x), ops.Drop(snap1, 6, 0.3))
    assert torch.equal(a, a2)                                     # same (seed, step, site): same mask (what backward relies on)
    for other in (b, c):                                          # another step / another site: independent masks
        agree = float(((a > 0) == (other > 0)).float().mean())
        assert abs(agree - (0.7 * 0.7 + 0.3 * 0.3)) < 0.02


def _ref_linear(x, w, bias, act, alpha):
    h = x.float() @ w.float().t()
    if bias is not None:
        h = h + bias
    a = h
    if act == 1:
        a = torch.relu(h)
    elif act == 2:
        a = h * torch.sigmoid(h)
    return h, alpha * a


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("m,n,k,mode", [
    (300, 2048, 256, "swish_aux"),   # FFN fc1: Swish + saved pre-activation with markers (TMA-store epilogue in bf16)
    (300, 512, 256, "relu"),         # decoder FFN fc1
    (300, 256, 512, "res"),          # fc2 / linear_o / pointwise_conv2: fp32 out + residual, alpha 0.5 (column-phase epilogue)
    (300, 256, 256, "plain_f32"),    # sub-sampling Linear (alpha sqrt(d)) and the CTC-head input gradient
    (70, 250, 72, "plain_ragged"),   # ragged N and K: element-wise epilogue
])
def test_gemm_epilogue_dropout_matches_reference_under_same_mask(dtype, m, n, k, mode):
    from liteasr_b200 import ops
    from oracle import philox_oracle as P
    g = torch.Generator(device=DEV).manual_seed(m + n)
    x = (torch.randn(m, k, generator=g, device=DEV) * 0.5).to(dtype)
    w = (torch.randn(n, k, generator=g, device=DEV) * 0.1).to(dtype)
    bias = torch.randn(n, generator=g, device=DEV) * 0.1
    res = torch.randn(m, n, generator=g, device=DEV)
    st = _state(77, 3)
    p = 0.25
    d = ops.Drop(st, 0x02000104, p)
    keep = torch.from_numpy(P.keep_mask(m, n, d.site, 77, 3, p)).to(DEV)
    lo = dtype
    ld = (n + 7) // 8 * 8
    if mode == "swish_aux":
        c = torch.empty(m, ld, device=DEV, dtype=lo)[:, :n]
        aux = torch.empty(m, ld, device=DEV, dtype=lo)[:, :n]
        # (a) the pre-activation variant: dropped elements of the saved tensor carry the marker whose swish'() is exactly 0
        ops.gemm(x, w, c, m, n, k, lda=k, ldb=k, ldc=ld, bias=bias, aux=aux, act=2, drop=d, drop_mark_aux=True)
        h, a = _ref_linear(x, w, bias, 2, 1.0)
        want = torch.where(keep, a * d.scale, torch.zeros_like(a))
        assert (aux.float()[~keep] < -1e29).all(), "dropped elements of the saved pre-activation must carry the marker"
        tol = 2e-2 if dtype == torch.bfloat16 else 1e-5
        assert torch.allclose(aux.float()[keep], h[keep], rtol=tol, atol=tol)
        dy = (torch.randn(m, k, generator=g, device=DEV) * 0.5).to(dtype)
        wt = (torch.randn(k, n, generator=g, device=DEV) * 0.1).to(dtype)   # W2 (K_in = n columns): dh = dy @ W2
        dh = torch.empty(m, ld, device=DEV, dtype=lo)[:, :n]
        ops.gemm(dy, wt, dh, m, n, k, lda=k, ldb=n, ldc=ld, tb=True, alpha=0.5 * d.scale, dact=aux, act=2)
        assert (dh.float()[~keep] == 0).all() and torch.isfinite(dh.float()).all()
        hs = h[keep]
        sg = torch.sigmoid(hs)
        want_dh = (dy.float() @ wt.float())[keep] * (0.5 * d.scale) * (sg * (1 + hs * (1 - sg)))
        btol = 3e-2 if dtype == torch.bfloat16 else 1e-4
        assert torch.allclose(dh.float()[keep], want_dh, rtol=btol, atol=btol * float(want_dh.abs().max()))
        # (b) the derivative variant the engine uses (aux_deriv): aux = swish'(h), 0 where dropped; backward = one multiply
        c2 = torch.empty(m, ld, device=DEV, dtype=lo)[:, :n]
        aux2 = torch.empty(m, ld, device=DEV, dtype=lo)[:, :n]
        ops.gemm(x, w, c2, m, n, k, lda=k, ldb=k, ldc=ld, bias=bias, aux=aux2, act=2, drop=d, drop_mark_aux=True, aux_deriv=True)
        assert torch.equal(c2, c)
        sga = torch.sigmoid(h)
        assert (aux2.float()[~keep] == 0).all()
        assert torch.allclose(aux2.float()[keep], (sga * (1 + h * (1 - sga)))[keep], rtol=tol, atol=tol)
        dh2 = torch.empty(m, ld, device=DEV, dtype=lo)[:, :n]
        ops.gemm(dy, wt, dh2, m, n, k, lda=k, ldb=n, ldc=ld, tb=True, alpha=0.5 * d.scale, dact=aux2, act=3)
        assert (dh2.float()[~keep] == 0).all()
        assert torch.allclose(dh2.float()[keep], want_dh, rtol=btol, atol=btol * float(want_dh.abs().max()))
    elif mode == "relu":
        c = torch.empty(m, ld, device=DEV, dtype=lo)[:, :n]
        ops.gemm(x, w, c, m, n, k, lda=k, ldb=k, ldc=ld, bias=bias, act=1, drop=d)
        _, a = _ref_linear(x, w, bias, 1, 1.0)
        want = torch.where(keep, a * d.scale, torch.zeros_like(a))
    elif mode == "res":
        c = torch.empty(m, n, device=DEV)
        ops.gemm(x, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias, res=res, ldres=n, alpha=0.5, drop=d)
        _, a = _ref_linear(x, w, bias, 0, 0.5)
        want = torch.where(keep, a * d.scale, torch.zeros_like(a)) + res   # the residual is added AFTER the mask
    elif mode == "plain_f32":
        c = torch.empty(m, n, device=DEV)
        ops.gemm(x, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias, alpha=16.0, drop=d)
        _, a = _ref_linear(x, w, bias, 0, 16.0)
        want = torch.where(keep, a * d.scale, torch.zeros_like(a))
    else:
        c = torch.empty(m, ld, device=DEV, dtype=lo)[:, :n]
        ops.gemm(x, w, c, m, n, k, lda=k, ldb=k, ldc=ld, bias=bias, drop=d)
        _, a = _ref_linear(x, w, bias, 0, 1.0)
        want = torch.where(keep, a * d.scale, torch.zeros_like(a))
    if mode == "res":
        assert torch.allclose(c, want, rtol=2e-2 if dtype == torch.bfloat16 else 1e-5, atol=2e-2 if dtype == torch.bfloat16 else 1e-5)
    else:
        assert (c.float()[~keep] == 0).all()
        tol = 3e-2 if dtype == torch.bfloat16 else 1e-5
        assert torch.allclose(c.float(), want, rtol=tol, atol=tol * float(want.abs().max()))


@pytest.mark.parametrize("lo_dtype", [torch.bfloat16, torch.float32])
def test_layernorm_bwd_masks_only_the_fused_outputs(lo_dtype):
    from liteasr_b200 import ops
    from oracle import philox_oracle as P
    rows, d = 333, 256
    g = torch.Generator(device=DEV).manual_seed(4)
    x = torch.randn(rows, d, generator=g, device=DEV)
    dy = torch.randn(rows, d, generator=g, device=DEV).bfloat16()
    gamma = 1 + 0.1 * torch.randn(d, generator=g, device=DEV)
    beta = torch.zeros(d, device=DEV)
    y = torch.empty(rows, d, device=DEV, dtype=torch.bfloat16)
    mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
    ops.layernorm_fwd(x, gamma, beta, y, mean, rstd, 1e-12)
    old = torch.randn(rows, d, generator=g, device=DEV)

    def run(drop):
        dx = old.clone()
        lo = torch.empty(rows, d, device=DEV, dtype=lo_dtype)
        cs = torch.zeros(d, device=DEV)
        dg, db = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
        ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, dg, db, True, dx_lo=lo, colsum=cs, colsum_scale=0.5, drop=drop)
        return dx, lo, cs, dg, db

    dx0, lo0, cs0, dg0, db0 = run(None)
    st = _state(5, 11)
    dr = ops.Drop(st, 0x01000304, 0.1)
    dx1, lo1, cs1, dg1, db1 = run(dr)
    keep = torch.from_numpy(P.keep_mask(rows, d, dr.site, 5, 11, 0.1)).to(DEV)
    assert torch.equal(dx0, dx1)   # the residual-stream gradient is NOT masked
    assert torch.allclose(dg0, dg1, rtol=1e-5, atol=1e-4) and torch.allclose(db0, db1, rtol=1e-5, atol=1e-4)  # red.add order varies
    want_lo = torch.where(keep, dx1 * dr.scale, torch.zeros_like(dx1))
    assert torch.equal(lo1, want_lo.to(lo_dtype))
    assert torch.allclose(cs1, 0.5 * want_lo.sum(0), rtol=1e-4, atol=1e-3)
    assert torch.equal(lo0, dx0.to(lo_dtype))


def _tiny(rates_kw, precision, seed=21, **over):
    from liteasr_b200.criterions.hybrid_ctc_attn import HybridCTCLoss, HybridCTCLossConfig
    from liteasr_b200.models.u2 import U2, U2Config
    from liteasr_b200.schema import U2Dims
    from liteasr_b200.utils.synthetic import synth_batch, synth_state_dict
    kw = dict(input_dim=80, vocab_size=60, enc_dim=64, enc_ff_dim=128, enc_attn_heads=2, enc_layers=2, dec_dim=64, dec_ff_dim=128,
              dec_attn_heads=2, dec_layers=2)
    kw.update(over)
    dims = U2Dims(**kw)
    batch = synth_batch(4, 140, 7, dims.vocab_size, seed=seed)
    sd = synth_state_dict(dims, seed=seed)
    model = U2(U2Config(**kw, precision=precision, **rates_kw))
    model.load_state_dict(sd)
    crit = HybridCTCLoss(HybridCTCLossConfig(vocab_size=dims.vocab_size, smoothing=0.1, ctc_weight=0.3))
    return kw, batch, sd, model.cuda(), crit


def _oracle(kw, batch, sd, rates, seed, step, training):
    from oracle import u2_oracle as O
    xs, xlens, ys, ylens = batch
    sd64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k and ".pe.pe" not in k
                else (v.double() if v.is_floating_point() else v)) for k, v in sd.items()}
    dp = O.Dropper(rates, seed, step, training=training)
    out = O.hybrid_loss(sd64, O.U2Shape(**kw), xs.double(), xlens, ys, ylens, 0.3, 0.1, training, {}, dp)
    if training:
        out["loss"].backward()
    return sd64, out


RATES = dict(dropout_rate=0.1, enc_attn_dropout_rate=0.0, dec_self_attn_dropout_rate=0.0, dec_src_attn_dropout_rate=0.0)   # my_U2.yaml
RATES_ALL = dict(dropout_rate=0.1, enc_attn_dropout_rate=0.15, dec_self_attn_dropout_rate=0.2, dec_src_attn_dropout_rate=0.05,
                 enc_pos_dropout_rate=0.2, dec_pos_dropout_rate=0.3, enc_ff_dropout_rate=0.25, dec_ff_dropout_rate=0.05)


def _oracle_rates(model_cfg_kw):
    from oracle import u2_oracle as O
    p = model_cfg_kw.get("dropout_rate", 0.0)
    r = O.DropRates.uniform(p, p)
    for k, v in model_cfg_kw.items():
        setattr(r, k, v)
    return r


@pytest.mark.parametrize("rates", [RATES, RATES_ALL], ids=["my_U2_yaml", "all_ten_sites"])
def test_train_step_with_dropout_fp32_matches_oracle_under_same_masks(rates):
    """fp32 mode, every site live: loss rel 1e-5, gradients max-abs 1e-4 * max|g| (SURVEY 8d tolerances) against the float64
    oracle replaying the product's masks.  Also proves forward/backward mask agreement at every site: a backward kernel that
    regenerated a different mask than its forward twin would miss these bounds by orders of magnitude."""
    from liteasr_b200 import functions as F
    kw, batch, sd, model, crit = _tiny(rates, "fp32")
    model.train()
    st, _, _ = F.bind(model, torch.device("cuda:0"))
    st.rng.seed(4242, 10)
    gb = [t.cuda() for t in batch]
    loss = crit(model, *gb)
    loss.backward()
    seed, step = st.rng.host()
    assert (seed, step) == (4242, 11)
    sd64, out = _oracle(kw, batch, sd, _oracle_rates(rates), seed, step, True)
    assert math.isclose(float(loss), float(out["loss"]), rel_tol=1e-5), (float(loss), float(out["loss"]))
    gmax = max(float(p.grad.abs().max()) for p in sd64.values() if getattr(p, "grad", None) is not None)
    for n, p in model.named_parameters():
        err = float((p.grad.double().cpu() - sd64[n].grad).abs().max())
        assert err <= 1e-4 * gmax, (n, err, gmax)
    # a second call draws new masks (step advanced) and therefore another loss
    model.load_state_dict(sd)
    loss2 = crit(model, *gb)
    assert st.rng.host()[1] == 12 and abs(float(loss2) - float(loss)) > 1e-5 * abs(float(loss))  # same masks would give the same bits


def test_train_step_with_dropout_bf16_within_tolerance():
    from liteasr_b200 import functions as F
    kw, batch, sd, model, crit = _tiny(RATES_ALL, "bf16")
    model.train()
    st, _, _ = F.bind(model, torch.device("cuda:0"))
    st.rng.seed(7, 0)
    loss = crit(model, *[t.cuda() for t in batch])
    loss.backward()
    sd64, out = _oracle(kw, batch, sd, _oracle_rates(RATES_ALL), *st.rng.host(), True)
    assert math.isclose(float(loss), float(out["loss"]), rel_tol=5e-3)
    rels = []
    for n, p in model.named_parameters():
        b = sd64[n].grad
        if float(b.abs().max()) > 1e-6:
            rels.append(float((p.grad.double().cpu() - b).norm() / b.norm()))
    rels.sort()
    assert rels[len(rels) // 2] < 3e-2 and rels[-1] < 0.3, (rels[len(rels) // 2], rels[-1])


def test_eval_mode_keeps_only_the_ctc_head_dropout():
    """Quirk Q3: ``F.dropout`` in nets/ctc.py:29 has no ``training=`` argument, so valid() still drops the CTC head's input;
    every nn.Dropout is the identity in eval()."""
    from liteasr_b200 import functions as F
    kw, batch, sd, model, crit = _tiny(RATES, "fp32")
    model.eval()
    st, _, _ = F.bind(model, torch.device("cuda:0"))
    st.rng.seed(31, 5)
    with torch.no_grad():
        loss = crit(model, *[t.cuda() for t in batch])
    _, out = _oracle(kw, batch, sd, _oracle_rates(RATES), *st.rng.host(), False)
    assert math.isclose(float(loss), float(out["loss"]), rel_tol=1e-5)
    _, out0 = _oracle(kw, batch, sd, _oracle_rates(dict(dropout_rate=0.0)), 0, 0, False)
    assert abs(float(out0["loss"]) - float(loss)) > 1e-4 * abs(float(loss))   # the CTC-head mask did change the loss


def test_zero_rates_launch_no_rng_kernels_and_match_the_dropout_free_model():
    from liteasr_b200 import _lib
    from liteasr_b200 import functions as F
    kw, batch, sd, model, crit = _tiny({}, "fp32")
    model.train()
    st, _, _ = F.bind(model, torch.device("cuda:0"))
    gb = [t.cuda() for t in batch]
    l0 = float(crit(model, *gb))
    assert st.rng.host()[1] == 0, "no pass may advance the stream when every rate is 0"
    kw2, _, _, model2, _ = _tiny(dict(dropout_rate=0.0, enc_ff_dropout_rate=0.0), "fp32")
    model2.train()
    assert float(crit(model2, *gb)) == l0


def test_graph_step_draws_fresh_masks_on_every_replay():
    from liteasr_b200.optims import FusedAdam, AdamConfig
    from liteasr_b200.trainer import TrainStep
    kw, batch, sd, model, crit = _tiny(RATES, "bf16")
    model.train()
    step = TrainStep(model, crit, device=torch.device("cuda:0"))
    step.optimizer = FusedAdam(step.store, AdamConfig(lr=0.0))  # frozen weights: the loss varies only through the masks
    gb = [t.cuda() for t in batch]
    losses = [float(step(*gb)) for _ in range(4)]
    assert len(set(losses)) == 4, losses
    s0 = step.store.rng.host()[1]
    step(*gb)
    assert step.store.rng.host()[1] == s0 + 1
