"""GPU: oracle parity AT THE BENCH SHAPE -- the whole BASELINE configs[1] model (12 Conformer layers d = 256 H = 4 + 6 decoder
layers, V = 4233), Tmax = 1200 (T' = 299), Lmax = 40, per-GPU batch 126 (float64 oracle) and 252 (the bench's batch; float32
oracle, see the second test): the step ``bench.py`` times, through the call it times
(``HybridCTCLoss.direct_step`` -> ``functions.hybrid_direct_step``: the fused forward + hand-written backward).  At this size the float64 CPU oracle would take minutes, so the SAME oracle functions (oracle/u2_oracle.py:
plain torch ops, pinned against the unmodified reference in tests/test_oracle_golden.py) are evaluated in float64 ON THE GPU,
with the reference's own CTC call (log_softmax + torch ctc_loss(sum), criterions/hybrid_ctc_attn.py:67-75) in float64 as in
tests/test_ctc_gpu.py::test_ctc_c4_grid_vs_torch_float64.  Dropout 0 (mask parity: tests/test_dropout_gpu.py).

Tolerances (stated here; measured values on a B200 in brackets, printed with -s):
  fp32 mode : loss rel 1e-5 [3e-8]; gradient max-abs error <= 1e-4 * max|g| over the whole flat gradient [2.8e-6] and
              whole-gradient rel-L2 <= 1e-4 [3.4e-6] (12 layers deep, 37 674 rows reduced per weight gradient);
  bf16 mode : loss rel 3e-3 [1.9e-5], whole-gradient rel-L2 <= 3e-2 [3.3e-3], median per-parameter rel-L2 <= 2e-2 [3.6e-3]
              (bf16 operands, fp32 accumulate).
The float64 oracle needs 100 GB of HBM for this shape (freed before the product runs); the whole file takes 10 s on a B200."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

C2 = dict(input_dim=80, vocab_size=4233, enc_dim=256, enc_ff_dim=2048, enc_attn_heads=4, enc_layers=12, dec_dim=256,
          dec_ff_dim=2048, dec_attn_heads=4, dec_layers=6)
TMAX, LMAX, W, SM = 1200, 40, 0.3, 0.1


def _oracle_on_gpu(BATCH: int, odt: torch.dtype):
    """loss and per-parameter gradients of the oracle evaluated in ``odt`` on the GPU, moved to the host (its activations --
    100 GB in float64 at batch 126 -- are released before the product runs)."""
    from liteasr_b200.schema import U2Dims
    from liteasr_b200.utils.synthetic import synth_batch, synth_state_dict
    from oracle import u2_oracle as O
    dev = torch.device("cuda:0")
    torch.cuda.reset_peak_memory_stats(0)
    dims = U2Dims(**C2)
    batch = synth_batch(BATCH, TMAX, LMAX, dims.vocab_size, seed=42)
    sd = synth_state_dict(dims, seed=7)
    sd64 = {k: (v.to(odt).to(dev).requires_grad_(True) if v.is_floating_point() and "running" not in k and ".pe.pe" not in k
                else (v.to(odt).to(dev) if v.is_floating_point() else v.to(dev))) for k, v in sd.items()}
    xs, xlens, ys, ylens = (t.to(dev) for t in batch)
    cfg = O.U2Shape(**C2)
    with torch.device(dev):  # the oracle's factory calls (arange / zeros / full: masks, position tables) follow its inputs to the GPU
        h_attn, h_ctc, _ = O.u2_forward(sd64, cfg, xs.to(odt), xlens, ys, ylens, True, {}, None)
        la = O.label_smoothing_kl(h_attn, O.attention_targets(ys, ylens, cfg.vocab_size), SM) / BATCH
        lp = torch.log_softmax(h_ctc.transpose(0, 1).double(), dim=-1)  # the CTC itself always in float64 (torch's float32 CTC is the noisy one)
        lc = torch.nn.functional.ctc_loss(lp, ys.clamp(min=0), O.subsampled_len(xlens), ylens, blank=0, reduction="sum") / BATCH
        loss = W * lc + (1 - W) * la
        loss.backward()
    print(f"\n{odt} oracle on the GPU, batch {BATCH}: peak memory {torch.cuda.max_memory_allocated(dev) / 1e9:.1f} GB")
    out = dict(loss=float(loss), ctc=float(lc), att=float(la),
               grads={k: v.grad.detach().double().cpu() for k, v in sd64.items() if getattr(v, "grad", None) is not None})
    del sd64, h_attn, h_ctc, la, lp, lc, loss
    torch.cuda.empty_cache()
    return sd, batch, out


def _or_skip(batch, odt):
    import gc
    gc.collect()
    torch.cuda.init()
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info(0)
    if free < 125e9:  # the oracle's autograd graph: 100 GB (float64, batch 126) / 103 GB (float32, batch 252), then the product
        pytest.skip(f"needs 125 GB of free HBM for the oracle at the bench shape ({free / 1e9:.0f} GB free)")
    try:
        return _oracle_on_gpu(batch, odt)
    except torch.cuda.OutOfMemoryError:
        torch.cuda.empty_cache()
        pytest.skip("out of device memory while evaluating the oracle at the bench shape")


@pytest.fixture(scope="module")
def oracle_f64():
    return _or_skip(126, torch.float64)


@pytest.fixture(scope="module")
def oracle_f32_b252():
    return _or_skip(252, torch.float32)


def _run_product(case, precision, BATCH):
    from liteasr_b200.criterions.hybrid_ctc_attn import HybridCTCLoss, HybridCTCLossConfig
    from liteasr_b200.models.u2 import U2, U2Config
    from liteasr_b200.trainer import TrainStep
    sd, batch, ref = case
    dev = torch.device("cuda:0")
    model = U2(U2Config(**{**C2, "precision": precision}))
    model.load_state_dict(sd)
    model = model.to(dev).train()
    crit = HybridCTCLoss(HybridCTCLossConfig(vocab_size=C2["vocab_size"], smoothing=SM, ctc_weight=W))
    step = TrainStep(model, crit, use_graph=False, device=dev)  # direct-gradient store: the path bench.py times
    b = tuple(t.to(dev) for t in batch)
    step.store.zero_grads()
    loss = crit.direct_step(model, *b)
    assert loss is not None, "the fused step did not take the direct path"
    torch.cuda.synchronize()
    parts = model.last_losses.tolist()
    gmax = max(float(g.abs().max()) for g in ref["grads"].values())
    num = den = 0.0
    worst_abs, rels = (0.0, ""), []
    for n, p in model.named_parameters():
        want = ref["grads"][n]
        got = p.grad.double().cpu()
        assert torch.isfinite(got).all(), n
        d = got - want
        num += float(d.pow(2).sum())
        den += float(want.pow(2).sum())
        m = float(d.abs().max())
        if m > worst_abs[0]:
            worst_abs = (m, n)
        if float(want.abs().max()) > 1e-6 * gmax:
            rels.append((float(d.norm() / (want.norm() + 1e-300)), n))
    rels.sort()
    rel_all, median = math.sqrt(num / den), rels[len(rels) // 2][0]
    print(f"\n{precision} @ C2 12L+6L, B={BATCH}, T'=299: loss {float(loss):.6f} (ctc {parts[1]:.4f} att {parts[2]:.4f}) vs the oracle "
          f"{ref['loss']:.6f} (ctc {ref['ctc']:.4f} att {ref['att']:.4f}); gradient rel-L2 {rel_all:.3e}, median per-parameter "
          f"{median:.3e}, worst {rels[-1]}, max-abs {worst_abs[0]:.3e} in {worst_abs[1]} (max|g| {gmax:.3e})")
    return float(loss), rel_all, median, worst_abs[0] / gmax


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_bench_shape_step_matches_float64_oracle(oracle_f64, precision):
    loss, rel_all, median, worst = _run_product(oracle_f64, precision, 126)
    ref = oracle_f64[2]
    if precision == "fp32":
        assert math.isclose(loss, ref["loss"], rel_tol=1e-5), (loss, ref["loss"])
        assert worst <= 1e-4, worst
        assert rel_all <= 1e-4, rel_all
    else:
        assert math.isclose(loss, ref["loss"], rel_tol=3e-3), (loss, ref["loss"])
        assert rel_all <= 3e-2, rel_all
        assert median <= 2e-2, median


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_batch_252_step_matches_float32_oracle(oracle_f32_b252, precision):
    """The bench's per-GPU batch (252 x 299 = 75 348 rows = four waves of row tiles; the parity planes of the front end hold
    1.55e9 elements, 72 % of the int32 range): a float64 oracle of this batch does not fit the GPU, so the oracle runs in
    float32 here -- an index-arithmetic guard at the full size (an overflow gives garbage, not rounding noise), with tolerances
    that absorb the float32 oracle's own rounding: fp32 mode loss rel 1e-4, whole-gradient rel-L2 1e-3; bf16 mode as above."""
    loss, rel_all, median, worst = _run_product(oracle_f32_b252, precision, 252)
    ref = oracle_f32_b252[2]
    if precision == "fp32":
        assert math.isclose(loss, ref["loss"], rel_tol=1e-4), (loss, ref["loss"])
        assert rel_all <= 1e-3, rel_all
    else:
        assert math.isclose(loss, ref["loss"], rel_tol=3e-3), (loss, ref["loss"])
        assert rel_all <= 3e-2, rel_all
        assert median <= 2e-2, median
