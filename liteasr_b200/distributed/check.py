"""Numeric check of the multi-GPU exchange step on real devices (one process per GPU, NCCL), SURVEY 8e / VERDICT r1 item 2b.

``ddp_numeric_check(device)`` runs under ``torch.distributed`` with any world size (1 included) and verifies, on a small U2 in
fp32 parity mode:
  * after ``FlatDDP.finish_backward`` the flat gradient buffer times the returned multiplier (1/world) equals the MEAN of the
    single-GPU gradients of every rank's batch -- each rank recomputes all ``world`` single-GPU gradients itself with a
    non-distributed copy of the model, so no second collective is involved in the yardstick (the reference semantics:
    torch DDP averages, trainer.py:76-88);
  * BatchNorm running statistics that a rank > 0 has perturbed equal rank 0's after ``broadcast_buffers``
    (``DistributedDataParallel(broadcast_buffers=True)`` default, distributed/ddp_model_wrapper.py:8-57);
  * every bucket was launched exactly once and the buckets tile the buffer.
It returns a dict (``ok``, errors, bucket count) that ``bench.py`` attaches to its JSON line at N > 1."""
from __future__ import annotations

import torch
import torch.distributed as dist


def ddp_numeric_check(device, precision: str = "fp32") -> dict:
    from ..criterions.hybrid_ctc_attn import HybridCTCLoss, HybridCTCLossConfig
    from ..models.u2 import U2, U2Config
    from ..schema import U2Dims
    from ..utils.synthetic import synth_batch, synth_state_dict
    from .. import functions as F
    from .flat_ddp import FlatDDP

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    dims = U2Dims(input_dim=80, vocab_size=200, enc_dim=128, enc_ff_dim=256, enc_attn_heads=2, enc_layers=2, dec_dim=128,
                  dec_ff_dim=256, dec_attn_heads=2, dec_layers=1)
    sd = synth_state_dict(dims, seed=7)
    crit = HybridCTCLoss(HybridCTCLossConfig(vocab_size=dims.vocab_size, smoothing=0.1, ctc_weight=0.3))

    def fresh():
        m = U2(U2Config(**dims.__dict__, precision=precision))
        m.load_state_dict(sd)
        return m.to(device).train()

    def batch_of(r):
        return tuple(t.to(device) for t in synth_batch(6, 260, 12, dims.vocab_size, seed=500 + r))

    # yardstick: single-GPU gradients of every rank's batch, computed locally (BatchNorm statistics are per rank in the
    # reference -- plain BatchNorm1d -- so the per-rank gradient is exactly the single-GPU gradient of that rank's batch)
    mean = None
    for r in range(world):
        m = fresh()
        st, _, _ = F.bind(m, device)
        st.enable_direct_grads()
        st.zero_grads()
        crit.direct_step(m, *batch_of(r))
        g = st.gflat.double().clone()
        mean = g if mean is None else mean + g
    mean /= world

    model = fresh()
    st, _, _ = F.bind(model, device)
    st.enable_direct_grads()
    ddp = FlatDDP(model, st, bucket_bytes=256 << 10)  # small buckets: several all-reduces overlap the backward
    # BatchNorm buffers: every rank > 0 perturbs its running statistics, the broadcast must restore rank 0's
    ref_bn = ddp.bn_flat.clone() if ddp.bn_flat is not None else None
    if rank > 0 and ddp.bn_flat is not None:
        ddp.bn_flat.add_(float(rank))
    st.zero_grads()
    ddp.broadcast_buffers()
    bn_err = float((ddp.bn_flat - ref_bn).abs().max()) if ref_bn is not None else 0.0
    ddp.sync_grads = True
    ddp.begin_backward()
    crit.direct_step(model, *batch_of(rank))
    mult = ddp.finish_backward()
    torch.cuda.synchronize(device)
    got = st.gflat.double() * mult
    scale = float(mean.abs().max())
    err = float((got - mean).abs().max())
    cover = sorted(ddp.launched)
    tiled = (world == 1) or (bool(cover) and cover[0][0] == 0 and cover[-1][1] == st.numel and
                             all(a[1] == b[0] for a, b in zip(cover, cover[1:])))
    # fp32 parity mode: split-K partial sums land through red.add in a run-dependent order and the ring all-reduce sums the ranks
    # in its own order: a few fp32 ulp of the largest gradient entry
    tol = (2e-5 if precision == "fp32" else 2e-2) * scale
    res = {"world": world, "max_abs_err": err, "grad_max": scale, "tolerance": tol, "bn_buffer_err": bn_err,
           "buckets": len(cover), "buckets_tile_buffer": bool(tiled), "mult": mult,
           "ok": bool(err <= tol and bn_err == 0.0 and tiled and abs(mult - 1.0 / world) < 1e-12)}
    if dist.is_initialized() and world > 1:
        flag = torch.tensor([1.0 if res["ok"] else 0.0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        res["ok_all_ranks"] = bool(flag.item() == 1.0)
    st.grad_ready_hook = None
    return res


def ddp_varshape_check(device, steps: int = 6) -> dict:
    """ranks that see DIFFERENT batch shapes capture their CUDA graphs at different steps.  Every rank runs
    `steps` optimizer steps of a small model through ``TrainStep`` (graphs on, capture at the second occurrence of a shape) with
    a rank-dependent shape schedule, so that in most steps some ranks replay, some capture and some run eagerly; the run must
    neither hang nor diverge: after every step the parameters of all ranks are identical (same all-reduced gradients)."""
    from ..criterions.hybrid_ctc_attn import HybridCTCLoss, HybridCTCLossConfig
    from ..models.u2 import U2, U2Config
    from ..schema import U2Dims
    from ..trainer import TrainStep
    from ..utils.synthetic import synth_batch, synth_state_dict

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    dims = U2Dims(input_dim=80, vocab_size=200, enc_dim=128, enc_ff_dim=256, enc_attn_heads=2, enc_layers=2, dec_dim=128,
                  dec_ff_dim=256, dec_attn_heads=2, dec_layers=1)
    model = U2(U2Config(**dims.__dict__, precision="bf16"))
    model.load_state_dict(synth_state_dict(dims, seed=7))
    model = model.to(device).train()
    crit = HybridCTCLoss(HybridCTCLossConfig(vocab_size=dims.vocab_size, smoothing=0.1, ctc_weight=0.3))
    step = TrainStep(model, crit, device=device, use_graph=True, graph_min_hits=2, max_graphs=2)
    max_dev = 0.0
    for i in range(steps):
        # rank r alternates between two shapes with period r + 2: the ranks repeat (and therefore capture) at different steps
        tmax = 200 + 16 * ((i % (rank + 2)) == 0) + 8 * rank
        batch = tuple(t.to(device) for t in synth_batch(4, tmax, 10, dims.vocab_size, seed=900 + 17 * i + rank))
        step(*batch)
        torch.cuda.synchronize(device)
        if world > 1:
            flat = step.store.flat
            ref = flat.clone()
            dist.broadcast(ref, src=0)
            max_dev = max(max_dev, float((flat - ref).abs().max()))
    res = {"world": world, "steps": steps, "max_param_deviation_from_rank0": max_dev, "stats": dict(step.stats), "ok": bool(max_dev == 0.0)}
    if dist.is_initialized() and world > 1:
        flag = torch.tensor([1.0 if res["ok"] else 0.0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        res["ok_all_ranks"] = bool(flag.item() == 1.0)
    return res
