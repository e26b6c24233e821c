#!/usr/bin/env python
"""Benchmark of the hot path: one U2 Conformer training step (fwd + bwd + hybrid CTC/attention loss + gradient all-reduce +
fused clip/Adam) on synthetic 80-dim fbank batches.  Metric: audio-seconds per second (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c1|c3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One JSON line on rank 0.  `value` = whole-job audio-s/s with the batch resident in HBM (CUDA-graph replay, CUDA-event timed,
max over ranks); `e2e` = the same step through the public API with the batch in pinned HOST memory (H2D inside the timed
region, loss read back every step).  `roofline` = the dominant kernel family (tcgen05 GEMM) timed live with CUDA events (the family's
launches of one step replayed from a CUDA graph); `cpu_baseline` = the CPU oracle port of the reference's path on the box's host cores (bounded
sample).  `--impl reference` runs only that CPU arm and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME_SHIFT_S = 0.01  # Kaldi fbank default frame shift (the reference never states it)

WORKLOADS = {
    # BASELINE.json configs[1]: AISHELL-1 shape
    "c2": dict(dims=(80, 4233, 256, 2048, 4, 12, 256, 2048, 4, 6), batch=126, tmax=1200, lmax=40, ctc_weight=0.3, smoothing=0.1,
               desc="C2: U2 Conformer 12L d256 H4 f2048 + 6L Transformer decoder, V=4233 (AISHELL-1 shape), Tmax=1200 (T'=299), "
                    "per-GPU batch 126 (builder's choice, SURVEY 8: 126 x 299 rows = 295 row tiles = two full waves of 148 SMs; "
                    "the reference default 32 is `--batch 32`), Lmax=40, hybrid ctc_weight 0.3, smoothing 0.1, dropout 0 (U2Config default)"),
    # configs[0]: the reference's CPU-runnable case
    "c1": dict(dims=(80, 500, 256, 2048, 4, 4, 256, 2048, 4, 6), batch=8, tmax=500, lmax=30, ctc_weight=0.3, smoothing=0.1,
               desc="C1: U2 Conformer 4L d256 H4 + 6L decoder, V=500, batch 8, Tmax=500"),
    # configs[2]: LibriSpeech-960 shape
    "c3": dict(dims=(80, 5000, 512, 2048, 8, 12, 512, 2048, 8, 6), batch=16, tmax=1600, lmax=100, ctc_weight=0.3, smoothing=0.1,
               desc="C3: Conformer-large 12L d512 H8 + 6L decoder d512, V=5000, Tmax=1600 (T'=399), per-GPU batch 16, Lmax=100"),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tc_burst=p["bf16_tflops"], tc_sustained=p["bf16_tflops_sustained"], src="measured")
    except Exception:  # noqa: BLE001
        return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_sm = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv is not None:
            try:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.05)

    def result(self):
        s = sorted(self.sm)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_sm, reasons=sorted(self.reasons), samples=len(s))


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's path, on the host cores, on a bounded sample
# --------------------------------------------------------------------------------------------------
def cpu_oracle_run(wl, steps: int, warmup: int, sample_batch: int = 2):
    import torch
    from liteasr_b200.schema import U2Dims
    from liteasr_b200.utils.synthetic import synth_batch, synth_state_dict
    from oracle import u2_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O.build_c_oracle()
    dims = U2Dims(*wl["dims"])
    xs, xlens, ys, ylens = synth_batch(sample_batch, wl["tmax"], wl["lmax"], dims.vocab_size, seed=42)
    sd = synth_state_dict(dims, seed=42)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and ".pe.pe" not in k else v)
          for k, v in sd.items()}
    cfg = O.U2Shape(**dims.__dict__)
    params = [v for v in sd.values() if getattr(v, "requires_grad", False)]
    opt = torch.optim.Adam(params, lr=1e-4)

    def one():
        opt.zero_grad(set_to_none=True)
        out = O.hybrid_loss(sd, cfg, xs, xlens, ys, ylens, wl["ctc_weight"], wl["smoothing"], True, {})
        out["loss"].backward()
        torch.nn.utils.clip_grad_norm_(params, 5.0)
        opt.step()
        return float(out["loss"])

    for _ in range(max(0, warmup)):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    audio = float(xlens.sum()) * FRAME_SHIFT_S
    return dict(value=audio / dt, ms_per_step=dt * 1e3, cores=cores, threads=torch.get_num_threads(),
                sample=f"{sample_batch} of {wl['batch']} utterances of the same workload (Tmax={wl['tmax']}), fwd+bwd+clip+Adam, "
                       f"{steps} timed steps after {warmup} warm-up, torch {torch.__version__} CPU fp32")


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 10))
    r = cpu_oracle_run(wl, steps, max(1, min(args.warmup, 2)), sample_batch=8)
    line = {
        "impl": "reference", "metric": "audio-sec/sec train (Conformer fwd+bwd+CTC)", "value": r["value"], "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "parallelism": "cpu", "frame_shift_ms": 10},
        "cpu_baseline": {"value": r["value"], "unit": "audio-s/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def gemm_family_replay(step_fn, batch, reps: int = 5):
    """The dominant kernel family (every tcgen05 GEMM of one training step) timed live and alone: one eager step records each
    `ops.gemm` call (operands stay alive), a CUDA graph replays exactly those launches back to back on the real operands, CUDA
    events time `reps` replays.  No host launch gaps and no other kernels inside the timed region, so
    flops / time is the family's own tensor-pipe throughput.  -> (flops per step, ms per step, launches per step)."""
    import torch
    from liteasr_b200 import ops
    rec = []
    orig = ops.gemm
    conv_names = ("conv2_fwd", "conv2_dgrad", "conv2_wgrad")  # the implicit-GEMM convolutions run on the same kernel
    conv_orig = {n: getattr(ops, n) for n in conv_names}

    def recording(a, b, c, m, n, k, **kw):
        bt = kw.get("batch", (1, 1))
        rec.append((orig, (a, b, c, m, n, k), kw, 2.0 * m * n * k * bt[0] * bt[1]))
        orig(a, b, c, m, n, k, **kw)

    def conv_recorder(name):
        fn = conv_orig[name]

        def wrapped(*args):
            B, T, F = args[-3:]
            d = {"conv2_fwd": args[0], "conv2_dgrad": args[2], "conv2_wgrad": args[1]}[name].shape[-1]  # the h1p operand
            _, _, _, _, T2, F2 = ops.plane_dims(T, F)
            rec.append((fn, args, {}, 2.0 * B * T2 * F2 * d * 9 * d))
            fn(*args)
        return wrapped

    ops.gemm = recording
    for n in conv_names:
        setattr(ops, n, conv_recorder(n))
    try:
        step_fn.step_eager(*batch)
        torch.cuda.synchronize()
    finally:
        ops.gemm = orig
        for n in conv_names:
            setattr(ops, n, conv_orig[n])
    flops = sum(r[3] for r in rec)

    def replay():
        for fn, args, kw, _ in rec:
            fn(*args, **kw)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        replay()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        replay()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return flops, e0.elapsed_time(e1) / reps, len(rec)


def ctc_standalone(pk):
    """BASELINE config 4 (largest point): standalone fused CTC fwd+bwd, algorithmic GB/s = T*B*V*(4+4) bytes / time."""
    import torch
    from liteasr_b200 import ops
    T, B, V, L = 1600, 64, 5000, 200
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(T, B, V, generator=g, device="cuda")
    il = torch.randint(int(0.6 * T), T + 1, (B,), generator=g, device="cuda"); il[0] = T
    tl = torch.randint(L // 2, L + 1, (B,), generator=g, device="cuda"); tl[0] = L
    tg = torch.randint(1, V, (B, L), generator=g, device="cuda")
    grad = torch.empty_like(x)
    ws = torch.empty(ops.ctc_workspace_bytes(T, B, L), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ops.ctc_fwdbwd(x, tg, il, tl, time_major=True, grad=grad, workspace=ws)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        ops.ctc_fwdbwd(x, tg, il, tl, time_major=True, grad=grad, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gbs = T * B * V * 8 / (ms * 1e-3) / 1e9
    return {"workload": f"standalone CTC fwd+bwd B={B} T={T} V={V} L={L} fp32 (2.05 GB logits > L2)", "ms": ms, "bound": "hbm",
            "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"], "traffic": None}


def run_gpu(args, wl):
    import torch
    import torch.distributed as dist
    from liteasr_b200 import _lib
    from liteasr_b200.criterions.hybrid_ctc_attn import HybridCTCLoss, HybridCTCLossConfig
    from liteasr_b200.distributed.utils import distributed_init
    from liteasr_b200.models.u2 import U2, U2Config
    from liteasr_b200.optims import FusedNoam, NoamConfig
    from liteasr_b200.schema import U2Dims
    from liteasr_b200.trainer import TrainStep
    from liteasr_b200.utils.synthetic import synth_batch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (liteasr_b200 has no CPU fallback); use --impl reference for the CPU arm")
    local = distributed_init("nccl")
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    pk = peaks()
    dims = U2Dims(*wl["dims"])
    torch.manual_seed(42)
    rates = {}
    if args.dropout > 0:  # config/model/my_U2.yaml: one rate everywhere, the three attention-probability rates 0.0
        rates = dict(dropout_rate=args.dropout, enc_attn_dropout_rate=0.0, dec_self_attn_dropout_rate=0.0, dec_src_attn_dropout_rate=0.0)
    model = U2(U2Config(**dims.__dict__, precision=args.precision, **rates)).to(dev).train()
    crit = HybridCTCLoss(HybridCTCLossConfig(vocab_size=dims.vocab_size, smoothing=wl["smoothing"], ctc_weight=wl["ctc_weight"]))
    step = TrainStep(model, crit, None, clip_grad_norm=5.0, use_graph=not args.no_graph, device=dev)
    step.optimizer = FusedNoam(step.store, NoamConfig(model_dim=dims.enc_dim))
    host = synth_batch(wl["batch"], wl["tmax"], wl["lmax"], dims.vocab_size, seed=42 + rank)
    host = tuple(t.pin_memory() for t in host)
    batch = tuple(t.to(dev) for t in host)
    audio_local = float(host[1].sum()) * FRAME_SHIFT_S

    # launches per step (eager, before capture) -- counts kernels launched by liblasr only
    step.step_eager(*batch)
    torch.cuda.synchronize()
    c0 = _lib.launch_count()
    step.step_eager(*batch)
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - c0

    static = step.static_inputs(*batch) if step.use_graph else batch
    for _ in range(max(3, args.warmup)):
        loss = step(*static)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    def timed(fn, k):
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms = e0.elapsed_time(e1) / k
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(lambda: step(*static), args.steps)

    # end-to-end: pinned host batch -> H2D -> step through the public TrainStep call -> loss read back on the host
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    from liteasr_b200.trainer import Prefetcher
    pf = Prefetcher(dev)
    pf.put(host)

    def e2e_step():
        # the public training API: Prefetcher (pinned host -> device on a copy stream, one step ahead) + TrainStep.__call__
        # (device-to-device into the graph's static inputs, graph replay); every step starts one H2D copy of a full batch
        # and reads the loss back on the host
        dev_batch = pf.get()
        pf.put(host)
        l = step(*dev_batch)
        loss_host.copy_(l, non_blocking=False)  # D2H + host sync: the loss is read every step

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    audio = audio_local
    if world > 1:
        t = torch.tensor([audio_local], device=dev, dtype=torch.float64)
        dist.all_reduce(t)
        audio = float(t)
    h2d = sum(t.numel() * t.element_size() for t in host)

    # roofline of the dominant kernel family (tcgen05 GEMM), measured live with CUDA events
    flops, gemm_ms, n_gemm = gemm_family_replay(step, batch)
    tflops = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    roof = {"kernel": "gemm_tc_kernel (tcgen05.mma bf16, all GEMMs of one step)" if args.precision == "bf16" else "gemm_simt_kernel",
            "bound": "tensor", "achieved": tflops, "peak": pk["tc_sustained"], "unit": "TFLOP/s", "frac": tflops / pk["tc_sustained"],
            "traffic": None, "peak_source": pk["src"] + " (sustained bf16: the family's launches of one step replayed back to back from a CUDA graph)",
            "launches": n_gemm, "kernel_ms_per_step": gemm_ms, "share_of_step": gemm_ms / ms, "algorithmic_tflop_per_step": flops / 1e12}

    tp = (((wl["tmax"] - 3) // 2 + 1) - 3) // 2 + 1
    ctc_logit_mb = wl["batch"] * tp * ((dims.vocab_size + 7) // 8 * 8) * 2 // 1000000
    line = {
        "metric": "audio-sec/sec train (Conformer fwd+bwd+CTC)", "value": audio / (ms * 1e-3), "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "global_batch": wl["batch"] * world, "audio_s_per_step": audio, "parallelism": f"dp{world}",
                   "frame_shift_ms": 10, "cuda_graph": bool(step.use_graph),
                   "l2": f"per-step working set (several GB of activations, {ctc_logit_mb} MB of CTC logits) exceeds the 126 MB L2; no explicit flush"},
        "clocks": sampler.result(),
        "e2e": {"value": audio / (ms_e2e * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches_per_step * args.steps),
        "gpu_launches_per_step": int(launches_per_step),
        "roofline": roof,
        "loss": float(loss),
    }
    if rank == 0 and world == 1:
        try:
            line["roofline_ctc"] = ctc_standalone(pk)
        except Exception as e:  # noqa: BLE001
            line["roofline_ctc"] = {"error": str(e)}
        if not args.no_cpu_baseline:
            r = cpu_oracle_run(wl, 3, 1, sample_batch=8)
            line["cpu_baseline"] = {"value": r["value"], "unit": "audio-s/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
                                    "ms_per_step": r["ms_per_step"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # the captured graph holds NCCL kernels: tearing the communicator down under it can block, so leave without destructors
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (0 = the workload's default)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.0, help="model.dropout_rate (the reference's my_U2.yaml trains with 0.1; attention rates stay 0.0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.dropout > 0:
        wl["desc"] = wl["desc"].replace("dropout 0 (U2Config default)", f"dropout {args.dropout} (my_U2.yaml rates, own Philox stream)")
    if args.batch > 0:
        wl["desc"] = wl["desc"].replace(f"per-GPU batch {wl['batch']}", f"per-GPU batch {args.batch}").replace(f"batch {wl['batch']},", f"batch {args.batch},")
        wl["batch"] = args.batch
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_gpu(args, wl)


if __name__ == "__main__":
    main()
