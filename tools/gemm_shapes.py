"""Developer tool (GPU): every tcgen05 GEMM of one training step grouped by shape signature; each group's launches are replayed
back to back from a CUDA graph on the real operands and timed with CUDA events (no host gaps).

    python tools/gemm_shapes.py [workload] [batch]
"""
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from liteasr_b200 import ops  # noqa: E402
from liteasr_b200.criterions.hybrid_ctc_attn import HybridCTCLoss, HybridCTCLossConfig  # noqa: E402
from liteasr_b200.models.u2 import U2, U2Config  # noqa: E402
from liteasr_b200.schema import U2Dims  # noqa: E402
from liteasr_b200.trainer import TrainStep  # noqa: E402
from liteasr_b200.utils.synthetic import synth_batch  # noqa: E402


def sig(a, b, c, m, n, k, kw):
    bt = kw.get("batch", (1, 1))
    return (f"m={m} n={n} k={k} b={bt[0] * bt[1]} ta={int(kw.get('ta', False))} tb={int(kw.get('tb', False))} sk={kw.get('split_k', 1)} "
            f"c={str(c.dtype)[6:]} aux={int(kw.get('aux') is not None)} res={int(kw.get('res') is not None)} act={kw.get('act', 0)} "
            f"bias={int(kw.get('bias') is not None)} acc={int(kw.get('accumulate', False))} dact={int(kw.get('dact') is not None)} "
            f"cs={int(kw.get('colsum') is not None)}")


def main(workload="c2", batch="0"):
    wl = dict(bench.WORKLOADS[workload])
    if int(batch) > 0:
        wl["batch"] = int(batch)
    dims = U2Dims(*wl["dims"])
    dev = torch.device("cuda:0")
    model = U2(U2Config(**dims.__dict__, precision="bf16")).to(dev).train()
    crit = HybridCTCLoss(HybridCTCLossConfig(vocab_size=dims.vocab_size, smoothing=wl["smoothing"], ctc_weight=wl["ctc_weight"]))
    step = TrainStep(model, crit, use_graph=False, device=dev)
    b = tuple(t.to(dev) for t in synth_batch(wl["batch"], wl["tmax"], wl["lmax"], dims.vocab_size, seed=42))
    step.step_eager(*b)
    rec = []
    orig = ops.gemm

    def recording(a, b_, c, m, n, k, **kw):
        rec.append((a, b_, c, m, n, k, kw))
        orig(a, b_, c, m, n, k, **kw)

    ops.gemm = recording
    step.step_eager(*b)
    torch.cuda.synchronize()
    ops.gemm = orig
    groups = OrderedDict()
    for r in rec:
        groups.setdefault(sig(*r), []).append(r)
    rows = []
    for s, calls in groups.items():
        def replay():
            for a, b_, c, m, n, k, kw in calls:
                orig(a, b_, c, m, n, k, **kw)
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
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        a, b_, c, m, n, k, kw = calls[0]
        bt = kw.get("batch", (1, 1))
        fl = 2.0 * m * n * k * bt[0] * bt[1]
        by = (m * k + n * k) * 2 * bt[0] * bt[1] + m * n * bt[0] * bt[1] * c.element_size() * (2 if kw.get("aux") is not None else 1)
        if kw.get("res") is not None:
            by += m * n * 4
        if kw.get("dact") is not None:
            by += m * n * 2
        rows.append((ms, len(calls), s, fl * len(calls) / ms / 1e9, by * len(calls) / ms / 1e6))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"workload {workload} batch {wl['batch']}: {len(rec)} GEMM launches, {tot:.3f} ms replayed group by group")
    for ms, n, s, tf, gb in rows:
        print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}% {n:4d}x {1e3 * ms / n:8.1f} us  {tf:7.1f} TF/s {gb:7.0f} GB/s  {s}")


if __name__ == "__main__":
    main(*sys.argv[1:])
