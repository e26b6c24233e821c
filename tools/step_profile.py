"""Developer tool (GPU): CUDA-event time of every ``ops.*`` call in one eager training step, aggregated by (op, signature).

    python tools/step_profile.py [workload] [batch] [top]
"""
import os
import sys
from collections import defaultdict

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

SKIP = {"dtype_code", "ctc_workspace_bytes", "linear"}


def sig_of(name, args, kw):
    if name == "gemm":
        a, b, c, m, n, k = args[:6]
        bt = kw.get("batch", (1, 1))
        return (f"m={m} n={n} k={k} b={bt[0] * bt[1]} ta={int(kw.get('ta', False))} tb={int(kw.get('tb', False))} sk={kw.get('split_k', 1)} "
                f"c={str(c.dtype)[6:]} aux={int(kw.get('aux') is not None)} res={int(kw.get('res') is not None)} act={kw.get('act', 0)} "
                f"bias={int(kw.get('bias') is not None)} acc={int(kw.get('accumulate', False))}"), 2.0 * m * n * k * bt[0] * bt[1]
    parts = []
    for a in list(args) + list(kw.values()):
        if isinstance(a, torch.Tensor) and a.dim() > 0:
            parts.append("x".join(map(str, a.shape)) + ":" + str(a.dtype)[6:9])
            if len(parts) == 2:
                break
    return " ".join(parts), 0.0


def main(workload="c2", batch="0", top="60", dropout="0"):
    wl = dict(bench.WORKLOADS[workload])
    if int(batch) > 0:
        wl["batch"] = int(batch)
    dims = U2Dims(*wl["dims"])
    dev = torch.device("cuda:0")
    model = U2(U2Config(**dims.__dict__, precision="bf16", **bench.my_u2_rates(float(dropout)))).to(dev).train()
    crit = HybridCTCLoss(HybridCTCLossConfig(vocab_size=dims.vocab_size, smoothing=wl["smoothing"], ctc_weight=wl["ctc_weight"]))
    step = TrainStep(model, crit, use_graph=False, device=dev)
    b = tuple(t.to(dev) for t in synth_batch(wl["batch"], wl["tmax"], wl["lmax"], dims.vocab_size, seed=42))
    for _ in range(2):
        step.step_eager(*b)
    torch.cuda.synchronize()
    rec = []
    saved = {}
    for name in dir(ops):
        fn = getattr(ops, name)
        if name.startswith("_") or name in SKIP or not callable(fn) or getattr(fn, "__module__", "") != ops.__name__:
            continue
        saved[name] = fn

        def make(name, fn):
            def timed(*args, **kw):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn(*args, **kw)
                e1.record()
                s, fl = sig_of(name, args, kw)
                rec.append((name, s, fl, e0, e1))
                return r
            return timed
        setattr(ops, name, make(name, fn))
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    step.step_eager(*b)
    eb.record()
    torch.cuda.synchronize()
    for name, fn in saved.items():
        setattr(ops, name, fn)
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    byop = defaultdict(lambda: [0, 0.0])
    for name, s, fl, e0, e1 in rec:
        ms = e0.elapsed_time(e1)
        a = agg[(name, s)]
        a[0] += 1; a[1] += ms; a[2] += fl
        byop[name][0] += 1; byop[name][1] += ms
    tot = sum(v[1] for v in byop.values())
    print(f"workload {workload} batch {wl['batch']}: eager step {ea.elapsed_time(eb):.2f} ms, sum of op times {tot:.2f} ms over {len(rec)} calls")
    for name, (cnt, ms) in sorted(byop.items(), key=lambda kv: -kv[1][1]):
        print(f"  {name:24s} {cnt:5d} calls {ms:8.3f} ms {100 * ms / tot:5.1f}%")
    print()
    for (name, s), (cnt, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(top)]:
        tf = f"{fl / ms / 1e9:7.1f} TF/s" if fl else ""
        print(f"{ms:8.3f} ms {cnt:4d}x {ms / cnt * 1e3:8.1f} us  {name:20s} {s} {tf}")


if __name__ == "__main__":
    main(*sys.argv[1:])
