"""Developer tool (GPU): per-shape time of every lasr_gemm launch in one eager training step of a bench workload."""
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


def main(workload="c2", precision="bf16"):
    wl = bench.WORKLOADS[workload]
    dims = U2Dims(*wl["dims"])
    dev = torch.device("cuda:0")
    model = U2(U2Config(**dims.__dict__, precision=precision)).to(dev).train()
    crit = HybridCTCLoss(HybridCTCLossConfig(vocab_size=dims.vocab_size, smoothing=wl["smoothing"], ctc_weight=wl["ctc_weight"]))
    step = TrainStep(model, crit, use_graph=False, device=dev)
    batch = tuple(t.to(dev) for t in synth_batch(wl["batch"], wl["tmax"], wl["lmax"], dims.vocab_size, seed=42))
    for _ in range(2):
        step.step_eager(*batch)
    torch.cuda.synchronize()
    rec = []
    orig = ops.gemm

    def timed(a, b, c, m, n, k, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(a, b, c, m, n, k, **kw)
        e1.record()
        bt = kw.get("batch", (1, 1))
        rec.append(((m, n, k, bt[0] * bt[1], int(kw.get("ta", False)), int(kw.get("tb", False)), kw.get("split_k", 1),
                     str(c.dtype)[6:], int(kw.get("aux") is not None), int(kw.get("res") is not None)), e0, e1))

    ops.gemm = timed
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_all0.record()
    step.step_eager(*batch)
    e_all1.record()
    torch.cuda.synchronize()
    ops.gemm = orig
    agg = defaultdict(lambda: [0, 0.0])
    for sig, e0, e1 in rec:
        agg[sig][0] += 1
        agg[sig][1] += e0.elapsed_time(e1)
    tot = sum(v[1] for v in agg.values())
    print(f"eager step {e_all0.elapsed_time(e_all1):.2f} ms, gemm total {tot:.2f} ms over {len(rec)} launches")
    print(f"{'m':>7} {'n':>6} {'k':>7} {'batch':>5} ta tb sk {'cdt':>8} aux res {'cnt':>4} {'ms':>8} {'us/call':>8} {'TFLOP/s':>8}")
    for sig, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        m, n, k, bt, ta, tb, sk, cdt, aux, res = sig
        fl = 2.0 * m * n * k * bt * cnt
        print(f"{m:7d} {n:6d} {k:7d} {bt:5d} {ta:2d} {tb:2d} {sk:2d} {cdt:>8} {aux:3d} {res:3d} {cnt:4d} {ms:8.3f} {ms / cnt * 1e3:8.1f} {fl / ms / 1e9:8.1f}")


if __name__ == "__main__":
    main(*sys.argv[1:])
