"""Developer tool (GPU): the fused feed-forward backward kernel against the two GEMMs it replaces, at the bench shape.

    python tools/ffn_bench.py [M] [d] [f]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from liteasr_b200 import ops  # noqa: E402


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main(m=37674, d=256, f=2048, colsum="1", only=""):
    m, d, f = int(m), int(d), int(f)
    dev = "cuda"
    dy = (torch.randn(m, d, device=dev) * 0.5).bfloat16()
    w2 = (torch.randn(d, f, device=dev) * 0.05).bfloat16()
    w1 = (torch.randn(f, d, device=dev) * 0.05).bfloat16()
    gd = torch.rand(m, f, device=dev).bfloat16()
    # rotate over enough buffer sets to exceed the 126 MB L2 like the real step does
    n = 3
    gs = [gd.clone() for _ in range(n)]
    dhs = [torch.empty(m, f, device=dev, dtype=torch.bfloat16) for _ in range(n)]
    dlns = [torch.empty(m, d, device=dev, dtype=torch.bfloat16) for _ in range(n)]
    cs = torch.zeros(f, device=dev) if int(colsum) else None
    it = {"i": 0}
    if only == "fused":  # one launch sequence for ncu
        for i in range(3):
            ops.ffn_bwd(dy, gs[i], w2, w1, dhs[i], dlns[i], colsum=cs, alpha=0.5)
        torch.cuda.synchronize()
        return

    def fused():
        i = it["i"] % n
        it["i"] += 1
        ops.ffn_bwd(dy, gs[i], w2, w1, dhs[i], dlns[i], colsum=cs, alpha=0.5)

    def unfused():
        i = it["i"] % n
        it["i"] += 1
        ops.gemm(dy, w2, dhs[i], m, f, d, lda=d, ldb=f, ldc=f, tb=True, alpha=0.5, dact=gs[i], act=ops.ACT_MUL, colsum=cs)
        ops.gemm(dhs[i], w1, dlns[i], m, d, f, lda=f, ldb=d, ldc=d, tb=True)

    fl = 2 * 2.0 * m * d * f
    by = (m * d * 2) * 2 + m * f * 2 * 2
    tf = timed(fused)
    tu = timed(unfused)
    print(f"M={m} d={d} f={f}: fused {tf:8.1f} us ({fl / tf / 1e6:6.1f} TFLOP/s, {by / tf / 1e3:6.0f} GB/s algorithmic)   "
          f"unfused pair {tu:8.1f} us ({fl / tu / 1e6:6.1f} TFLOP/s)")


if __name__ == "__main__":
    main(*sys.argv[1:])
