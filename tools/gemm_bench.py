"""Developer tool (GPU): time individual lasr_gemm shapes of the C2 step (CUDA events, L2-flushing rotation of buffers).

    python tools/gemm_bench.py [case ...]      cases: fc1 fc2 qkv scores pv wgrad dgrad out all
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from liteasr_b200 import ops  # noqa: E402

dev = "cuda"
bf, f32 = torch.bfloat16, torch.float32


def rnd(*shape, dtype=bf):
    return (torch.randn(*shape, device=dev) * 0.1).to(dtype)


USE_GRAPH = True


def run(name, fn, flops, nbytes, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if USE_GRAPH:  # GPU time only: the Python/ctypes launch path costs ~15 us per call, more than the small GEMMs
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if USE_GRAPH:
        g.replay()
    else:
        for _ in range(iters):
            fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"{name:34s} {us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s  {nbytes / us / 1e3:8.1f} GB/s (algorithmic)", flush=True)


def main(cases):
    N, d, f = 9568, 256, 2048
    allc = not cases or "all" in cases
    x = rnd(N, d)
    w1, b1 = rnd(f, d), torch.randn(f, device=dev)
    w2, b2 = rnd(d, f), torch.randn(d, device=dev)
    a = rnd(N, f)
    res = torch.randn(N, d, device=dev)
    if allc or "fc1" in cases:
        out, aux = torch.empty(N, f, device=dev, dtype=bf), torch.empty(N, f, device=dev, dtype=bf)
        run("fc1 swish+aux 9568x2048x256", lambda: ops.linear(x, w1, out, bias=b1, aux=aux, act=ops.ACT_SWISH), 2 * N * d * f, N * d * 2 + 2 * N * f * 2)
        run("fc1 bias only 9568x2048x256", lambda: ops.linear(x, w1, out, bias=b1), 2 * N * d * f, N * d * 2 + N * f * 2)
        run("fc1 plain     9568x2048x256", lambda: ops.linear(x, w1, out), 2 * N * d * f, N * d * 2 + N * f * 2)
    if allc or "fc2" in cases:
        o32 = torch.empty(N, d, device=dev)
        run("fc2 res f32   9568x256x2048", lambda: ops.linear(a, w2, o32, bias=b2, res=res, alpha=0.5), 2 * N * d * f, N * f * 2 + 2 * N * d * 4)
        run("fc2 plain f32 9568x256x2048", lambda: ops.linear(a, w2, o32), 2 * N * d * f, N * f * 2 + N * d * 4)
    if allc or "out" in cases:
        wo = rnd(d, d)
        o32 = torch.empty(N, d, device=dev)
        run("out res f32   9568x256x256", lambda: ops.linear(x, wo, o32, bias=b2, res=res), 2 * N * d * d, N * d * 2 + 2 * N * d * 4)
        run("out plain f32 9568x256x256", lambda: ops.linear(x, wo, o32), 2 * N * d * d, N * d * 2 + N * d * 4)
    if allc or "qkv" in cases:
        wq, bq = rnd(3 * d, d), torch.randn(3 * d, device=dev)
        o = torch.empty(N, 3 * d, device=dev, dtype=bf)
        run("qkv bias bf16 9568x768x256", lambda: ops.linear(x, wq, o, bias=bq), 2 * N * d * 3 * d, N * d * 2 + N * 3 * d * 2)
    if allc or "scores" in cases:
        B, H, T, dk, ld = 32, 4, 299, 64, 304
        q, k = rnd(B * T, d), rnd(B * T, d)
        ac = torch.empty(B, H, T, ld, device=dev)
        run("scores f32 128x(299x299x64)", lambda: ops.gemm(q, k, ac, T, T, dk, lda=d, ldb=d, ldc=ld, batch=(B, H), sa=(T * d, dk), sb=(T * d, dk),
                                                           sc=(H * T * ld, T * ld)), 2 * B * H * T * T * dk, 2 * N * d * 2 + B * H * T * ld * 4)
    if allc or "pv" in cases:
        B, H, T, dk, ld = 32, 4, 299, 64, 304
        pr, v = rnd(B, H, T, ld), rnd(B * T, d)
        o = torch.empty(B * T, d, device=dev, dtype=bf)
        run("p.v bf16 128x(299x64x299)", lambda: ops.gemm(pr, v, o, T, dk, T, lda=ld, ldb=d, ldc=d, tb=True, batch=(B, H), sa=(H * T * ld, T * ld),
                                                         sb=(T * d, dk), sc=(T * d, dk)), 2 * B * H * T * T * dk, B * H * T * ld * 2 + 2 * N * d * 2)
    if allc or "dgrad" in cases:
        dy = rnd(N, d)
        o = torch.empty(N, f, device=dev, dtype=bf)
        run("dgrad tb bf16 9568x2048x256", lambda: ops.gemm(dy, w2, o, N, f, d, lda=d, ldb=f, ldc=f, tb=True), 2 * N * d * f, N * d * 2 + N * f * 2)
    if allc or "dswish" in cases:
        dy = rnd(N, d)
        h = rnd(N, f)
        o = torch.empty(N, f, device=dev, dtype=bf)
        cs = torch.zeros(f, device=dev)
        run("dgrad dswish+colsum 9568x2048x256", lambda: ops.gemm(dy, w2, o, N, f, d, lda=d, ldb=f, ldc=f, tb=True, alpha=0.5, dact=h, act=ops.ACT_SWISH, colsum=cs),
            2 * N * d * f, N * d * 2 + 2 * N * f * 2)
        run("dgrad dswish        9568x2048x256", lambda: ops.gemm(dy, w2, o, N, f, d, lda=d, ldb=f, ldc=f, tb=True, alpha=0.5, dact=h, act=ops.ACT_SWISH),
            2 * N * d * f, N * d * 2 + 2 * N * f * 2)
        run("dgrad colsum        9568x2048x256", lambda: ops.gemm(dy, w2, o, N, f, d, lda=d, ldb=f, ldc=f, tb=True, colsum=cs),
            2 * N * d * f, N * d * 2 + N * f * 2)
    if allc or "scores" in cases:
        B, H, T, dk, ld = 32, 4, 299, 64, 304
        q, k = rnd(B * T, d), rnd(B * T, d)
        ac = torch.empty(B, H, T, ld, device=dev)
        run("scores f32 n_store=304", lambda: ops.gemm(q, k, ac, T, T, dk, lda=d, ldb=d, ldc=ld, batch=(B, H), sa=(T * d, dk), sb=(T * d, dk),
                                                      sc=(H * T * ld, T * ld), n_store=ld), 2 * B * H * T * T * dk, 2 * N * d * 2 + B * H * T * ld * 4)
    if allc or "wgrad" in cases:
        dy = rnd(N, f)
        gw = torch.zeros(f, d, device=dev)
        run("wgrad sk4 f32 2048x256x9568", lambda: ops.gemm(dy, x, gw, f, d, N, lda=f, ldb=d, ldc=d, ta=True, tb=True, accumulate=True, split_k=4),
            2 * N * d * f, N * f * 2 + N * d * 2 + f * d * 4)
        dy2 = rnd(N, d)
        gw2 = torch.zeros(d, d, device=dev)
        run("wgrad sk32 f32 256x256x9568", lambda: ops.gemm(dy2, x, gw2, d, d, N, lda=d, ldb=d, ldc=d, ta=True, tb=True, accumulate=True, split_k=32),
            2 * N * d * d, 2 * N * d * 2 + d * d * 4)


if __name__ == "__main__":
    args = sys.argv[1:]
    if "--eager" in args:
        USE_GRAPH = False
        args.remove("--eager")
    main(args)
