"""Developer tool (GPU): the CTA-pair weight-gradient kernel (lasr_wgrad2) against lasr_gemm's split-K wgrad: result and time."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from liteasr_b200 import ops  # noqa: E402
from tools.gemm_bench import run  # noqa: E402

dev, bf = "cuda", torch.bfloat16


def main(K=37674, m=256, n=2048, quick="0"):
    K, m, n = int(K), int(m), int(n)
    nb = 3
    dy = [(torch.randn(K, m, device=dev) * 0.1).to(bf) for _ in range(nb)]
    x = [(torch.randn(K, n, device=dev) * 0.1).to(bf) for _ in range(nb)]
    ref = dy[0].float().t() @ x[0].float()
    tiles = (m // 256) * (n // 256)
    for sk in sorted({max(1, 74 // tiles), max(1, 148 // tiles), max(1, 37 // tiles)}):
        gw = torch.zeros(m, n, device=dev)
        ops.wgrad2(dy[0], x[0], gw, alpha=1.0, split_k=sk)
        torch.cuda.synchronize()
        err = float((gw - ref).abs().max()) / float(ref.abs().max())
        print(f"wgrad2 {m}x{n}x{K} sk={sk}: max rel err {err:.2e}", flush=True)
        assert err < 2e-3, err
    if int(quick):
        return
    gw = torch.zeros(m, n, device=dev)
    it = {"i": 0}
    fl, byt = 2.0 * m * n * K, (m + n) * K * 2 + m * n * 4

    def pair(sk):
        def f():
            i = it["i"] % nb
            it["i"] += 1
            ops.wgrad2(dy[i], x[i], gw, split_k=sk)
        return f

    def single(sk):
        def f():
            i = it["i"] % nb
            it["i"] += 1
            ops.gemm(dy[i], x[i], gw, m, n, K, lda=m, ldb=n, ldc=n, ta=True, tb=True, accumulate=True, split_k=sk)
        return f

    for sk in sorted({max(1, 74 // tiles), max(1, 148 // tiles)}):
        run(f"wgrad2 (CTA pairs) {m}x{n}x{K} sk={sk}", pair(sk), fl, byt)
    for sk in sorted({max(1, 74 // tiles), max(1, 148 // (2 * tiles))}):
        run(f"lasr_gemm wgrad     {m}x{n}x{K} sk={sk}", single(sk), fl, byt)


if __name__ == "__main__":
    main(*sys.argv[1:])
