"""Developer tool (GPU): is the weight-gradient GEMM (MN-major A and B, split-K, red.add epilogue) slower than the same contraction
with K-major operands?  python tools/wgrad_probe.py [K] [m] [n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from liteasr_b200 import ops  # noqa: E402
from tools.gemm_bench import run  # noqa: E402

dev, bf = "cuda", torch.bfloat16


def main(K=37674, m=256, n=2048):
    K, m, n = int(K), int(m), int(n)
    nb = 3  # rotate buffers past the L2
    dy = [(torch.randn(K, m, device=dev) * 0.1).to(bf) for _ in range(nb)]
    x = [(torch.randn(K, n, device=dev) * 0.1).to(bf) for _ in range(nb)]
    Kp = (K + 7) // 8 * 8
    dyt = [torch.zeros(m, Kp, device=dev, dtype=bf) for _ in range(nb)]
    xt = [torch.zeros(n, Kp, device=dev, dtype=bf) for _ in range(nb)]
    for i in range(nb):
        dyt[i][:, :K] = dy[i].t()
        xt[i][:, :K] = x[i].t()
    gw = torch.zeros(m, n, device=dev)
    it = {"i": 0}
    fl, byt = 2.0 * m * n * K, (m + n) * K * 2 + m * n * 4

    def mn(sk):
        def f():
            i = it["i"] % nb
            it["i"] += 1
            ops.gemm(dy[i], x[i], gw, m, n, K, lda=m, ldb=n, ldc=n, ta=True, tb=True, accumulate=True, split_k=sk)
        return f

    def km(sk):
        def f():
            i = it["i"] % nb
            it["i"] += 1
            ops.gemm(dyt[i], xt[i], gw, m, n, K, lda=Kp, ldb=Kp, ldc=n, accumulate=True, split_k=sk)
        return f

    for sk in (9, 18, 37):
        run(f"wgrad MN-major {m}x{n}x{K} sk={sk}", mn(sk), fl, byt)
        run(f"same, K-major operands   sk={sk}", km(sk), fl, byt)
    # correctness of the comparison
    gw.zero_(); mn(9)(); a = gw.clone(); gw.zero_(); km(9)(); torch.cuda.synchronize()
    print("max |MN - K-major| =", float((a - gw).abs().max()), "of", float(a.abs().max()))


if __name__ == "__main__":
    main(*sys.argv[1:])
