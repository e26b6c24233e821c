"""Developer tool (GPU): per-launch time of small GEMMs inside a CUDA graph (what the decoder's 5166-row launches cost)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from liteasr_b200 import ops  # noqa: E402
from tools.gemm_bench import run, rnd  # noqa: E402

dev, bf = "cuda", torch.bfloat16
for (m, n, k) in [(128, 64, 64), (128, 256, 256), (5166, 256, 256), (5166, 768, 256), (5166, 2048, 256), (5166, 256, 2048), (37674, 256, 256)]:
    a, w = rnd(m, k), rnd(n, k)
    c = torch.empty(m, n, device=dev, dtype=bf)
    c32 = torch.empty(m, n, device=dev)
    res = torch.randn(m, n, device=dev)
    bias = torch.randn(n, device=dev)
    run(f"plain bf16 {m}x{n}x{k}", lambda: ops.gemm(a, w, c, m, n, k, lda=k, ldb=k, ldc=n, bias=bias), 2.0 * m * n * k, (m * k + n * k + m * n) * 2)
    run(f"res fp32   {m}x{n}x{k}", lambda: ops.gemm(a, w, c32, m, n, k, lda=k, ldb=k, ldc=n, bias=bias, res=res, ldres=n), 2.0 * m * n * k, (m * k + n * k) * 2 + m * n * 8)
x = torch.randn(5166, 256, device=dev)
y = torch.empty(5166, 256, device=dev, dtype=bf)
g, b = torch.ones(256, device=dev), torch.zeros(256, device=dev)
mean, rstd = torch.empty(5166, device=dev), torch.empty(5166, device=dev)
run("layernorm_fwd 5166x256", lambda: ops.layernorm_fwd(x, g, b, y, mean, rstd), 0.0, 5166 * 256 * 6)
