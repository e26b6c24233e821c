"""Developer tool (GPU): the dP = dO . V^T batched GEMM of the attention backward (m = n = 299, k = 64, 504 units, bf16 out)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from liteasr_b200 import ops  # noqa: E402
from tools.gemm_bench import run  # noqa: E402

dev, bf = "cuda", torch.bfloat16
B, H, T, dk, ld = 126, 4, 299, 64, 320
d = H * dk
do = (torch.randn(B * T, d, device=dev) * 0.1).to(bf)
qkv = (torch.randn(B * T, 3 * d, device=dev) * 0.1).to(bf)
v = qkv[:, 2 * d:]
dp = [torch.empty(B, H, T, ld, device=dev, dtype=bf) for _ in range(2)]
bs = (H * T * ld, T * ld)
it = {"i": 0}


def f():
    i = it["i"] % 2
    it["i"] += 1
    ops.gemm(do, v, dp[i], T, T, dk, lda=d, ldb=v.stride(0), ldc=ld, batch=(B, H), sa=(T * d, dk), sb=(T * v.stride(0), dk), sc=bs, n_store=ld)


if len(sys.argv) > 1 and sys.argv[1] == "once":
    for _ in range(4):
        f()
    torch.cuda.synchronize()
else:
    run("dP gemm 504 x (299x299x64) bf16", f, 2.0 * B * H * T * T * dk, B * H * T * ld * 2 + 2 * B * T * d * 2)
