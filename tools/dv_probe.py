"""Developer tool (GPU): the dV = P^T . dO batched GEMM of the attention backward with and without its bias-gradient column sums
(1512 same-address atomics per column and launch)."""
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
nb = 2
P = [(torch.rand(B, H, T, ld, device=dev) * 0.01).to(bf) for _ in range(nb)]
do = [(torch.randn(B * T, d, device=dev) * 0.1).to(bf) for _ in range(nb)]
dqkv = torch.empty(B * T, 3 * d, device=dev, dtype=bf)
dv = dqkv[:, 2 * d:]
bv = torch.zeros(d, device=dev)
bs = (H * T * ld, T * ld)
it = {"i": 0}


def f(cs):
    def g():
        i = it["i"] % nb
        it["i"] += 1
        ops.gemm(P[i], do[i], dv, T, dk, T, lda=ld, ldb=d, ldc=dv.stride(0), ta=True, tb=True, batch=(B, H), sa=bs, sb=(T * d, dk),
                 sc=(T * dv.stride(0), dk), colsum=(bv if cs else None), cs=(0, dk))
    return g


fl, byt = 2.0 * B * H * T * T * dk, B * H * T * ld * 2 + 2 * B * T * d * 2
run("dV gemm with colsum", f(True), fl, byt)
run("dV gemm without colsum", f(False), fl, byt)
run("dV gemm with colsum", f(True), fl, byt)
run("dV gemm without colsum", f(False), fl, byt)
