"""Developer tool (GPU): clock64 timeline of the fused feed-forward backward kernel (CTA 0, first 32 chunks).
Columns per chunk (cycles relative to the first stamp): MMA thread: W2 landed, acc1 buffer free, slab written (all warps);
epilogue warp: acc1 ready, TMEM read done, math done, slab free, slab written."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from liteasr_b200 import _lib, ops  # noqa: E402

m, d, f = 37674, 256, 2048
dev = "cuda"
dy = (torch.randn(m, d, device=dev) * 0.5).bfloat16()
w2 = (torch.randn(d, f, device=dev) * 0.05).bfloat16()
w1 = (torch.randn(f, d, device=dev) * 0.05).bfloat16()
gd = torch.rand(m, f, device=dev).bfloat16()
dh = torch.empty(m, f, device=dev, dtype=torch.bfloat16)
dln = torch.empty(m, d, device=dev, dtype=torch.bfloat16)
cs = torch.zeros(f, device=dev)
for _ in range(2):
    ops.ffn_bwd(dy, gd, w2, w1, dh, dln, colsum=cs, alpha=0.5)
torch.cuda.synchronize()
buf = torch.zeros(32 * 8 + 2 * 16 * 8, dtype=torch.int64, device=dev)
_lib.lib().lasr_ffn_bwd_set_trace(C.c_void_p(buf.data_ptr()))
ops.ffn_bwd(dy, gd, w2, w1, dh, dln, colsum=cs, alpha=0.5)
torch.cuda.synchronize()
_lib.lib().lasr_ffn_bwd_set_trace(C.c_void_p(0))
full = buf.cpu()
t = full[:256].view(32, 8)
t0 = int(t[t > 0].min())
names = ["M:W2full", "M:acc1free", "M:slabfull", "E:acc1full", "E:ld done", "E:math", "E:slabfree", "E:slabdone"]
print("chunk " + " ".join(f"{n:>11s}" for n in names))
for c in range(32):
    print(f"{c:5d} " + " ".join(f"{int(x) - t0:11d}" for x in t[c]))

w = full[256:].view(2, 16, 8)
if int(w.max()) > 0:
    print("per-warp view of chunks 8 and 9 (acc1 ready, TMEM read done, math done, slab free, slab written + arrived, column sums done):")
    for ci in range(2):
        for wi in range(16):
            print(f"chunk {8 + ci} warp {wi + 3:2d} (q={(wi + 3) & 3} part={wi >> 2}): " + " ".join(f"{int(x) - t0:8d}" for x in w[ci, wi, :6]))
